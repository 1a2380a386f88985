// fasta.h — indexed FASTA -> G/C bit mask per contig (rsq_reference), for --fasta runs.
//
// Replaces the reference's Fasta class on this path (src/Fasta.h:44-70, src/Fasta.cpp:77-197, bioio.hpp:291-331): the loop
// only ever asks for gc(getSeq(...)) (src/Expression.cpp:473-475, src/Metrics.cpp:300-302), so the sequence is reduced once
// to one bit per base and no page cache is needed.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "annotation.h"
#include "rsq_b200.h"

namespace rsqh {

struct FaiEntry { std::string name; uint64_t length = 0, offset = 0, line_len = 0, line_bytes = 0; };

struct ReferenceModel {
    std::string fasta_path;
    std::vector<FaiEntry> entries;                 // .fai order
    std::vector<uint32_t> entry_contig;            // contig id of each entry
    // per contig id (sized by finish())
    std::vector<uint8_t> present;
    std::vector<uint64_t> length, word_off;
    std::vector<uint32_t> bits;
    bool open = false;

    rsq_reference view() const;
    // gc(getSeq(contig, start, end)) with 0-based end-exclusive coordinates (src/Fasta.cpp:66-75,116-150): the window is
    // clipped to the contig; -1 for an empty sequence.  The sum is the reference's: one 1.0/size per G/C base.
    double gc_content(uint32_t contig, int64_t start, int64_t end) const;
};

// Fasta::open (src/Fasta.cpp:77-99): locates the index (path with the extension replaced by .fai, else path + ".fai"),
// registers every contig name IN INDEX ORDER in the shared contig map — before the GTF is parsed, as main() does
// (src/RNASeQC.cpp:117-122), which fixes the chromosome id order — and keeps the index.  HostError(10) when the FASTA or
// its index cannot be opened or the index is empty.
void open_reference(const std::string& fasta_path, ContigMap& contigs, ReferenceModel& ref);
// Reads the sequences and builds the bit mask over an index space of n_contigs ids.
void load_reference_bits(ReferenceModel& ref, size_t n_contigs);

}  // namespace rsqh
