// annotation.h — host-side annotation model: GTF/BED text -> flattened interval arrays for the device.
//
// Re-states, quirk for quirk, what the reference builds while parsing (src/GTF.cpp:30-143,
// src/BED.cpp:18-43, src/RNASeQC.cpp:108-187) but stores it as integer-indexed SoA arrays instead of
// string-keyed maps and a list<Feature> per chromosome.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "rsq_b200.h"

namespace rsqh {

// Mirrors the reference's exception -> exit-code table (src/RNASeQC.cpp:678-764).
struct HostError : public std::runtime_error {
    int exit_code;
    HostError(int code, const std::string& msg) : std::runtime_error(msg), exit_code(code) {}
};

// First-seen contig naming, shared by GTF, BED and BAM header (chromosomeMap, src/Fasta.cpp:17-25).
struct ContigMap {
    std::vector<std::string> names;
    std::unordered_map<std::string, uint32_t> index;
    uint32_t get(const std::string& name) {
        auto it = index.find(name);
        if (it != index.end()) return it->second;
        uint32_t id = (uint32_t)names.size();
        names.push_back(name);
        index.emplace(name, id);
        return id;
    }
    int find(const std::string& name) const {
        auto it = index.find(name);
        return it == index.end() ? -1 : (int)it->second;
    }
};

struct GeneFeature {      // one kept line of type "gene" (src/RNASeQC.cpp:137-139)
    uint32_t contig;
    int64_t start, end;
    uint8_t strand;       // RSQ_STRAND_*
    bool ribosomal;
    uint32_t gene;        // gene-universe index of feature_id
    uint32_t line_order;  // position among kept lines (ties of the stable sort)
    uint32_t gid = 0;     // gene-universe index of the line's gene_id (== gene unless the attribute was carried over)
    uint32_t flat = 0;    // index of this gene line in the rsq_transcripts request (contig order, then start order)
};

struct ExonFeature {      // one kept line of type "exon"
    uint32_t contig;
    int64_t start, end;
    uint8_t strand;
    bool ribosomal;
    uint32_t gene;        // gene-universe index of gene_id (src/GTF.cpp:90, carried over when absent)
    uint32_t slot;        // index into exon_ids (exonList order, src/GTF.cpp:106)
    uint32_t line_order;
};

struct Annotation {
    ContigMap contigs;                       // id space shared with BED and BAM header names
    // gene universe: every distinct string used as a gene key (gene feature_id or exon gene_id)
    std::vector<std::string> gene_ids;
    std::unordered_map<std::string, uint32_t> gene_index;
    std::vector<uint32_t> gene_list;         // geneList (src/GTF.cpp:87): universe indices in GTF order
    std::vector<std::string> gene_desc;      // per universe gene: geneNames[id] ("" when absent)
    std::vector<int64_t> gene_coding_len;    // geneCodingLengths (src/GTF.cpp:107), per universe gene
    std::vector<uint8_t> gene_flags;         // RSQ_GENE_GLOBIN
    // exons in exonList order
    std::vector<std::string> exon_ids;
    std::vector<std::string> exon_desc;      // geneNames[exon_id]
    std::vector<int64_t> exon_len;           // exonLengths[id].length (src/GTF.cpp:108)
    std::vector<int64_t> exon_start;         // exonLengths[id].start
    std::vector<uint32_t> exon_contig;       // exonLengths[id].chromosome
    std::vector<uint64_t> exon_cov_off;      // per slot: first base in the coverage array
    uint64_t cov_len = 0;
    // kept features per contig, stable-sorted by start (src/RNASeQC.cpp:150-152)
    std::vector<std::vector<GeneFeature>> genes_by_contig;
    std::vector<std::vector<ExonFeature>> exons_by_contig;
    // exonsForGene (src/RNASeQC.cpp:153-154): per universe gene, exon slots in stitched order
    std::vector<std::vector<uint32_t>> exons_for_gene;

    // flattened device arrays (what rsq_annotation points at)
    std::vector<uint32_t> f_exon_off, f_exon_gene, f_exon_slot, f_seg_off;
    std::vector<int32_t> f_exon_start, f_exon_end, f_seg_start;
    std::vector<uint8_t> f_exon_flags, f_seg_flags;
    std::vector<uint64_t> f_exon_cov_off;
    // gene lines as intervals + list-order ranks (read by the --legacy rules only)
    std::vector<uint32_t> f_gline_off, f_gline_gene, f_gline_order, f_exon_order, f_gline_fid;
    std::vector<int32_t> f_gline_start, f_gline_end;
    std::vector<uint8_t> f_gline_flags;

    uint32_t n_feature_contigs() const { return (uint32_t)genes_by_contig.size(); }
    // Extends the contig index space (BED-only contigs) with empty slices.
    void pad_contigs(size_t n_contigs);
    rsq_annotation view() const;
    // request for rsq_coverage_statistics: one entry per kept gene line (see GeneFeature::flat)
    std::vector<uint64_t> t_gene_cov_off;
    std::vector<uint32_t> t_gene_exon_off, t_exon_len, t_exon_slot;
    std::vector<uint8_t> t_gene_strand;
    void build_transcripts();
    rsq_transcripts transcripts(unsigned coverage_mask, int bias_offset, int bias_window, unsigned long bias_gene_length) const;

    uint32_t gene_universe(const std::string& id);
};

// Parses a GTF the way operator>>(ifstream&, Feature&) does (src/GTF.cpp:30-143) and builds the model.
// Throws HostError(10) when the file cannot be opened, HostError(11) for the reference's gtfException
// cases and for "no genes or no exons" (src/RNASeQC.cpp:158-164).
// legacy: single-base features are dropped from the interval lists after being recorded in the name lists, and a
// single-base exon takes its base back from geneCodingLengths (src/RNASeQC.cpp:129-135).
void load_gtf(const std::string& path, Annotation& out, bool legacy = false);

struct BedModel {
    std::vector<std::vector<int64_t>> starts, ends;   // per contig id (sized to contigs.names.size())
    std::vector<std::vector<uint32_t>> name_ids;
    std::vector<uint8_t> present;                      // bedFeatures->find(chr) != end
    std::vector<std::string> names;                    // name_id -> whole BED line
    bool was_sorted = true;
    // flattened
    std::vector<uint32_t> f_off, f_name;
    std::vector<int32_t> f_start, f_end;
    rsq_bed view() const;
};

// extractBED (src/BED.cpp:18-43): start+1, end+1, feature_id = the whole line.  Contigs are added to
// `contigs` in first-seen order exactly as chromosomeMap would.
void load_bed(const std::string& path, ContigMap& contigs, BedModel& out);
// (Re)builds the flattened arrays after all contigs are known.
void flatten_bed(BedModel& bed, size_t n_contigs);

// ---- coordinate partitioning across GPUs ----------------------------------------------------------------
// "Islands": per contig, the merged union of (a) the hull of every gene — its gene line and all of its exon lines, wherever
// they lie — and (b) every BED interval.  A record stream may be cut between two islands and nowhere else: all records that
// count toward a gene (every aligned block inside exons of that gene, src/Expression.cpp:363-377), both mates the fragment
// tracker and the two mate maps pair up, and the gene's per-base coverage then belong to one partition.
struct Islands {
    std::vector<std::vector<std::pair<int64_t, int64_t>>> by_contig;   // annotation contig id -> sorted, disjoint [start, end] (1-based closed)
    // first coordinate >= pos1 (1-based) on the contig that lies in no island: a legal cut (records with pos+1 >= cut go right)
    int64_t next_cut(uint32_t contig, int64_t pos1) const;
};
void build_islands(const Annotation& A, const BedModel* bed, Islands& out);
// One cut of a partitioned stream: records at (stream rank of the contig, 1-based start) >= (contig_rank, pos1) lie right of it.
struct StreamCut { int64_t contig_rank; int64_t pos1; };
// rsq_transcripts.gene_owned of partition `rank`: gene line i (the request's order) is owned by the partition its start falls
// into; contig_rank maps an annotation contig id to its rank in the stream's contig order (-1: no records -> partition 0).
void gene_ownership(const Annotation& A, const std::vector<StreamCut>& cuts, const std::vector<int64_t>& contig_rank, uint32_t rank,
                    std::vector<uint8_t>& owned);

}  // namespace rsqh
