// alnreader.h — SAM/BAM/CRAM decode into the device record format (rsq_read + rsq_block).
//
// Own implementation (BGZF via zlib + BAM record parse; SAM text parse): the reference decodes through
// SeqLib::BamReader -> htslib sam_read1 (src/BamReader.cpp:12-27).  Only the fields the hot loop touches
// are extracted.  CRAM: cramreader.cpp.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "annotation.h"
#include "rsq_b200.h"

namespace rsqh {

struct AlnHeader {
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_lens;
};

struct AuxField {
    char tag[2];
    char type;      // BAM type letter: A c C s S i I f d Z H B
    int64_t ival;   // value for integer types; the character for 'A'
};

struct RawAln {
    int32_t tid = -1, pos = -1, mtid = -1, mpos = -1, isize = 0, l_qseq = 0;
    uint16_t flag = 0;
    uint8_t mapq = 0;
    std::string qname;
    std::vector<uint32_t> cigar;   // BAM encoding: len << 4 | op, op index into "MIDNSHP=XB"
    std::vector<AuxField> aux;
};

class AlnReader {
public:
    virtual ~AlnReader() {}
    virtual const AlnHeader& header() const = 0;
    virtual bool next(RawAln& out) = 0;   // false at EOF; throws HostError(-1) on corrupt input
    // throws HostError(10) if unreadable; `fasta`: the reference genome, needed for CRAM input only (NM is regenerated);
    // `threads`: CRAM containers decoded concurrently (records still come out in file order); 0 or 1 = on the caller's thread
    static std::unique_ptr<AlnReader> open(const std::string& path, const std::string& fasta = std::string(), int threads = 0);
};

std::unique_ptr<AlnReader> open_cram(const std::string& path, const std::string& fasta, int threads);   // cramreader.cpp

// 64-bit read-name identity (FNV-1a with a splitmix64 finaliser).
uint64_t hash_qname(const char* s, size_t n);

struct DecodeOptions {
    std::string chimeric_tag = "ch";          // --chimeric-tag (src/RNASeQC.cpp:98)
    std::vector<std::string> filter_tags;     // --tag (src/RNASeQC.cpp:97)
    bool exclude_chimeric = false;
    int chimeric_distance = 2000000;
    int verbosity = 0;
    bool legacy = false;                      // --legacy gates of the loop (src/RNASeQC.cpp:258,276,279,287)
};

// Converts records and carries the order-dependent scalar state of the reference loop that cannot be a
// commutative device reduction: readLength (src/RNASeQC.cpp:277-278), current_chrom / dropFeatures
// (:346-353), the unsorted warning (:354-356), alignmentCount (:245) and the contig visit order that fixes
// the row order of coverage.tsv.
class Decoder {
public:
    Decoder(const AlnHeader& hdr, ContigMap& contigs, size_t n_annotation_contigs, const DecodeOptions& opt);
    // Appends one record (and its blocks) to the batch vectors.  Throws HostError(7) on an unknown CIGAR op
    // (src/Expression.cpp:61-63).
    void convert(const RawAln& aln, std::vector<rsq_read>& reads, std::vector<rsq_block>& blocks);
    // The two halves of convert() for the multi-threaded BAM path (bamparallel.cpp):
    //  - convert_bam_bytes: PURE (thread-safe) conversion of one raw BAM record (the bytes after block_size);
    //    r.block_off is relative to `blocks`; tid receives core.tid
    //  - observe: the order-dependent scalar state, must be called in file order
    void convert_bam_bytes(const uint8_t* rec, size_t len, rsq_read& r, std::vector<rsq_block>& blocks, int32_t& tid) const;
    void observe(rsq_read& r, int32_t tid);

    int read_length() const { return read_length_; }
    uint64_t alignment_count() const { return alignment_count_; }
    // contigs in the order their feature lists were dropped during the stream (dropFeatures calls)
    const std::vector<uint32_t>& dropped_order() const { return dropped_order_; }
    bool any_overlap_with(const std::vector<uint8_t>& contig_has_features) const;
    const std::vector<int>& tid_to_contig() const { return tid_to_contig_; }
    // current_chrom at this point of the stream, and the position of the last read that reached trimFeatures on
    // it (INT32_MIN when none did): fixes which genes of the final contig were evicted before the end-of-run sweep
    int64_t current_contig() const { return current_chrom_; }
    int32_t trim_position() const { return last_metrics_chrom_ == current_chrom_ ? last_position_ : INT32_MIN; }

private:
    DecodeOptions opt_;
    std::vector<int> tid_to_contig_;        // BAM tid -> ContigMap id (always valid: names are added)
    size_t n_annotation_contigs_;
    int read_length_ = 0;
    uint64_t alignment_count_ = 0;
    int64_t current_chrom_ = -1;             // ContigMap id; -1 plays the role of the initial 0
    int32_t last_position_ = 0;
    int64_t last_metrics_chrom_ = -2;
    std::vector<uint8_t> dropped_;           // per ContigMap id
    std::vector<uint32_t> dropped_order_;
    bool warned_unsorted_ = false;
};

// In the header: the multi-threaded ingest runs it once per record on its single consumer thread (bamparallel.cpp), where the
// call and the reloads of the run constants were a third of that thread's time.
inline void Decoder::observe(rsq_read& r, int32_t tid) {
    ++alignment_count_;
    const int n_refs = (int)tid_to_contig_.size();
    const bool tid_ok = tid >= 0 && tid < n_refs;
    const int chr = tid_ok ? tid_to_contig_[tid] : -1;
    // ---- order-dependent scalar state: the same gates as the loop, in the same order -----------------------
    const bool has_ch = r.aux & RSQ_AUX_CHIMERIC_TAG;
    bool reaches_metrics = false;
    do {
        if ((r.flag & RSQ_F_SUPPLEMENTARY) && !(opt_.legacy || has_ch) && opt_.exclude_chimeric) break;   // :258-262
        if (r.flag & (RSQ_F_SECONDARY | RSQ_F_QCFAIL | RSQ_F_SUPPLEMENTARY)) break;               // :263
        if (r.flag & RSQ_F_UNMAP) break;                                                           // :268
        const unsigned alignment_size = (unsigned)(r.endpos - r.pos);                              // :275
        if (opt_.legacy && alignment_size > 100000u) break;                                        // :276 LEGACY_MAX_READ_LENGTH
        if (!read_length_ && tid_ok) current_chrom_ = chr;                                         // :277
        if (alignment_size > (unsigned)read_length_) read_length_ = r.l_qseq;                      // :278
        if (!opt_.legacy && has_ch && opt_.exclude_chimeric) break;                                // :279-283
        if ((r.flag & RSQ_F_PAIRED) && !(r.flag & RSQ_F_MUNMAP)) {                                 // :284-292
            const bool chim = (r.aux & RSQ_AUX_MATE_OTHER_CONTIG) || std::abs(r.pos - r.mpos) > opt_.chimeric_distance ||
                              (opt_.legacy && tid > 127);
            if (chim && opt_.exclude_chimeric) break;
        }
        if (r.filter_mask) break;                                                                  // :319-328
        if (!tid_ok) {                                                                             // :333-337
            if (opt_.verbosity) std::cerr << "Unrecognized RefID on alignment #" << alignment_count_ << std::endl;
            break;
        }
        reaches_metrics = true;
    } while (false);
    if (reaches_metrics) {
        if ((size_t)chr >= dropped_.size()) dropped_.resize(chr + 1, 0);
        if (chr != current_chrom_) {                                                               // :346-353
            if (current_chrom_ >= 0) {
                if ((size_t)current_chrom_ >= dropped_.size()) dropped_.resize(current_chrom_ + 1, 0);
                if (!dropped_[current_chrom_]) dropped_order_.push_back((uint32_t)current_chrom_);
                dropped_[current_chrom_] = 1;
            }
            current_chrom_ = chr;
        } else if (last_position_ > r.pos && !warned_unsorted_) {                                  // :354-355
            std::cerr << "Warning: The input bam does not appear to be sorted. An unsorted bam will yield incorrect results" << std::endl;
            warned_unsorted_ = true;
        }
        last_position_ = r.pos;
        last_metrics_chrom_ = chr;
        if (dropped_[chr]) r.aux |= RSQ_AUX_CONTIG_DROPPED;
    }
}

}  // namespace rsqh
