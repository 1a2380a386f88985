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
struct RunSummary {
    size_t n = 0;                        // records in the run
    bool any_read_length_line = false;   // some record gets to src/RNASeQC.cpp:277-278 ...
    bool l_qseq_uniform = true;          // ... and all of those have the same l_qseq
    int32_t l_qseq = 0;
    bool bad_tid = false;                // some record stops at the reference-id check (a message under -v)
    size_t n_reaching = 0;               // records that reach trimFeatures / dropFeatures
    bool single_chr = true, sorted = true;
    int chr = -1;                        // as the BAM reference id (the workers digest jobs before the annotation's contig ids exist)
    int32_t first_pos = 0, last_pos = 0;
};

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
    //    ... and its two parts: everything that needs only the command line and the BAM header (the ingest starts converting while
    //    the GTF is still being parsed: no Decoder exists yet), and the contig id of the annotation (bind_contig, by the consumer)
    static void convert_bam_record(const DecodeOptions& opt, int n_refs, const uint8_t* rec, size_t len, rsq_read& r, std::vector<rsq_block>& blocks, int32_t& tid);
    void bind_contig(rsq_read& r, int32_t tid) const {
        const int chr = (tid >= 0 && tid < (int)tid_to_contig_.size()) ? tid_to_contig_[tid] : -1;
        r.contig = (chr >= 0 && (size_t)chr < n_annotation_contigs_ && chr <= INT16_MAX) ? (int16_t)chr : (int16_t)-1;
    }
    void observe(rsq_read& r, int32_t tid);
    //  - summarise (pure, static: command line + header only) + observe_run: observe() for a whole run of records in O(1) when its
    //    digest allows (see below)
    static RunSummary summarise(const DecodeOptions& opt, int n_refs, const rsq_read* r, const int32_t* tids, size_t n);
    bool observe_run(const RunSummary& s);
    const DecodeOptions& options() const { return opt_; }
    // bind_contig as a table over the BAM reference ids (for rsq_pack_compact_segments, which binds while it packs)
    std::vector<int16_t> contig_of_tid() const {
        std::vector<int16_t> t(tid_to_contig_.size());
        for (size_t k = 0; k < t.size(); ++k) { rsq_read r{}; bind_contig(r, (int32_t)k); t[k] = r.contig; }
        return t;
    }

    bool saw_unsorted() const { return warned_unsorted_; }
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
    static unsigned gates(const DecodeOptions& opt_, const rsq_read& r, int32_t tid, bool tid_ok);
    unsigned gates(const rsq_read& r, int32_t tid, bool tid_ok) const { return gates(opt_, r, tid, tid_ok); }
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

// The gates of the loop in front of the order-dependent state, as a pure function of one record (same order as the loop; none of
// them reads that state): bit 0 = the record gets to the read-length lines (:275-278), bit 1 = it reaches the metrics
// (trimFeatures / dropFeatures, :346-356), bit 2 = it stopped at the reference-id check (:333-337).
inline unsigned Decoder::gates(const DecodeOptions& opt_, const rsq_read& r, int32_t tid, bool tid_ok) {
    const bool has_ch = r.aux & RSQ_AUX_CHIMERIC_TAG;
    if ((r.flag & RSQ_F_SUPPLEMENTARY) && !(opt_.legacy || has_ch) && opt_.exclude_chimeric) return 0;   // :258-262
    if (r.flag & (RSQ_F_SECONDARY | RSQ_F_QCFAIL | RSQ_F_SUPPLEMENTARY)) return 0;               // :263
    if (r.flag & RSQ_F_UNMAP) return 0;                                                           // :268
    const unsigned alignment_size = (unsigned)(r.endpos - r.pos);                                 // :275
    if (opt_.legacy && alignment_size > 100000u) return 0;                                        // :276 LEGACY_MAX_READ_LENGTH
    if (!opt_.legacy && has_ch && opt_.exclude_chimeric) return 1;                                // :279-283
    if ((r.flag & RSQ_F_PAIRED) && !(r.flag & RSQ_F_MUNMAP)) {                                    // :284-292
        const bool chim = (r.aux & RSQ_AUX_MATE_OTHER_CONTIG) || std::abs(r.pos - r.mpos) > opt_.chimeric_distance ||
                          (opt_.legacy && tid > 127);
        if (chim && opt_.exclude_chimeric) return 1;
    }
    if (r.filter_mask) return 1;                                                                  // :319-328
    if (!tid_ok) return 1 | 4;                                                                    // :333-337
    return 1 | 2;
}

// In the header: the multi-threaded ingest runs it on its single consumer thread (bamparallel.cpp).
inline void Decoder::observe(rsq_read& r, int32_t tid) {
    ++alignment_count_;
    const int n_refs = (int)tid_to_contig_.size();
    const bool tid_ok = tid >= 0 && tid < n_refs;
    const int chr = tid_ok ? tid_to_contig_[tid] : -1;
    // ---- order-dependent scalar state: the same gates as the loop, in the same order -----------------------
    const unsigned g = gates(r, tid, tid_ok);
    if (g & 1) {
        const unsigned alignment_size = (unsigned)(r.endpos - r.pos);                              // :275
        if (!read_length_ && tid_ok) current_chrom_ = chr;                                         // :277
        if (alignment_size > (unsigned)read_length_) read_length_ = r.l_qseq;                      // :278
    }
    if ((g & 4) && opt_.verbosity) std::cerr << "Unrecognized RefID on alignment #" << alignment_count_ << std::endl;
    if (g & 2) {
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

// Order-independent digest of a run of consecutive records (pure; the ingest workers compute it per job).
inline RunSummary Decoder::summarise(const DecodeOptions& opt, int n_refs, const rsq_read* r, const int32_t* tids, size_t n) {
    RunSummary s;
    s.n = n;
    for (size_t k = 0; k < n; ++k) {
        const int32_t tid = tids[k];
        const bool tid_ok = tid >= 0 && tid < n_refs;
        const unsigned g = gates(opt, r[k], tid, tid_ok);
        if (g & 1) {
            if (!s.any_read_length_line) { s.any_read_length_line = true; s.l_qseq = r[k].l_qseq; }
            else if (r[k].l_qseq != s.l_qseq) s.l_qseq_uniform = false;
        }
        if (g & 4) s.bad_tid = true;
        if (g & 2) {
            const int chr = tid;                 // two ids of one contig name (a header with a repeated @SQ) only make the digest unusable
            if (!s.n_reaching) { s.chr = chr; s.first_pos = r[k].pos; }
            else { if (chr != s.chr) s.single_chr = false; if (r[k].pos < s.last_pos) s.sorted = false; }
            s.last_pos = r[k].pos;
            ++s.n_reaching;
        }
    }
    return s;
}

// observe() for every record of the run at once, when the digest proves that none of them changes the order-dependent state
// except the position of the last one (and none gets an aux bit): same contig as the stream is on, not a dropped one, positions
// non-decreasing from the last seen, read length already settled at the run's only l_qseq.  false = nothing done, observe() each.
inline bool Decoder::observe_run(const RunSummary& s) {
    if (s.any_read_length_line && !(s.l_qseq_uniform && read_length_ != 0 && s.l_qseq == read_length_)) return false;
    if (s.bad_tid && opt_.verbosity) return false;
    if (s.n_reaching) {
        const int chr = tid_to_contig_[s.chr];                     // a record reaches the metrics only with a reference id in range
        if (!s.single_chr || (int64_t)chr != current_chrom_) return false;
        if ((size_t)chr < dropped_.size() && dropped_[chr]) return false;
        if (!warned_unsorted_ && !(s.sorted && last_position_ <= s.first_pos)) return false;
        if ((size_t)chr >= dropped_.size()) dropped_.resize(chr + 1, 0);
        last_position_ = s.last_pos;
        last_metrics_chrom_ = chr;
    }
    alignment_count_ += s.n;
    return true;
}

}  // namespace rsqh
