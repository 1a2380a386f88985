// host_capi.cpp — C entry points of the host-side helper library (librsq_host.so) for the Python binding:
// GTF/BED -> flattened annotation, SAM/BAM -> record arrays + stream-order scalars, results -> report files.
// Pure host code; it never touches the hot path (that is librsq_b200.so behind include/rsq_b200.h).
#include <cstring>
#include <memory>
#include <string>

#include "alnreader.h"
#include "bamparallel.h"
#include "fastinflate.h"
#include "annotation.h"
#include "fasta.h"
#include "pipeline.h"
#include "report.h"

using namespace rsqh;

namespace {
int g_decode_threads = 0;
bool g_legacy = false;
std::string g_fasta;
struct Handle {
    Annotation ann;
    BedModel bed;
    bool have_bed = false;
    rsq_annotation ann_view;
    rsq_bed bed_view;
    ReferenceModel ref;
    rsq_reference ref_view;
    StreamState state;
};
void set_err(char* err, size_t n, const std::string& s) {
    if (err && n) { strncpy(err, s.c_str(), n - 1); err[n - 1] = 0; }
}
}  // namespace

extern "C" {

struct rsqh_run_options {
    const char* sample_name;
    int sample_given;
    const char* output_dir;
    int use_rpkm, write_coverage;
    unsigned coverage_mask;
    int bias_offset, bias_window;
    unsigned long bias_gene_length;
    unsigned detection_threshold;
    const char* const* filter_tags;
    int n_filter_tags;
    int exact_complexity_scan;
};

struct rsqh_stream_state {
    int read_length;
    unsigned long long alignment_count;
    long long final_contig;
    int final_position;
    unsigned n_dropped;
    unsigned dropped[1024];
};

int rsqh_load_annotation(const char* gtf, const char* bed, void** out, char* err, size_t errlen) {
    try {
        std::unique_ptr<Handle> h(new Handle());
        if (!g_fasta.empty()) open_reference(g_fasta, h->ann.contigs, h->ref);   // before the GTF, src/RNASeQC.cpp:117-122
        load_gtf(gtf, h->ann, g_legacy);
        if (bed && *bed) { load_bed(bed, h->ann.contigs, h->bed); h->have_bed = true; }
        const size_t nc = h->ann.contigs.names.size();
        h->ann.pad_contigs(nc);
        if (h->have_bed) { flatten_bed(h->bed, nc); h->bed_view = h->bed.view(); }
        h->ann_view = h->ann.view();
        if (h->ref.open) { load_reference_bits(h->ref, nc); h->ref_view = h->ref.view(); }
        *out = h.release();
        return 0;
    } catch (HostError& e) {
        set_err(err, errlen, e.what());
        return e.exit_code ? e.exit_code : -1;
    } catch (std::exception& e) {
        set_err(err, errlen, e.what());
        return -1;
    }
}
void rsqh_free_annotation(void* h) { delete static_cast<Handle*>(h); }
const rsq_annotation* rsqh_annotation_view(void* h) { return &static_cast<Handle*>(h)->ann_view; }
const rsq_bed* rsqh_bed_view(void* h) { Handle* H = static_cast<Handle*>(h); return H->have_bed ? &H->bed_view : nullptr; }
void rsqh_transcripts(void* h, unsigned mask, int bias_offset, int bias_window, unsigned long bias_gene_length, rsq_transcripts* out) {
    *out = static_cast<Handle*>(h)->ann.transcripts(mask, bias_offset, bias_window, bias_gene_length);
}
// gene lines in the order of the rsq_transcripts request: annotation contig id, 1-based start and end (arrays of n_genes entries)
void rsqh_gene_lines(void* h, uint32_t* contig, int64_t* start, int64_t* end) {
    const rsqh::Annotation& A = static_cast<Handle*>(h)->ann;
    for (size_t c = 0; c < A.genes_by_contig.size(); ++c)
        for (const rsqh::GeneFeature& g : A.genes_by_contig[c]) { contig[g.flat] = (uint32_t)c; start[g.flat] = g.start; end[g.flat] = g.end; }
}
// islands of the loaded annotation (+ BED): call with NULL arrays for the count, then with arrays of that size
size_t rsqh_islands(void* h, uint32_t* contig, int64_t* start, int64_t* end) {
    Handle* H = static_cast<Handle*>(h);
    rsqh::Islands I;
    rsqh::build_islands(H->ann, H->have_bed ? &H->bed : nullptr, I);
    size_t n = 0;
    for (size_t c = 0; c < I.by_contig.size(); ++c)
        for (auto& iv : I.by_contig[c]) {
            if (contig) { contig[n] = (uint32_t)c; start[n] = iv.first; end[n] = iv.second; }
            ++n;
        }
    return n;
}
// rsq_pack_compact (include/rsq_b200.h) with its own scratch; out: 16-byte aligned, rsqh_compact_bound bytes
size_t rsqh_compact_bound(size_t n_reads, size_t n_blocks) { return rsq_compact_bound(n_reads, n_blocks); }
int rsqh_pack_compact(const rsq_read* reads, size_t n_reads, const rsq_block* blocks, size_t n_blocks, void* out, size_t cap, size_t* bytes_out) {
    std::vector<uint32_t> scratch(131072);
    return rsq_pack_compact(reads, n_reads, blocks, n_blocks, out, cap, scratch.data(), bytes_out);
}
unsigned rsqh_n_gene_list(void* h) { return (unsigned)static_cast<Handle*>(h)->ann.gene_list.size(); }
unsigned rsqh_gene_list_index(void* h, unsigned i) { return static_cast<Handle*>(h)->ann.gene_list[i]; }
const char* rsqh_gene_id(void* h, unsigned g) { return static_cast<Handle*>(h)->ann.gene_ids[g].c_str(); }
const char* rsqh_exon_id(void* h, unsigned e) { return static_cast<Handle*>(h)->ann.exon_ids[e].c_str(); }
unsigned rsqh_n_contigs(void* h) { return (unsigned)static_cast<Handle*>(h)->ann.contigs.names.size(); }
const char* rsqh_contig_name(void* h, unsigned c) { return static_cast<Handle*>(h)->ann.contigs.names[c].c_str(); }
int rsqh_contig_index(void* h, const char* name) { return static_cast<Handle*>(h)->ann.contigs.find(name); }

// Decodes a whole SAM/BAM into malloc'd arrays (free with rsqh_free).  Contig indices refer to the annotation
// loaded in `h`; contigs first seen in the alignment header are appended to its name table like chromosomeMap does.
int rsqh_decode_file(void* hv, const char* path, const char* chimeric_tag, const char* const* filter_tags, int n_tags,
                     int exclude_chimeric, int chimeric_distance, rsq_read** reads, size_t* n_reads, rsq_block** blocks,
                     size_t* n_blocks, rsqh_stream_state* st, char* err, size_t errlen) {
    Handle* h = static_cast<Handle*>(hv);
    try {
        DecodeOptions opt;
        if (chimeric_tag) opt.chimeric_tag = chimeric_tag;
        for (int k = 0; k < n_tags; ++k) opt.filter_tags.push_back(filter_tags[k]);
        opt.exclude_chimeric = exclude_chimeric != 0;
        opt.chimeric_distance = chimeric_distance;
        opt.legacy = g_legacy;
        const size_t n_ann_contigs = h->ann.genes_by_contig.size();
        std::vector<rsq_read> R;
        std::vector<rsq_block> B;
        std::unique_ptr<AlnReader> rd;
        std::unique_ptr<ParallelBamDecoder> pb;
        const int threads = g_decode_threads;
        if (threads > 0) {
            FILE* fp = std::fopen(path, "rb");
            unsigned char m[2] = {0, 0};
            if (fp) { if (std::fread(m, 1, 2, fp) != 2) m[0] = 0; std::fclose(fp); }
            if (m[0] == 0x1f && m[1] == 0x8b) pb.reset(new ParallelBamDecoder(path, threads));
        }
        if (!pb) rd = AlnReader::open(path, g_fasta, threads);
        Decoder dec(pb ? pb->header() : rd->header(), h->ann.contigs, n_ann_contigs, opt);
        if (pb) {
            std::vector<rsq_read> r1;
            std::vector<rsq_block> b1;
            while (pb->next_batch(dec, 1u << 20, r1, b1)) {
                const uint32_t base = (uint32_t)B.size();
                for (rsq_read& r : r1) r.block_off += base;
                R.insert(R.end(), r1.begin(), r1.end());
                B.insert(B.end(), b1.begin(), b1.end());
            }
        } else {
            RawAln a;
            while (rd->next(a)) dec.convert(a, R, B);
        }
        *n_reads = R.size();
        *n_blocks = B.size();
        *reads = (rsq_read*)malloc(std::max<size_t>(R.size(), 1) * sizeof(rsq_read));
        *blocks = (rsq_block*)malloc(std::max<size_t>(B.size(), 1) * sizeof(rsq_block));
        if (!R.empty()) memcpy(*reads, R.data(), R.size() * sizeof(rsq_read));
        if (!B.empty()) memcpy(*blocks, B.data(), B.size() * sizeof(rsq_block));
        if (st) {
            st->read_length = dec.read_length();
            st->alignment_count = dec.alignment_count();
            st->final_contig = dec.current_contig();
            st->final_position = dec.trim_position();
            st->n_dropped = (unsigned)std::min<size_t>(dec.dropped_order().size(), 1024);
            for (unsigned k = 0; k < st->n_dropped; ++k) st->dropped[k] = dec.dropped_order()[k];
        }
        return 0;
    } catch (HostError& e) {
        set_err(err, errlen, e.what());
        return e.exit_code ? e.exit_code : -1;
    } catch (std::exception& e) {
        set_err(err, errlen, e.what());
        return -1;
    }
}
void rsqh_free(void* p) { free(p); }
// rsq_pack_records (include/rsq_b200.h) for the Python binding
int rsqh_pack_records(const rsq_read* reads, size_t n_reads, const rsq_block* blocks, size_t n_blocks, rsq_read_packed* out_reads,
                      rsq_block* out_blocks, size_t* n_out_blocks) {
    return rsq_pack_records(reads, n_reads, blocks, n_blocks, out_reads, out_blocks, n_out_blocks);
}
// worker threads for BAM decode in rsqh_decode_file (0 = sequential reader)
void rsqh_set_decode_threads(int n) { g_decode_threads = n; }
// jobs of the last finished multi-threaded BAM decode that were framed on the speculated record chain / by the sequential hop
unsigned long long rsqh_last_ingest_runs_summarised(void) { return rsqh::ParallelBamDecoder::last_runs_summarised(); }
void rsqh_last_ingest_framing(unsigned long long* on_chain, unsigned long long* by_hop) {
    *on_chain = rsqh::ParallelBamDecoder::last_chain_hits();
    *by_hop = rsqh::ParallelBamDecoder::last_chain_misses();
}
// --legacy for the calls that follow: GTF filter (src/RNASeQC.cpp:129-135) and the decoder's loop gates
void rsqh_set_legacy(int on) { g_legacy = on != 0; }
// --fasta for the rsqh_load_annotation calls that follow (NULL or "" = none)
void rsqh_set_fasta(const char* path) { g_fasta = path ? path : ""; }
const rsq_reference* rsqh_reference_view(void* h) { Handle* H = static_cast<Handle*>(h); return H->ref.open ? &H->ref_view : nullptr; }

// Test hook for rsq_pack_compact_segments: the batch cut at the given record indices into pieces with their own (rebased) block arrays;
// with `use_tids` the pieces carry the contig as a reference id (identity table) and a scrambled contig field, as the ingest workers' do.
int rsqh_pack_compact_split(const rsq_read* reads, size_t n_reads, const rsq_block* blocks, size_t n_blocks, const size_t* cuts, size_t n_cuts,
                            int use_tids, void* out, size_t cap, size_t* bytes_out) {
    std::vector<std::vector<rsq_read>> pr;
    std::vector<std::vector<rsq_block>> pb;
    std::vector<std::vector<int32_t>> pt;
    std::vector<rsq_segment> segs;
    int16_t max_contig = 0;
    for (size_t i = 0; i < n_reads; ++i) max_contig = std::max(max_contig, reads[i].contig);
    std::vector<int16_t> table((size_t)max_contig + 1);
    for (size_t k = 0; k < table.size(); ++k) table[k] = (int16_t)k;
    size_t lo = 0;
    for (size_t c = 0; c <= n_cuts; ++c) {
        const size_t hi = c < n_cuts ? cuts[c] : n_reads;
        if (hi < lo || hi > n_reads) return RSQ_ERR_INVALID;
        pr.emplace_back(reads + lo, reads + hi); pb.emplace_back(); pt.emplace_back();
        for (rsq_read& r : pr.back()) {
            if ((size_t)r.block_off + r.n_blocks > n_blocks) return RSQ_ERR_INVALID;
            const uint32_t at = (uint32_t)pb.back().size();
            pb.back().insert(pb.back().end(), blocks + r.block_off, blocks + r.block_off + r.n_blocks);
            r.block_off = at;
            if (use_tids) { pt.back().push_back(r.contig); r.contig = (int16_t)0x5a5a; }      // -1 stays -1: out of the table's range
        }
        lo = hi;
    }
    for (size_t k = 0; k < pr.size(); ++k)
        segs.push_back(rsq_segment{pr[k].data(), use_tids ? pt[k].data() : nullptr, pr[k].size(), pb[k].data(), pb[k].size()});
    std::vector<uint32_t> scratch(131072);
    return rsq_pack_compact_segments(segs.data(), segs.size(), table.data(), table.size(), out, cap, scratch.data(), bytes_out);
}
uint64_t rsqh_hash_qname(const char* s, size_t n) { return hash_qname(s, n); }
// the ingest workers' whole-buffer DEFLATE decoder (fastinflate.h): 1 = inflated exactly out_len bytes, 0 = refused (callers fall back to zlib)
int rsqh_fast_inflate(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) { return rsqh::fast_inflate(in, in_len, out, out_len) ? 1 : 0; }
// hashes of the read names "<prefix><id>" for id = first .. first + n - 1 (tools/name_hash_collisions.py)
void rsqh_hash_qname_series(const char* prefix, uint64_t first, size_t n, uint64_t* out) {
    char buf[96];
    const size_t lp = strnlen(prefix, 64);
    memcpy(buf, prefix, lp);
    for (size_t i = 0; i < n; ++i) {
        const int l = snprintf(buf + lp, sizeof(buf) - lp, "%llu", (unsigned long long)(first + i));
        out[i] = hash_qname(buf, lp + (size_t)l);
    }
}

// Writes the reference's output files from finalised results.
int rsqh_write_reports(void* hv, const rsq_results* res, const double* exon_seq_or_null, const rsq_coverage_stats* cs, const rsqh_run_options* o,
                       const rsqh_stream_state* st, char* err, size_t errlen) {
    Handle* h = static_cast<Handle*>(hv);
    try {
        const Annotation& A = h->ann;
        HostResults R;
        memcpy(R.counters, res->counters, sizeof(R.counters));
        const size_t ng = A.gene_ids.size(), ne = A.exon_ids.size();
        R.gene_reads.assign(res->gene_reads, res->gene_reads + ng);
        R.gene_unique.assign(res->gene_unique, res->gene_unique + ng);
        R.gene_fragments.assign(res->gene_fragments, res->gene_fragments + ng);
        R.exon_whole.assign(res->exon_whole, res->exon_whole + ne);
        R.exon_frac.assign(res->exon_frac, res->exon_frac + ne);
        if (exon_seq_or_null) R.exon_seq.assign(exon_seq_or_null, exon_seq_or_null + ne);
        if (cs) {                       // what rsq_coverage_statistics returned for the request of rsqh_transcripts
            const size_t nl = A.t_gene_strand.size();
            R.cs_mean.assign(cs->gene_mean, cs->gene_mean + nl); R.cs_std.assign(cs->gene_std, cs->gene_std + nl);
            R.cs_cv.assign(cs->gene_cv, cs->gene_cv + nl); R.cs_five.assign(cs->five_end, cs->five_end + nl);
            R.cs_three.assign(cs->three_end, cs->three_end + nl); R.cs_exon_cv.assign(cs->exon_cv, cs->exon_cv + ne);
            R.have_cov_stats = true;
        }
        if (res->fragment_hist) R.fragment_hist.assign(res->fragment_hist, res->fragment_hist + res->fragment_hist_len);
        R.fragment_samples_taken = res->fragment_samples_taken;
        if (res->gc_hist) memcpy(R.gc_hist, res->gc_hist, sizeof(R.gc_hist));
        RunOptions opt;
        opt.sample_name = o->sample_name; opt.sample_given = o->sample_given != 0; opt.output_dir = o->output_dir;
        opt.use_rpkm = o->use_rpkm != 0; opt.write_coverage = o->write_coverage != 0; opt.coverage_mask = o->coverage_mask;
        opt.bias_offset = o->bias_offset; opt.bias_window = o->bias_window; opt.bias_gene_length = o->bias_gene_length;
        opt.detection_threshold = o->detection_threshold; opt.exact_complexity_scan = o->exact_complexity_scan != 0;
        for (int k = 0; k < o->n_filter_tags; ++k) opt.filter_tags.push_back(o->filter_tags[k]);
        StreamState S;
        S.read_length = st->read_length; S.alignment_count = st->alignment_count;
        S.final_contig = st->final_contig; S.final_position = st->final_position;
        S.dropped_order.assign(st->dropped, st->dropped + st->n_dropped);
        write_reports(A, R, opt, S, h->ref.open ? &h->ref : nullptr);
        return 0;
    } catch (HostError& e) {
        set_err(err, errlen, e.what());
        return e.exit_code ? e.exit_code : -1;
    } catch (std::exception& e) {
        set_err(err, errlen, e.what());
        return -1;
    }
}

}  // extern "C"

// ---- BAM writer for synthetic records (bench / tests): BGZF blocks of raw-deflate data, standard BAM records ----
#include <zlib.h>

#include <cstdio>
#include <vector>

namespace {
class BgzfWriter {
public:
    BgzfWriter(const char* path, int level) : level_(level) { fp_ = std::fopen(path, "wb"); }
    bool ok() const { return fp_ != nullptr; }
    void write(const void* p, size_t n) {
        const uint8_t* s = (const uint8_t*)p;
        while (n) {
            size_t k = std::min(n, kBlock - buf_.size());
            buf_.insert(buf_.end(), s, s + k);
            s += k; n -= k;
            if (buf_.size() == kBlock) flush();
        }
    }
    void close() {
        if (!fp_) return;
        if (!buf_.empty()) flush();
        static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        std::fwrite(eof, 1, 28, fp_);
        std::fclose(fp_);
        fp_ = nullptr;
    }
private:
    void flush() {
        std::vector<uint8_t> out(kBlock + 1024);
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        deflateInit2(&zs, level_, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        zs.next_in = buf_.data(); zs.avail_in = (uInt)buf_.size();
        zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
        deflate(&zs, Z_FINISH);
        const size_t clen = zs.total_out;
        deflateEnd(&zs);
        const uint32_t crc = (uint32_t)crc32(crc32(0L, nullptr, 0), buf_.data(), (uInt)buf_.size());
        const uint16_t bsize = (uint16_t)(clen + 25);
        uint8_t h[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (uint8_t)(bsize & 0xff), (uint8_t)(bsize >> 8)};
        std::fwrite(h, 1, 18, fp_);
        std::fwrite(out.data(), 1, clen, fp_);
        const uint32_t isize = (uint32_t)buf_.size();
        std::fwrite(&crc, 4, 1, fp_);
        std::fwrite(&isize, 4, 1, fp_);
        buf_.clear();
    }
    static constexpr size_t kBlock = 0xff00;
    FILE* fp_ = nullptr;
    int level_;
    std::vector<uint8_t> buf_;
};

int reg2bin(int beg, int end) {
    --end;
    if (beg >> 14 == end >> 14) return ((1 << 15) - 1) / 7 + (beg >> 14);
    if (beg >> 17 == end >> 17) return ((1 << 12) - 1) / 7 + (beg >> 17);
    if (beg >> 20 == end >> 20) return ((1 << 9) - 1) / 7 + (beg >> 20);
    if (beg >> 23 == end >> 23) return ((1 << 6) - 1) / 7 + (beg >> 23);
    if (beg >> 26 == end >> 26) return ((1 << 3) - 1) / 7 + (beg >> 26);
    return 0;
}
}  // namespace

extern "C" int rsqh_write_bam(const char* path, int n_ref, const char* const* ref_names, const int64_t* ref_lens, size_t n,
                              const int32_t* tid, const int32_t* pos, const int32_t* mtid, const int32_t* mpos,
                              const int32_t* isize, const int32_t* flag, const int32_t* mapq, const int32_t* nm,
                              const int32_t* aux, const int64_t* name_id, const int32_t* n_blocks, const int32_t* blk_start,
                              const int32_t* blk_len, int stride, const int32_t* lead_soft, const int32_t* ins_len,
                              int l_qseq, int level) {
    BgzfWriter w(path, level);
    if (!w.ok()) return 10;
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (int i = 0; i < n_ref; ++i) text += std::string("@SQ\tSN:") + ref_names[i] + "\tLN:" + std::to_string(ref_lens[i]) + "\n";
    w.write("BAM\1", 4);
    int32_t v = (int32_t)text.size();
    w.write(&v, 4);
    w.write(text.data(), text.size());
    v = n_ref; w.write(&v, 4);
    for (int i = 0; i < n_ref; ++i) {
        v = (int32_t)strlen(ref_names[i]) + 1; w.write(&v, 4);
        w.write(ref_names[i], v);
        v = (int32_t)ref_lens[i]; w.write(&v, 4);
    }
    const std::vector<uint8_t> seq((l_qseq + 1) / 2, 0x11), qual(l_qseq, (uint8_t)('F' - 33));
    std::vector<uint8_t> rec;
    std::vector<uint32_t> cig;
    for (size_t i = 0; i < n; ++i) {
        cig.clear();
        if (lead_soft[i]) cig.push_back(((uint32_t)lead_soft[i] << 4) | 4);
        const int32_t* bs = blk_start + i * stride;
        const int32_t* bl = blk_len + i * stride;
        int ref_end = pos[i];
        for (int j = 0; j < n_blocks[i]; ++j) {
            cig.push_back(((uint32_t)bl[j] << 4) | 0);
            ref_end = bs[j] - 1 + bl[j];
            if (j + 1 < n_blocks[i]) {
                const int gap = bs[j + 1] - (bs[j] + bl[j]);
                if (gap == 0) cig.push_back(((uint32_t)ins_len[i] << 4) | 1);
                else if (gap <= 3) cig.push_back(((uint32_t)gap << 4) | 2);
                else cig.push_back(((uint32_t)gap << 4) | 3);
            }
        }
        char name[32];
        const int l_name = snprintf(name, sizeof(name), "r%lld", (long long)name_id[i]) + 1;
        rec.clear();
        auto put32 = [&](int32_t x) { rec.insert(rec.end(), (uint8_t*)&x, (uint8_t*)&x + 4); };
        auto put16 = [&](uint16_t x) { rec.insert(rec.end(), (uint8_t*)&x, (uint8_t*)&x + 2); };
        put32(tid[i]); put32(pos[i]);
        rec.push_back((uint8_t)l_name); rec.push_back((uint8_t)mapq[i]);
        put16((uint16_t)reg2bin(pos[i], std::max(ref_end, pos[i] + 1)));
        put16((uint16_t)cig.size()); put16((uint16_t)flag[i]);
        put32(l_qseq); put32(mtid[i]); put32(mpos[i]); put32(isize[i]);
        rec.insert(rec.end(), name, name + l_name);
        rec.insert(rec.end(), (uint8_t*)cig.data(), (uint8_t*)cig.data() + 4 * cig.size());
        rec.insert(rec.end(), seq.begin(), seq.end());
        rec.insert(rec.end(), qual.begin(), qual.end());
        if (aux[i] & RSQ_AUX_HAS_NM) { rec.push_back('N'); rec.push_back('M'); rec.push_back('C'); rec.push_back((uint8_t)nm[i]); }
        if (aux[i] & RSQ_AUX_CHIMERIC_TAG) { rec.push_back('c'); rec.push_back('h'); rec.push_back('A'); rec.push_back('1'); }
        const int32_t block_size = (int32_t)rec.size();
        w.write(&block_size, 4);
        w.write(rec.data(), rec.size());
    }
    w.close();
    return 0;
}
