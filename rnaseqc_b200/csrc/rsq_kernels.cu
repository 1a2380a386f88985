// rsq_kernels.cu — hand-written sm_100a kernels of the RNA-SeQC per-read hot path.
//
// K_hot  : one pass over a batch of decoded alignment records (rsq_read + rsq_block): the flag/quality/
//          mismatch filter cascade of src/RNASeQC.cpp:254-342, interval intersection of every aligned block
//          with the flattened exon index and gene-segment index (src/Expression.cpp:106-117, src/GTF.cpp:171-186),
//          exon/intron classification and count accumulation (src/Expression.cpp:308-458, src/Metrics.cpp:49-66),
//          per-base coverage as a difference array (src/Metrics.cpp:96-124,257-262), fragment-tracker set
//          (src/Metrics.cpp:22, src/Expression.cpp:383-387) and, in BED mode, fragment-size candidates
//          (src/Expression.cpp:482-516).
// There is no dense contraction anywhere on this path: no tensor cores.  The work is HBM/L2-bound integer work.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "rsq_device.h"

namespace rsqd {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kHotThreads = 256;
constexpr uint32_t kNoGene = 0xffffffffu;

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

// first index in [lo, hi) whose start > key
__device__ __forceinline__ int upper_bound_start(const int2* __restrict__ se, int lo, int hi, int key) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(&se[mid].x) > key) hi = mid; else lo = mid + 1;
    }
    return lo;
}
// first index in [lo, hi) whose running-max end >= key
__device__ __forceinline__ int lower_bound_pmax(const int32_t* __restrict__ pmax, int lo, int hi, int key) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(&pmax[mid]) >= key) hi = mid; else lo = mid + 1;
    }
    return lo;
}
// last index in [lo, hi) whose start <= key (the slice begins with an INT32_MIN sentinel, so it exists)
__device__ __forceinline__ int seg_find(const int32_t* __restrict__ ss, int lo, int hi, int key) {
    int a = lo, b = hi;
    while (a < b) {
        const int mid = (a + b) >> 1;
        if (__ldg(&ss[mid]) > key) b = mid; else a = mid + 1;
    }
    return a - 1;
}

// floor(isz * 2^64 / length) for isz <= length (isz == length gives whole = 1)
__device__ __forceinline__ void frac_q64(uint32_t isz, uint32_t length, unsigned long long& whole, unsigned long long& lo) {
    if (isz >= length) { whole = 1ull; lo = 0ull; return; }
    unsigned long long r = isz;
    const unsigned long long t1 = (r << 32) / length;
    r = (r << 32) % length;
    const unsigned long long t0 = (r << 32) / length;
    whole = 0ull;
    lo = (t1 << 32) | t0;
}

__device__ __forceinline__ bool set_insert(unsigned long long* __restrict__ tab, unsigned long long mask, unsigned long long key,
                                           unsigned int* err, unsigned int err_bit) {
    unsigned long long slot = key & mask;
    for (int probe = 0; probe < 4096; ++probe) {
        unsigned long long cur = tab[slot];
        if (cur == key) return false;
        if (cur == 0ull) {
            cur = atomicCAS(&tab[slot], 0ull, key);
            if (cur == 0ull) return true;
            if (cur == key) return false;
        }
        slot = (slot + 1) & mask;
    }
    atomicOr(err, err_bit);
    return false;
}

struct Window { int lo, hi; };

// Exon window of a block through the direct-address bucket index: two table reads (independent, issued together)
// and two short forward scans replace two ~17-step dependent binary searches.
__device__ __forceinline__ Window exon_window(const AnnView& ann, int c_lo, int c_hi, int bk_lo, int bk_n, int bs, int be) {
    const int b0 = min(max(bs, 0) >> kBucketShift, bk_n - 1), b1 = min(max(be, 0) >> kBucketShift, bk_n - 1);
    int lo = __ldg(&ann.buckets[bk_lo + b0].x);
    int hi = __ldg(&ann.buckets[bk_lo + b1].y);
    while (hi < c_hi && __ldg(&ann.exon_se[hi].x) <= be) ++hi;        // every i < hi has start <= block.end   (Expression.cpp:111)
    while (lo < hi && __ldg(&ann.exon_pmax[lo]) < bs) ++lo;           // nothing before lo can reach block.start (GTF.cpp:171-179)
    Window w; w.lo = lo; w.hi = hi;
    return w;
}
// last gene segment whose start <= bs
__device__ __forceinline__ int seg_lookup(const AnnView& ann, int s_hi, int bk_lo, int bk_n, int bs) {
    const int b0 = min(max(bs, 0) >> kBucketShift, bk_n - 1);
    int k = __ldg(&ann.buckets[bk_lo + b0].z);
    while (k + 1 < s_hi && __ldg(&ann.seg_start[k + 1]) <= bs) ++k;
    return k;
}

// does block (bs, be) lie fully inside an exon of `gene` (strand-admissible)?  src/Expression.cpp:341-344
__device__ __forceinline__ bool block_has_gene(const AnnView& ann, Window w, int bs, int be, uint32_t gene, int rstrand) {
    for (int i = w.lo; i < w.hi; ++i) {
        const int2 x = __ldg(&ann.exon_se[i]);
        if (x.x <= bs && be - 1 <= x.y) {
            const uint32_t gf = __ldg(&ann.exon_meta[i].gene_flags);
            const int fs = (gf >> 24) & RSQ_EXON_STRAND_MASK;
            if (rstrand != RSQ_STRAND_UNKNOWN && rstrand != fs) continue;
            if ((gf & 0xffffffu) == gene) return true;
        }
    }
    return false;
}

template <bool kBed>
__global__ void __launch_bounds__(kHotThreads)
hot_path_kernel(const rsq_read* __restrict__ reads, const rsq_block* __restrict__ blocks, size_t n_reads, size_t n_blocks,
                const AnnView ann, const BedView bed, const Accum acc, const Params prm) {
    __shared__ unsigned long long s_cnt[RSQ_N_COUNTERS];
    __shared__ unsigned long long s_bytes;
    for (int k = threadIdx.x; k < RSQ_N_COUNTERS; k += blockDim.x) s_cnt[k] = 0ull;
    if (threadIdx.x == 0) s_bytes = 0ull;
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    const size_t warps_total = (size_t)gridDim.x * (blockDim.x >> 5);
    const size_t n_tiles = (n_reads + 31) >> 5;
    // Persistent warps: warp w of the grid takes 32-record tiles w, w + W, w + 2W, ... (records are coordinate
    // sorted, so a tile touches a handful of neighbouring index entries).  No CTA barrier inside the loop.
    size_t tile = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int4 q0 = make_int4(0, 0, 0, 0), q1 = q0, q2 = q0;
    if (tile < n_tiles && tile * 32 + lane < n_reads) {
        const int4* rp = reinterpret_cast<const int4*>(reads + tile * 32 + lane);
        q0 = __ldg(rp); q1 = __ldg(rp + 1); q2 = __ldg(rp + 2);
    }
  for (; tile < n_tiles; tile += warps_total) {
    const size_t i = tile * 32 + lane;
    // software pipeline: the next tile's records are requested before this tile is processed
    int4 p0 = make_int4(0, 0, 0, 0), p1 = p0, p2 = p0;
    {
        const size_t nx = (tile + warps_total) * 32 + lane;
        if (nx < n_reads) {
            const int4* rp = reinterpret_cast<const int4*>(reads + nx);
            p0 = __ldg(rp); p1 = __ldg(rp + 1); p2 = __ldg(rp + 2);
        }
    }

    unsigned long long cm = 0ull;       // bit k set: counter k += 1
    int v_mm1 = 0, v_b1 = 0, v_mm2 = 0, v_b2 = 0, v_mm = 0, v_tb = 0, v_nb = 0;
    unsigned int my_bytes = 0;
    uint32_t agg_gene = kNoGene;        // first gene this read counts toward (aggregated across the warp below)
    bool agg_unique = false, agg_frag = false;

#define CNT(k) (cm |= 1ull << (k))
    if (i < n_reads) {
        // ---- record: three 128-bit loads (issued one tile ahead) ------------------------------------------------
        const int pos = q0.x, mpos = q0.y, endpos = q0.z, isize = q0.w;
        const int l_qseq = q1.x, nm_raw = q1.y;
        const uint32_t w6 = (uint32_t)q1.z, w7 = (uint32_t)q1.w;
        const uint32_t f = w6 & 0xffffu;
        const uint32_t mapq = (w6 >> 16) & 0xffu;
        const uint32_t aux = w6 >> 24;
        const int contig = (int)(int16_t)(w7 & 0xffffu);
        const uint32_t filter_mask = (w7 >> 16) & 0xffu;
        const unsigned long long name_hash = ((unsigned long long)(uint32_t)q2.y << 32) | (uint32_t)q2.x;
        const uint32_t block_off = (uint32_t)q2.z;
        uint32_t nb = (uint32_t)q2.w & 0xffffu;
        if ((size_t)block_off + nb > n_blocks) nb = 0;  // defensive: never read out of the submitted array
        (void)endpos; (void)isize;

        CNT(RSQ_C_TOTAL_ALIGNMENTS);                                                  // RNASeQC.cpp:245
        my_bytes = 40u + 8u * nb;
        // ---- filter cascade, RNASeQC.cpp:254-342 -----------------------------------------------------------------
        if (f & RSQ_F_SECONDARY) CNT(RSQ_C_ALTERNATIVE_ALIGNMENTS);                   // :254
        if (f & RSQ_F_SUPPLEMENTARY) CNT(RSQ_C_SUPPLEMENTARY_ALIGNMENTS);             // :255
        else if (f & RSQ_F_QCFAIL) CNT(RSQ_C_FAILED_VENDOR_QC);
        else if (mapq < prm.mapq_threshold) CNT(RSQ_C_LOW_MAPPING_QUALITY);
        const bool has_ch = aux & RSQ_AUX_CHIMERIC_TAG;
        bool go = true;
        if ((f & RSQ_F_SUPPLEMENTARY) && !has_ch) {                                    // :258-262
            CNT(RSQ_C_CHIMERIC_AUTO);
            if (prm.exclude_chimeric) go = false;
        }
        if (f & (RSQ_F_SECONDARY | RSQ_F_QCFAIL | RSQ_F_SUPPLEMENTARY)) go = false;   // :263
        if (go) {
            CNT(RSQ_C_UMVQ);
            if (!(f & RSQ_F_PAIRED)) CNT(RSQ_C_UNPAIRED_READS);
            if (f & RSQ_F_UNMAP) go = false;                                          // :268
        }
        if (go) {
            CNT(RSQ_C_MAPPED_READS);
            if (f & RSQ_F_DUP) CNT(RSQ_C_MAPPED_DUPLICATE_READS); else CNT(RSQ_C_MAPPED_UNIQUE_READS);
            if (has_ch) {                                                              // :279-283
                if (f & RSQ_F_READ1) CNT(RSQ_C_CHIMERIC_TAG);
                if (prm.exclude_chimeric) go = false;
            }
        }
        if (go && (f & RSQ_F_PAIRED) && !(f & RSQ_F_MUNMAP)) {                         // :284-292
            if (f & RSQ_F_READ1) CNT(RSQ_C_TOTAL_MAPPED_PAIRS);
            int d = pos - mpos; d = d < 0 ? -d : d;
            if ((aux & RSQ_AUX_MATE_OTHER_CONTIG) || d > prm.chimeric_distance) {
                if (f & RSQ_F_READ1) CNT(RSQ_C_CHIMERIC_AUTO);
                if (prm.exclude_chimeric) go = false;
            }
        }
        bool hq = false;
        if (go) {
            int mismatches = 0;
            if (aux & RSQ_AUX_HAS_NM) {                                                // :294-316
                mismatches = nm_raw;
                if (f & RSQ_F_PAIRED) {
                    if (f & RSQ_F_READ1) {
                        CNT(RSQ_C_END1_MAPPED_READS);
                        v_mm1 = mismatches; v_b1 = l_qseq;
                        if (f & RSQ_F_DUP) CNT(RSQ_C_DUPLICATE_PAIRS); else CNT(RSQ_C_UNIQUE_FRAGMENTS);
                    } else {
                        CNT(RSQ_C_END2_MAPPED_READS);
                        v_mm2 = mismatches; v_b2 = l_qseq;
                    }
                }
                v_mm = mismatches;
            }
            v_tb = l_qseq;                                                             // :317
            if (filter_mask) {                                                         // :319-328
                for (uint32_t k = 0; k < prm.n_filter_tags && k < 8; ++k)
                    if (filter_mask & (1u << k)) { CNT(RSQ_C_FILTERED_BY_TAG0 + k); go = false; }
            }
            hq = (uint32_t)mismatches <= prm.base_mismatch_threshold && (prm.unpaired || (f & RSQ_F_PROPER)) &&
                 mapq >= prm.mapq_threshold;                                           // :330
            if (aux & RSQ_AUX_TID_OUT_OF_RANGE) go = false;                            // :333-337
        }
        if (go) {
            if (hq) CNT(RSQ_C_HIGH_QUALITY_READS); else CNT(RSQ_C_LOW_QUALITY_READS);  // :340-342
            CNT(RSQ_C_READS_USED);
            v_nb = (int)nb;                                                            // :360

            // ---- exonAlignmentMetrics, Expression.cpp:308-458 ----------------------------------------------------
            const bool cleared = (aux & RSQ_AUX_CONTIG_DROPPED) || contig < 0 || (uint32_t)contig >= ann.n_contigs;
            int rstrand = RSQ_STRAND_UNKNOWN;                                          // feature_strand, :119-125
            if (prm.strand_mode) {
                bool target = f & RSQ_F_REVERSE;
                if ((prm.strand_mode == 1) != ((f & RSQ_F_READ1) != 0)) target = !target;
                rstrand = target ? RSQ_STRAND_REV : RSQ_STRAND_FWD;
            }
            int c_lo = 0, c_hi = 0, s_hi = 0, bk_lo = 0, bk_n = 1;
            if (!cleared) {
                c_lo = (int)__ldg(&ann.exon_off[contig]); c_hi = (int)__ldg(&ann.exon_off[contig + 1]);
                s_hi = (int)__ldg(&ann.seg_off[contig + 1]);
                bk_lo = (int)__ldg(&ann.bucket_off[contig]); bk_n = (int)__ldg(&ann.bucket_off[contig + 1]) - bk_lo;
            }
            bool intragenic = false, plus = false, minus = false, ribosomal = false, exonic = false, do_exon = false;
            uint32_t length = 0;                                                       // extractBlocks' return, :359
            Window w0 = {0, 0}, w1 = {0, 0};
            int bs0 = 0, be0 = 0, bs1 = 0, be1 = 0;
            const rsq_block* bp = blocks + block_off;
            for (uint32_t b = 0; b < nb; ++b) {                                        // :325-361
                const int2 blk = __ldg(reinterpret_cast<const int2*>(bp + b));
                const int bs = blk.x, be = blk.x + blk.y;                              // closed end = start+len, :45
                length += (uint32_t)blk.y;
                if (b == 0) { bs0 = bs; be0 = be; } else if (b == 1) { bs1 = bs; be1 = be; }
                if (cleared) continue;
                // Gene features: elementary segments overlapping [bs, be]  (:352-358)
                for (int k = seg_lookup(ann, s_hi, bk_lo, bk_n, bs); k < s_hi && __ldg(&ann.seg_start[k]) <= be; ++k) {
                    uint32_t fl = __ldg(&ann.seg_flags[k]);
                    if (rstrand != RSQ_STRAND_UNKNOWN) fl &= (RSQ_SEG_GENE_FWD | RSQ_SEG_RIBO_FWD) << rstrand;   // :331
                    if (fl & (RSQ_SEG_GENE_FWD | RSQ_SEG_GENE_REV | RSQ_SEG_GENE_UNK)) intragenic = true;
                    if (fl & RSQ_SEG_GENE_FWD) plus = true;
                    if (fl & RSQ_SEG_GENE_REV) minus = true;
                    if (fl & (RSQ_SEG_RIBO_FWD | RSQ_SEG_RIBO_REV | RSQ_SEG_RIBO_UNK)) ribosomal = true;
                }
                // Exon features (:335-351)
                const Window w = exon_window(ann, c_lo, c_hi, bk_lo, bk_n, bs, be);
                if (b == 0) w0 = w; else if (b == 1) w1 = w;
                for (int e = w.lo; e < w.hi; ++e) {
                    const int2 x = __ldg(&ann.exon_se[e]);
                    if (x.y < bs) continue;                                            // intersectInterval, GTF.cpp:176-179
                    const uint32_t gf = __ldg(&ann.exon_meta[e].gene_flags);
                    const int fs = (gf >> 24) & RSQ_EXON_STRAND_MASK;
                    if (rstrand != RSQ_STRAND_UNKNOWN && rstrand != fs) continue;      // :331
                    if (fs == RSQ_STRAND_FWD) plus = true; else if (fs == RSQ_STRAND_REV) minus = true;
                    exonic = true;                                                     // :337
                    if ((gf >> 24) & RSQ_EXON_RIBOSOMAL) ribosomal = true;             // :358
                }
            }

            bool globin = false;
            if (nb >= 1 && !cleared) {                                                 // :363-394
                // `last` = intersection over blocks of {genes with an exon that fully contains the block}.  Enumerate
                // the candidates of block 0 (deduplicated by re-scanning the window: no per-thread set storage) and
                // test each against every other block.
                for (int e0 = w0.lo; e0 < w0.hi; ++e0) {
                    const int2 x0 = __ldg(&ann.exon_se[e0]);
                    if (!(x0.x <= bs0 && be0 - 1 <= x0.y)) continue;                   // isz == block length, :341
                    const uint32_t gf0 = __ldg(&ann.exon_meta[e0].gene_flags);
                    if (rstrand != RSQ_STRAND_UNKNOWN && rstrand != (int)((gf0 >> 24) & RSQ_EXON_STRAND_MASK)) continue;
                    const uint32_t g = gf0 & 0xffffffu;
                    if (e0 > w0.lo && block_has_gene(ann, Window{w0.lo, e0}, bs0, be0, g, rstrand)) continue;  // seen already
                    bool in_all = true;
                    for (uint32_t b = 1; b < nb && in_all; ++b) {
                        int bs, be; Window w;
                        if (b == 1) { bs = bs1; be = be1; w = w1; }
                        else {
                            const int2 blk = __ldg(reinterpret_cast<const int2*>(bp + b));
                            bs = blk.x; be = blk.x + blk.y;
                            w = exon_window(ann, c_lo, c_hi, bk_lo, bk_n, bs, be);
                        }
                        in_all = block_has_gene(ann, w, bs, be, g, rstrand);
                    }
                    if (!in_all) continue;
                    // g is in `last` (:377)
                    do_exon = true;                                                    // :393
                    if (__ldg(&ann.gene_flags[g]) & RSQ_GENE_GLOBIN) globin = true;    // :396-398
                    if (!hq) continue;
                    if (length > 0) {                                                  // queryGene: some coverage > 0, :380
                        const unsigned long long key0 = mix64(name_hash ^ ((unsigned long long)(g + 1u) * 0x9E3779B97F4A7C15ull));
                        const bool fresh = set_insert(acc.frag_set, acc.frag_mask, key0 ? key0 : 1ull, acc.error_flag, 1u);  // :383-387
                        if (agg_gene == kNoGene) { agg_gene = g; agg_unique = !(f & RSQ_F_DUP); agg_frag = fresh; }
                        else {
                            atomicAdd(&acc.gene_reads[g], 1ull);
                            if (fresh) atomicAdd(&acc.gene_fragments[g], 1ull);
                            if (!(f & RSQ_F_DUP)) atomicAdd(&acc.gene_unique[g], 1ull);
                        }
                        my_bytes += 8u;
                    }
                    // Collector::collect + BaseCoverage::commit for gene g: every (block, containing exon of g) pair
                    for (uint32_t b = 0; b < nb; ++b) {
                        int bs, be; Window w;
                        if (b == 0) { bs = bs0; be = be0; w = w0; }
                        else if (b == 1) { bs = bs1; be = be1; w = w1; }
                        else {
                            const int2 blk = __ldg(reinterpret_cast<const int2*>(bp + b));
                            bs = blk.x; be = blk.x + blk.y;
                            w = exon_window(ann, c_lo, c_hi, bk_lo, bk_n, bs, be);
                        }
                        const uint32_t len = (uint32_t)(be - bs);
                        if (len == 0) continue;                                        // Collector::add keeps coverage > 0, Metrics.cpp:51
                        for (int e = w.lo; e < w.hi; ++e) {
                            const int2 x = __ldg(&ann.exon_se[e]);
                            if (!(x.x <= bs && be - 1 <= x.y)) continue;
                            const ExonMeta m = ann.exon_meta[e];
                            if ((m.gene_flags & 0xffffffu) != g) continue;
                            if (rstrand != RSQ_STRAND_UNKNOWN && rstrand != (int)((m.gene_flags >> 24) & RSQ_EXON_STRAND_MASK)) continue;
                            unsigned long long whole, lo;
                            frac_q64(len, length, whole, lo);                          // isz/length, :345
                            if (lo) {
                                const unsigned long long old = atomicAdd(&acc.exon_frac[m.slot], lo);
                                if (old + lo < old) whole += 1ull;                     // carry out of the 2^-64 word
                            }
                            if (whole) atomicAdd(&acc.exon_whole[m.slot], whole);
                            my_bytes += 8u;
                            if (prm.collect_coverage) {                                // add_range as +1/-1, Metrics.cpp:257-262
                                const unsigned long long o = m.cov_off + (unsigned long long)(bs - x.x);
                                atomicAdd(&acc.cov_diff[o], 1u);
                                atomicAdd(&acc.cov_diff[o + len], 0xffffffffu);
                                my_bytes += 8u;
                            }
                        }
                    }
                }
            }
            if (nb >= 1 && !globin) {                                                  // :399-404
                CNT(RSQ_C_NON_GLOBIN_READS);
                if (f & RSQ_F_DUP) CNT(RSQ_C_NON_GLOBIN_DUPLICATE_READS);
            }
            if (!exonic) {                                                             // :407-423
                if (intragenic) {
                    CNT(RSQ_C_INTRONIC_READS); CNT(RSQ_C_INTRAGENIC_READS);
                    if (hq) { CNT(RSQ_C_HQ_INTRONIC_READS); CNT(RSQ_C_HQ_INTRAGENIC_READS); }
                } else {
                    CNT(RSQ_C_INTERGENIC_READS);
                    if (hq) CNT(RSQ_C_HQ_INTERGENIC_READS);
                }
            } else if (do_exon) {                                                      // :424-433
                CNT(RSQ_C_EXONIC_READS); CNT(RSQ_C_INTRAGENIC_READS);
                if (hq) { CNT(RSQ_C_HQ_EXONIC_READS); CNT(RSQ_C_HQ_INTRAGENIC_READS); }
            } else {                                                                   // :434-441
                CNT(RSQ_C_AMBIGUOUS_READS);
                if (hq) CNT(RSQ_C_HQ_AMBIGUOUS_READS);
            }
            if (ribosomal) CNT(RSQ_C_RRNA_READS);                                      // :442
            if ((minus != plus) && (prm.unpaired || (f & RSQ_F_PAIRED))) {             // :445-457
                const bool sense = (f & RSQ_F_REVERSE) ? minus : plus;
                if (prm.unpaired || (f & RSQ_F_READ1)) { if (sense) CNT(RSQ_C_END1_SENSE); else CNT(RSQ_C_END1_ANTISENSE); }
                else { if (sense) CNT(RSQ_C_END2_SENSE); else CNT(RSQ_C_END2_ANTISENSE); }
            }

            // ---- fragmentSizeMetrics candidate, RNASeQC.cpp:372 + Expression.cpp:482-516 ----------------------------
            if (kBed) {
                if (hq && (f & RSQ_F_PAIRED) && contig >= 0 && __ldg(&bed.present[contig])) {
                    const int b_lo = (int)__ldg(&bed.off[contig]), b_hi = (int)__ldg(&bed.off[contig + 1]);
                    bool same = nb > 0;
                    uint32_t name = 0;
                    for (uint32_t b = 0; b < nb && same; ++b) {                        // :490-507
                        const int2 blk = __ldg(reinterpret_cast<const int2*>(bp + b));
                        const int bs = blk.x, be = blk.x + blk.y;
                        const int hi = upper_bound_start(bed.se, b_lo, b_hi, be);
                        const int lo = lower_bound_pmax(bed.pmax, b_lo, hi, bs);
                        int hits = 0, hit = -1;
                        for (int e = lo; e < hi; ++e)
                            if (__ldg(&bed.se[e].y) >= bs) { if (++hits == 1) hit = e; }
                        if (hits == 1) {
                            const int2 x = __ldg(&bed.se[hit]);
                            const uint32_t nid = __ldg(&bed.name_id[hit]);
                            if (!(x.x <= bs && be - 1 <= x.y)) same = false;           // partialIntersect == block length, :494
                            else if (b == 0) name = nid;
                            else if (nid != name) same = false;                        // :497
                        } else same = false;
                    }
                    if (same) {                                                        // :508-516: enters the mate map
                        CNT(RSQ_C_FRAGMENT_CANDIDATES);
                        // claim the table slot of this read name and record the earliest candidate
                        unsigned long long key = name_hash ? name_hash : 1ull;
                        unsigned long long slot = mix64(key) & acc.pair_mask;
                        bool placed = false;
                        for (int probe = 0; probe < 4096 && !placed; ++probe) {
                            unsigned long long cur = acc.pair_keys[slot];
                            if (cur == 0ull) cur = atomicCAS(&acc.pair_keys[slot], 0ull, key);
                            if (cur == 0ull || cur == key) placed = true; else slot = (slot + 1) & acc.pair_mask;
                        }
                        if (!placed) atomicOr(acc.error_flag, 2u);
                        else {
                            atomicMin(&acc.pair_first[slot], prm.record_base + i);
                            const unsigned int ci = atomicAdd(acc.cand_count, 1u);
                            if (ci < acc.cand_capacity) {
                                PairCandidate c;
                                c.name_hash = key; c.local_idx = (uint32_t)i; c.name_id = name; c.endpos = endpos;
                                c.abs_isize = (uint32_t)(isize < 0 ? -(long long)isize : (long long)isize);
                                c.cond_ok = (!(f & RSQ_F_MREVERSE) && (f & RSQ_F_REVERSE) && pos != mpos) ? 1u : 0u;   // :528
                                c.slot = (uint32_t)slot;
                                acc.cand[ci] = c;
                            } else atomicOr(acc.error_flag, 4u);
                        }
                    }
                }
            }
        }
    }
#undef CNT

    // ---- warp-converged epilogue: aggregate before touching HBM -------------------------------------------------------
    // gene counters: lanes that count toward the same gene elect one leader (sorted input => usually the whole warp)
    {
        const unsigned peers = __match_any_sync(kFull, agg_gene);
        const unsigned bu = __ballot_sync(kFull, agg_unique);
        const unsigned bf = __ballot_sync(kFull, agg_frag);
        if (agg_gene != kNoGene && lane == (unsigned)(__ffs(peers) - 1)) {
            atomicAdd(&acc.gene_reads[agg_gene], (unsigned long long)__popc(peers));
            const int nu = __popc(peers & bu), nf = __popc(peers & bf);
            if (nu) atomicAdd(&acc.gene_unique[agg_gene], (unsigned long long)nu);
            if (nf) atomicAdd(&acc.gene_fragments[agg_gene], (unsigned long long)nf);
        }
    }
    // +1 counters: one ballot per counter that any lane of the warp touched
    {
        unsigned any_lo = __reduce_or_sync(kFull, (unsigned)(cm & 0xffffffffull));
        unsigned any_hi = __reduce_or_sync(kFull, (unsigned)(cm >> 32));
        while (any_lo) {
            const int k = __ffs(any_lo) - 1;
            any_lo &= any_lo - 1;
            const unsigned b = __ballot_sync(kFull, (cm >> k) & 1ull);
            if (lane == 0) atomicAdd(&s_cnt[k], (unsigned long long)__popc(b));
        }
        while (any_hi) {
            const int k = __ffs(any_hi) - 1;
            any_hi &= any_hi - 1;
            const unsigned b = __ballot_sync(kFull, (cm >> (32 + k)) & 1ull);
            if (lane == 0) atomicAdd(&s_cnt[32 + k], (unsigned long long)__popc(b));
        }
    }
    // valued counters (Metrics::increment(key, int), Metrics.cpp:33-36): signed 32-bit sums, sign-extended
    {
        const int s_mm1 = __reduce_add_sync(kFull, v_mm1), s_b1 = __reduce_add_sync(kFull, v_b1);
        const int s_mm2 = __reduce_add_sync(kFull, v_mm2), s_b2 = __reduce_add_sync(kFull, v_b2);
        const int s_mm = __reduce_add_sync(kFull, v_mm), s_tb = __reduce_add_sync(kFull, v_tb);
        const int s_nb = __reduce_add_sync(kFull, v_nb);
        const unsigned s_bytes_w = __reduce_add_sync(kFull, my_bytes);
        if (lane == 0) {
            if (s_mm1) atomicAdd(&s_cnt[RSQ_C_END1_MISMATCHES], (unsigned long long)(long long)s_mm1);
            if (s_b1) atomicAdd(&s_cnt[RSQ_C_END1_BASES], (unsigned long long)(long long)s_b1);
            if (s_mm2) atomicAdd(&s_cnt[RSQ_C_END2_MISMATCHES], (unsigned long long)(long long)s_mm2);
            if (s_b2) atomicAdd(&s_cnt[RSQ_C_END2_BASES], (unsigned long long)(long long)s_b2);
            if (s_mm) atomicAdd(&s_cnt[RSQ_C_MISMATCHED_BASES], (unsigned long long)(long long)s_mm);
            if (s_tb) atomicAdd(&s_cnt[RSQ_C_TOTAL_BASES], (unsigned long long)(long long)s_tb);
            if (s_nb) atomicAdd(&s_cnt[RSQ_C_ALIGNMENT_BLOCKS], (unsigned long long)(long long)s_nb);
            if (s_bytes_w) atomicAdd(&s_bytes, (unsigned long long)s_bytes_w);
        }
    }
    q0 = p0; q1 = p1; q2 = p2;
  }  // tile loop
    __syncthreads();
    // one global atomic per counter per CTA
    for (int k = threadIdx.x; k < RSQ_N_COUNTERS; k += blockDim.x)
        if (s_cnt[k]) atomicAdd(&acc.counters[k], s_cnt[k]);
    if (threadIdx.x == 0 && s_bytes) atomicAdd(acc.algo_bytes, s_bytes);
}

// ---- annotation preprocessing -----------------------------------------------------------------------------------------
// running max of exon end inside each contig (one thread per contig: a few hundred thousand elements, once per run)
__global__ void prefix_max_kernel(const int2* __restrict__ se, const uint32_t* __restrict__ off, uint32_t n_contigs, int32_t* __restrict__ pmax) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_contigs) return;
    int m = INT32_MIN;
    for (uint32_t i = off[c]; i < off[c + 1]; ++i) {
        m = max(m, se[i].y);
        pmax[i] = m;
    }
}

// direct-address bucket index (one thread per bucket, binary searches once per run)
__global__ void build_buckets_kernel(const AnnView ann, int4* __restrict__ buckets, uint32_t n_buckets) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_buckets) return;
    int a = 0, b = (int)ann.n_contigs;           // contig of this bucket: last c with bucket_off[c] <= g
    while (a < b) { const int mid = (a + b) >> 1; if (ann.bucket_off[mid] > g) b = mid; else a = mid + 1; }
    const int c = a - 1;
    const long long x0 = (long long)(g - ann.bucket_off[c]) << kBucketShift;
    const int key = (int)min(x0, (long long)INT32_MAX);
    const int c_lo = (int)ann.exon_off[c], c_hi = (int)ann.exon_off[c + 1];
    int4 e;
    e.x = lower_bound_pmax(ann.exon_pmax, c_lo, c_hi, key);
    e.y = upper_bound_start(ann.exon_se, c_lo, c_hi, key - 1);
    e.z = seg_find(ann.seg_start, (int)ann.seg_off[c], (int)ann.seg_off[c + 1], key);
    e.w = 0;
    buckets[g] = e;
}

// mate-map growth: re-insert (key, first index, payload) triples into a larger table
__global__ void pair_rehash_kernel(const Accum old_acc, unsigned long long old_cap, unsigned long long* __restrict__ nk,
                                   unsigned long long* __restrict__ nf, unsigned long long* __restrict__ np, unsigned long long new_mask,
                                   unsigned int* err) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= old_cap) return;
    const unsigned long long key = old_acc.pair_keys[i];
    if (!key) return;
    unsigned long long slot = mix64(key) & new_mask;
    for (int probe = 0; probe < 4096; ++probe) {
        const unsigned long long cur = atomicCAS(&nk[slot], 0ull, key);
        if (cur == 0ull) { nf[slot] = old_acc.pair_first[i]; np[slot] = old_acc.pair_payload[i]; return; }
        slot = (slot + 1) & new_mask;
    }
    atomicOr(err, 2u);
}

// ---- coverage: integrate the difference array (three-phase scan) ----------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;                       // per thread
constexpr int kScanTile = kScanThreads * kScanItems; // 4096 elements per CTA

__device__ __forceinline__ unsigned int block_exclusive_scan(unsigned int v, unsigned int* s_warp, unsigned int& total) {
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    unsigned int inc = v;
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int t = __shfl_up_sync(kFull, inc, d);
        if ((int)lane >= d) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned int w = lane < (blockDim.x >> 5) ? s_warp[lane] : 0u;
        unsigned int winc = w;
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int t = __shfl_up_sync(kFull, winc, d);
            if ((int)lane >= d) winc += t;
        }
        s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    total = s_warp[32];
    const unsigned int r = s_warp[wid] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) tile_sum_kernel(const unsigned int* __restrict__ in, unsigned long long n, unsigned int* __restrict__ sums) {
    __shared__ unsigned int s_warp[33];
    const unsigned long long base = (unsigned long long)blockIdx.x * kScanTile;
    unsigned int v = 0;
    for (int k = 0; k < kScanItems; ++k) {
        const unsigned long long idx = base + (unsigned long long)k * kScanThreads + threadIdx.x;
        if (idx < n) v += in[idx];
    }
    unsigned int total;
    block_exclusive_scan(v, s_warp, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of the tile sums in place (sequential over chunks of blockDim)
__global__ void __launch_bounds__(1024) scan_sums_kernel(unsigned int* __restrict__ sums, unsigned long long n_tiles) {
    __shared__ unsigned int s_warp[33];
    unsigned int carry = 0;
    for (unsigned long long base = 0; base < n_tiles; base += blockDim.x) {
        const unsigned long long idx = base + threadIdx.x;
        const unsigned int v = idx < n_tiles ? sums[idx] : 0u;
        unsigned int total;
        const unsigned int ex = block_exclusive_scan(v, s_warp, total);
        if (idx < n_tiles) sums[idx] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(kScanThreads) tile_scan_kernel(const unsigned int* __restrict__ in, unsigned int* __restrict__ out, unsigned long long n,
                                                                 const unsigned int* __restrict__ sums) {
    __shared__ unsigned int s_warp[33];
    const unsigned long long base = (unsigned long long)blockIdx.x * kScanTile + (unsigned long long)threadIdx.x * kScanItems;
    unsigned int vals[kScanItems];
    unsigned int v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const unsigned long long idx = base + k;
        vals[k] = idx < n ? in[idx] : 0u;
        v += vals[k];
    }
    unsigned int total;
    unsigned int run = block_exclusive_scan(v, s_warp, total) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        run += vals[k];                      // inclusive: coverage[i] = sum of diff[0..i]
        const unsigned long long idx = base + k;
        if (idx < n) out[idx] = run;
    }
}

// ---- fragment-tracker set growth ----------------------------------------------------------------------------------------------
__global__ void rehash_kernel(const unsigned long long* __restrict__ old_tab, unsigned long long old_cap, unsigned long long* __restrict__ new_tab,
                              unsigned long long new_mask, unsigned int* err) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= old_cap) return;
    const unsigned long long key = old_tab[i];
    if (key) set_insert(new_tab, new_mask, key, err, 1u);
}

// ---- count vector export for the NVLink all-reduce: counters | gene_reads | gene_unique | gene_fragments | exon_whole |
//      exon_frac >> 32 | exon_frac & 0xffffffff   (the split keeps the u64 sum exact across ranks)
__global__ void pack_counts_kernel(const Accum acc, uint32_t n_genes, uint32_t n_exons, unsigned long long* __restrict__ out) {
    const unsigned long long total = RSQ_N_COUNTERS + 3ull * n_genes + 3ull * n_exons;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long j = i, v;
        if (j < RSQ_N_COUNTERS) v = acc.counters[j];
        else if ((j -= RSQ_N_COUNTERS) < n_genes) v = acc.gene_reads[j];
        else if ((j -= n_genes) < n_genes) v = acc.gene_unique[j];
        else if ((j -= n_genes) < n_genes) v = acc.gene_fragments[j];
        else if ((j -= n_genes) < n_exons) v = acc.exon_whole[j];
        else if ((j -= n_exons) < n_exons) v = acc.exon_frac[j] >> 32;
        else { j -= n_exons; v = acc.exon_frac[j] & 0xffffffffull; }
        out[i] = v;
    }
}
__global__ void unpack_counts_kernel(const Accum acc, uint32_t n_genes, uint32_t n_exons, const unsigned long long* __restrict__ in) {
    const unsigned long long n_simple = RSQ_N_COUNTERS + 3ull * n_genes;
    const unsigned long long total = n_simple + n_exons;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long j = i;
        if (j < RSQ_N_COUNTERS) acc.counters[j] = in[i];
        else if ((j -= RSQ_N_COUNTERS) < n_genes) acc.gene_reads[j] = in[i];
        else if ((j -= n_genes) < n_genes) acc.gene_unique[j] = in[i];
        else if ((j -= n_genes) < n_genes) acc.gene_fragments[j] = in[i];
        else {
            j -= n_genes;
            const unsigned long long whole = in[n_simple + j], hi = in[n_simple + n_exons + j], lo = in[n_simple + 2ull * n_exons + j];
            // whole*2^64 + hi*2^32 + lo with carries
            const unsigned long long mid = hi + (lo >> 32);
            acc.exon_frac[j] = (mid << 32) | (lo & 0xffffffffull);
            acc.exon_whole[j] = whole + (mid >> 32);
        }
    }
}

// ---- fragment-size pairing (BED mode) ----------------------------------------------------------------------------------------------
// owner: the earliest candidate of a read name publishes (interval, PositionEnd)   (Expression.cpp:512-516)
__global__ void pair_owner_kernel(const Accum acc, const Params prm) {
    const unsigned int n = min(*acc.cand_count, acc.cand_capacity);
    const unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const PairCandidate c = acc.cand[k];
    if (acc.pair_first[c.slot] == prm.record_base + c.local_idx)
        acc.pair_payload[c.slot] = ((unsigned long long)c.name_id << 32) | (unsigned long long)(uint32_t)c.endpos;
}
// second: a later candidate with the same name tests the four conditions   (Expression.cpp:517-533)
__global__ void pair_check_kernel(const Accum acc, const Params prm) {
    const unsigned int n = min(*acc.cand_count, acc.cand_capacity);
    const unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const PairCandidate c = acc.cand[k];
    if (acc.pair_first[c.slot] >= prm.record_base + c.local_idx) return;
    const unsigned long long p = acc.pair_payload[c.slot];
    if ((uint32_t)(p >> 32) != c.name_id) return;                       // exonName == stored exon, :517
    if (!c.cond_ok || c.endpos <= (int32_t)(uint32_t)p) return;         // :528
    acc.sample_flag[c.local_idx] = 1;
    acc.sample_isize[c.local_idx] = c.abs_isize;                        // fragmentSizes[abs(InsertSize())], :530
}
// rank samples in stream order and keep the first fragment_samples of the whole run   (RNASeQC.cpp:372, Expression.cpp:532)
__global__ void __launch_bounds__(kScanThreads) sample_count_kernel(const unsigned char* __restrict__ flag, unsigned long long n, unsigned int* __restrict__ sums) {
    __shared__ unsigned int s_warp[33];
    const unsigned long long base = (unsigned long long)blockIdx.x * kScanTile;
    unsigned int v = 0;
    for (int k = 0; k < kScanItems; ++k) {
        const unsigned long long idx = base + (unsigned long long)k * kScanThreads + threadIdx.x;
        if (idx < n) v += flag[idx];
    }
    unsigned int total;
    block_exclusive_scan(v, s_warp, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kScanThreads) sample_rank_kernel(const unsigned char* __restrict__ flag, const unsigned int* __restrict__ isize, unsigned long long n,
                                                                   const unsigned int* __restrict__ sums, const Accum acc, const Params prm) {
    __shared__ unsigned int s_warp[33];
    const unsigned long long base = (unsigned long long)blockIdx.x * kScanTile + (unsigned long long)threadIdx.x * kScanItems;
    unsigned int v = 0;
    unsigned char fl[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const unsigned long long idx = base + k;
        fl[k] = idx < n ? flag[idx] : (unsigned char)0;
        v += fl[k];
    }
    unsigned int total;
    unsigned long long rank = *acc.sample_count + sums[blockIdx.x] + block_exclusive_scan(v, s_warp, total);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (fl[k]) {
            if (rank < prm.fragment_samples) acc.sample_vals[rank] = isize[base + k];
            ++rank;
        }
}
__global__ void sample_commit_kernel(const unsigned int* __restrict__ sums, unsigned long long n_tiles, const unsigned int* __restrict__ last_tile_total,
                                     const Accum acc) {
    // sums[] is exclusive-scanned: batch total = sums[last] + count of the last tile
    if (threadIdx.x == 0 && blockIdx.x == 0 && n_tiles) *acc.sample_count += (unsigned long long)sums[n_tiles - 1] + *last_tile_total;
}
__global__ void save_last_kernel(const unsigned int* __restrict__ sums, unsigned long long n_tiles, unsigned int* __restrict__ dst) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && n_tiles) *dst = sums[n_tiles - 1];
}

__global__ void sample_hist_kernel(const unsigned int* __restrict__ vals, unsigned long long n, unsigned long long* __restrict__ hist, unsigned long long hist_len) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long b = vals[i];
        if (b >= hist_len) b = hist_len - 1;
        atomicAdd(&hist[b], 1ull);
    }
}

}  // namespace

// -------------------------------------------------------------------------------------------------------------------------------------
int launch_hot_path(const rsq_read* reads, const rsq_block* blocks, size_t n_reads, size_t n_blocks, const AnnView& ann, const BedView& bed,
                    const Accum& acc, const Params& prm, cudaStream_t stream) {
    if (!n_reads) return 0;
    // persistent grid: every SM holds exactly as many CTAs as its register file admits
    static int per_sm[2] = {0, 0}, sms = 0;
    const int which = prm.bed_enabled ? 1 : 0;
    if (!per_sm[which]) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (which) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[1], hot_path_kernel<true>, kHotThreads, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[0], hot_path_kernel<false>, kHotThreads, 0);
        if (per_sm[which] < 1) per_sm[which] = 1;
    }
    const size_t tiles = (n_reads + 31) / 32;
    const unsigned grid = (unsigned)std::min<size_t>((size_t)sms * per_sm[which], (tiles + 7) / 8);
    if (prm.bed_enabled) hot_path_kernel<true><<<grid, kHotThreads, 0, stream>>>(reads, blocks, n_reads, n_blocks, ann, bed, acc, prm);
    else hot_path_kernel<false><<<grid, kHotThreads, 0, stream>>>(reads, blocks, n_reads, n_blocks, ann, bed, acc, prm);
    return 1;
}

int launch_prefix_max(const int2* se, const uint32_t* off, uint32_t n_contigs, int32_t* pmax, cudaStream_t stream) {
    if (!n_contigs) return 0;
    prefix_max_kernel<<<(n_contigs + 63) / 64, 64, 0, stream>>>(se, off, n_contigs, pmax);
    return 1;
}

int launch_build_buckets(const AnnView& ann, int4* buckets, uint32_t n_buckets, cudaStream_t stream) {
    if (!n_buckets) return 0;
    build_buckets_kernel<<<(n_buckets + 255) / 256, 256, 0, stream>>>(ann, buckets, n_buckets);
    return 1;
}

int launch_pair_rehash(const Accum& old_acc, unsigned long long old_cap, unsigned long long* nk, unsigned long long* nf,
                       unsigned long long* np, unsigned long long new_mask, unsigned int* error_flag, cudaStream_t stream) {
    pair_rehash_kernel<<<(unsigned)((old_cap + 255) / 256), 256, 0, stream>>>(old_acc, old_cap, nk, nf, np, new_mask, error_flag);
    return 1;
}

size_t coverage_scan_scratch_elems(unsigned long long n) { return (size_t)((n + kScanTile - 1) / kScanTile) + 1; }

int launch_coverage_scan(const unsigned int* diff, unsigned int* cov, unsigned long long n, unsigned int* block_sums, cudaStream_t stream) {
    if (!n) return 0;
    const unsigned long long tiles = (n + kScanTile - 1) / kScanTile;
    tile_sum_kernel<<<(unsigned)tiles, kScanThreads, 0, stream>>>(diff, n, block_sums);
    scan_sums_kernel<<<1, 1024, 0, stream>>>(block_sums, tiles);
    tile_scan_kernel<<<(unsigned)tiles, kScanThreads, 0, stream>>>(diff, cov, n, block_sums);
    return 3;
}

int launch_rehash(const unsigned long long* old_tab, unsigned long long old_cap, unsigned long long* new_tab, unsigned long long new_mask,
                  unsigned int* error_flag, cudaStream_t stream) {
    rehash_kernel<<<(unsigned)((old_cap + 255) / 256), 256, 0, stream>>>(old_tab, old_cap, new_tab, new_mask, error_flag);
    return 1;
}

int launch_pack_counts(const Accum& acc, uint32_t n_genes, uint32_t n_exons, unsigned long long* out, cudaStream_t stream) {
    pack_counts_kernel<<<148 * 4, 256, 0, stream>>>(acc, n_genes, n_exons, out);
    return 1;
}
int launch_unpack_counts(const Accum& acc, uint32_t n_genes, uint32_t n_exons, const unsigned long long* in, cudaStream_t stream) {
    unpack_counts_kernel<<<148 * 4, 256, 0, stream>>>(acc, n_genes, n_exons, in);
    return 1;
}

size_t pair_scan_scratch_elems(size_t n_reads) { return (n_reads + kScanTile - 1) / kScanTile + 2; }

int launch_pair_stages(size_t n_reads, const Accum& acc, const Params& prm, unsigned int* scratch, cudaStream_t stream) {
    if (!n_reads) return 0;
    const unsigned cgrid = (unsigned)((std::min<size_t>(n_reads, acc.cand_capacity) + 255) / 256);
    pair_owner_kernel<<<cgrid, 256, 0, stream>>>(acc, prm);
    pair_check_kernel<<<cgrid, 256, 0, stream>>>(acc, prm);
    const unsigned long long tiles = (n_reads + kScanTile - 1) / kScanTile;
    unsigned int* sums = scratch;
    unsigned int* last = scratch + tiles + 1;
    sample_count_kernel<<<(unsigned)tiles, kScanThreads, 0, stream>>>(acc.sample_flag, n_reads, sums);
    save_last_kernel<<<1, 32, 0, stream>>>(sums, tiles, last);            // count of the last tile, before the scan overwrites it
    scan_sums_kernel<<<1, 1024, 0, stream>>>(sums, tiles);
    sample_rank_kernel<<<(unsigned)tiles, kScanThreads, 0, stream>>>(acc.sample_flag, acc.sample_isize, n_reads, sums, acc, prm);
    sample_commit_kernel<<<1, 32, 0, stream>>>(sums, tiles, last, acc);
    return 7;
}

int launch_sample_histogram(const unsigned int* vals, unsigned long long n, unsigned long long* hist, unsigned long long hist_len, cudaStream_t stream) {
    if (!n) return 0;
    sample_hist_kernel<<<148, 256, 0, stream>>>(vals, n, hist, hist_len);
    return 1;
}

}  // namespace rsqd
