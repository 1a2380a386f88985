// rsq_hot_core.cuh — the per-record logic of the hot path, written once for the device kernel (rsq_kernels.cu) and for
// the host-side emulation that the CPU tests run against the oracle (tests/hostcore/): the filter cascade of
// src/RNASeQC.cpp:254-342 and exonAlignmentMetrics of src/Expression.cpp:308-458 over the elementary-interval index
// (rsq_eit.h).  Everything that needs a warp (staging, aggregation, counters) stays in the kernel; memory side effects go
// through a `Mem` policy (device: RED/CAS atomics; host emulation: plain adds).
#pragma once
#include <stdint.h>

#include "rsq_device.h"

#if defined(__CUDACC__)
#define RSQ_HD __device__ __forceinline__      // the CUDA build only ever needs the device instantiation
#define RSQ_HDX __host__ __device__ __forceinline__   // ... except for the pure functions the library's host side calls too
#else
#define RSQ_HD inline                          // host emulation (tests/hostcore), plain g++
#define RSQ_HDX inline
#endif

namespace rsqd {

template <class T> RSQ_HD T ldro(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

constexpr uint32_t kNoGene = 0xffffffffu;
RSQ_HD int imin(int a, int b) { return a < b ? a : b; }
RSQ_HD int imax(int a, int b) { return a > b ? a : b; }

RSQ_HD unsigned long long umulhi64(unsigned long long a, unsigned long long b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
#endif
}

RSQ_HD unsigned long long mix64(unsigned long long x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

// floor(isz * 2^64 / length) for isz <= length (isz == length gives whole = 1).  For aligned lengths below 65536 the
// table R[length] = floor(2^64 / length) (rsq_eit.h, 512 KB, one hot line per read length) leaves one 32-bit divide:
// 2^64 = R*L + r0  =>  floor(isz * 2^64 / L) = isz*R + floor(isz*r0 / L), and isz*r0 < L^2 < 2^32.
RSQ_HD void frac_q64(const unsigned long long* recip64, uint32_t isz, uint32_t length, unsigned long long& whole, unsigned long long& lo) {
    if (isz >= length) { whole = 1ull; lo = 0ull; return; }
    whole = 0ull;
    if (length < 65536u) {
        const unsigned long long R = ldro(recip64 + length);
        const uint32_t r0 = (uint32_t)(0ull - R * (unsigned long long)length);
#ifdef RSQ_FRAC_DIVIDE
        lo = (unsigned long long)isz * R + (unsigned long long)((isz * r0) / length);
#else
        // floor(x / L) for x = isz * r0 < L^2 < 2^32 without a divide: with c = R + 1 = 2^64 / L + d, 0 < d <= 1, the product
        // x * c / 2^64 = x / L + x * d / 2^64 exceeds x / L by less than 2^-32, and x / L lies at least 1 / L >= 2^-16 below the
        // next integer unless it is one: the high word is the exact quotient
        lo = (unsigned long long)isz * R + umulhi64((unsigned long long)(isz * r0), R + 1ull);
#endif
        return;
    }
    unsigned long long r = isz;
    const unsigned long long t1 = (r << 32) / length;
    r = (r << 32) % length;
    const unsigned long long t0 = (r << 32) / length;
    lo = (t1 << 32) | t0;
}

// Warp-uniform description of the slice of ei_start the warp has staged in shared memory for its current tile.
struct TileWin {
    const int32_t* st;   // staged copy of ei_start[base .. ), shared memory on the device
#ifdef RSQ_STAGE_DESC
    const uint2* sd;     // staged copy of ei_desc[base .. ) (experiment: the descriptor read leaves the dependent-load chain)
#endif
    int base;            // global interval index of st[0]
    int k0;              // interval containing the first coordinate served (>= base)
    int hi;              // one past the last staged interval that may be used (global index)
    int wmin;            // blocks starting before this coordinate are not served
    int lastv;           // start of interval hi-1, or INT32_MAX when hi is the end of the contig: blocks must end before it
    bool on;
};

// last index in [lo, hi_incl] whose start <= key; starts[lo] <= key is given.  `starts` is indexed by global interval index.
template <class P>
RSQ_HD int ei_search(P starts, int lo, int hi_incl, int key) {
    int len = hi_incl - lo + 1;
    while (len > 1) {
        const int half = len >> 1;
        if (starts[lo + half] <= key) lo += half;
        len -= half;
    }
    return lo;
}

struct StagedStarts {     // view of the staged slice under global indices
    const int32_t* st; int base;
    RSQ_HD int operator[](int k) const { return st[k - base]; }
};
struct GlobalStarts {
    const int32_t* p;
    RSQ_HD int operator[](int k) const { return ldro(p + k); }
};

// descriptor of a STAGED interval (base <= k < base + kStageN)
RSQ_HD uint2 staged_desc(const AnnView& ann, const TileWin& w, int k) {
#ifdef RSQ_STAGE_DESC
    (void)ann; return w.sd[k - w.base];
#else
    (void)w; return ldro(ann.ei_desc + k);
#endif
}
RSQ_HD uint32_t staged_desc_flags(const AnnView& ann, const TileWin& w, int k) {
#ifdef RSQ_STAGE_DESC
    (void)ann; return w.sd[k - w.base].y;
#else
    (void)w; return ldro(&ann.ei_desc[k].y);
#endif
}

// Interval of a block's start and the OR of the class bits of every interval the closed block [bs, be] touches
// (intersectBlock + intersectInterval, src/Expression.cpp:106-117, src/GTF.cpp:171-179).  Returns the descriptor of k(bs).
RSQ_HD uint2 block_lookup(const AnnView& ann, const TileWin& w, int contig, int bs, int be, uint32_t& cls) {
    uint2 d;
    if (w.on && bs >= w.wmin && be < w.lastv) {
        const StagedStarts s{w.st, w.base};
        const int k = ei_search(s, w.k0, w.hi - 1, bs);
        d = staged_desc(ann, w, k);
        cls |= d.y & 0xffffu;
        for (int kk = k + 1; kk < w.hi && s[kk] <= be; ++kk) cls |= staged_desc_flags(ann, w, kk) & 0xffffu;
    } else {
        const int c_hi = (int)ldro(ann.ei_off + contig + 1);
        const int bk_lo = (int)ldro(ann.eib_off + contig), bk_n = (int)ldro(ann.eib_off + contig + 1) - bk_lo - 1;
        const int b = imin(imax(bs, 0) >> kEiShift, bk_n - 1);
        const int lo = (int)ldro(ann.eib + bk_lo + b), hi = (int)ldro(ann.eib + bk_lo + b + 1);
        const GlobalStarts s{ann.ei_start};
        const int k = ei_search(s, lo, hi, bs);
        d = ldro(ann.ei_desc + k);
        cls |= d.y & 0xffffu;
        for (int kk = k + 1; kk < c_hi && s[kk] <= be; ++kk) cls |= ldro(&ann.ei_desc[kk].y) & 0xffffu;
    }
    return d;
}

// What one record contributes outside the exon tables; the kernel aggregates these across the warp.
struct RecOut {
    unsigned long long cm;   // bit k set: counter k += 1
    unsigned vflags;         // bit0: End 1 sums, bit1: End 2 sums, bit2: Mismatched Bases, bit3: Total Bases, bit4: blocks
    unsigned my_bytes;       // algorithmic bytes of this record
    uint32_t agg_gene;       // first gene this read counts toward
    bool agg_unique;
    bool fast;               // took the straight-line path of exon_metrics (statistics of the host emulation)
    unsigned long long frag_key;   // fragment-tracker key of (read name, agg_gene); 0 = none.  Inserted one tile later
    int n_fresh;             // new fragment-set entries of this read
    bool go, hq;
};

#define RSQ_CNT(o, k) ((o).cm |= 1ull << (k))

// ---- filter cascade, src/RNASeQC.cpp:254-342 ---------------------------------------------------------------------------
// Everything the cascade reads of a record fits 17 bits: nine flag bits and eight predicates against the run constants.
// cascade_index packs them; cascade_entry is the cascade as a pure function of that index (and of three run flags), so the
// kernel can read it from a table built once per context (Params::cascade_lut, 2^17 x 8 bytes) — one load instead of ~170
// branchy instructions — while the host emulation and the table builder call the very same function.
enum : uint32_t { CI_PAIRED = 1u << 0, CI_PROPER = 1u << 1, CI_UNMAP = 1u << 2, CI_MUNMAP = 1u << 3, CI_READ1 = 1u << 4, CI_SECONDARY = 1u << 5,
                  CI_QCFAIL = 1u << 6, CI_DUP = 1u << 7, CI_SUPP = 1u << 8, CI_LOW_MAPQ = 1u << 9, CI_HAS_CH = 1u << 10, CI_CHIMERIC = 1u << 11,
                  CI_HAS_NM = 1u << 12, CI_NM_OK = 1u << 13, CI_LONG_SPAN = 1u << 14, CI_FILTERED = 1u << 15, CI_TID_OOR = 1u << 16 };
constexpr uint32_t kCascadeEntries = 1u << 17;
// entry: counter bits 0..51 | vflags << 56 | (a filter tag fired) << 61 | go << 62 | hq << 63
constexpr int kCeVflagsShift = 56;
constexpr unsigned long long kCeTagged = 1ull << 61, kCeGo = 1ull << 62, kCeHq = 1ull << 63, kCeCounterMask = (1ull << 52) - 1;

RSQ_HDX uint32_t cascade_index(const Params& prm, uint32_t f, uint32_t mapq, uint32_t aux, uint32_t filter_mask, int pos, int mpos, int endpos, int nm_raw) {
    static_assert(RSQ_F_PAIRED == 1 && RSQ_F_PROPER == 2 && RSQ_F_UNMAP == 4 && RSQ_F_MUNMAP == 8 && RSQ_F_READ1 == 0x40 && RSQ_F_SECONDARY == 0x100 &&
                  RSQ_F_QCFAIL == 0x200 && RSQ_F_DUP == 0x400 && RSQ_F_SUPPLEMENTARY == 0x800, "flag bits move by shifts");
    uint32_t ix = (f & 0xfu) | ((f >> 2) & 0x10u) | ((f >> 3) & 0x1e0u);
    if (mapq < prm.mapq_threshold) ix |= CI_LOW_MAPQ;
    if (aux & RSQ_AUX_CHIMERIC_TAG) ix |= CI_HAS_CH;
    int d = pos - mpos; d = d < 0 ? -d : d;
    if ((aux & RSQ_AUX_MATE_OTHER_CONTIG) || d > prm.chimeric_distance || (prm.legacy && (aux & RSQ_AUX_TID_GT_127))) ix |= CI_CHIMERIC;   // :287
    const bool has_nm = aux & RSQ_AUX_HAS_NM;
    if (has_nm) ix |= CI_HAS_NM;
    if ((uint32_t)(has_nm ? nm_raw : 0) <= prm.base_mismatch_threshold) ix |= CI_NM_OK;                                                    // :294,330
    if ((unsigned)(endpos - pos) > 100000u) ix |= CI_LONG_SPAN;                                                                            // :275-276
    const uint32_t n_tags = prm.n_filter_tags < 8u ? prm.n_filter_tags : 8u;
    if (filter_mask & ((1u << n_tags) - 1u)) ix |= CI_FILTERED;                                                                            // :319-328
    if (aux & RSQ_AUX_TID_OUT_OF_RANGE) ix |= CI_TID_OOR;                                                                                  // :333
    return ix;
}

RSQ_HDX unsigned long long cascade_entry(bool legacy, bool exclude_chimeric, bool unpaired, uint32_t ix) {
    unsigned long long cm = 0ull;
    unsigned vflags = 0u;
    bool go = true, hq = false, tagged = false;
#define CE_CNT(k) (cm |= 1ull << (k))
    CE_CNT(RSQ_C_TOTAL_ALIGNMENTS);                                                      // RNASeQC.cpp:245
    if (ix & CI_SECONDARY) CE_CNT(RSQ_C_ALTERNATIVE_ALIGNMENTS);                         // :254
    if (ix & CI_SUPP) CE_CNT(RSQ_C_SUPPLEMENTARY_ALIGNMENTS);                            // :255
    else if (ix & CI_QCFAIL) CE_CNT(RSQ_C_FAILED_VENDOR_QC);
    else if (ix & CI_LOW_MAPQ) CE_CNT(RSQ_C_LOW_MAPPING_QUALITY);
    const bool has_ch = ix & CI_HAS_CH;
    if ((ix & CI_SUPP) && !(legacy || has_ch)) {                                         // :258-262
        CE_CNT(RSQ_C_CHIMERIC_AUTO);
        if (exclude_chimeric) go = false;
    }
    if (ix & (CI_SECONDARY | CI_QCFAIL | CI_SUPP)) go = false;                           // :263
    if (go) {
        CE_CNT(RSQ_C_UMVQ);
        if (!(ix & CI_PAIRED)) CE_CNT(RSQ_C_UNPAIRED_READS);
        if (ix & CI_UNMAP) go = false;                                                   // :268
    }
    if (go) {
        CE_CNT(RSQ_C_MAPPED_READS);
        if (ix & CI_DUP) CE_CNT(RSQ_C_MAPPED_DUPLICATE_READS); else CE_CNT(RSQ_C_MAPPED_UNIQUE_READS);
        if (legacy && (ix & CI_LONG_SPAN)) go = false;                                   // :275-276 LEGACY_MAX_READ_LENGTH
        if (go && !legacy && has_ch) {                                                   // :279-283
            if (ix & CI_READ1) CE_CNT(RSQ_C_CHIMERIC_TAG);
            if (exclude_chimeric) go = false;
        }
    }
    if (go && (ix & CI_PAIRED) && !(ix & CI_MUNMAP)) {                                   // :284-292
        if (ix & CI_READ1) CE_CNT(RSQ_C_TOTAL_MAPPED_PAIRS);
        if (ix & CI_CHIMERIC) {
            if (ix & CI_READ1) CE_CNT(RSQ_C_CHIMERIC_AUTO);
            if (exclude_chimeric) go = false;
        }
    }
    if (go) {
        if (ix & CI_HAS_NM) {                                                            // :294-316
            if (ix & CI_PAIRED) {
                if (ix & CI_READ1) {
                    CE_CNT(RSQ_C_END1_MAPPED_READS);
                    vflags |= 1u;
                    if (ix & CI_DUP) CE_CNT(RSQ_C_DUPLICATE_PAIRS); else CE_CNT(RSQ_C_UNIQUE_FRAGMENTS);
                } else {
                    CE_CNT(RSQ_C_END2_MAPPED_READS);
                    vflags |= 2u;
                }
            }
            vflags |= 4u;
        }
        vflags |= 8u;                                                                    // :317
        if (ix & CI_FILTERED) { tagged = true; go = false; }                             // :319-328 (the caller adds the tag counters)
        hq = (ix & CI_NM_OK) && (unpaired || (ix & CI_PROPER)) && !(ix & CI_LOW_MAPQ);   // :330
        if (ix & CI_TID_OOR) go = false;                                                 // :333-337
    }
    if (go) {
        if (hq) CE_CNT(RSQ_C_HIGH_QUALITY_READS); else CE_CNT(RSQ_C_LOW_QUALITY_READS);  // :340-342
        CE_CNT(RSQ_C_READS_USED);
        vflags |= 16u;                                                                   // :360
    }
#undef CE_CNT
    return cm | ((unsigned long long)vflags << kCeVflagsShift) | (tagged ? kCeTagged : 0ull) | (go ? kCeGo : 0ull) | (hq ? kCeHq : 0ull);
}

// The cascade of one record: table look-up when the context carries one (the CUDA library always does), else computed.
RSQ_HD void filter_cascade(const Params& prm, bool valid, uint32_t f, uint32_t mapq, uint32_t aux, uint32_t filter_mask, int pos, int mpos,
                           int endpos, int nm_raw, uint32_t nb, RecOut& o) {
    o.cm = 0ull; o.vflags = 0u; o.my_bytes = 0u; o.agg_gene = kNoGene; o.agg_unique = false; o.frag_key = 0ull; o.n_fresh = 0; o.fast = false;
    o.go = false; o.hq = false;
    if (!valid) return;
    const uint32_t ix = cascade_index(prm, f, mapq, aux, filter_mask, pos, mpos, endpos, nm_raw);
    const unsigned long long e = prm.cascade_lut ? ldro(prm.cascade_lut + ix) : cascade_entry(prm.legacy != 0, prm.exclude_chimeric != 0, prm.unpaired != 0, ix);
    o.cm = e & kCeCounterMask;
    if (e & kCeTagged) {
        const uint32_t n_tags = prm.n_filter_tags < 8u ? prm.n_filter_tags : 8u;
        o.cm |= (unsigned long long)(filter_mask & ((1u << n_tags) - 1u)) << RSQ_C_FILTERED_BY_TAG0;
    }
    o.vflags = (unsigned)(e >> kCeVflagsShift) & 31u;
    o.my_bytes = 40u + 8u * nb;
    o.go = (e & kCeGo) != 0ull; o.hq = (e & kCeHq) != 0ull;
}

// A block and the covering-exon list of the interval its start falls into.
struct BlockRef { int bs, be; uint32_t coff, cn; };

// first 16 bytes of a CoverEntry
struct CoverHead { int start, end; uint32_t gene_flags, slot; };
RSQ_HD CoverHead cover_head(const AnnView& ann, uint32_t j) {
    const int4 v = ldro(reinterpret_cast<const int4*>(ann.cover) + 2 * (size_t)j);
    CoverHead h; h.start = v.x; h.end = v.y; h.gene_flags = (uint32_t)v.z; h.slot = (uint32_t)v.w;
    return h;
}
// second 16 bytes of a CoverEntry: where the exon's bases start in the coverage array, and the coordinate beyond which reads find
// the exon's gene already evicted from the reference's sliding window (rsq_eit.h, gene_eviction_bounds)
struct CoverTail { unsigned long long cov_off; int late_after; };
RSQ_HD CoverTail cover_tail(const AnnView& ann, uint32_t j) {
    const int4 v = ldro(reinterpret_cast<const int4*>(ann.cover) + 2 * (size_t)j + 1);
    CoverTail t; t.cov_off = ((unsigned long long)(uint32_t)v.y << 32) | (uint32_t)v.x; t.late_after = v.z;
    return t;
}
// does the exon fully contain the block (it covers the block's start by construction) and pass the strand filter?
// partialIntersect == block length, src/Expression.cpp:341; strand test :331
RSQ_HD bool cover_admits(const CoverHead& h, const BlockRef& r, int rstrand) {
    if (h.end < r.be - 1) return false;
    return rstrand == RSQ_STRAND_UNKNOWN || rstrand == (int)((h.gene_flags >> 24) & RSQ_EXON_STRAND_MASK);
}
// is the block inside an exon of `gene` among entries [0, upto) of its covering list?
RSQ_HD bool ref_has_gene(const AnnView& ann, const BlockRef& r, uint32_t upto, uint32_t gene, int rstrand) {
    for (uint32_t j = 0; j < upto; ++j) {
        const CoverHead h = cover_head(ann, r.coff + j);
        if (cover_admits(h, r, rstrand) && (h.gene_flags & 0xffffffu) == gene) return true;
    }
    return false;
}

struct ExonOut { bool exonic, do_exon; BlockRef r0; };

// Collector::collect + BaseCoverage::commit for one (block, containing exon) pair of a gene in `last`
// (src/Expression.cpp:345-347,390-391; Metrics.cpp:59-66,106-124,257-262).
// `late`: the gene has left the reference's window (BaseCoverage::commit returns at "Gene encountered after computing coverage",
// src/Metrics.cpp:108-112): the exon count is still collected, the coverage is not.
template <class Mem>
RSQ_HD void accumulate_hit(Mem& mem, const AnnView& ann, const Params& prm, const Accum& acc, const CoverHead& h, unsigned long long cov_off,
                           bool late, int bs, uint32_t len, uint32_t length, RecOut& o) {
    unsigned long long whole, lo;
    frac_q64(ann.recip64, len, length, whole, lo);                                          // isz/length, :345
    // exact Q64.64 sum kept as three carry-free words (whole, frac>>32, frac&0xffffffff): every update is a fire-and-forget
    // RED, nothing waits for an atomic's return value
    if (whole) mem.exon_one(acc, h.slot);
    else mem.exon_frac(acc, h.slot, (uint32_t)(lo >> 32), (uint32_t)lo);
    o.my_bytes += 8u;
    if (prm.collect_coverage && !late) {                                                   // add_range as +1/-1, Metrics.cpp:257-262
        mem.cov_range(&acc.cov_diff[cov_off + (unsigned long long)(bs - h.start)], len);   // +1 at the first base, -1 behind the last
        o.my_bytes += 8u;
    }
}

// A high-quality read counts toward gene g (src/Expression.cpp:379-388).  The FIRST gene of a read is not inserted into the
// fragment tracker here: its key travels in RecOut; the kernel parks it for one tile while the table line is prefetched
// into L2 and inserts it then (the host emulation inserts it right away) — two dependent HBM round trips that no longer sit
// in this thread's chain.  Further genes of the same read (overlapping annotations, rare) are inserted on the spot.
// `late`: the gene's fragmentTracker set was erased when its gene line left the window (trimFeatures, src/Expression.cpp:86-90); the
// reads that still reach the gene afterwards fill a second set, so the key carries that epoch.
template <class Mem>
RSQ_HD void count_gene(Mem& mem, const Accum& acc, uint32_t f, uint32_t g, bool late, unsigned long long name_hash, RecOut& o) {
    // name_hash is a mixed 64-bit hash already (rsq_read.name_hash); one odd multiply is a bijection that carries every bit of
    // (name, gene, epoch) into the high half, where the table takes its home slot from
    const unsigned long long key0 = (name_hash ^ ((unsigned long long)(g + 1u) << 40) ^ (late ? 1ull << 39 : 0ull)) * 0x9E3779B97F4A7C15ull;
    const unsigned long long key = key0 ? key0 : 1ull;
    if (o.agg_gene == kNoGene) { o.agg_gene = g; o.agg_unique = !(f & RSQ_F_DUP); o.frag_key = key; }
    else {
        const bool fresh = mem.frag_insert(key);                                           // :383-387
        o.n_fresh += fresh ? 1 : 0;
        mem.add64(&acc.gene_reads[g], 1ull);
        if (fresh) mem.add64(&acc.gene_fragments[g], 1ull);
        if (!(f & RSQ_F_DUP)) mem.add64(&acc.gene_unique[g], 1ull);
    }
    o.my_bytes += 8u;
}

// ---- exonAlignmentMetrics, src/Expression.cpp:308-458 --------------------------------------------------------------------
// Mem: exon_one(acc, slot), exon_frac(acc, slot, hi, lo), cov_range(ptr, len), add64(ptr, v) for the rare direct gene-table adds,
// frag_insert(key) -> bool (true when the key was not in the fragment tracker yet).
template <class Mem>
RSQ_HD void exon_metrics(Mem& mem, const AnnView& ann, const TileWin& w, const Params& prm, const Accum& acc, uint32_t f, bool hq, bool cleared,
                         int contig, int pos0, int rstrand, const rsq_block* bp, uint32_t nb, unsigned long long name_hash, RecOut& o, ExonOut& x) {
    uint32_t length = 0, cls = 0;                                                          // extractBlocks' return, :359
    BlockRef r0 = {0, 0, 0u, 0u}, r1 = r0;
    bool globin = false, do_exon = false, fast = false;

    // ---- fast path: every block lies inside the staged slice, runs over at most one cut, and starts in an elementary
    // interval that at most ONE exon covers (collapsed annotations: practically every read).  Blocks 0 and 1 are searched in
    // lock step (two independent chains of shared-memory loads) and kept in registers; blocks 2.. (reads over three or more
    // exons) are checked one at a time and looked up once more when their turn to accumulate comes.
    if (nb >= 1 && !cleared && w.on) {
        const int2 k0 = ldro(reinterpret_cast<const int2*>(bp));
        const int2 k1 = nb >= 2 ? ldro(reinterpret_cast<const int2*>(bp + 1)) : k0;
        const int bs0 = k0.x, be0 = k0.x + k0.y, bs1 = k1.x, be1 = k1.x + k1.y;
        if (bs0 >= w.wmin && be0 < w.lastv && bs1 >= w.wmin && be1 < w.lastv) {
            const StagedStarts s{w.st, w.base};
            int lo0 = w.k0, lo1 = w.k0, len = w.hi - w.k0;
            while (len > 1) {
                const int half = len >> 1;
                if (s[lo0 + half] <= bs0) lo0 += half;
                if (s[lo1 + half] <= bs1) lo1 += half;
                len -= half;
            }
            const uint2 d0 = staged_desc(ann, w, lo0), d1 = staged_desc(ann, w, lo1);
            // a block may run over ONE cut (a read that ends on the last base of its exon always does: the reference's closed
            // end is one past it): the class bits of the next interval join in; two cuts or more take the general path
            const int na0 = lo0 + 1 < w.hi ? s[lo0 + 1] : INT32_MAX, na1 = lo1 + 1 < w.hi ? s[lo1 + 1] : INT32_MAX;
            const int nb0 = lo0 + 2 < w.hi ? s[lo0 + 2] : INT32_MAX, nb1 = lo1 + 2 < w.hi ? s[lo1 + 2] : INT32_MAX;
            uint32_t over = 0;
            if (na0 <= be0) over |= staged_desc_flags(ann, w, lo0 + 1);
            if (na1 <= be1) over |= staged_desc_flags(ann, w, lo1 + 1);
            const uint32_t cn0 = d0.y >> 16, cn1 = d1.y >> 16;
            if (nb0 > be0 && nb1 > be1 && cn0 <= 1u && cn1 <= 1u) {
                // the cover array carries one spare entry, so these loads are unconditional (used only when cn == 1)
                const CoverHead h0 = cover_head(ann, d0.x), h1 = cover_head(ann, d1.x);
                const CoverTail t0 = cover_tail(ann, d0.x), t1 = cover_tail(ann, d1.x);
                length = (uint32_t)k0.y + (nb >= 2 ? (uint32_t)k1.y : 0u);
                cls = (d0.y | d1.y | over) & 0xffffu;
                r0.bs = bs0; r0.be = be0; r0.coff = d0.x; r0.cn = cn0;
                r1.bs = bs1; r1.be = be1; r1.coff = d1.x; r1.cn = cn1;
                const bool a0 = cn0 && cover_admits(h0, r0, rstrand), a1 = cn1 && cover_admits(h1, r1, rstrand);
                const uint32_t g = h0.gene_flags & 0xffffffu;
                bool in_last = a0 && (nb == 1 || (a1 && (h1.gene_flags & 0xffffffu) == g));
                fast = true;
                // one staged block: its interval, covering exon and class bits; false when it needs the general path
                auto one_block = [&](uint32_t b, BlockRef& r, CoverHead& h, uint32_t& flags) -> bool {
                    const int2 k = ldro(reinterpret_cast<const int2*>(bp + b));
                    r.bs = k.x; r.be = k.x + k.y;
                    if (!(r.bs >= w.wmin && r.be < w.lastv)) return false;
                    const int lo = ei_search(s, w.k0, w.hi - 1, r.bs);
                    const uint2 d = staged_desc(ann, w, lo);
                    const int na = lo + 1 < w.hi ? s[lo + 1] : INT32_MAX, nn = lo + 2 < w.hi ? s[lo + 2] : INT32_MAX;
                    flags = d.y & 0xffffu;
                    if (na <= r.be) flags |= staged_desc_flags(ann, w, lo + 1) & 0xffffu;
                    r.coff = d.x; r.cn = d.y >> 16;
                    if (nn <= r.be || r.cn > 1u) return false;
                    h = cover_head(ann, d.x);
                    return true;
                };
                for (uint32_t b = 2; b < nb && fast; ++b) {
                    BlockRef r; CoverHead h; uint32_t flags = 0;
                    fast = one_block(b, r, h, flags);
                    if (!fast) break;
                    length += (uint32_t)(r.be - r.bs);
                    cls |= flags;
                    in_last = in_last && r.cn && cover_admits(h, r, rstrand) && (h.gene_flags & 0xffffffu) == g;
                }
                if (fast && in_last) {                                                     // g is in `last` (:368-377)
                    do_exon = true;                                                        // :393
                    if ((h0.gene_flags >> 24) & kCoverGlobin) globin = true;               // :396-398
                    if (hq) {
                        const bool late = pos0 > t0.late_after;                            // a property of the gene: the same for every block
                        if (length > 0) count_gene(mem, acc, f, g, late, name_hash, o);    // queryGene: some coverage > 0, :380
                        if (k0.y) accumulate_hit(mem, ann, prm, acc, h0, t0.cov_off, late, bs0, (uint32_t)k0.y, length, o);
                        if (nb >= 2 && k1.y) accumulate_hit(mem, ann, prm, acc, h1, t1.cov_off, late, bs1, (uint32_t)k1.y, length, o);
                        for (uint32_t b = 2; b < nb; ++b) {
                            BlockRef r; CoverHead h; uint32_t flags = 0;
                            one_block(b, r, h, flags);
                            if (r.be != r.bs) accumulate_hit(mem, ann, prm, acc, h, cover_tail(ann, r.coff).cov_off, late, r.bs, (uint32_t)(r.be - r.bs), length, o);
                        }
                    }
                }
                if (!fast) { length = 0; cls = 0; }                                        // the general path starts over
            }
        }
    }

    // ---- general path: any number of blocks, intervals and covering exons ------------------------------------------------
    if (!fast) [[unlikely]] {
        auto ref_of = [&](uint32_t b) -> BlockRef {
            const int2 blk = ldro(reinterpret_cast<const int2*>(bp + b));
            BlockRef r; r.bs = blk.x; r.be = blk.x + blk.y;
            uint32_t dummy = 0;
            const uint2 d = block_lookup(ann, w, contig, r.bs, r.be, dummy);
            r.coff = d.x; r.cn = d.y >> 16;
            return r;
        };
        for (uint32_t b = 0; b < nb; ++b) {                                                // :325-361
            const int2 blk = ldro(reinterpret_cast<const int2*>(bp + b));
            const int bs = blk.x, be = blk.x + blk.y;                                      // closed end = start+len, :45
            length += (uint32_t)blk.y;
            if (cleared) continue;
            const uint2 d = block_lookup(ann, w, contig, bs, be, cls);
            if (b == 0) { r0.bs = bs; r0.be = be; r0.coff = d.x; r0.cn = d.y >> 16; }
            else if (b == 1) { r1.bs = bs; r1.be = be; r1.coff = d.x; r1.cn = d.y >> 16; }
        }
        if (nb >= 1 && !cleared) {                                                         // :363-394
            // `last` = intersection over blocks of {genes with an exon that fully contains the block}: enumerate the genes of
            // block 0 (deduplicated by re-scanning its short covering list) and test each against every other block.
            for (uint32_t j = 0; j < r0.cn; ++j) {
                const CoverHead h0 = cover_head(ann, r0.coff + j);
                if (!cover_admits(h0, r0, rstrand)) continue;                              // isz == block length, :341
                const uint32_t g = h0.gene_flags & 0xffffffu;
                if (ref_has_gene(ann, r0, j, g, rstrand)) continue;                        // an earlier entry already gave this gene
                bool in_all = true;
                for (uint32_t b = 1; b < nb && in_all; ++b) {
                    if (b == 1) in_all = ref_has_gene(ann, r1, r1.cn, g, rstrand);
                    else { const BlockRef r = ref_of(b); in_all = ref_has_gene(ann, r, r.cn, g, rstrand); }
                }
                if (!in_all) continue;
                // g is in `last` (:377)
                do_exon = true;                                                            // :393
                if ((h0.gene_flags >> 24) & kCoverGlobin) globin = true;                   // :396-398
                if (!hq) continue;
                const bool late = pos0 > cover_tail(ann, r0.coff + j).late_after;
                if (length > 0) count_gene(mem, acc, f, g, late, name_hash, o);            // queryGene: some coverage > 0, :380
                // Collector::collect + BaseCoverage::commit for gene g: every (block, containing exon of g) pair
                for (uint32_t b = 0; b < nb; ++b) {
                    const BlockRef r = b == 0 ? r0 : (b == 1 ? r1 : ref_of(b));
                    const uint32_t len = (uint32_t)(r.be - r.bs);
                    if (len == 0) continue;                                                // Collector::add keeps coverage > 0, Metrics.cpp:51
                    for (uint32_t jj = 0; jj < r.cn; ++jj) {
                        const CoverHead h = cover_head(ann, r.coff + jj);
                        if (!cover_admits(h, r, rstrand) || (h.gene_flags & 0xffffffu) != g) continue;
                        accumulate_hit(mem, ann, prm, acc, h, cover_tail(ann, r.coff + jj).cov_off, late, r.bs, len, length, o);
                    }
                }
            }
        }
    }

    if (rstrand != RSQ_STRAND_UNKNOWN) cls &= (EI_GENE | EI_GENE_RIBO | EI_EXON | EI_EXON_RIBO) << rstrand;   // :331
    const bool intragenic = cls & EI_ANY_GENE;                                              // :352-354
    const bool exonic = cls & EI_ANY_EXON;                                                  // :335-337
    const bool plus = cls & ((EI_GENE | EI_EXON) << RSQ_STRAND_FWD);                        // :332
    const bool minus = cls & ((EI_GENE | EI_EXON) << RSQ_STRAND_REV);                       // :333
    const bool ribosomal = cls & (EI_ANY_GENE_RIBO | EI_ANY_EXON_RIBO);                     // :358
    if (nb >= 1 && !globin) {                                                              // :399-404
        RSQ_CNT(o, RSQ_C_NON_GLOBIN_READS);
        if (f & RSQ_F_DUP) RSQ_CNT(o, RSQ_C_NON_GLOBIN_DUPLICATE_READS);
    }
    if (!exonic) {                                                                         // :407-423
        if (intragenic) {
            RSQ_CNT(o, RSQ_C_INTRONIC_READS); RSQ_CNT(o, RSQ_C_INTRAGENIC_READS);
            if (hq) { RSQ_CNT(o, RSQ_C_HQ_INTRONIC_READS); RSQ_CNT(o, RSQ_C_HQ_INTRAGENIC_READS); }
        } else {
            RSQ_CNT(o, RSQ_C_INTERGENIC_READS);
            if (hq) RSQ_CNT(o, RSQ_C_HQ_INTERGENIC_READS);
        }
    } else if (do_exon) {                                                                  // :424-433
        RSQ_CNT(o, RSQ_C_EXONIC_READS); RSQ_CNT(o, RSQ_C_INTRAGENIC_READS);
        if (hq) { RSQ_CNT(o, RSQ_C_HQ_EXONIC_READS); RSQ_CNT(o, RSQ_C_HQ_INTRAGENIC_READS); }
    } else {                                                                               // :434-441
        RSQ_CNT(o, RSQ_C_AMBIGUOUS_READS);
        if (hq) RSQ_CNT(o, RSQ_C_HQ_AMBIGUOUS_READS);
    }
    if (ribosomal) RSQ_CNT(o, RSQ_C_RRNA_READS);                                           // :442
    if ((minus != plus) && (prm.unpaired || (f & RSQ_F_PAIRED))) {                         // :445-457
        const bool sense = (f & RSQ_F_REVERSE) ? minus : plus;
        if (prm.unpaired || (f & RSQ_F_READ1)) { if (sense) RSQ_CNT(o, RSQ_C_END1_SENSE); else RSQ_CNT(o, RSQ_C_END1_ANTISENSE); }
        else { if (sense) RSQ_CNT(o, RSQ_C_END2_SENSE); else RSQ_CNT(o, RSQ_C_END2_ANTISENSE); }
    }
    x.exonic = exonic; x.do_exon = do_exon; x.r0 = r0;
    o.fast = fast;
}

RSQ_HD int read_strand(const Params& prm, uint32_t f) {                                    // feature_strand, src/Expression.cpp:119-125
    if (!prm.strand_mode) return RSQ_STRAND_UNKNOWN;
    bool target = f & RSQ_F_REVERSE;
    if ((prm.strand_mode == 1) != ((f & RSQ_F_READ1) != 0)) target = !target;
    return target ? RSQ_STRAND_REV : RSQ_STRAND_FWD;
}

}  // namespace rsqd
