// rsq_kernels.cu — hand-written sm_100a kernels of the RNA-SeQC per-read hot path.
//
// Compile-time ablation switches (never defined in the product build; used for the time breakdown in DESIGN.md §4):
//   RSQ_AB_NO_FRAGSET, RSQ_AB_NO_ACCUM, RSQ_AB_NO_EPILOGUE, RSQ_AB_CASCADE_ONLY, RSQ_AB_NO_STAGE, RSQ_AB_NO_PREFETCH,
//   RSQ_AB_SCATTER, RSQ_AB_NO_EXON, RSQ_AB_NO_COV, RSQ_NO_AGG_EXON, RSQ_NO_AGG_COV, RSQ_COV_TWO_MATCHES
//
// K_hot  : one pass over a batch of decoded alignment records (rsq_read + rsq_block): the flag/quality/
//          mismatch filter cascade of src/RNASeQC.cpp:254-342, interval intersection of every aligned block
//          with the elementary-interval index of the annotation (src/Expression.cpp:106-117, src/GTF.cpp:171-186),
//          exon/intron classification and count accumulation (src/Expression.cpp:308-458, src/Metrics.cpp:49-66),
//          per-base coverage as a difference array (src/Metrics.cpp:96-124,257-262), fragment-tracker set
//          (src/Metrics.cpp:22, src/Expression.cpp:383-387) and, in BED mode, fragment-size candidates
//          (src/Expression.cpp:482-516).  The per-record logic lives in rsq_hot_core.cuh (shared with the host emulation
//          the CPU tests check against the oracle); this file holds what needs a warp: staging, aggregation, counters.
// There is no dense contraction anywhere on this path: no tensor cores.  The work is HBM/L2-bound integer work.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "rsq_device.h"
#include "rsq_hot_core.cuh"

namespace rsqd {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kHotThreads = 256;
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
// Open-addressing insert; returns true when the key was not present.  The table has ANY capacity (not a power of two: it is
// zeroed on every reset, so it is sized 2 x the records, not rounded up to the next power): home slot = high half of
// key x capacity, linear probing with wrap-around.  (A bucketed variant — four slots = one 32-byte sector fetched and examined
// together — was measured slower: 14.2 vs 12.6 ms per 200 M records.)
__device__ __forceinline__ unsigned long long set_home(unsigned long long key, unsigned long long cap) { return __umul64hi(key, cap); }
// `seen`: what an EARLIER load of the key's home slot returned (the kernel fetches it asynchronously one tile ahead).  A stale value
// is harmless: slots only ever change from empty to a key, so a stale "empty" is corrected by the CAS and a key is never stale.
__device__ __forceinline__ bool set_insert_seen(unsigned long long* __restrict__ tab, unsigned long long cap, unsigned long long key,
                                                unsigned long long seen, unsigned int* err, unsigned int err_bit) {
    unsigned long long slot = set_home(key, cap);
    unsigned long long cur = seen;
    for (int probe = 0; probe < 65536; ++probe) {       // the host keeps the table below 80 % (rsq_capi.cu reserve_frag_capacity)
        if (cur == key) return false;
        if (cur == 0ull) {
            cur = atomicCAS(&tab[slot], 0ull, key);
            if (cur == 0ull) return true;
            if (cur == key) return false;
        }
        slot = slot + 1 == cap ? 0ull : slot + 1;
        cur = tab[slot];
    }
    atomicOr(err, err_bit);
    return false;
}
__device__ __forceinline__ bool set_insert(unsigned long long* __restrict__ tab, unsigned long long cap, unsigned long long key,
                                           unsigned int* err, unsigned int err_bit) {
    // plain load first: measured 30 % faster than CAS-first on B200 (an L2 atomic on a missing line is far more
    // expensive than a load, and second mates find their key without any atomic)
    return set_insert_seen(tab, cap, key, tab[set_home(key, cap)], err, err_bit);
}

// ---- per-warp staging of the interval starts into shared memory (TMA bulk copy) ----------------------------------------
// A tile is 32 coordinate-sorted records, so their blocks fall into a short run of elementary intervals: the starts of the
// kStageN intervals from the tile's lowest coordinate on are copied ONCE per tile into the warp's private slice of shared
// memory by cp.async.bulk (1-D TMA, completion on an mbarrier; the filter cascade runs while the copy flies).  Every block
// that ends before the last staged start is then located by a fixed-depth binary search in shared memory; blocks beyond it
// (the far side of a long intron) go through the 1 kb direct-address table and search one bucket in L2.
#ifndef RSQ_STAGE_N
#define RSQ_STAGE_N 64
#endif
constexpr int kStageN = RSQ_STAGE_N;    // a multiple of 4 (16-byte bulk-copy granularity), at most kEiPad
struct __align__(16) WarpStage {
    int32_t starts[kStageN];
#ifdef RSQ_STAGE_DESC
    uint2 descs[kStageN];             // experiment: the descriptors of the staged intervals ride along (second bulk copy, same barrier)
#endif
    unsigned long long fkey[32];      // fragment-tracker keys parked for one tile (0 = none), one slot per lane
    uint32_t fgene[32];               // the gene each parked key counts toward
    uint32_t fpeers[32];              // the lanes of that tile that counted toward the same gene (MATCH.ANY of the epilogue)
    unsigned long long fval[32];      // what the home slot of each parked key held (cp.async, lands during the next tile)
    unsigned long long mbar;
    unsigned long long pad;
};

#if !defined(RSQ_CUDA_EMU)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mbar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    } while (!ok);
}
// the staged slice of one tile: issued by ONE lane, completion on the warp's mbarrier
__device__ __forceinline__ void stage_issue(WarpStage* st, uint32_t mbar, const int32_t* starts, const uint2* descs) {
#ifdef RSQ_STAGE_DESC
    mbar_expect_tx(mbar, (uint32_t)(kStageN * 12));
    tma_bulk_g2s(smem_u32(st->starts), starts, (uint32_t)(kStageN * 4), mbar);   // both arrays are padded
    tma_bulk_g2s(smem_u32(st->descs), descs, (uint32_t)(kStageN * 8), mbar);
#else
    (void)descs;
    mbar_expect_tx(mbar, (uint32_t)(kStageN * 4));
    tma_bulk_g2s(smem_u32(st->starts), starts, (uint32_t)(kStageN * 4), mbar);   // the array is padded
#endif
}
__device__ __forceinline__ void stage_wait(uint32_t mbar, uint32_t parity) { mbar_wait(mbar, parity); }
// 8 bytes of global memory into shared memory without a register in between (LDGSTS); home_wait before they are read
__device__ __forceinline__ void home_fetch(unsigned long long* dst_smem, const unsigned long long* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void home_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void home_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t peek_u32(const void* p) { uint32_t v; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
// CPU test harness (tests/cudaemu): the bulk copy is a memcpy by the issuing lane and its mbarrier a warp barrier; the asynchronous
// 8-byte copy is a load; prefetches are nothing
inline uint32_t smem_u32(const void*) { return 0u; }
inline void mbar_init(uint32_t) {}
inline void stage_issue(WarpStage* st, uint32_t, const int32_t* starts, const uint2* descs) {
    memcpy(st->starts, starts, kStageN * 4);
#ifdef RSQ_STAGE_DESC
    memcpy(st->descs, descs, kStageN * 8);
#else
    (void)descs;
#endif
}
inline void stage_wait(uint32_t, uint32_t) { __syncwarp(); }
inline void home_fetch(unsigned long long* dst, const unsigned long long* src) { *dst = __atomic_load_n(src, __ATOMIC_SEQ_CST); }
inline void home_commit() {}
inline void home_wait() {}
inline uint32_t peek_u32(const void* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline void prefetch_l2(const void*) {}
#endif

// Inserts the keys a warp parked one tile ago (their table lines were prefetched into L2 then) and counts the distinct
// (read name, gene) pairs: geneFragmentCounts, src/Expression.cpp:383-387.  Coordinate-sorted input => the lanes mostly share
// one gene: one RED per distinct gene.  Returns the number of new entries (warp-uniform).
template <class Stage>
__device__ __forceinline__ int resolve_parked(Stage* st, const Accum& acc, unsigned lane) {
    const unsigned long long key = st->fkey[lane];
    const uint32_t g = st->fgene[lane];
    const unsigned peers = st->fpeers[lane];
#ifdef RSQ_AB_NO_FRAGSET
    const bool fresh = (key & 2ull) != 0ull;   // ablation only: no table traffic (results are wrong)
#else
#ifdef RSQ_FRAG_PREFETCH_ONLY
    const bool fresh = key != 0ull && set_insert(acc.frag_set, acc.frag_cap, key, acc.error_flag, 1u);
#else
    home_wait();                                                        // the home slots fetched one tile ago have landed
    const bool fresh = key != 0ull && set_insert_seen(acc.frag_set, acc.frag_cap, key, st->fval[lane], acc.error_flag, 1u);
#endif
#endif
    const unsigned all_fresh = __ballot_sync(kFull, fresh);
    // the gene groups of that tile are known from its epilogue: the lowest lane of each group adds the group's new entries
    if (g != kNoGene && lane == (unsigned)(__ffs(peers) - 1)) {
        const int n = __popc(peers & all_fresh);
        if (n) atomicAdd(&acc.gene_fragments[g], (unsigned long long)n);
    }
    return __popc(all_fresh);
}

// One slot of a list for every lane converged at the call: the lowest lane adds the group's size to the counter (one atomic with
// a return value per warp instead of one per lane on a single hot address), every lane takes base + its rank.
__device__ __forceinline__ unsigned int warp_claim(unsigned int* counter) {
    const unsigned m = __activemask();
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(m) - 1;
    unsigned int base = 0;
    if (lane == (unsigned)leader) base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + (unsigned)__popc(m & ((1u << lane) - 1u));
}

// Memory side effects of the per-record logic on the device: fire-and-forget REDs and the fragment-tracker CAS.
// RSQ_AB_SCATTER is a diagnostic (results wrong by construction): same instruction stream, but every RED goes
// to a thread-unique address.
struct DevMem {
    const Accum& acc;
    unsigned long long n_exons, cov_len;
    // Opportunistic warp aggregation: coordinate-sorted input makes the lanes that reach an accumulation together hit the same
    // exon words and the same few coverage words, and the L2 atomic units serialise per address (measured: the REDs were 65 %
    // of the kernel).  The lanes converged at the call (any subset is correct) group by target with one MATCH.ANY; one lane
    // per group issues the RED with the group's sum (a population count, or REDUX over the group's mask).
    // exonCounts[slot] += 1 (a read that lies in one exon with all its aligned bases): the group's size is the sum
    __device__ __forceinline__ void exon_one(const Accum& a, uint32_t slot) const {
#if defined(RSQ_AB_SCATTER)       // needs n_exons >= 2^18 (the GENCODE-scale bench annotation)
        atomicAdd(a.exon_whole + ((slot + (blockIdx.x * blockDim.x + threadIdx.x) * 37u) & 0x3ffffu), 1ull);
#elif defined(RSQ_AB_NO_ACCUM) || defined(RSQ_AB_NO_EXON)
#elif defined(RSQ_NO_AGG_EXON)
        atomicAdd(&a.exon_whole[slot], 1ull);
#else
        const unsigned peers = __match_any_sync(__activemask(), slot);
        if ((threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&a.exon_whole[slot], (unsigned long long)__popc(peers));
#endif
    }
    // exonCounts[slot] += frac as the two 32-bit halves of its Q0.64 image: one MATCH.ANY on the slot serves both words; the
    // halves are summed as 16-bit quarters so that no REDUX can overflow
    __device__ __forceinline__ void exon_frac(const Accum& a, uint32_t slot, uint32_t hi, uint32_t lo) const {
#if defined(RSQ_AB_SCATTER)
        const uint32_t s2 = (slot + (blockIdx.x * blockDim.x + threadIdx.x) * 37u) & 0x3ffffu;
        atomicAdd(&a.exon_fhi[s2], (unsigned long long)hi); atomicAdd(&a.exon_flo[s2], (unsigned long long)lo);
#elif defined(RSQ_AB_NO_ACCUM) || defined(RSQ_AB_NO_EXON)
#elif defined(RSQ_NO_AGG_EXON)
        atomicAdd(&a.exon_fhi[slot], (unsigned long long)hi); atomicAdd(&a.exon_flo[slot], (unsigned long long)lo);
#else
        const unsigned peers = __match_any_sync(__activemask(), slot);
        if ((peers & (peers - 1u)) == 0u) { atomicAdd(&a.exon_fhi[slot], (unsigned long long)hi); atomicAdd(&a.exon_flo[slot], (unsigned long long)lo); return; }
        const unsigned h0 = __reduce_add_sync(peers, hi & 0xffffu), h1 = __reduce_add_sync(peers, hi >> 16);
        const unsigned l0 = __reduce_add_sync(peers, lo & 0xffffu), l1 = __reduce_add_sync(peers, lo >> 16);
        if ((threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1)) {
            atomicAdd(&a.exon_fhi[slot], ((unsigned long long)h1 << 16) + h0);
            atomicAdd(&a.exon_flo[slot], ((unsigned long long)l1 << 16) + l0);
        }
#endif
    }
    // direct adds to the gene tables (second and further genes of one read: rare)
    __device__ __forceinline__ void add64(unsigned long long* p, unsigned long long v) const {
#if !defined(RSQ_AB_NO_ACCUM)
        atomicAdd(p, v);
#endif
    }
    // +1 at p, -1 at p + len (add_range as a difference array).  Reads with the same start AND length share both words: one
    // MATCH.ANY on the pair groups them (RSQ_COV_TWO_MATCHES: one per word, merges more but costs a second match).
    __device__ __forceinline__ void cov_range(unsigned int* p, unsigned int len) const {
#if defined(RSQ_AB_SCATTER)       // needs cov_len >= 2^25
        const unsigned gt = blockIdx.x * blockDim.x + threadIdx.x;
        unsigned int* q = acc.cov_diff + (((unsigned)(p - acc.cov_diff) + gt * 4099u) & 0x1ffffffu);
        atomicAdd(q, 1u); atomicAdd(q + len, 0xffffffffu);
#elif defined(RSQ_AB_NO_ACCUM) || defined(RSQ_AB_NO_COV)
#elif defined(RSQ_NO_AGG_COV)
        atomicAdd(p, 1u); atomicAdd(p + len, 0xffffffffu);
#elif defined(RSQ_COV_TWO_MATCHES)
        const unsigned m = __activemask();
        const unsigned pa = __match_any_sync(m, (unsigned long long)p), pb = __match_any_sync(m, (unsigned long long)(p + len));
        const unsigned lane = threadIdx.x & 31u;
        if (lane == (unsigned)(__ffs(pa) - 1)) atomicAdd(p, (unsigned)__popc(pa));
        if (lane == (unsigned)(__ffs(pb) - 1)) atomicAdd(p + len, 0u - (unsigned)__popc(pb));
#else
        if (len >> 24 || (unsigned long long)(p - acc.cov_diff) >> 40) { atomicAdd(p, 1u); atomicAdd(p + len, 0xffffffffu); return; }   // the pair does not fit one key
        const unsigned m = __activemask();
        const unsigned long long key = (unsigned long long)(p - acc.cov_diff) | ((unsigned long long)len << 40);
        const unsigned peers = __match_any_sync(m, key);
        if ((threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1)) {
            const unsigned n = (unsigned)__popc(peers);
            atomicAdd(p, n); atomicAdd(p + len, 0u - n);
        }
#endif
    }
    __device__ __forceinline__ bool frag_insert(unsigned long long key) const {
#ifdef RSQ_AB_NO_FRAGSET
        return (key & 1ull) != 0;   // ablation only: no hash-set traffic (results are wrong)
#else
        return set_insert(acc.frag_set, acc.frag_cap, key, acc.error_flag, 1u);
#endif
    }
};

#ifndef RSQ_HOT_MIN_BLOCKS
#define RSQ_HOT_MIN_BLOCKS 4
#endif
}  // namespace
}  // namespace rsqd
#include "rsq_legacy.cuh"
namespace rsqd {
namespace {

template <bool kBed, bool kLegacy, bool kGc>
__global__ void __launch_bounds__(kHotThreads, RSQ_HOT_MIN_BLOCKS)
hot_path_kernel(const rsq_read* __restrict__ reads, const rsq_block* __restrict__ blocks, uint32_t n_reads, uint32_t n_blocks,
                const AnnView ann, const BedView bed, const Accum acc, const Params prm) {   // 32-bit indices: a launch is one sub-batch
    __shared__ unsigned long long s_cnt[RSQ_N_COUNTERS];
    __shared__ unsigned long long s_bytes;
    __shared__ WarpStage s_stage[kHotThreads / 32];
    for (int k = threadIdx.x; k < RSQ_N_COUNTERS; k += blockDim.x) s_cnt[k] = 0ull;
    if (threadIdx.x == 0) s_bytes = 0ull;

    const unsigned lane = threadIdx.x & 31u;
    WarpStage* st = &s_stage[threadIdx.x >> 5];
    const uint32_t mbar = smem_u32(&st->mbar);
    if (lane == 0) mbar_init(mbar);
    st->fkey[lane] = 0ull; st->fgene[lane] = kNoGene; st->fpeers[lane] = 0u;
    __syncthreads();
    uint32_t phase = 0;

    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    const uint32_t n_tiles = (n_reads + 31u) >> 5;
    // Persistent warps: warp w of the grid takes 32-record tiles w, w + W, w + 2W, ... No CTA barrier inside the loop.
    uint32_t tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    unsigned long long own_lo = 0ull, own_hi = 0ull, own_bytes = 0ull;   // lane-owned counters (see the epilogue)
#if defined(RSQ_PEEK2) && !defined(RSQ_AB_NO_PREFETCH)
    uint32_t peek_boff = 0;                                              // block offset of this lane's record in the NEXT tile
    if ((size_t)(tile + warps_total) * 32u + lane < n_reads)
        peek_boff = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(reads + ((size_t)(tile + warps_total) * 32u + lane)) + 40));
#endif
  for (; tile < n_tiles; tile += warps_total) {
#define RSQ_TILE_NEXT (tile + warps_total)
    const uint32_t i = tile * 32u + lane;
    const bool valid = i < n_reads;
    int4 q0 = make_int4(0, 0, 0, 0), q1 = q0, q2 = q0;
    // records stream through once: evict-first loads; the NEXT tile of this warp is prefetched into L2 meanwhile
    if (valid) {
        const int4* rp = reinterpret_cast<const int4*>(reads + i);
        q0 = __ldcs(rp); q1 = __ldcs(rp + 1); q2 = __ldcs(rp + 2);
    }
    // The demand load of the NEXT tile's block offsets pulls that tile's record lines toward the SM now (it is consumed at
    // the end of this iteration, one whole tile of work later) and tells where its aligned blocks start: they are
    // prefetched into L2 below, so that neither array is first touched by a waiting warp.
    uint32_t next_boff = 0;
    const bool has_next = RSQ_TILE_NEXT * 32u + lane < n_reads;               // no wrap: n_reads < 2^31 (launcher)
#ifndef RSQ_AB_NO_PREFETCH
#ifdef RSQ_PEEK2
    // Two tiles ahead: the block offsets asked for during the PREVIOUS tile are used now (they had a whole tile of work to
    // arrive: the prefetch below never waits), and the records of the tile after the next are pulled into L2 whole (twelve
    // 128-byte lines per tile) instead of the two sectors out of three that the offset loads happen to touch.
    next_boff = peek_boff;
    {
        const size_t t2 = (size_t)tile + 2u * (size_t)warps_total;
        if (t2 * 32u + lane < n_reads) peek_boff = peek_u32(reinterpret_cast<const char*>(reads + (t2 * 32u + lane)) + 40);
        if (lane < 12u && t2 * 32u < n_reads) prefetch_l2(reinterpret_cast<const char*>(reads + t2 * 32u) + 128u * lane);
    }
#else
    if (has_next) next_boff = peek_u32(reinterpret_cast<const char*>(reads + ((size_t)RSQ_TILE_NEXT * 32u + lane)) + 40);
#endif
#endif

    // ---- stage this tile's run of interval starts (warp-uniform) ----------------------------------------------------------
    TileWin w;
    w.st = st->starts; w.base = 0; w.k0 = 0; w.hi = 0; w.wmin = 0; w.lastv = 0; w.on = false;
#ifdef RSQ_STAGE_DESC
    w.sd = st->descs;
#endif
    bool to_end = false;
#ifndef RSQ_AB_NO_STAGE
    if constexpr (!kLegacy) {
        const int my_contig = (int)(int16_t)((uint32_t)q1.w & 0xffffu);
        const int pmin = __reduce_min_sync(kFull, valid ? q0.x : INT32_MAX);
        const int cmin = __reduce_min_sync(kFull, valid ? my_contig : INT32_MAX);
        const int cmax = __reduce_max_sync(kFull, valid ? my_contig : INT32_MIN);
        if (cmin == cmax && cmin >= 0 && (uint32_t)cmin < ann.n_contigs && pmin >= 0 && pmin < INT32_MAX) {
            const int c_hi = (int)__ldg(&ann.ei_off[cmin + 1]);
            const int bk_lo = (int)__ldg(&ann.eib_off[cmin]), bk_n = (int)__ldg(&ann.eib_off[cmin + 1]) - bk_lo - 1;
            w.wmin = pmin + 1;
            w.k0 = (int)__ldg(&ann.eib[bk_lo + min(w.wmin >> kEiShift, bk_n - 1)]);
            w.base = w.k0 & ~3;                                          // bulk copies are 16-byte granular
            w.hi = min(w.base + kStageN, c_hi);
            to_end = w.hi == c_hi;
            w.on = true;
            __syncwarp();                                                // every lane is done with the previous tile's copy
            if (lane == 0) stage_issue(st, mbar, ann.ei_start + w.base, ann.ei_desc + w.base);
        }
    }
#endif

    // ---- filter cascade, RNASeQC.cpp:254-342 (independent of the annotation: overlaps the bulk copy) ------------------------
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
    if ((unsigned long long)block_off + nb > n_blocks) nb = 0;  // defensive: never read out of the submitted array
    (void)isize; (void)l_qseq;
    RecOut o;
    filter_cascade(prm, valid, f, mapq, aux, filter_mask, pos, mpos, endpos, nm_raw, nb, o);

    if (w.on) {                                                                        // the starts have landed in shared memory
        stage_wait(mbar, phase); phase ^= 1u;
        w.lastv = to_end ? INT32_MAX : st->starts[w.hi - 1 - w.base];
    }

#ifdef RSQ_AB_CASCADE_ONLY
    o.go = false;   // ablation only
#endif
    const bool hq = o.hq;
    if (o.go) {
        {
            // ---- exonAlignmentMetrics, Expression.cpp:308-458 ----------------------------------------------------
            const bool cleared = (aux & RSQ_AUX_CONTIG_DROPPED) || contig < 0 || (uint32_t)contig >= ann.n_contigs;
            const int rstrand = read_strand(prm, f);                                   // feature_strand, :119-125
            const rsq_block* bp = blocks + block_off;
            if constexpr (kLegacy) {                                                   // :364
                LegacyRead lr;
                lr.contig = contig; lr.pos = pos; lr.endpos = endpos; lr.l_qseq = l_qseq; lr.rstrand = rstrand;
                lr.flag = f; lr.nb = nb; lr.cleared = cleared; lr.hq = hq; lr.name_hash = name_hash; lr.bp = bp;
                legacy_metrics(ann, acc, prm, lr, o.cm, o.n_fresh, o.my_bytes);
            } else {
            const DevMem mem{acc, (unsigned long long)ann.n_exons, prm.cov_len_diag};
            ExonOut x;
            exon_metrics(mem, ann, w, prm, acc, f, hq, cleared, contig, pos, rstrand, bp, nb, name_hash, o, x);

            // ---- fragment GC content candidate, Expression.cpp:459-467 --------------------------------------------------
            if constexpr (kGc) {
                const unsigned aisz = (unsigned)(isize < 0 ? -(long long)isize : (long long)isize);
                if (hq && x.exonic && x.do_exon && nb == 1 && aisz > 100u && aisz < 1000u && !cleared && __ldg(&prm.ref_present[contig])) {
                    // alignedExons (:319,348): exons of any gene that fully contain the block, read strand respected
                    int n_full = 0;
                    uint32_t the_slot = 0;
                    for (uint32_t j = 0; j < x.r0.cn; ++j) {
                        const CoverHead h = cover_head(ann, x.r0.coff + j);
                        if (!cover_admits(h, x.r0, rstrand)) continue;
                        ++n_full; the_slot = h.slot;
                    }
                    if (n_full == 1) {
                        const unsigned long long key = name_hash ? name_hash : 1ull;
                        unsigned long long slot = mix64(key) & acc.gc_mask;
                        bool placed = false;
                        for (int probe = 0; probe < 4096 && !placed; ++probe) {
                            unsigned long long cur = acc.gc_keys[slot];
                            if (cur == 0ull) cur = atomicCAS(&acc.gc_keys[slot], 0ull, key);
                            if (cur == 0ull || cur == key) placed = true; else slot = (slot + 1) & acc.gc_mask;
                        }
                        if (!placed) atomicOr(acc.error_flag, 2u);
                        else {
                            atomicMax(&acc.gc_first[slot], mate_tag(prm.mate_round, 1u, i));             // round 1: the earliest candidate of the name
                            const unsigned int ci = warp_claim(acc.gc_cand_count);
                            if (ci < acc.gc_cand_capacity) {
                                GcCandidate c;
                                c.name_hash = key; c.local_idx = (uint32_t)i; c.exon_slot = the_slot; c.endpos = endpos; c.l_qseq = l_qseq;
                                c.contig = contig; c.cond_ok = pos != mpos ? 1u : 0u; c.slot = (uint32_t)slot; c.pad = 0u;
                                acc.gc_cand[ci] = c;
                            } else atomicOr(acc.error_flag, 4u);
                        }
                    }
                }
            }
            }  // !kLegacy

            // ---- fragmentSizeMetrics candidate, RNASeQC.cpp:372 + Expression.cpp:482-516 ----------------------------
            if (kBed) {
                if (hq && (f & RSQ_F_PAIRED) && contig >= 0 && __ldg(&bed.present[contig])) {
                    const int b_lo = (int)__ldg(&bed.off[contig]), b_hi = (int)__ldg(&bed.off[contig + 1]);
                    bool same = nb > 0;
                    uint32_t name = 0;
                    const int bc_hi = (int)__ldg(&bed.ei_off[contig + 1]);
                    const int bbk_lo = (int)__ldg(&bed.eib_off[contig]), bbk_n = (int)__ldg(&bed.eib_off[contig + 1]) - bbk_lo - 1;
                    for (uint32_t b = 0; b < nb && same; ++b) {                        // :490-507
                        const int2 blk = __ldg(reinterpret_cast<const int2*>(bp + b));
                        const int bs = blk.x, be = blk.x + blk.y;
                        int hits = 0; int2 x = make_int2(0, -1); uint32_t nid = 0;
                        // through the BED's elementary-interval index: a block between two cuts intersects exactly the intervals that
                        // cover it (one table read, a search inside a 1 kb bucket, one descriptor, one entry)
                        const int bb = min(max(bs, 0) >> kEiShift, bbk_n - 1);
                        const GlobalStarts gs{bed.ei_start};
                        const int k = ei_search(gs, (int)__ldg(&bed.eib[bbk_lo + bb]), (int)__ldg(&bed.eib[bbk_lo + bb + 1]), bs);
                        if (k + 1 >= bc_hi || gs[k + 1] > be) {
                            const uint2 d = __ldg(&bed.ei_desc[k]);
                            hits = (int)(d.y >> 16);
                            if (hits == 1) {
                                const int4 ce = __ldg(reinterpret_cast<const int4*>(bed.cover) + 2 * (size_t)d.x);
                                x = make_int2(ce.x, ce.y); nid = (uint32_t)ce.w;
                            }
                        } else {                                                       // a cut inside the block: count over the sorted list
                            const int hi = upper_bound_start(bed.se, b_lo, b_hi, be);
                            const int lo = lower_bound_pmax(bed.pmax, b_lo, hi, bs);
                            int hit = -1;
                            for (int e = lo; e < hi; ++e)
                                if (__ldg(&bed.se[e].y) >= bs) { if (++hits == 1) hit = e; }
                            if (hits == 1) { x = __ldg(&bed.se[hit]); nid = __ldg(&bed.name_id[hit]); }
                        }
                        if (hits == 1) {
                            if (!(x.x <= bs && be - 1 <= x.y)) same = false;           // partialIntersect == block length, :494
                            else if (b == 0) name = nid;
                            else if (nid != name) same = false;                        // :497
                        } else same = false;
                    }
                    if (same) {                                                        // :508-516: enters the mate map
                        RSQ_CNT(o, RSQ_C_FRAGMENT_CANDIDATES);
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
                            atomicMax(&acc.pair_first[slot], mate_tag(prm.mate_round, 1u, i));           // round 1: the earliest candidate of the name
                            const unsigned int ci = warp_claim(acc.cand_count);
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

#ifndef RSQ_AB_NO_PREFETCH
    if (has_next) prefetch_l2(blocks + next_boff);
#endif

    // ---- warp-converged epilogue: aggregate before touching HBM -------------------------------------------------------
    // gene counters: lanes that count toward the same gene elect one leader (sorted input => usually the whole warp)
    const unsigned peers = __match_any_sync(kFull, o.agg_gene);
    {
        const unsigned bu = __ballot_sync(kFull, o.agg_unique);
        if (o.agg_gene != kNoGene && lane == (unsigned)(__ffs(peers) - 1)) {
            atomicAdd(&acc.gene_reads[o.agg_gene], (unsigned long long)__popc(peers));
            const int nu = __popc(peers & bu);
            if (nu) atomicAdd(&acc.gene_unique[o.agg_gene], (unsigned long long)nu);
        }
    }
    // fragment tracker (src/Metrics.cpp:22, src/Expression.cpp:383-387).  A hash-set insert is two dependent round trips to a
    // multi-GB table: instead of waiting for HBM twice per tile, the warp parks this tile's keys in shared memory, starts an
    // asynchronous copy of their home slots, and inserts them one tile LATER.
    int s_late = 0;
    if constexpr (!kLegacy) {
        s_late = resolve_parked(st, acc, lane);
        st->fkey[lane] = o.frag_key; st->fgene[lane] = o.agg_gene; st->fpeers[lane] = peers;   // each lane reads back its own slot only
#if defined(RSQ_AB_NO_FRAGSET)
#elif defined(RSQ_FRAG_PREFETCH_ONLY)
        if (o.frag_key) prefetch_l2(acc.frag_set + set_home(o.frag_key, acc.frag_cap));
#else
        // asynchronous 8-byte copy of the key's home slot into the lane's parking place (LDGSTS): by the next tile the warp
        // knows without any further global access whether the key is already there (the upper mate of a pair), the slot is
        // free (one CAS on a line that is in L2 by then) or taken by another key (rare: probe on)
        if (o.frag_key) home_fetch(&st->fval[lane], acc.frag_set + set_home(o.frag_key, acc.frag_cap));
        home_commit();
#endif
    }
    // +1 counters: lane k of the warp OWNS counters k and k+32 in registers for the whole persistent loop.  The 64-bit mask of
    // a record is cut into 5-bit words; a multiply spreads the five bits into five 6-bit fields (a tile adds at most 32 to a
    // field), one REDUX per word sums the fields of all 32 records at once, and every lane picks its two fields: no loop, no
    // dependent shuffles — 11 independent reductions for the 55 counters in use.
    {
#ifdef RSQ_AB_NO_EPILOGUE
        own_lo += o.cm & 1ull;   // ablation only
#else
        static_assert(RSQ_C_FRAGSET_ENTRIES == 54, "11 words of 5 counters cover every counter in use");
        unsigned vlo = 0, vhi = 0;
#ifdef RSQ_SKIP_RARE_WORDS
        // words 9 and 10 hold the counters 45..53 (filter tags 1..7, Split Reads, Fragment Candidates): zero for whole tiles
        // on most inputs — one vote instead of two reductions
        const bool rare = __any_sync(kFull, (o.cm >> 45) != 0ull);
#else
        constexpr bool rare = true;
#endif
#pragma unroll
        for (int wd = 0; wd < 11; ++wd) {
            if (wd >= 9 && !rare) break;
            const unsigned x = (unsigned)(o.cm >> (5 * wd)) & 31u;
            const unsigned sum = __reduce_add_sync(kFull, (x * 0x108421u) & 0x1041041u);
            if (wd <= 6) vlo = (lane / 5u == (unsigned)wd) ? sum : vlo;
            if (wd >= 6) vhi = ((lane + 32u) / 5u == (unsigned)wd) ? sum : vhi;
        }
        own_lo += (vlo >> (6u * (lane % 5u))) & 63u;
        own_hi += (vhi >> (6u * ((lane + 32u) % 5u))) & 63u;
#endif
    }
    // valued counters (Metrics::increment(key, int), Metrics.cpp:33-36): signed 32-bit warp sums, sign-extended
    {
#ifdef RSQ_PACK_VALUED
        // short reads: every addend is below 2^11, so two 32-lane sums fit one word (16 bits each) — four reductions instead of
        // seven; one vote sends tiles with a longer read, a larger NM or a negative value down the exact general path
        int s_mm1, s_b1, s_mm2, s_b2, s_mm, s_tb, s_nb;
        if (!__any_sync(kFull, ((unsigned)l_qseq | (unsigned)nm_raw | nb) >> 11)) {
            const unsigned both = (unsigned)nm_raw | ((unsigned)l_qseq << 16);
            const unsigned a1 = __reduce_add_sync(kFull, (o.vflags & 1u) ? both : 0u), a2 = __reduce_add_sync(kFull, (o.vflags & 2u) ? both : 0u);
            const unsigned a3 = __reduce_add_sync(kFull, ((o.vflags & 4u) ? (unsigned)nm_raw : 0u) | ((o.vflags & 8u) ? (unsigned)l_qseq << 16 : 0u));
            s_nb = (int)__reduce_add_sync(kFull, (o.vflags & 16u) ? nb : 0u);
            s_mm1 = (int)(a1 & 0xffffu); s_b1 = (int)(a1 >> 16); s_mm2 = (int)(a2 & 0xffffu); s_b2 = (int)(a2 >> 16);
            s_mm = (int)(a3 & 0xffffu); s_tb = (int)(a3 >> 16);
        } else {
            s_mm1 = __reduce_add_sync(kFull, (o.vflags & 1u) ? nm_raw : 0); s_b1 = __reduce_add_sync(kFull, (o.vflags & 1u) ? l_qseq : 0);
            s_mm2 = __reduce_add_sync(kFull, (o.vflags & 2u) ? nm_raw : 0); s_b2 = __reduce_add_sync(kFull, (o.vflags & 2u) ? l_qseq : 0);
            s_mm = __reduce_add_sync(kFull, (o.vflags & 4u) ? nm_raw : 0); s_tb = __reduce_add_sync(kFull, (o.vflags & 8u) ? l_qseq : 0);
            s_nb = __reduce_add_sync(kFull, (o.vflags & 16u) ? (int)nb : 0);
        }
#else
        const int s_mm1 = __reduce_add_sync(kFull, (o.vflags & 1u) ? nm_raw : 0), s_b1 = __reduce_add_sync(kFull, (o.vflags & 1u) ? l_qseq : 0);
        const int s_mm2 = __reduce_add_sync(kFull, (o.vflags & 2u) ? nm_raw : 0), s_b2 = __reduce_add_sync(kFull, (o.vflags & 2u) ? l_qseq : 0);
        const int s_mm = __reduce_add_sync(kFull, (o.vflags & 4u) ? nm_raw : 0), s_tb = __reduce_add_sync(kFull, (o.vflags & 8u) ? l_qseq : 0);
        const int s_nb = __reduce_add_sync(kFull, (o.vflags & 16u) ? (int)nb : 0);
#endif
        // algorithmic bytes: every lane keeps its own sum; new fragment-set entries: the parked ones are warp-uniform, the
        // on-the-spot ones (second genes of a read) are rare enough for a shared-memory add
        own_bytes += o.my_bytes;
        static_assert(RSQ_C_FRAGSET_ENTRIES >= 32 && RSQ_C_FRAGSET_ENTRIES < 64, "lives in own_hi");
        if (lane == RSQ_C_FRAGSET_ENTRIES - 32) own_hi += (unsigned long long)s_late;
        if (o.n_fresh) atomicAdd(&s_cnt[RSQ_C_FRAGSET_ENTRIES], (unsigned long long)o.n_fresh);
        static_assert(RSQ_C_END1_MISMATCHES < 32 && RSQ_C_END1_BASES < 32 && RSQ_C_END2_MISMATCHES < 32 && RSQ_C_END2_BASES < 32 &&
                      RSQ_C_MISMATCHED_BASES < 32 && RSQ_C_TOTAL_BASES < 32 && RSQ_C_ALIGNMENT_BLOCKS < 32, "valued counters live in own_lo");
        int mine = 0;
        mine = lane == RSQ_C_END1_MISMATCHES ? s_mm1 : mine;
        mine = lane == RSQ_C_END1_BASES ? s_b1 : mine;
        mine = lane == RSQ_C_END2_MISMATCHES ? s_mm2 : mine;
        mine = lane == RSQ_C_END2_BASES ? s_b2 : mine;
        mine = lane == RSQ_C_MISMATCHED_BASES ? s_mm : mine;
        mine = lane == RSQ_C_TOTAL_BASES ? s_tb : mine;
        mine = lane == RSQ_C_ALIGNMENT_BLOCKS ? s_nb : mine;
        own_lo += (unsigned long long)(long long)mine;
    }
  }  // tile loop
    if constexpr (!kLegacy) {                                                        // the last tile's parked keys
        const int s_late = resolve_parked(st, acc, lane);
        if (lane == RSQ_C_FRAGSET_ENTRIES - 32) own_hi += (unsigned long long)s_late;
    }
    // one shared-memory add per counter per warp, then one global atomic per counter per CTA
    if (own_lo) atomicAdd(&s_cnt[lane], own_lo);
    if (own_hi) atomicAdd(&s_cnt[32 + lane], own_hi);
    if (own_bytes) atomicAdd(&s_bytes, own_bytes);
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

// mate-map growth: re-insert (key, pending mate) pairs into a larger table (the round scratch of the new table starts at zero)
__global__ void pair_rehash_kernel(const unsigned long long* __restrict__ old_keys, const unsigned long long* __restrict__ old_payload, unsigned long long old_cap,
                                   unsigned long long* __restrict__ nk, unsigned long long* __restrict__ np, unsigned long long new_mask, unsigned int* err) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= old_cap) return;
    const unsigned long long key = old_keys[i];
    if (!key) return;
    unsigned long long slot = mix64(key) & new_mask;
    for (int probe = 0; probe < 4096; ++probe) {
        const unsigned long long cur = atomicCAS(&nk[slot], 0ull, key);
        if (cur == 0ull) { np[slot] = old_payload[i]; return; }
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

// Phase 3: every warp owns 512 consecutive elements of the tile and reads them as four fully coalesced 128-bit rows (one row =
// 32 lanes x uint4 = 512 bytes); prefix inside the uint4, warp scan of the lane sums per row, carry from row to row; one
// block barrier hands every warp the total of the warps before it.  HBM-bound: reads the difference array once, writes the
// coverage once.
__global__ void __launch_bounds__(kScanThreads) tile_scan_kernel(const unsigned int* __restrict__ in, unsigned int* __restrict__ out, unsigned long long n,
                                                                 const unsigned int* __restrict__ sums) {
    static_assert(kScanThreads == 256 && kScanItems == 16, "8 warps x 4 rows x 32 lanes x 4 elements");
    __shared__ unsigned int s_tot[8];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const unsigned long long seg = (unsigned long long)blockIdx.x * kScanTile + (unsigned long long)wid * 512u;
    uint4 x[4];
    unsigned int carry = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned long long idx = seg + (unsigned long long)j * 128u + lane * 4u;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (idx + 3 < n) v = *reinterpret_cast<const uint4*>(in + idx);          // the arrays are 256-byte aligned, idx a multiple of 4
        else {
            if (idx < n) v.x = in[idx];
            if (idx + 1 < n) v.y = in[idx + 1];
            if (idx + 2 < n) v.z = in[idx + 2];
        }
        v.y += v.x; v.z += v.y; v.w += v.z;                                      // inclusive inside the lane
        unsigned int inc = v.w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int t = __shfl_up_sync(kFull, inc, d);
            if ((int)lane >= d) inc += t;
        }
        const unsigned int off = carry + inc - v.w;                              // everything before this lane in the warp's segment
        v.x += off; v.y += off; v.z += off; v.w += off;
        x[j] = v;
        carry += __shfl_sync(kFull, inc, 31);
    }
    if (lane == 0) s_tot[wid] = carry;
    __syncthreads();
    unsigned int before = sums[blockIdx.x];
    for (unsigned w = 0; w < wid; ++w) before += s_tot[w];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned long long idx = seg + (unsigned long long)j * 128u + lane * 4u;
        uint4 v = x[j];
        v.x += before; v.y += before; v.z += before; v.w += before;              // inclusive: coverage[i] = sum of diff[0..i]
        if (idx + 3 < n) *reinterpret_cast<uint4*>(out + idx) = v;
        else {
            if (idx < n) out[idx] = v.x;
            if (idx + 1 < n) out[idx + 1] = v.y;
            if (idx + 2 < n) out[idx + 2] = v.z;
        }
    }
}

// ---- fragment-tracker set growth ----------------------------------------------------------------------------------------------
__global__ void rehash_kernel(const unsigned long long* __restrict__ old_tab, unsigned long long old_cap, unsigned long long* __restrict__ new_tab,
                              unsigned long long new_cap, unsigned int* err) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= old_cap) return;
    const unsigned long long key = old_tab[i];
    if (key) set_insert(new_tab, new_cap, key, err, 1u);
}

// ---- count vector export for the NVLink all-reduce: counters | gene_reads | gene_unique | gene_fragments | exon_whole |
//      exon_frac >> 32 | exon_frac & 0xffffffff | gc_hist[100] | samples found | kept fragment-size samples, two per word
//      (the split keeps the u64 sum exact across ranks)
// Fragment-size samples (first-N cut across partitions, src/RNASeQC.cpp:372, src/Expression.cpp:532): `found` holds the sample
// counts of all partitions (all-reduced phase-1 buffer; null = one partition).  This partition's k-th sample has global rank
// before + k, before = samples of the partitions ahead of it in the stream; it is kept when that rank is below the quota.
__device__ __forceinline__ void sample_window(const unsigned long long* __restrict__ found, uint32_t rank, uint32_t world, unsigned long long own,
                                              unsigned long long quota, unsigned long long& before, unsigned long long& kept) {
    before = 0ull;
    if (found) for (uint32_t r = 0; r < rank && r < world; ++r) before += found[r];
    const unsigned long long have = own < quota ? own : quota;            // sample_vals holds at most `quota` of them
    kept = before >= quota ? 0ull : (quota - before < have ? quota - before : have);
}
__global__ void pack_sample_counts_kernel(const Accum acc, uint32_t rank, uint32_t world, unsigned long long* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < world) out[i] = (i == rank && acc.sample_count) ? *acc.sample_count : 0ull;
}
__global__ void pack_counts_kernel(const Accum acc, uint32_t n_genes, uint32_t n_exons, unsigned long long quota, const unsigned long long* __restrict__ found,
                                   uint32_t rank, uint32_t world, unsigned long long* __restrict__ out) {
    const unsigned long long fixed = RSQ_N_COUNTERS + 3ull * n_genes + 3ull * n_exons + 100ull;
    const unsigned long long total = fixed + 1ull + (quota + 1ull) / 2ull;
    unsigned long long before = 0ull, kept = 0ull;
    const unsigned long long own = (quota && acc.sample_count) ? *acc.sample_count : 0ull;
    if (quota) sample_window(found, rank, world, own, quota, before, kept);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long j = i, v;
        if (j < RSQ_N_COUNTERS) v = acc.counters[j];
        else if ((j -= RSQ_N_COUNTERS) < n_genes) v = acc.gene_reads[j];
        else if ((j -= n_genes) < n_genes) v = acc.gene_unique[j];
        else if ((j -= n_genes) < n_genes) v = acc.gene_fragments[j];
        else if ((j -= n_genes) < n_exons) v = acc.exon_whole[j];
        else if ((j -= n_exons) < n_exons) v = acc.exon_fhi[j];
        else if ((j -= n_exons) < n_exons) v = acc.exon_flo[j];
        else if ((j -= n_exons) < 100ull) v = acc.gc_hist ? acc.gc_hist[j] : 0ull;      // 100 GC-content bins (zeros without a reference)
        else if ((j -= 100ull) == 0ull) v = own;                                           // samples found (may exceed the quota)
        else {
            j -= 1ull;
            const unsigned long long k0 = 2ull * j, k1 = k0 + 1ull;                        // global ranks of the word's two halves
            const unsigned long long a = (k0 >= before && k0 - before < kept) ? acc.sample_vals[k0 - before] : 0u;
            const unsigned long long b = (k1 < quota && k1 >= before && k1 - before < kept) ? acc.sample_vals[k1 - before] : 0u;
            v = a | (b << 32);
        }
        out[i] = v;
    }
}
__global__ void unpack_counts_kernel(const Accum acc, uint32_t n_genes, uint32_t n_exons, unsigned long long quota, const unsigned long long* __restrict__ in) {
    const unsigned long long fixed = RSQ_N_COUNTERS + 3ull * n_genes + 3ull * n_exons + 100ull;
    const unsigned long long total = fixed + 1ull + (quota + 1ull) / 2ull;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long j = i;
        if (j < RSQ_N_COUNTERS) acc.counters[j] = in[i];
        else if ((j -= RSQ_N_COUNTERS) < n_genes) acc.gene_reads[j] = in[i];
        else if ((j -= n_genes) < n_genes) acc.gene_unique[j] = in[i];
        else if ((j -= n_genes) < n_genes) acc.gene_fragments[j] = in[i];
        else if ((j -= n_genes) < n_exons) acc.exon_whole[j] = in[i];
        else if ((j -= n_exons) < n_exons) acc.exon_fhi[j] = in[i];
        else if ((j -= n_exons) < n_exons) acc.exon_flo[j] = in[i];
        else if ((j -= n_exons) < 100ull) { if (acc.gc_hist) acc.gc_hist[j] = in[i]; }
        else if ((j -= 100ull) == 0ull) { if (quota && acc.sample_count) *acc.sample_count = in[i]; }
        else if (quota) {
            j -= 1ull;
            acc.sample_vals[2ull * j] = (unsigned int)in[i];
            if (2ull * j + 1ull < quota) acc.sample_vals[2ull * j + 1ull] = (unsigned int)(in[i] >> 32);
        }
    }
}
// (whole, frac>>32 sums, frac&0xffffffff sums) -> canonical (whole, frac) with carries resolved
__global__ void combine_exon_kernel(const Accum acc, uint32_t n_exons, unsigned long long* __restrict__ whole_out, unsigned long long* __restrict__ frac_out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_exons) return;
    const unsigned long long hi = acc.exon_fhi[j], lo = acc.exon_flo[j];
    const unsigned long long mid = hi + (lo >> 32);
    frac_out[j] = (mid << 32) | (lo & 0xffffffffull);
    whole_out[j] = acc.exon_whole[j] + (mid >> 32);
}

// ---- mate maps: fragment-size pairing (BED mode) and fragment GC content (--fasta) ---------------------------------------------------
// Both replay, per read name, what the reference does to its map entry when the candidates arrive in stream order
// (fragmentSizeMetrics, src/Expression.cpp:508-533; the FASTA tail of exonAlignmentMetrics, :459-476):
//   no entry                      -> store (interval, PositionEnd)
//   entry with the same interval  -> if the pair's conditions hold: take the sample and ERASE the entry; else leave it
//   entry with another interval   -> leave it
// The hot kernel has given every name's earliest candidate of this launch the round-1 tag (atomicMax).  mate_first: those apply
// their transition; every other candidate bids for round 2.  mate_second: the round-2 winners apply theirs; what is left (a
// third record of one name inside one launch) queues up and is replayed by one CTA, one candidate per name and round.
struct BedMap {
    using Cand = PairCandidate;
    static constexpr int kWhich = 0;
    __device__ static const Cand* list(const Accum& a) { return a.cand; }
    __device__ static unsigned int count(const Accum& a) { return min(*a.cand_count, a.cand_capacity); }
    __device__ static unsigned long long* first(const Accum& a) { return a.pair_first; }
    __device__ static unsigned long long* second(const Accum& a) { return a.pair_second; }
    __device__ static unsigned long long* payload(const Accum& a) { return a.pair_payload; }
    __device__ static uint32_t where(const Cand& c) { return c.name_id; }
    // this candidate meets the pending mate (same name, same interval) whose read ended at stored_end: the four conditions of :528
    __device__ static bool complete(const Accum& a, const RefView&, const Cand& c, int stored_end) {
        if (!c.cond_ok || c.endpos <= stored_end) return false;
        a.sample_flag[c.local_idx] = 1;
        a.sample_isize[c.local_idx] = c.abs_isize;                       // fragmentSizes[abs(InsertSize())], :530
        return true;
    }
};
struct GcMap {
    using Cand = GcCandidate;
    static constexpr int kWhich = 1;
    __device__ static const Cand* list(const Accum& a) { return a.gc_cand; }
    __device__ static unsigned int count(const Accum& a) { return min(*a.gc_cand_count, a.gc_cand_capacity); }
    __device__ static unsigned long long* first(const Accum& a) { return a.gc_first; }
    __device__ static unsigned long long* second(const Accum& a) { return a.gc_second; }
    __device__ static unsigned long long* payload(const Accum& a) { return a.gc_payload; }
    __device__ static uint32_t where(const Cand& c) { return c.exon_slot; }
    // :471-475: the end test and pos != mpos, then gc(getSeq(chr, stored end - Length(), PositionEnd())); the entry is erased even
    // when the window yields no sequence
    __device__ static bool complete(const Accum& a, const RefView& ref, const Cand& c, int stored_end) {
        if (!c.cond_ok || c.endpos <= stored_end) return false;
        // 0-based, end-exclusive window, clipped to the contig (Fasta.cpp:116-150, bioio.hpp:300-309)
        const long long len = (long long)ref.length[c.contig];
        const long long start = (long long)stored_end - c.l_qseq;
        if (start < 0 || start >= len) return true;
        const long long stop = min((long long)c.endpos, len);
        if (stop <= start) return true;
        const uint32_t* w = ref.bits + ref.word_off[c.contig];
        int gcs = 0;
        for (long long wi = start >> 5; wi <= (stop - 1) >> 5; ++wi) {
            uint32_t m = __ldg(&w[wi]);
            if (wi == (start >> 5)) m &= 0xffffffffu << (start & 31);
            if (wi == ((stop - 1) >> 5) && (stop & 31)) m &= 0xffffffffu >> (32 - (int)(stop & 31));
            gcs += __popc(m);
        }
        // gc() adds 1.0/size once per G/C base (Fasta.cpp:66-75): the same sequence of double additions
        const double inv = 1.0 / (double)(stop - start);
        double content = 0.0;
        for (int t = 0; t < gcs; ++t) content += inv;
        const unsigned bin = (unsigned)(content * 100.0);                      // RNASeQC.cpp:368
        if (bin < 100u) atomicAdd(&a.gc_hist[bin], 1ull);
        return true;
    }
};

template <class M>
__device__ __forceinline__ void mate_apply(const Accum& acc, const RefView& ref, const typename M::Cand& c) {
    unsigned long long* pay = M::payload(acc) + c.slot;
    const unsigned long long s = *pay;
    if (s == kMateEmpty) { *pay = ((unsigned long long)M::where(c) << 32) | (unsigned long long)(uint32_t)c.endpos; return; }   // :512-516
    if ((uint32_t)(s >> 32) != M::where(c)) return;                                                                          // :517
    if (M::complete(acc, ref, c, (int32_t)(uint32_t)s)) *pay = kMateEmpty;                                                     // fragments.erase
}
template <class M>
__global__ void mate_first_kernel(const Accum acc, const Params prm, const RefView ref) {
    const unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) atomicAdd(&acc.mate_totals[M::kWhich], (unsigned long long)M::count(acc));
    if (k >= M::count(acc)) return;
    const typename M::Cand c = M::list(acc)[k];
    if (M::first(acc)[c.slot] == mate_tag(prm.mate_round, 1u, c.local_idx)) mate_apply<M>(acc, ref, c);
    else atomicMax(&M::second(acc)[c.slot], mate_tag(prm.mate_round, 2u, c.local_idx));
}
template <class M>
__global__ void mate_second_kernel(const Accum acc, const Params prm, const RefView ref) {
    const unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M::count(acc)) return;
    const typename M::Cand c = M::list(acc)[k];
    if (M::first(acc)[c.slot] == mate_tag(prm.mate_round, 1u, c.local_idx)) return;                 // applied in round 1
    if (M::second(acc)[c.slot] == mate_tag(prm.mate_round, 2u, c.local_idx)) { mate_apply<M>(acc, ref, c); return; }
    const unsigned int at = atomicAdd(&acc.mate_left_count[M::kWhich], 1u);
    if (at < kMateLeftCap) acc.mate_left[M::kWhich * kMateLeftCap + at] = k;
    else atomicOr(acc.error_flag, 16u);
}
// one CTA: the queued candidates, one per name and round, until none is left (rounds 3, 4, ...)
template <class M>
__global__ void __launch_bounds__(1024) mate_rest_kernel(const Accum acc, const Params prm, const RefView ref) {
    __shared__ unsigned int s_left;
    const unsigned int n = min(acc.mate_left_count[M::kWhich], kMateLeftCap);
    if (n == 0) return;
    unsigned int* left = acc.mate_left + M::kWhich * kMateLeftCap;
    if (threadIdx.x == 0) s_left = n;
    __syncthreads();
    for (unsigned r = 3; r < 3u + kMateLeftCap + 1u; ++r) {
        for (unsigned j = threadIdx.x; j < n; j += blockDim.x) {
            const unsigned k = left[j];
            if (k == 0xffffffffu) continue;
            const typename M::Cand c = M::list(acc)[k];
            atomicMax(&M::first(acc)[c.slot], mate_tag(prm.mate_round, r, c.local_idx));
        }
        __threadfence();
        __syncthreads();
        for (unsigned j = threadIdx.x; j < n; j += blockDim.x) {
            const unsigned k = left[j];
            if (k == 0xffffffffu) continue;
            const typename M::Cand c = M::list(acc)[k];
            if (*(volatile unsigned long long*)&M::first(acc)[c.slot] != mate_tag(prm.mate_round, r, c.local_idx)) continue;
            mate_apply<M>(acc, ref, c);
            left[j] = 0xffffffffu;
            atomicSub(&s_left, 1u);
        }
        __threadfence();
        __syncthreads();
        if (s_left == 0) break;
        __syncthreads();
    }
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


// ---- compact wire format (rsq_read_packed) -> device layout -------------------------------------------------------------------
__global__ void unpack_counts_of_kernel(const rsq_read_packed* __restrict__ packed, size_t n, unsigned int* __restrict__ tot, unsigned int* __restrict__ expl) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(packed + i) + 1);     // name_hash | contig,l_qseq | nm,n_blocks,filter
        const unsigned nb = (q.w >> 16) & 0xffu;
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(packed + i));
        const bool implicit = (h.w >> 24) & RSQ_AUX_IMPLICIT_BLOCK;
        tot[i] = nb;
        expl[i] = implicit ? 0u : nb;
    }
}
__global__ void unpack_records_kernel(const rsq_read_packed* __restrict__ packed, size_t n, const rsq_block* __restrict__ xblocks, size_t n_xblocks,
                                      const unsigned int* __restrict__ incl_tot, const unsigned int* __restrict__ incl_expl,
                                      rsq_read* __restrict__ reads, rsq_block* __restrict__ blocks, unsigned int* __restrict__ err) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 h = __ldcs(reinterpret_cast<const uint4*>(packed + i));        // pos | mpos | isize | flag,mapq,aux
        const uint4 q = __ldcs(reinterpret_cast<const uint4*>(packed + i) + 1);
        const int pos = (int)h.x;
        const unsigned aux = h.w >> 24;
        const unsigned nb = (q.w >> 16) & 0xffu;
        const int l_qseq = (int)(q.z >> 16);
        const bool implicit = aux & RSQ_AUX_IMPLICIT_BLOCK;
        const unsigned block_off = incl_tot[i] - nb;
        int endpos = pos + 1;
        if (implicit) {
            endpos = pos + l_qseq;
            blocks[block_off] = rsq_block{pos + 1, l_qseq};
        } else if (nb) {
            const size_t src = (size_t)incl_expl[i] - nb;
            if (src + nb > n_xblocks) { atomicOr(err, 8u); }
            else {
                rsq_block last{0, 0};
                for (unsigned b = 0; b < nb; ++b) { last = xblocks[src + b]; blocks[block_off + b] = last; }
                endpos = last.start - 1 + last.len;
            }
        }
        int4 o0, o1, o2;
        o0.x = pos; o0.y = (int)h.y; o0.z = endpos; o0.w = (int)h.z;
        o1.x = l_qseq; o1.y = (int)(q.w & 0xffffu);                                 // nm
        o1.z = (int)((h.w & 0x00ffffffu) | ((aux & ~(unsigned)RSQ_AUX_IMPLICIT_BLOCK) << 24));   // flag | mapq << 16 | aux << 24
        o1.w = (int)((q.z & 0xffffu) | ((q.w >> 24) << 16));                        // contig | filter_mask << 16
        o2.x = (int)q.x; o2.y = (int)q.y; o2.z = (int)block_off; o2.w = (int)nb;
        int4* out = reinterpret_cast<int4*>(reads + i);
        out[0] = o0; out[1] = o1; out[2] = o2;
    }
}

// ---- 16-byte wire format (rsq_batch16) -> device layout: one warp per tile of 32 records ----------------------------------------------
// Every tile header says where the tile's variable-length parts start (explicit blocks, exceptions, expanded blocks), so tiles
// expand independently: two warp scans give each record its offsets.  HBM-bound: reads ~21 bytes, writes ~60 bytes per record.
__global__ void __launch_bounds__(256) expand_compact_kernel(const unsigned char* __restrict__ buf, const rsq_batch16 H, rsq_read* __restrict__ reads,
                                                             rsq_block* __restrict__ blocks, unsigned int* __restrict__ err) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= H.n_tiles) return;
    const rsq_tile16* T = reinterpret_cast<const rsq_tile16*>(buf + H.off_tiles) + tile;
    const int base_pos = __ldg(&T->base_pos);
    const uint32_t block_base = __ldg(&T->block_base), out_base = __ldg(&T->out_block_base), exc_base = __ldg(&T->exc_base);
    const int contig = (int)__ldg(&T->contig);
    const uint32_t i = tile * 32u + lane;
    const bool valid = i < H.n_reads;
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (valid) r = __ldcs(reinterpret_cast<const uint4*>(buf + H.off_reads) + i);
    const uint32_t combo = r.w >> 16;
    const bool exc = valid && combo == RSQ_COMBO_EXCEPTION;
    bool bad = valid && !exc && combo >= H.n_combos;
    uint4 c = make_uint4(0u, 0u, 0u, 0u);
    if (valid && !exc && !bad) c = __ldg(reinterpret_cast<const uint4*>(buf + H.off_combos) + combo);
    // exception records: the full rsq_read, ranked inside the tile by a ballot
    const unsigned excm = __ballot_sync(kFull, exc);
    int4 e0 = make_int4(0, 0, 0, 0), e1 = e0, e2 = e0;
    if (exc) {
        const uint32_t e = exc_base + (uint32_t)__popc(excm & ((1u << lane) - 1u));
        if (e >= H.n_exc) bad = true;
        else {
            const int4* ep = reinterpret_cast<const int4*>(buf + H.off_exc) + 3 * (size_t)e;
            e0 = __ldg(ep); e1 = __ldg(ep + 1); e2 = __ldg(ep + 2);
        }
    }
    const uint32_t aux = exc ? ((uint32_t)e1.z >> 24) : ((c.x >> 24) & 0xffu);
    const bool implicit = !exc && (aux & RSQ_AUX_IMPLICIT_BLOCK);
    uint32_t nb = exc ? ((uint32_t)e2.w & 0xffffu) : (c.z & 0xffu);
    if (bad) nb = 0;
    const uint32_t expl = (exc || implicit) ? 0u : nb;
    // exclusive warp scans of the block counts
    uint32_t s_tot = nb, s_expl = expl;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(kFull, s_tot, d), b = __shfl_up_sync(kFull, s_expl, d);
        if ((int)lane >= d) { s_tot += a; s_expl += b; }
    }
    const uint32_t out_off = out_base + s_tot - nb, x_off = block_base + s_expl - expl;
    if (valid && (unsigned long long)out_off + nb > H.total_blocks) { bad = true; nb = 0; }
    int4 o0, o1, o2;
    if (exc && !bad) {
        const uint32_t src = (uint32_t)e2.z;                                      // block_off into the exception blocks
        if ((unsigned long long)src + nb > H.n_exc_blocks) { bad = true; nb = 0; }
        const rsq_block* eb = reinterpret_cast<const rsq_block*>(buf + H.off_exc_blocks) + src;
        for (uint32_t b = 0; b < nb; ++b) blocks[out_off + b] = eb[b];
        o0 = e0; o1 = e1; o2 = e2;
        o2.z = (int)out_off; o2.w = (int)nb;
    } else {
        const int pos = base_pos + (int)(r.z & 0xffffu);
        const int dmate = (int)(int16_t)(r.z >> 16), isize = (int)(int16_t)(r.w & 0xffffu);
        const int l_qseq = (int)(c.y & 0xffffu), nm = (int)(c.y >> 16);
        int endpos = pos + 1;
        if (implicit) {
            endpos = pos + l_qseq;
            if (!bad) blocks[out_off] = rsq_block{pos + 1, l_qseq};
        } else if (nb) {
            if ((unsigned long long)x_off + nb > H.n_blocks) { bad = true; nb = 0; }
            const uint32_t* xb = reinterpret_cast<const uint32_t*>(buf + H.off_blocks) + x_off;
            rsq_block last{0, 0};
            for (uint32_t b = 0; b < nb; ++b) {
                const uint32_t w = __ldg(xb + b);
                last = rsq_block{pos + 1 + (int)(w >> RSQ_BLOCK16_LEN_BITS), (int)(w & ((1u << RSQ_BLOCK16_LEN_BITS) - 1u))};
                blocks[out_off + b] = last;
            }
            if (nb) endpos = last.start - 1 + last.len;
        }
        o0.x = pos; o0.y = pos + dmate; o0.z = endpos; o0.w = isize;
        o1.x = l_qseq; o1.y = nm;
        o1.z = (int)((c.x & 0x00ffffffu) | ((aux & ~(uint32_t)RSQ_AUX_IMPLICIT_BLOCK) << 24));   // flag | mapq << 16 | aux << 24
        o1.w = (int)(((uint32_t)contig & 0xffffu) | (((c.z >> 8) & 0xffu) << 16));               // contig | filter_mask << 16
        o2.x = (int)r.x; o2.y = (int)r.y; o2.z = (int)out_off; o2.w = (int)nb;
    }
    if (bad) {                      // a batch that contradicts itself: flag it, and leave a record that touches nothing
        atomicOr(err, 8u);
        o0 = make_int4(0, 0, 1, 0); o1 = make_int4(0, 0, (int)RSQ_F_UNMAP, 0xffff); o2 = make_int4(0, 0, 0, 0);
    }
    if (valid) {
        int4* out = reinterpret_cast<int4*>(reads + i);
        out[0] = o0; out[1] = o1; out[2] = o2;
    }
}

}  // namespace

// -------------------------------------------------------------------------------------------------------------------------------------
int launch_hot_path(const rsq_read* reads, const rsq_block* blocks, size_t n_reads, size_t n_blocks, const AnnView& ann, const BedView& bed,
                    const Accum& acc, const Params& prm, cudaStream_t stream) {
    if (!n_reads) return 0;
    // persistent grid: every SM holds exactly as many CTAs as its register file admits
    static int per_sm[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sms = 0;
    const int which = (prm.bed_enabled ? 1 : 0) | (prm.legacy ? 2 : 0) | ((prm.gc_enabled && !prm.legacy) ? 4 : 0);
    using Kernel = void (*)(const rsq_read*, const rsq_block*, uint32_t, uint32_t, const AnnView, const BedView, const Accum, const Params);
    if (n_reads >= (1ull << 31) - 64 * 1024 * 32 || n_blocks > 0xffffffffull) return -1;   // callers pass sub-batches (rsq_capi.cu kSubBatch)
    static const Kernel kernels[8] = {hot_path_kernel<false, false, false>, hot_path_kernel<true, false, false>,
                                      hot_path_kernel<false, true, false>,  hot_path_kernel<true, true, false>,
                                      hot_path_kernel<false, false, true>,  hot_path_kernel<true, false, true>,
                                      nullptr, nullptr};
    if (!per_sm[which]) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[which], kernels[which], kHotThreads, 0);
        if (per_sm[which] < 1) per_sm[which] = 1;
    }
    const size_t tiles = (n_reads + 31) / 32;
    const unsigned grid = (unsigned)std::min<size_t>((size_t)sms * per_sm[which], (tiles + 7) / 8);
    RSQ_LAUNCH(kernels[which], grid, kHotThreads, 0, stream, reads, blocks, (uint32_t)n_reads, (uint32_t)n_blocks, ann, bed, acc, prm);
    return 1;
}

int launch_prefix_max(const int2* se, const uint32_t* off, uint32_t n_contigs, int32_t* pmax, cudaStream_t stream) {
    if (!n_contigs) return 0;
    RSQ_LAUNCH(prefix_max_kernel, (n_contigs + 63) / 64, 64, 0, stream, se, off, n_contigs, pmax);
    return 1;
}


int launch_pair_rehash(const unsigned long long* old_keys, const unsigned long long* old_payload, unsigned long long old_cap, unsigned long long* nk,
                       unsigned long long* np, unsigned long long new_mask, unsigned int* error_flag, cudaStream_t stream) {
    RSQ_LAUNCH(pair_rehash_kernel, (unsigned)((old_cap + 255) / 256), 256, 0, stream, old_keys, old_payload, old_cap, nk, np, new_mask, error_flag);
    return 1;
}

size_t coverage_scan_scratch_elems(unsigned long long n) { return (size_t)((n + kScanTile - 1) / kScanTile) + 1; }

int launch_coverage_scan(const unsigned int* diff, unsigned int* cov, unsigned long long n, unsigned int* block_sums, cudaStream_t stream) {
    if (!n) return 0;
    const unsigned long long tiles = (n + kScanTile - 1) / kScanTile;
    RSQ_LAUNCH(tile_sum_kernel, (unsigned)tiles, kScanThreads, 0, stream, diff, n, block_sums);
    RSQ_LAUNCH(scan_sums_kernel, 1, 1024, 0, stream, block_sums, tiles);
    RSQ_LAUNCH(tile_scan_kernel, (unsigned)tiles, kScanThreads, 0, stream, diff, cov, n, block_sums);
    return 3;
}

int launch_rehash(const unsigned long long* old_tab, unsigned long long old_cap, unsigned long long* new_tab, unsigned long long new_cap,
                  unsigned int* error_flag, cudaStream_t stream) {
    RSQ_LAUNCH(rehash_kernel, (unsigned)((old_cap + 255) / 256), 256, 0, stream, old_tab, old_cap, new_tab, new_cap, error_flag);
    return 1;
}

size_t export_count_words(uint32_t n_genes, uint32_t n_exons, unsigned long long quota) {
    return (size_t)(RSQ_N_COUNTERS + 3ull * n_genes + 3ull * n_exons + 100ull + 1ull + (quota + 1ull) / 2ull);
}
int launch_pack_sample_counts(const Accum& acc, uint32_t rank, uint32_t world, unsigned long long* out, cudaStream_t stream) {
    RSQ_LAUNCH(pack_sample_counts_kernel, (world + 63) / 64, 64, 0, stream, acc, rank, world, out);
    return 1;
}
int launch_pack_counts(const Accum& acc, uint32_t n_genes, uint32_t n_exons, unsigned long long quota, const unsigned long long* found, uint32_t rank,
                       uint32_t world, unsigned long long* out, cudaStream_t stream) {
    RSQ_LAUNCH(pack_counts_kernel, 148 * 4, 256, 0, stream, acc, n_genes, n_exons, quota, found, rank, world, out);
    return 1;
}
int launch_unpack_counts(const Accum& acc, uint32_t n_genes, uint32_t n_exons, unsigned long long quota, const unsigned long long* in, cudaStream_t stream) {
    RSQ_LAUNCH(unpack_counts_kernel, 148 * 4, 256, 0, stream, acc, n_genes, n_exons, quota, in);
    return 1;
}

int launch_combine_exon(const Accum& acc, uint32_t n_exons, unsigned long long* whole_out, unsigned long long* frac_out, cudaStream_t stream) {
    if (!n_exons) return 0;
    RSQ_LAUNCH(combine_exon_kernel, (n_exons + 255) / 256, 256, 0, stream, acc, n_exons, whole_out, frac_out);
    return 1;
}

size_t pair_scan_scratch_elems(size_t n_reads) { return (n_reads + kScanTile - 1) / kScanTile + 2; }

int launch_pair_stages(size_t n_reads, const Accum& acc, const Params& prm, unsigned int* scratch, cudaStream_t stream) {
    if (!n_reads) return 0;
    const unsigned cgrid = (unsigned)((std::min<size_t>(n_reads, acc.cand_capacity) + 255) / 256);
    const RefView no_ref{};
    RSQ_LAUNCH(mate_first_kernel<BedMap>, cgrid, 256, 0, stream, acc, prm, no_ref);
    RSQ_LAUNCH(mate_second_kernel<BedMap>, cgrid, 256, 0, stream, acc, prm, no_ref);
    RSQ_LAUNCH(mate_rest_kernel<BedMap>, 1, 1024, 0, stream, acc, prm, no_ref);
    const unsigned long long tiles = (n_reads + kScanTile - 1) / kScanTile;
    unsigned int* sums = scratch;
    unsigned int* last = scratch + tiles + 1;
    RSQ_LAUNCH(sample_count_kernel, (unsigned)tiles, kScanThreads, 0, stream, acc.sample_flag, n_reads, sums);
    RSQ_LAUNCH(save_last_kernel, 1, 32, 0, stream, sums, tiles, last);            // count of the last tile, before the scan overwrites it
    RSQ_LAUNCH(scan_sums_kernel, 1, 1024, 0, stream, sums, tiles);
    RSQ_LAUNCH(sample_rank_kernel, (unsigned)tiles, kScanThreads, 0, stream, acc.sample_flag, acc.sample_isize, n_reads, sums, acc, prm);
    RSQ_LAUNCH(sample_commit_kernel, 1, 32, 0, stream, sums, tiles, last, acc);
    return 8;
}

int launch_unpack_counts_of(const rsq_read_packed* packed, size_t n, unsigned int* tot, unsigned int* expl, cudaStream_t stream) {
    if (!n) return 0;
    RSQ_LAUNCH(unpack_counts_of_kernel, (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, stream, packed, n, tot, expl);
    return 1;
}
int launch_unpack_records(const rsq_read_packed* packed, size_t n, const rsq_block* xblocks, size_t n_xblocks, const unsigned int* incl_tot,
                          const unsigned int* incl_expl, rsq_read* reads, rsq_block* blocks, unsigned int* error_flag, cudaStream_t stream) {
    if (!n) return 0;
    RSQ_LAUNCH(unpack_records_kernel, (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, stream, packed, n, xblocks, n_xblocks, incl_tot, incl_expl,
                                                                                                        reads, blocks, error_flag);
    return 1;
}

int launch_expand_compact(const unsigned char* buf, const rsq_batch16& hdr, rsq_read* reads, rsq_block* blocks, unsigned int* error_flag, cudaStream_t stream) {
    if (!hdr.n_tiles) return 0;
    RSQ_LAUNCH(expand_compact_kernel, (hdr.n_tiles + 7) / 8, 256, 0, stream, buf, hdr, reads, blocks, error_flag);
    return 1;
}

int launch_gc_stages(size_t n_reads, const RefView& ref, const Accum& acc, const Params& prm, cudaStream_t stream) {
    if (!n_reads) return 0;
    const unsigned cgrid = (unsigned)((std::min<size_t>(n_reads, acc.gc_cand_capacity) + 255) / 256);
    RSQ_LAUNCH(mate_first_kernel<GcMap>, cgrid, 256, 0, stream, acc, prm, ref);
    RSQ_LAUNCH(mate_second_kernel<GcMap>, cgrid, 256, 0, stream, acc, prm, ref);
    RSQ_LAUNCH(mate_rest_kernel<GcMap>, 1, 1024, 0, stream, acc, prm, ref);
    return 3;
}

int launch_sample_histogram(const unsigned int* vals, unsigned long long n, unsigned long long* hist, unsigned long long hist_len, cudaStream_t stream) {
    if (!n) return 0;
    RSQ_LAUNCH(sample_hist_kernel, 148, 256, 0, stream, vals, n, hist, hist_len);
    return 1;
}

}  // namespace rsqd
