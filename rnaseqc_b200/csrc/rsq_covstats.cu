// rsq_covstats.cu — per-gene coverage statistics on the device (K5).
//
// One CTA per kept "gene" line.  Replaces, for that gene, BaseCoverage::compute -> computeCoverage
// (src/Metrics.cpp:132-151,265-337: per-exon masked mean/std/CV, stitched transcript, masked transcript mean/std/CV)
// and BiasCounter::computeBias (src/Metrics.cpp:160-235: first-maximum peak, the raw "median" read out of the unsorted
// vector, 5th-percentile-of-non-zero flank trimming, medians of the sorted 3'/5' windows, truncation into u64).
// The transcript is a contiguous slice of the integrated coverage array, so every vector erase of the reference is an
// index-range adjustment here.  Sums are formed as exact integers (or block-reduced doubles) and rounded once: the
// results agree with the reference's sequential fp64 sums to ~1e-15 relative (tests: 1e-9), the u64 bias ends exactly.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>

#include "rsq_device.h"

namespace rsqd {

namespace {

constexpr int kThreads = 128;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxWindow = 8192;          // bias windows up to this are sorted in shared memory (2 x 8 bytes x the next power of two), wider ones in global scratch
constexpr int kWarpWindow = 128;          // ... the warp-per-gene kernel keeps its windows below this
constexpr unsigned long long kWarpGeneMax = 6144;   // stitched transcripts up to this length go one warp each (no block barriers)

// A team works on one gene: a whole CTA (kTeam == kThreads; barriers are __syncthreads) or one warp (kTeam == 32; shuffles and
// __syncwarp only — most genes are a few kb, and twenty block barriers per gene were the old kernel's main stall).
template <int kTeam>
struct Team {
    __device__ static unsigned tid() { return kTeam == 32 ? (threadIdx.x & 31u) : threadIdx.x; }
    __device__ static void sync() { if (kTeam == 32) __syncwarp(); else __syncthreads(); }
    template <typename T, typename Op>
    __device__ static T reduce(T v, Op op, T identity, T* s_tmp) {
        for (int d = 16; d > 0; d >>= 1) v = op(v, __shfl_xor_sync(kFull, v, d));
        if (kTeam == 32) return v;
        __syncthreads();                       // s_tmp may still be read from a previous reduction
        if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = v;
        __syncthreads();
        T r = identity;
        for (int w = 0; w < kThreads / 32; ++w) r = op(r, s_tmp[w]);
        return r;
    }
};

struct SumU64 { __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a + b; } };
struct MinU64 { __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a < b ? a : b; } };
struct MaxU64 { __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a > b ? a : b; } };

// population variance from exact integer moments: (n*sum(x^2) - sum(x)^2) / n^2, rounded once; exactly 0 for a flat profile
__device__ __forceinline__ double exact_variance(unsigned long long n, unsigned long long s1, unsigned __int128 s2) {
    const unsigned __int128 num = (unsigned __int128)n * s2 - (unsigned __int128)s1 * s1;
    const double hi = (double)(unsigned long long)(num >> 64), lo = (double)(unsigned long long)num;
    return (hi * 18446744073709551616.0 + lo) / ((double)n * (double)n);
}

// computeMedian (src/Metrics.h:147-160) over a sorted shared-memory array; NOTE the inverted parity
__device__ __forceinline__ double median_quirk(const double* a, unsigned n) {
    if (n == 1) return a[0];
    const unsigned mid = (n - 1) / 2;
    if (n % 2) return (a[mid] + a[mid + 1]) / 2.0;
    return a[mid];
}

// in-place bitonic sort of n doubles in shared memory (padded to a power of two with +inf)
template <int kTeam>
__device__ void sort_window(double* a, unsigned n) {
    unsigned p = 1;
    while (p < n) p <<= 1;
    for (unsigned i = n + Team<kTeam>::tid(); i < p; i += kTeam) a[i] = CUDART_INF;
    Team<kTeam>::sync();
    for (unsigned k = 2; k <= p; k <<= 1)
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = Team<kTeam>::tid(); i < p; i += kTeam) {
                const unsigned l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const double x = a[i], y = a[l];
                    if ((x > y) == up) { a[i] = y; a[l] = x; }
                }
            }
            Team<kTeam>::sync();
        }
}

// masked mean / std / CV of coverage [a, b) by one warp (exact integer moments)
__device__ __forceinline__ void warp_moments(const unsigned int* tc, unsigned long long a, unsigned long long b, unsigned lane, unsigned long long& s,
                                             unsigned __int128& s2) {
    unsigned long long q_lo = 0, q_hi = 0;
    s = 0;
    for (unsigned long long j = a + lane; j < b; j += 32) {
        const unsigned long long x = tc[j], xx = x * x;
        s += x; q_lo += xx; q_hi += q_lo < xx;
    }
    for (int d = 16; d > 0; d >>= 1) {
        s += __shfl_xor_sync(kFull, s, d);
        const unsigned long long ol = __shfl_xor_sync(kFull, q_lo, d), oh = __shfl_xor_sync(kFull, q_hi, d);
        q_lo += ol; q_hi += oh + (q_lo < ol);
    }
    s2 = ((unsigned __int128)q_hi << 64) | q_lo;
}

// Everything the reference computes for one gene line when it leaves the window.  kTeam threads work on gene g; s_win holds two
// windows of win_cap doubles for this team.
template <int kTeam>
__device__ void gene_statistics(const unsigned int* __restrict__ cov, uint32_t g, const unsigned long long* __restrict__ gene_cov_off,
                                const uint32_t* __restrict__ gene_exon_off, const uint8_t* __restrict__ gene_strand, const uint32_t* __restrict__ exon_len,
                                const uint32_t* __restrict__ exon_slot, uint32_t mask, int bias_offset, int bias_window, uint32_t bias_gene_length,
                                double* __restrict__ gene_mean, double* __restrict__ gene_std, double* __restrict__ gene_cv,
                                unsigned long long* __restrict__ five_end, unsigned long long* __restrict__ three_end, double* __restrict__ exon_cv,
                                unsigned long long T, unsigned long long* s_u64, double* s_win, unsigned win_cap) {
    const uint32_t e0 = gene_exon_off[g], e1 = gene_exon_off[g + 1];
    const unsigned int* tc = cov + gene_cov_off[g];
    const unsigned tid = Team<kTeam>::tid(), lane = threadIdx.x & 31u, warp = kTeam == 32 ? 0u : (threadIdx.x >> 5);
    constexpr unsigned kWarps = kTeam / 32;

    // ---- per-exon masked statistics (one warp per exon), src/Metrics.cpp:280-305 ---------------------------------------
    const unsigned long long lo = mask, hi = T > mask ? T - mask : 0;       // unmasked transcript positions are [lo, hi)
    {
        unsigned long long o = 0;
        for (uint32_t e = e0; e < e1; ++e) {
            const unsigned long long L = exon_len[e];
            if ((e - e0) % kWarps == warp) {
                const unsigned long long a = o > lo ? o : lo, b = (o + L) < hi ? (o + L) : hi;
                if (b > a) {
                    unsigned long long s; unsigned __int128 s2;
                    warp_moments(tc, a, b, lane, s, s2);
                    const double mean = (double)s / (double)(b - a);
                    const double cv = sqrt(exact_variance(b - a, s, s2)) / mean;
                    if (lane == 0 && !(isnan(cv) || isinf(cv))) exon_cv[exon_slot[e]] = cv;
                }
            }
            o += L;
        }
    }

    // ---- 3'/5' bias, src/Metrics.cpp:160-235 ------------------------------------------------------------------------------
    unsigned long long ta = 0, tlen = T;      // the (possibly flank-trimmed) vector handed on to the masked statistics
    unsigned long long five = 0, three = 0;
    if (T >= bias_gene_length) {
        // first maximum: max value, then the smallest index holding it
        unsigned long long best = 0;          // (value << 32) | (0xffffffff - index): max picks the largest value, lowest index
        for (unsigned long long j = tid; j < T; j += kTeam) {
            const unsigned long long key = ((unsigned long long)tc[j] << 32) | (0xffffffffull - j);
            best = key > best ? key : best;
        }
        best = Team<kTeam>::reduce(best, MaxU64(), 0ull, s_u64);
        const unsigned long long peak = best >> 32;
        const unsigned long long peak_pos = peak ? 0xffffffffull - (best & 0xffffffffull) : 0ull;     // all zero: index 0, :164-170
        const unsigned long long p1 = (peak_pos + (unsigned)(bias_window / 2)) < T ? peak_pos + (unsigned)(bias_window / 2) : T;
        const unsigned long long n = (unsigned long long)bias_window < p1 ? (unsigned long long)bias_window : p1;
        const unsigned long long it = p1 - n;
        auto at = [&](unsigned long long i) { return (it + i) < T ? (double)tc[it + i] : 0.0; };
        double pm = 0.0;
        if (n == 1) pm = at(0);
        else if (n > 1) { const unsigned long long mid = (n - 1) / 2; pm = (n % 2) ? (at(mid) + at(mid + 1)) / 2.0 : at(mid); }
        if (n > 0 && pm >= 100.0) {                                                           // :181
            unsigned long long nnz = 0;
            for (unsigned long long j = tid; j < T; j += kTeam) nnz += tc[j] != 0u;
            nnz = Team<kTeam>::reduce(nnz, SumU64(), 0ull, s_u64);
            const unsigned long long k = (unsigned long long)((double)nnz * 0.05);            // pct[size*0.05], :189
            // k-th smallest non-zero value: smallest v with #(0 < cov <= v) > k
            unsigned long long vlo = 1, vhi = peak;
            while (vlo < vhi) {
                const unsigned long long vm = (vlo + vhi) >> 1;
                unsigned long long cnt = 0;
                for (unsigned long long j = tid; j < T; j += kTeam) { const unsigned x = tc[j]; cnt += (x != 0u && x <= vm); }
                cnt = Team<kTeam>::reduce(cnt, SumU64(), 0ull, s_u64);
                if (cnt > k) vhi = vm; else vlo = vm + 1;
            }
            const unsigned long long lower = vlo;
            unsigned long long first = ~0ull, last = 0ull;                                    // flank trimming, :193-205
            for (unsigned long long j = tid; j < T; j += kTeam)
                if (tc[j] > lower) { first = j < first ? j : first; last = (j + 1) > last ? (j + 1) : last; }
            first = Team<kTeam>::reduce(first, MinU64(), ~0ull, s_u64);
            last = Team<kTeam>::reduce(last, MaxU64(), 0ull, s_u64);
            if (first == ~0ull) { ta = 0; tlen = 0; } else { ta = first; tlen = last - first; }
            if (tlen >= bias_gene_length) {                                                   // :208-229
                const unsigned long long off = (unsigned)bias_offset, win = (unsigned)bias_window;
                const unsigned long long l_end = (off + win) < tlen ? (off + win) : tlen;
                const unsigned nl = l_end > off ? (unsigned)(l_end - off) : 0u;
                unsigned nr = 0;
                unsigned long long r_beg = 0;
                if (tlen >= win + off) { r_beg = tlen - (win + off); nr = (unsigned)((tlen - off) - r_beg); }
                double* wl = s_win; double* wr = s_win + win_cap;
                for (unsigned i = tid; i < nl; i += kTeam) wl[i] = (double)tc[ta + off + i];
                for (unsigned i = tid; i < nr; i += kTeam) wr[i] = (double)tc[ta + r_beg + i];
                Team<kTeam>::sync();
                if (nl && nr) {
                    sort_window<kTeam>(wl, nl);
                    sort_window<kTeam>(wr, nr);
                    const double ml = median_quirk(wl, nl), mr = median_quirk(wr, nr);
                    if (gene_strand[g] == RSQ_STRAND_FWD) { three = (unsigned long long)mr; five = (unsigned long long)ml; }
                    else { three = (unsigned long long)ml; five = (unsigned long long)mr; }
                }
                Team<kTeam>::sync();          // the windows are reused by the team's next gene
            }
        }
    }

    // ---- masked transcript mean / std / CV, src/Metrics.cpp:312-336 ----------------------------------------------------------
    unsigned long long L1 = tlen, start = ta;
    if (mask) {
        L1 = mask > tlen ? 0 : tlen - mask;                       // erase the last `mask` bases (or everything)
        if (L1) { if (mask > L1) L1 = 0; else { L1 -= mask; start = ta + mask; } }   // then the first `mask`
    }
    double mean = -1.0, sd = -1.0, cv = -1.0;
    if (L1 > 0) {
        unsigned long long s = 0, q_lo = 0, q_hi = 0;
        for (unsigned long long j = tid; j < L1; j += kTeam) {
            const unsigned long long x = tc[start + j], xx = x * x;
            s += x; q_lo += xx; q_hi += q_lo < xx;
        }
        s = Team<kTeam>::reduce(s, SumU64(), 0ull, s_u64);
        // 128-bit sum of squares: reduce the low words in two 32-bit halves (carry-free), then the high words
        const unsigned long long l0 = Team<kTeam>::reduce(q_lo & 0xffffffffull, SumU64(), 0ull, s_u64);
        const unsigned long long l1 = Team<kTeam>::reduce(q_lo >> 32, SumU64(), 0ull, s_u64);
        const unsigned long long h = Team<kTeam>::reduce(q_hi, SumU64(), 0ull, s_u64);
        const unsigned __int128 s2 = ((unsigned __int128)h << 64) + ((unsigned __int128)l1 << 32) + l0;
        mean = (double)s / (double)L1;
        sd = sqrt(exact_variance(L1, s, s2));
        cv = sd / mean;
    }
    if (tid == 0) {
        gene_mean[g] = mean; gene_std[g] = sd; gene_cv[g] = cv;
        five_end[g] = five; three_end[g] = three;
    }
}

// One gene line by one team (a warp or a CTA): whose gene is it, is it owned, then the statistics.  `s_win`: two windows of win_cap
// doubles for this team.  kBig: bias windows beyond what shared memory holds — every gene goes to a CTA whose windows live in global
// memory (coverage_stats_big_kernel).
template <int kTeam, bool kBig>
__device__ __forceinline__ void gene_entry(uint32_t g, const unsigned int* __restrict__ cov, const unsigned long long* __restrict__ gene_cov_off,
                                           const uint32_t* __restrict__ gene_exon_off, const uint8_t* __restrict__ gene_strand,
                                           const uint32_t* __restrict__ exon_len, const uint32_t* __restrict__ exon_slot, uint32_t mask, int bias_offset,
                                           int bias_window, uint32_t bias_gene_length, double* __restrict__ gene_mean, double* __restrict__ gene_std,
                                           double* __restrict__ gene_cv, unsigned long long* __restrict__ five_end, unsigned long long* __restrict__ three_end,
                                           double* __restrict__ exon_cv, const uint8_t* __restrict__ gene_owned, unsigned long long* s_u64, double* s_win, unsigned win_cap) {
    const unsigned tid = Team<kTeam>::tid();
    const uint32_t e0 = gene_exon_off[g], e1 = gene_exon_off[g + 1];
    unsigned long long T = 0;
    for (uint32_t e = e0; e < e1; ++e) T += exon_len[e];
    if (!kBig) {
        const bool warp_sized = T <= kWarpGeneMax && bias_window <= kWarpWindow;
        if (warp_sized != (kTeam == 32)) return;                   // the other kernel's gene (uniform over the team)
    }
    if (gene_owned && !gene_owned[g]) {
        // another partition owns this gene's reads (rsq_transcripts.gene_owned): all-zero bits, so that the u64 sum over the
        // partitions is the owner's value
        for (uint32_t e = e0 + tid; e < e1; e += kTeam) exon_cv[exon_slot[e]] = 0.0;
        if (tid == 0) { gene_mean[g] = 0.0; gene_std[g] = 0.0; gene_cv[g] = 0.0; five_end[g] = 0ull; three_end[g] = 0ull; }
        return;
    }
    if (gene_owned) {   // the slots this gene does not write below must read "no entry" (NaN) whatever the prefill of this partition was
        for (uint32_t e = e0 + tid; e < e1; e += kTeam) exon_cv[exon_slot[e]] = CUDART_NAN;
        Team<kTeam>::sync();
    }
    gene_statistics<kTeam>(cov, g, gene_cov_off, gene_exon_off, gene_strand, exon_len, exon_slot, mask, bias_offset, bias_window, bias_gene_length,
                           gene_mean, gene_std, gene_cv, five_end, three_end, exon_cv, T, s_u64, s_win, win_cap);
}

// kTeam == 32: one warp per gene, the genes with short transcripts and bias windows up to kWarpWindow; kTeam == kThreads: one CTA per
// gene, the rest.  Both kernels run over all genes and skip the ones that are the other's.
template <int kTeam>
__global__ void __launch_bounds__(kThreads)
coverage_stats_kernel(const unsigned int* __restrict__ cov, uint32_t n_genes, const unsigned long long* __restrict__ gene_cov_off,
                      const uint32_t* __restrict__ gene_exon_off, const uint8_t* __restrict__ gene_strand,
                      const uint32_t* __restrict__ exon_len, const uint32_t* __restrict__ exon_slot, uint32_t mask, int bias_offset,
                      int bias_window, uint32_t bias_gene_length, double* __restrict__ gene_mean, double* __restrict__ gene_std,
                      double* __restrict__ gene_cv, unsigned long long* __restrict__ five_end, unsigned long long* __restrict__ three_end,
                      double* __restrict__ exon_cv, const uint8_t* __restrict__ gene_owned, unsigned win_cap) {
    RSQ_DYNAMIC_SMEM(double, s_dyn);
    __shared__ unsigned long long s_u64[kThreads / 32];
    constexpr unsigned kTeams = kThreads / kTeam;
    const uint32_t g = kTeam == 32 ? blockIdx.x * kTeams + (threadIdx.x >> 5) : blockIdx.x;
    if (g >= n_genes) return;
    double* s_win = s_dyn + (kTeam == 32 ? (threadIdx.x >> 5) * 2u * win_cap : 0u);
    gene_entry<kTeam, false>(g, cov, gene_cov_off, gene_exon_off, gene_strand, exon_len, exon_slot, mask, bias_offset, bias_window, bias_gene_length,
                             gene_mean, gene_std, gene_cv, five_end, three_end, exon_cv, gene_owned, s_u64, s_win, win_cap);
}

// Bias windows above kMaxWindow (the reference accepts any --window-size): a fixed number of CTAs walk over all genes, each with
// its two windows in a slice of a global scratch buffer (2 x win_cap doubles per CTA; win_cap = the next power of two of the
// smaller of the window and the longest transcript, rsqd::big_window_cap).  Rare by construction: no attempt at speed.
__global__ void __launch_bounds__(kThreads)
coverage_stats_big_kernel(const unsigned int* __restrict__ cov, uint32_t n_genes, const unsigned long long* __restrict__ gene_cov_off,
                          const uint32_t* __restrict__ gene_exon_off, const uint8_t* __restrict__ gene_strand,
                          const uint32_t* __restrict__ exon_len, const uint32_t* __restrict__ exon_slot, uint32_t mask, int bias_offset,
                          int bias_window, uint32_t bias_gene_length, double* __restrict__ gene_mean, double* __restrict__ gene_std,
                          double* __restrict__ gene_cv, unsigned long long* __restrict__ five_end, unsigned long long* __restrict__ three_end,
                          double* __restrict__ exon_cv, const uint8_t* __restrict__ gene_owned, double* __restrict__ scratch, unsigned win_cap) {
    __shared__ unsigned long long s_u64[kThreads / 32];
    double* s_win = scratch + (size_t)blockIdx.x * 2u * win_cap;
    for (uint32_t g = blockIdx.x; g < n_genes; g += gridDim.x) {
        gene_entry<kThreads, true>(g, cov, gene_cov_off, gene_exon_off, gene_strand, exon_len, exon_slot, mask, bias_offset, bias_window, bias_gene_length,
                                   gene_mean, gene_std, gene_cv, five_end, three_end, exon_cv, gene_owned, s_u64, s_win, win_cap);
        __syncthreads();                                               // the windows and s_u64 are the next gene's
    }
}

__global__ void fill_nan_kernel(double* p, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = CUDART_NAN;
}

}  // namespace

// RSQ_COVSTATS_SMEM_WINDOW (tests): a lower limit sends ordinary windows through the global-scratch kernel
static int smem_window_limit() {
    const char* e = getenv("RSQ_COVSTATS_SMEM_WINDOW");
    const int v = e ? atoi(e) : kMaxWindow;
    return v < 1 || v > kMaxWindow ? kMaxWindow : v;
}
bool coverage_stats_needs_scratch(int bias_window) { return bias_window > smem_window_limit(); }

int launch_coverage_stats(const unsigned int* cov, uint32_t n_genes, const unsigned long long* gene_cov_off, const uint32_t* gene_exon_off,
                          const uint8_t* gene_strand, const uint32_t* exon_len, const uint32_t* exon_slot, uint32_t mask, int bias_offset,
                          int bias_window, uint32_t bias_gene_length, double* gene_mean, double* gene_std, double* gene_cv,
                          unsigned long long* five_end, unsigned long long* three_end, double* exon_cv, uint32_t n_exons,
                          const uint8_t* gene_owned, bool nan_prefill, double* big_scratch, unsigned big_ctas, unsigned big_win_cap,
                          cudaStream_t stream) {
    int n = 0;
    if (n_exons) {
        if (nan_prefill) { RSQ_LAUNCH(fill_nan_kernel, (n_exons + 255) / 256, 256, 0, stream, exon_cv, n_exons); ++n; }
        else cudaMemsetAsync(exon_cv, 0, 8ull * n_exons, stream);       // partitions other than the first: zero bits
    }
    if (n_genes && coverage_stats_needs_scratch(bias_window)) {
        // windows that do not fit shared memory: global scratch, big_ctas CTAs over all genes (the caller sized both)
        RSQ_LAUNCH(coverage_stats_big_kernel, big_ctas, kThreads, 0, stream,
            cov, n_genes, gene_cov_off, gene_exon_off, gene_strand, exon_len, exon_slot, mask, bias_offset, bias_window, bias_gene_length, gene_mean,
            gene_std, gene_cv, five_end, three_end, exon_cv, gene_owned, big_scratch, big_win_cap);
        return n + 1;
    }
    if (n_genes) {
        unsigned p2 = 1;
        while (p2 < (unsigned)bias_window) p2 <<= 1;
        // short transcripts: one warp each (four genes per CTA), windows of up to kWarpWindow values per warp
        const unsigned wcap = std::min<unsigned>(p2, kWarpWindow);
        RSQ_LAUNCH(coverage_stats_kernel<32>, (n_genes + 3) / 4, kThreads, 4 * 2 * wcap * sizeof(double), stream, 
            cov, n_genes, gene_cov_off, gene_exon_off, gene_strand, exon_len, exon_slot, mask, bias_offset, bias_window, bias_gene_length, gene_mean,
            gene_std, gene_cv, five_end, three_end, exon_cv, gene_owned, wcap);
        // the rest: one CTA each; windows of up to kMaxWindow values need more than the default 48 KB of shared memory
        const size_t big = 2 * (size_t)p2 * sizeof(double);
        if (big > 48 * 1024) cudaFuncSetAttribute(coverage_stats_kernel<kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big);
        RSQ_LAUNCH(coverage_stats_kernel<kThreads>, n_genes, kThreads, big, stream, 
            cov, n_genes, gene_cov_off, gene_exon_off, gene_strand, exon_len, exon_slot, mask, bias_offset, bias_window, bias_gene_length, gene_mean,
            gene_std, gene_cv, five_end, three_end, exon_cv, gene_owned, p2);
        n += 2;
    }
    return n;
}

}  // namespace rsqd
