// rsq_eit.h — elementary-interval index over the annotation (host-side builder, plain C++; once per rsq_set_annotation).
//
// The reference answers "which features does this aligned block touch" by walking a start-sorted list per block
// (intersectBlock, src/Expression.cpp:106-117; closed-interval tests, src/GTF.cpp:171-186).  Every answer it can give is
// a function of the ELEMENTARY INTERVAL the block's ends fall into: cut each contig at every exon start, every
// exon end + 1 and every boundary of the flattened gene lines; between two neighbouring cuts the set of features covering
// a base is constant.  For a block [bs, be] (be = bs + len, the reference's closed end with its one-base overhang,
// src/Expression.cpp:45):
//   * features intersecting it            = union of the covering sets of the intervals k(bs) .. k(be)
//   * exons that contain it completely    = exons covering bs whose end >= be - 1      (src/Expression.cpp:341)
// so the index stores, per interval, 12 OR-ed class bits (gene / ribosomal gene / exon / ribosomal exon, by strand) and the
// list of covering exons in list order, with everything the accumulation needs inlined (32-byte CoverEntry).
// A 1 kb direct-address table gives the interval of a bucket start; a search inside one bucket finishes the look-up.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "rsq_device.h"

namespace rsqd {

struct EiDesc { uint32_t cover_off; uint32_t flags_n; };   // flags_n: class bits in the low 16, covering-exon count in the high 16

struct EiIndex {
    std::vector<uint32_t>   ei_off;     // [n_contigs+1] first interval of each contig (its INT32_MIN sentinel)
    std::vector<int32_t>    ei_start;   // [n_ei + kEiPad] interval starts, ascending inside a contig
    std::vector<EiDesc>     ei_desc;    // [n_ei + 1 + kEiPad]
    std::vector<CoverEntry> cover;      // covering exons, interval-major, list order inside an interval
    std::vector<uint32_t>   eib_off;    // [n_contigs+1] first bucket entry of each contig (each contig has n_b + 1 entries)
    std::vector<uint32_t>   eib;        // interval containing the bucket start; last entry of a contig = its last interval
    uint32_t                max_cover = 0;
};

// The reference keeps, per chromosome, ONE list of gene and exon lines, stable-sorted by start (src/RNASeQC.cpp:150-152), and pops
// its front while front.end < Position() for every read that reaches the metrics (trimFeatures, src/Expression.cpp:80-93).  A
// gene line therefore leaves the window at the first such read that starts beyond the largest end of everything at or before the
// line in that list — its prefix maximum.  From then on the gene is "seen": BaseCoverage::commit drops its coverage
// (src/Metrics.cpp:108-112) and its fragmentTracker set has been erased, so read names count again.  Only a gene with an exon
// that reaches past that bound (not a collapsed-annotation shape) can still be hit afterwards.  Per gene index: the bound, or
// INT32_MAX when the annotation carries no gene-line arrays or the gene has no line.
inline void gene_eviction_bounds(const rsq_annotation* a, std::vector<int32_t>& bound) {
    bound.assign(a->n_genes, INT32_MAX);
    if (!a->gline_off || !a->gline_fid || !a->gline_order || !a->exon_order || !a->n_gene_lines) return;
    struct Item { int32_t start, end; uint32_t order; uint32_t fid; };
    std::vector<Item> list;
    for (uint32_t c = 0; c < a->n_contigs; ++c) {
        list.clear();
        for (uint32_t i = a->gline_off[c]; i < a->gline_off[c + 1]; ++i) list.push_back(Item{a->gline_start[i], a->gline_end[i], a->gline_order[i], a->gline_fid[i]});
        for (uint32_t i = a->exon_off[c]; i < a->exon_off[c + 1]; ++i) list.push_back(Item{a->exon_start[i], a->exon_end[i], a->exon_order[i], 0xffffffffu});
        std::sort(list.begin(), list.end(), [](const Item& x, const Item& y) { return x.start != y.start ? x.start < y.start : x.order < y.order; });
        int32_t pmax = INT32_MIN;
        for (const Item& it : list) {
            pmax = std::max(pmax, it.end);
            if (it.fid != 0xffffffffu && it.fid < a->n_genes) bound[it.fid] = std::min(bound[it.fid], pmax);   // a second line of the id (another contig) cannot exist
        }
    }
}

// One contig's share of the index, with offsets relative to the contig (build_ei_index adds the bases when it joins the parts).
struct EiPart {
    std::vector<int32_t>    ei_start;
    std::vector<EiDesc>     ei_desc;
    std::vector<CoverEntry> cover;
    std::vector<uint32_t>   eib;
    uint32_t                max_cover = 0;
    std::string             error;
};

inline void build_ei_contig(const rsq_annotation* a, uint32_t c, const std::vector<int32_t>& late_after, EiPart& P) {
    const uint32_t e0 = a->exon_off[c], e1 = a->exon_off[c + 1], s0 = a->seg_off[c], s1 = a->seg_off[c + 1];
    std::vector<int64_t> cuts;
    cuts.reserve(2 * (size_t)(e1 - e0) + (s1 - s0) + 1);
    cuts.push_back(INT32_MIN);
    for (uint32_t i = s0; i < s1; ++i) cuts.push_back(a->seg_start[i]);
    for (uint32_t i = e0; i < e1; ++i) {
        if (a->exon_end[i] < a->exon_start[i]) continue;                 // covers nothing
        cuts.push_back(a->exon_start[i]);
        if (a->exon_end[i] < INT32_MAX) cuts.push_back((int64_t)a->exon_end[i] + 1);
    }
    std::sort(cuts.begin(), cuts.end());
    cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
    const size_t n = cuts.size();
    P.ei_start.resize(n);
    for (size_t k = 0; k < n; ++k) P.ei_start[k] = (int32_t)cuts[k];
    auto interval_of = [&](int64_t p) -> uint32_t {   // last cut <= p
        return (uint32_t)((size_t)(std::upper_bound(cuts.begin(), cuts.end(), p) - cuts.begin()) - 1);
    };
    // class bits: gene lines from the flattened segments, exons from their own spans
    std::vector<uint32_t> flags(n, 0), cnt(n + 1, 0);
    {
        uint32_t si = s0;
        for (size_t k = 0; k < n; ++k) {
            while (si + 1 < s1 && a->seg_start[si + 1] <= cuts[k]) ++si;
            if (si < s1 && a->seg_start[si] <= cuts[k]) flags[k] = a->seg_flags[si] & 0x3fu;   // RSQ_SEG_* share the low six bits
        }
    }
    std::vector<uint32_t> k_lo(e1 - e0), k_hi(e1 - e0);                  // the intervals each exon spans (looked up once)
    for (uint32_t i = e0; i < e1; ++i) {
        if (a->exon_end[i] < a->exon_start[i]) continue;
        const uint32_t k0 = interval_of(a->exon_start[i]), k1 = interval_of(a->exon_end[i]);
        k_lo[i - e0] = k0; k_hi[i - e0] = k1;
        const uint32_t s = a->exon_flags[i] & RSQ_EXON_STRAND_MASK;
        const uint32_t bits = (EI_EXON << s) | ((a->exon_flags[i] & RSQ_EXON_RIBOSOMAL) ? (EI_EXON_RIBO << s) : 0u);
        for (uint32_t k = k0; k <= k1; ++k) { flags[k] |= bits; ++cnt[k]; }
    }
    size_t total = 0;
    std::vector<size_t> first(n);
    for (size_t k = 0; k < n; ++k) {
        if (cnt[k] > 0xffffu) { P.error = "more than 65535 exons cover one base"; return; }
        first[k] = total; total += cnt[k];
        P.max_cover = std::max(P.max_cover, cnt[k]);
    }
    if (total >= 0xffffffffull) { P.error = "covering-exon lists exceed 2^32 entries"; return; }
    P.cover.resize(total);
    std::vector<uint32_t> fill(n, 0);
    for (uint32_t i = e0; i < e1; ++i) {             // list order inside every interval
        if (a->exon_end[i] < a->exon_start[i]) continue;
        CoverEntry ce;
        ce.start = a->exon_start[i]; ce.end = a->exon_end[i];
        const uint32_t g = a->exon_gene[i];
        uint32_t fl = a->exon_flags[i] & (RSQ_EXON_STRAND_MASK | RSQ_EXON_RIBOSOMAL);
        if (g < a->n_genes && (a->gene_flags[g] & RSQ_GENE_GLOBIN)) fl |= kCoverGlobin;
        ce.gene_flags = (g & 0xffffffu) | (fl << 24);
        ce.slot = a->exon_slot[i]; ce.cov_off = a->exon_cov_off[i]; ce.late_after = g < a->n_genes ? late_after[g] : INT32_MAX; ce.reserved = 0;
        for (uint32_t k = k_lo[i - e0]; k <= k_hi[i - e0]; ++k) P.cover[first[k] + fill[k]++] = ce;
    }
    P.ei_desc.resize(n);
    for (size_t k = 0; k < n; ++k) { P.ei_desc[k].cover_off = (uint32_t)first[k]; P.ei_desc[k].flags_n = flags[k] | (cnt[k] << 16); }
    // direct-address table: interval containing each bucket start, one extra entry = the contig's last interval
    const int64_t top = std::max<int64_t>(cuts.back(), 0);
    const uint32_t n_b = (uint32_t)(top >> kEiShift) + 2u;
    P.eib.reserve((size_t)n_b + 1);
    size_t k = 0;
    for (uint32_t b = 0; b < n_b; ++b) {
        const int64_t p = (int64_t)b << kEiShift;
        while (k + 1 < n && cuts[k + 1] <= p) ++k;
        P.eib.push_back((uint32_t)k);
    }
    P.eib.push_back((uint32_t)(n - 1));
}

// Returns an empty string on success, else the reason the index cannot be built.  The contigs are independent: they are built
// on a few threads (largest first) and joined in contig order — the index is on the start-up path of every run.
inline std::string build_ei_index(const rsq_annotation* a, EiIndex& X) {
    const uint32_t nc = a->n_contigs;
    X.ei_off.assign(nc + 1, 0); X.eib_off.assign(nc + 1, 0);
    X.ei_start.clear(); X.ei_desc.clear(); X.cover.clear(); X.eib.clear();
    X.max_cover = 0;
    std::vector<int32_t> late_after;
    gene_eviction_bounds(a, late_after);
    std::vector<EiPart> parts(nc);
    std::vector<uint32_t> order(nc);
    for (uint32_t c = 0; c < nc; ++c) order[c] = c;
    std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return a->exon_off[x + 1] - a->exon_off[x] > a->exon_off[y + 1] - a->exon_off[y]; });
    std::atomic<uint32_t> next{0};
    auto work = [&] { for (uint32_t k; (k = next.fetch_add(1)) < nc;) build_ei_contig(a, order[k], late_after, parts[order[k]]); };
    const unsigned n_threads = a->n_exons < 20000 ? 1u : std::min<unsigned>(std::min<unsigned>(8u, nc), std::max(1u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    size_t n_ei = 0, n_cover = 0, n_eib = 0;
    for (const EiPart& P : parts) {
        if (!P.error.empty()) return P.error;
        n_ei += P.ei_start.size(); n_cover += P.cover.size(); n_eib += P.eib.size();
    }
    if (n_cover >= 0xffffffffull) return "covering-exon lists exceed 2^32 entries";
    X.ei_start.reserve(n_ei + kEiPad); X.ei_desc.reserve(n_ei + 1 + kEiPad); X.cover.reserve(n_cover + 1); X.eib.reserve(n_eib);
    for (uint32_t c = 0; c < nc; ++c) {
        const EiPart& P = parts[c];
        const uint32_t base = (uint32_t)X.ei_start.size(), cover_base = (uint32_t)X.cover.size();
        X.ei_off[c] = base;
        X.eib_off[c] = (uint32_t)X.eib.size();
        X.ei_start.insert(X.ei_start.end(), P.ei_start.begin(), P.ei_start.end());
        for (const EiDesc& d : P.ei_desc) X.ei_desc.push_back(EiDesc{d.cover_off + cover_base, d.flags_n});
        X.cover.insert(X.cover.end(), P.cover.begin(), P.cover.end());
        for (uint32_t k : P.eib) X.eib.push_back(base + k);
        X.max_cover = std::max(X.max_cover, P.max_cover);
    }
    X.ei_off[nc] = (uint32_t)X.ei_start.size();
    X.eib_off[nc] = (uint32_t)X.eib.size();
    X.ei_desc.insert(X.ei_desc.end(), 1 + kEiPad, EiDesc{(uint32_t)X.cover.size(), 0u});   // padded like ei_start: staged by bulk copies too
    X.ei_start.insert(X.ei_start.end(), kEiPad, INT32_MAX);
    X.cover.push_back(CoverEntry{0, -1, 0, 0, 0, INT32_MAX, 0});   // spare entry: the fast path loads cover[cover_off] even when the count is 0
    return std::string();
}

// R[L] = floor(2^64 / L) for 2 <= L < 65536 (frac_q64 in rsq_hot_core.cuh); R[0] and R[1] are never read.
inline void build_recip64(std::vector<unsigned long long>& R) {
    R.assign(65536, 0ull);
    for (unsigned long long L = 2; L < 65536; ++L) R[L] = ~0ull / L + ((~0ull % L) == L - 1 ? 1ull : 0ull);
}

}  // namespace rsqd
