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
#include <string>
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

// Returns an empty string on success, else the reason the index cannot be built.
inline std::string build_ei_index(const rsq_annotation* a, EiIndex& X) {
    const uint32_t nc = a->n_contigs;
    X.ei_off.assign(nc + 1, 0); X.eib_off.assign(nc + 1, 0);
    X.ei_start.clear(); X.ei_desc.clear(); X.cover.clear(); X.eib.clear();
    std::vector<int64_t> cuts;
    std::vector<uint32_t> cnt, fill;
    std::vector<int32_t> late_after;
    gene_eviction_bounds(a, late_after);
    for (uint32_t c = 0; c < nc; ++c) {
        const uint32_t e0 = a->exon_off[c], e1 = a->exon_off[c + 1], s0 = a->seg_off[c], s1 = a->seg_off[c + 1];
        cuts.clear();
        cuts.push_back(INT32_MIN);
        for (uint32_t i = s0; i < s1; ++i) cuts.push_back(a->seg_start[i]);
        for (uint32_t i = e0; i < e1; ++i) {
            if (a->exon_end[i] < a->exon_start[i]) continue;                 // covers nothing
            cuts.push_back(a->exon_start[i]);
            if (a->exon_end[i] < INT32_MAX) cuts.push_back((int64_t)a->exon_end[i] + 1);
        }
        std::sort(cuts.begin(), cuts.end());
        cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
        const size_t base = X.ei_start.size(), n = cuts.size();
        X.ei_off[c] = (uint32_t)base;
        for (int64_t v : cuts) X.ei_start.push_back((int32_t)v);
        auto interval_of = [&](int64_t p) -> size_t {   // last cut <= p
            return (size_t)(std::upper_bound(cuts.begin(), cuts.end(), p) - cuts.begin()) - 1;
        };
        // class bits: gene lines from the flattened segments, exons from their own spans
        std::vector<uint32_t> flags(n, 0);
        {
            uint32_t si = s0;
            for (size_t k = 0; k < n; ++k) {
                while (si + 1 < s1 && a->seg_start[si + 1] <= cuts[k]) ++si;
                if (si < s1 && a->seg_start[si] <= cuts[k]) flags[k] = a->seg_flags[si] & 0x3fu;   // RSQ_SEG_* share the low six bits
            }
        }
        cnt.assign(n + 1, 0);
        for (uint32_t i = e0; i < e1; ++i) {
            if (a->exon_end[i] < a->exon_start[i]) continue;
            const size_t k0 = interval_of(a->exon_start[i]), k1 = interval_of(a->exon_end[i]);
            const uint32_t s = a->exon_flags[i] & RSQ_EXON_STRAND_MASK;
            const uint32_t bits = (EI_EXON << s) | ((a->exon_flags[i] & RSQ_EXON_RIBOSOMAL) ? (EI_EXON_RIBO << s) : 0u);
            for (size_t k = k0; k <= k1; ++k) { flags[k] |= bits; ++cnt[k]; }
        }
        size_t total = X.cover.size();
        std::vector<size_t> first(n);
        for (size_t k = 0; k < n; ++k) {
            if (cnt[k] > 0xffffu) return "more than 65535 exons cover one base";
            first[k] = total; total += cnt[k];
            X.max_cover = std::max(X.max_cover, cnt[k]);
        }
        if (total >= 0xffffffffull) return "covering-exon lists exceed 2^32 entries";
        X.cover.resize(total);
        fill.assign(n, 0);
        for (uint32_t i = e0; i < e1; ++i) {             // list order inside every interval
            if (a->exon_end[i] < a->exon_start[i]) continue;
            const size_t k0 = interval_of(a->exon_start[i]), k1 = interval_of(a->exon_end[i]);
            CoverEntry ce;
            ce.start = a->exon_start[i]; ce.end = a->exon_end[i];
            const uint32_t g = a->exon_gene[i];
            uint32_t fl = a->exon_flags[i] & (RSQ_EXON_STRAND_MASK | RSQ_EXON_RIBOSOMAL);
            if (g < a->n_genes && (a->gene_flags[g] & RSQ_GENE_GLOBIN)) fl |= kCoverGlobin;
            ce.gene_flags = (g & 0xffffffu) | (fl << 24);
            ce.slot = a->exon_slot[i]; ce.cov_off = a->exon_cov_off[i]; ce.late_after = g < a->n_genes ? late_after[g] : INT32_MAX; ce.reserved = 0;
            for (size_t k = k0; k <= k1; ++k) X.cover[first[k] + fill[k]++] = ce;
        }
        for (size_t k = 0; k < n; ++k) {
            EiDesc d; d.cover_off = (uint32_t)first[k]; d.flags_n = flags[k] | (cnt[k] << 16);
            X.ei_desc.push_back(d);
        }
        // direct-address table: interval containing each bucket start, one extra entry = the contig's last interval
        const int64_t top = std::max<int64_t>(cuts.back(), 0);
        const uint32_t n_b = (uint32_t)(top >> kEiShift) + 2u;
        X.eib_off[c] = (uint32_t)X.eib.size();
        size_t k = 0;
        for (uint32_t b = 0; b < n_b; ++b) {
            const int64_t p = (int64_t)b << kEiShift;
            while (k + 1 < n && cuts[k + 1] <= p) ++k;
            X.eib.push_back((uint32_t)(base + k));
        }
        X.eib.push_back((uint32_t)(base + n - 1));
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
