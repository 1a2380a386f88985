// hostcore.cpp — TEST INFRASTRUCTURE.  Host emulation of hot_path_kernel<false,false,false> (rnaseqc_b200/csrc/rsq_kernels.cu):
// the same elementary-interval index (rsq_eit.h) and the same per-record logic (rsq_hot_core.cuh, compiled here for the CPU),
// driven tile by tile exactly as a warp drives it — 32 records, a 128-entry slice of interval starts "staged" at a
// 4-entry-aligned base, blocks outside the slice through the 1 kb table — with plain adds in place of the atomics and a
// std::unordered_set in place of the device hash set.  tests/test_hostcore.py compares its tables with the CPU oracle, so
// the index builder and the record logic are parity-checked without a GPU; only the warp plumbing (TMA, mbarrier, shuffles)
// is left to the -m gpu tests.  Never linked into the product.
#include <cuda_runtime.h>   // vector types only (int2, int4, uint2); no CUDA call is made
#include <stdint.h>
#include <string.h>

#include <unordered_set>
#include <vector>

#include "rsq_b200.h"
#include "rsq_device.h"
#include "rsq_eit.h"
#include "rsq_hot_core.cuh"

using namespace rsqd;

namespace {
struct HostMem {
    std::unordered_set<unsigned long long>* set;
    void add64(unsigned long long* p, unsigned long long v) const { *p += v; }
    void exon_one(const Accum& acc, uint32_t slot) const { acc.exon_whole[slot] += 1ull; }
    void exon_frac(const Accum& acc, uint32_t slot, uint32_t hi, uint32_t lo) const { acc.exon_fhi[slot] += hi; acc.exon_flo[slot] += lo; }
    void cov_range(unsigned int* p, unsigned int len) const { p[0] += 1u; p[len] -= 1u; }
    bool frag_insert(unsigned long long key) const { return set->insert(key).second; }
};
#ifndef RSQ_STAGE_N
#define RSQ_STAGE_N 128
#endif
constexpr int kStageN = RSQ_STAGE_N;
}  // namespace

extern "C" int hostcore_run(const rsq_annotation* a, const rsq_config* cfg, const rsq_read* reads, size_t n_reads, const rsq_block* blocks,
                            size_t n_blocks, uint64_t* counters, uint64_t* gene_reads, uint64_t* gene_unique, uint64_t* gene_fragments,
                            uint64_t* exon_whole, uint64_t* exon_fhi, uint64_t* exon_flo, uint32_t* cov_diff, uint64_t* algo_bytes,
                            uint64_t* stats /* [8]: blocks served from the staged slice, blocks through the table, intervals, cover entries,
                                                records on the fast path, on the general path, tiles, tiles with a general-path record */) {
    EiIndex X;
    if (!build_ei_index(a, X).empty()) return 1;
    AnnView ann;
    memset(&ann, 0, sizeof ann);
    ann.n_contigs = a->n_contigs; ann.n_genes = a->n_genes; ann.n_exons = a->n_exons;
    ann.ei_off = X.ei_off.data(); ann.ei_start = X.ei_start.data(); ann.ei_desc = reinterpret_cast<const uint2*>(X.ei_desc.data());
    ann.cover = X.cover.data(); ann.eib_off = X.eib_off.data(); ann.eib = X.eib.data();
    std::vector<unsigned long long> recip;
    build_recip64(recip);
    ann.recip64 = recip.data();
    Params prm;
    memset(&prm, 0, sizeof prm);
    prm.mapq_threshold = cfg->mapq_threshold; prm.base_mismatch_threshold = cfg->base_mismatch_threshold;
    prm.chimeric_distance = cfg->chimeric_distance; prm.strand_mode = cfg->strand_mode; prm.unpaired = cfg->unpaired;
    prm.exclude_chimeric = cfg->exclude_chimeric; prm.n_filter_tags = cfg->n_filter_tags; prm.collect_coverage = cfg->collect_coverage;
    Accum acc;
    memset(&acc, 0, sizeof acc);
    acc.gene_reads = (unsigned long long*)gene_reads; acc.gene_unique = (unsigned long long*)gene_unique;
    acc.gene_fragments = (unsigned long long*)gene_fragments; acc.exon_whole = (unsigned long long*)exon_whole;
    acc.exon_fhi = (unsigned long long*)exon_fhi; acc.exon_flo = (unsigned long long*)exon_flo; acc.cov_diff = cov_diff;
    std::unordered_set<unsigned long long> fragset;
    HostMem mem{&fragset};
    stats[0] = stats[1] = stats[4] = stats[5] = stats[6] = stats[7] = 0; stats[2] = X.ei_start.size() - kEiPad; stats[3] = X.cover.size();

    int32_t staged[kStageN];
#ifdef RSQ_STAGE_DESC
    uint2 staged_descs[kStageN];
#endif
    const size_t n_tiles = (n_reads + 31) / 32;
    for (size_t tile = 0; tile < n_tiles; ++tile) {
        const size_t t0 = tile * 32, t1 = std::min(n_reads, t0 + 32);
        // the warp-uniform part of the kernel
        int pmin = INT32_MAX, cmin = INT32_MAX, cmax = INT32_MIN;
        for (size_t i = t0; i < t1; ++i) { pmin = std::min(pmin, reads[i].pos); cmin = std::min<int>(cmin, reads[i].contig); cmax = std::max<int>(cmax, reads[i].contig); }
        TileWin w;
#ifdef RSQ_STAGE_DESC
        w.sd = staged_descs;
#endif
        w.st = staged; w.base = 0; w.k0 = 0; w.hi = 0; w.wmin = 0; w.lastv = 0; w.on = false;
        if (cmin == cmax && cmin >= 0 && (uint32_t)cmin < ann.n_contigs && pmin >= 0 && pmin < INT32_MAX) {
            const int c_hi = (int)ann.ei_off[cmin + 1];
            const int bk_lo = (int)ann.eib_off[cmin], bk_n = (int)ann.eib_off[cmin + 1] - bk_lo - 1;
            w.wmin = pmin + 1;
            w.k0 = (int)ann.eib[bk_lo + std::min(w.wmin >> kEiShift, bk_n - 1)];
            w.base = w.k0 & ~3;
            w.hi = std::min(w.base + kStageN, c_hi);
            w.on = true;
            memcpy(staged, ann.ei_start + w.base, sizeof staged);          // the bulk copy
#ifdef RSQ_STAGE_DESC
            memcpy(staged_descs, ann.ei_desc + w.base, sizeof staged_descs);
#endif
            w.lastv = w.hi == c_hi ? INT32_MAX : staged[w.hi - 1 - w.base];
        }
        bool tile_general = false;
        ++stats[6];
        for (size_t i = t0; i < t1; ++i) {
            const rsq_read& r = reads[i];
            uint32_t nb = r.n_blocks;
            if ((size_t)r.block_off + nb > n_blocks) nb = 0;
            RecOut o;
            filter_cascade(prm, true, r.flag, r.mapq, r.aux, r.filter_mask, r.pos, r.mpos, r.endpos, r.nm, nb, o);
            if (o.go) {
                const bool cleared = (r.aux & RSQ_AUX_CONTIG_DROPPED) || r.contig < 0 || (uint32_t)r.contig >= ann.n_contigs;
                const int rstrand = read_strand(prm, r.flag);
                if (!cleared)
                    for (uint32_t b = 0; b < nb; ++b) {
                        const rsq_block& k = blocks[r.block_off + b];
                        ++stats[(w.on && k.start >= w.wmin && k.start + k.len < w.lastv) ? 0 : 1];
                    }
                ExonOut x;
                exon_metrics(mem, ann, w, prm, acc, r.flag, o.hq, cleared, r.contig, r.pos, rstrand, blocks + r.block_off, nb, r.name_hash, o, x);
                ++stats[o.fast ? 4 : 5];
                if (!o.fast) tile_general = true;
            }
            // the epilogue, one record at a time
            if (o.agg_gene != kNoGene) {
                gene_reads[o.agg_gene] += 1;
                if (o.agg_unique) gene_unique[o.agg_gene] += 1;
            }
            if (o.frag_key) {                                              // resolve_parked, without the one-tile delay
                const bool fresh = fragset.insert(o.frag_key).second;
                if (fresh) { gene_fragments[o.agg_gene] += 1; counters[RSQ_C_FRAGSET_ENTRIES] += 1; }
            }
            for (int k = 0; k < 64; ++k) if ((o.cm >> k) & 1ull) counters[k] += 1;
            auto addv = [&](int k, int v) { counters[k] += (uint64_t)(int64_t)v; };
            if (o.vflags & 1u) { addv(RSQ_C_END1_MISMATCHES, r.nm); addv(RSQ_C_END1_BASES, r.l_qseq); }
            if (o.vflags & 2u) { addv(RSQ_C_END2_MISMATCHES, r.nm); addv(RSQ_C_END2_BASES, r.l_qseq); }
            if (o.vflags & 4u) addv(RSQ_C_MISMATCHED_BASES, r.nm);
            if (o.vflags & 8u) addv(RSQ_C_TOTAL_BASES, r.l_qseq);
            if (o.vflags & 16u) addv(RSQ_C_ALIGNMENT_BLOCKS, (int)nb);
            counters[RSQ_C_FRAGSET_ENTRIES] += (uint64_t)o.n_fresh;
            *algo_bytes += o.my_bytes;
        }
        if (tile_general) ++stats[7];
    }
    return 0;
}
