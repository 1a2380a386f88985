// hostlegacy.cpp — TEST INFRASTRUCTURE.  Host emulation of hot_path_kernel<false, kLegacy=true, false>: the very source the CUDA
// kernel runs under --legacy (rnaseqc_b200/csrc/rsq_legacy.cuh, included here unchanged) compiled for the CPU behind a few shims
// (__ldg = load, atomicAdd = add, the correctly rounded float intrinsics = the host's IEEE operators, the device hash set = a
// std::unordered_set), driven record by record with the kernel's filter cascade (rsq_hot_core.cuh) in front and its counter
// epilogue behind.  tests/test_hostcore.py compares the tables with the CPU oracle, so the legacy device rules are parity-checked
// without a GPU on inputs the -m gpu suite never sees (fuzzed annotations and records).  Never linked into the product.
#include <cuda_runtime.h>   // vector types only
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <unordered_set>
#include <vector>

#include "rsq_b200.h"
#include "rsq_device.h"
#include "rsq_hot_core.cuh"

// ---- shims for the device-only constructs of rsq_legacy.cuh --------------------------------------------------------------------
#undef __device__
#undef __forceinline__
#undef __noinline__
#define __device__
#define __forceinline__ inline
#define __noinline__
namespace {
std::unordered_set<unsigned long long>* g_fragset = nullptr;
}
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T, class U> inline T atomicAdd(T* p, U v) { const T old = *p; *p = (T)(old + (T)v); return old; }
inline uint32_t __float_as_uint(float v) { uint32_t u; memcpy(&u, &v, 4); return u; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }   // IEEE single, round to nearest: what the intrinsic is
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
using std::max;
using std::min;
namespace rsqd {
namespace {
inline int upper_bound_start(const int2* se, int lo, int hi, int key) {
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (se[mid].x > key) hi = mid; else lo = mid + 1; }
    return lo;
}
inline int lower_bound_pmax(const int32_t* pmax, int lo, int hi, int key) {
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (pmax[mid] >= key) hi = mid; else lo = mid + 1; }
    return lo;
}
inline bool set_insert(unsigned long long*, unsigned long long, unsigned long long key, unsigned int*, unsigned) { return g_fragset->insert(key).second; }
}  // namespace
}  // namespace rsqd
#include "rsq_legacy.cuh"

using namespace rsqd;

extern "C" int hostlegacy_run(const rsq_annotation* a, const rsq_config* cfg, const rsq_read* reads, size_t n_reads, const rsq_block* blocks,
                              size_t n_blocks, uint64_t* counters, uint64_t* gene_reads, uint64_t* gene_unique, uint64_t* gene_fragments,
                              uint64_t* exon_whole, uint64_t* exon_fhi, uint64_t* exon_flo, uint32_t* cov_diff, uint64_t* algo_bytes) {
    if (!a->gline_off || !a->exon_order) return 1;
    // the device layout of rsq_set_annotation's legacy branch (rsq_capi.cu)
    std::vector<int2> se(a->n_exons), gse(a->n_gene_lines);
    std::vector<ExonMeta> meta(a->n_exons);
    std::vector<uint32_t> ggf(a->n_gene_lines);
    std::vector<int32_t> pmax(std::max<uint32_t>(a->n_exons, 1)), gpmax(std::max<uint32_t>(a->n_gene_lines, 1));
    for (uint32_t i = 0; i < a->n_exons; ++i) {
        se[i] = make_int2(a->exon_start[i], a->exon_end[i]);
        meta[i].gene_flags = a->exon_gene[i] | ((uint32_t)a->exon_flags[i] << 24);
        meta[i].slot = a->exon_slot[i];
        meta[i].cov_off = a->exon_cov_off[i];
    }
    for (uint32_t i = 0; i < a->n_gene_lines; ++i) {
        gse[i] = make_int2(a->gline_start[i], a->gline_end[i]);
        ggf[i] = (a->gline_gene[i] & 0xffffffu) | ((uint32_t)a->gline_flags[i] << 24);
    }
    for (uint32_t c = 0; c < a->n_contigs; ++c) {                                  // prefix_max_kernel
        int m = INT32_MIN;
        for (uint32_t i = a->exon_off[c]; i < a->exon_off[c + 1]; ++i) { m = std::max(m, se[i].y); pmax[i] = m; }
        m = INT32_MIN;
        for (uint32_t i = a->gline_off[c]; i < a->gline_off[c + 1]; ++i) { m = std::max(m, gse[i].y); gpmax[i] = m; }
    }
    AnnView ann;
    memset(&ann, 0, sizeof ann);
    ann.n_contigs = a->n_contigs; ann.n_genes = a->n_genes; ann.n_exons = a->n_exons;
    ann.exon_off = a->exon_off; ann.exon_se = se.data(); ann.exon_pmax = pmax.data(); ann.exon_meta = meta.data(); ann.gene_flags = a->gene_flags;
    ann.gline_off = a->gline_off; ann.gline_se = gse.data(); ann.gline_pmax = gpmax.data(); ann.gline_gf = ggf.data();
    ann.gline_order = a->gline_order; ann.exon_order = a->exon_order;
    Params prm;
    memset(&prm, 0, sizeof prm);
    prm.mapq_threshold = cfg->mapq_threshold; prm.base_mismatch_threshold = cfg->base_mismatch_threshold;
    prm.chimeric_distance = cfg->chimeric_distance; prm.strand_mode = cfg->strand_mode; prm.unpaired = cfg->unpaired;
    prm.exclude_chimeric = cfg->exclude_chimeric; prm.n_filter_tags = cfg->n_filter_tags; prm.collect_coverage = cfg->collect_coverage;
    prm.legacy = 1;
    Accum acc;
    memset(&acc, 0, sizeof acc);
    acc.gene_reads = (unsigned long long*)gene_reads; acc.gene_unique = (unsigned long long*)gene_unique;
    acc.gene_fragments = (unsigned long long*)gene_fragments; acc.exon_whole = (unsigned long long*)exon_whole;
    acc.exon_fhi = (unsigned long long*)exon_fhi; acc.exon_flo = (unsigned long long*)exon_flo; acc.cov_diff = cov_diff;
    std::unordered_set<unsigned long long> fragset;
    g_fragset = &fragset;

    for (size_t i = 0; i < n_reads; ++i) {
        const rsq_read& r = reads[i];
        uint32_t nb = r.n_blocks;
        if ((size_t)r.block_off + nb > n_blocks) nb = 0;
        RecOut o;
        filter_cascade(prm, true, r.flag, r.mapq, r.aux, r.filter_mask, r.pos, r.mpos, r.endpos, r.nm, nb, o);
        if (o.go) {                                                                // the kLegacy branch of hot_path_kernel
            LegacyRead lr;
            lr.contig = r.contig; lr.pos = r.pos; lr.endpos = r.endpos; lr.l_qseq = r.l_qseq; lr.rstrand = read_strand(prm, r.flag);
            lr.flag = r.flag; lr.nb = nb;
            lr.cleared = (r.aux & RSQ_AUX_CONTIG_DROPPED) || r.contig < 0 || (uint32_t)r.contig >= ann.n_contigs;
            lr.hq = o.hq; lr.name_hash = r.name_hash; lr.bp = blocks + r.block_off;
            legacy_metrics(ann, acc, prm, lr, o.cm, o.n_fresh, o.my_bytes);
        }
        for (int k = 0; k < 64; ++k) if ((o.cm >> k) & 1ull) counters[k] += 1;     // the epilogue, one record at a time
        auto addv = [&](int k, int v) { counters[k] += (uint64_t)(int64_t)v; };
        if (o.vflags & 1u) { addv(RSQ_C_END1_MISMATCHES, r.nm); addv(RSQ_C_END1_BASES, r.l_qseq); }
        if (o.vflags & 2u) { addv(RSQ_C_END2_MISMATCHES, r.nm); addv(RSQ_C_END2_BASES, r.l_qseq); }
        if (o.vflags & 4u) addv(RSQ_C_MISMATCHED_BASES, r.nm);
        if (o.vflags & 8u) addv(RSQ_C_TOTAL_BASES, r.l_qseq);
        if (o.vflags & 16u) addv(RSQ_C_ALIGNMENT_BLOCKS, (int)nb);
        counters[RSQ_C_FRAGSET_ENTRIES] += (uint64_t)o.n_fresh;
        *algo_bytes += o.my_bytes;
    }
    g_fragset = nullptr;
    return 0;
}
