// rsq_capi.cu — C-ABI (include/rsq_b200.h) over the sm_100a kernels: device memory, streams, staging.
// One rsq_ctx drives one GPU.  There is no CPU fallback: without a CUDA device rsq_create fails with
// RSQ_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>      // types and enums only: the library is bound at run time (dlopen), never linked

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "rsq_b200.h"
#include "rsq_device.h"
#include "rsq_eit.h"
#include "rsq_hot_core.cuh"

namespace {

using namespace rsqd;

std::string g_create_error;

struct Ctx {
    rsq_config cfg;
    std::string err;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    bool own_stream = true;
    // annotation
    bool have_ann = false, have_bed = false;
    uint32_t n_contigs = 0, n_genes = 0, n_exons = 0, n_segs = 0;
    uint64_t cov_len = 0;
    AnnView ann{};
    BedView bed{};
    std::vector<void*> ann_allocs, bed_allocs, ref_allocs;
    bool have_ref = false;
    RefView ref{};
    const uint8_t* d_ref_present = nullptr;
    const unsigned long long* d_cascade_lut = nullptr;   // the filter cascade as a table (rsq_hot_core.cuh), uploaded with the annotation
    unsigned long long gc_cap = 0;
    size_t gc_cand_cap = 0;
    // accumulators
    Accum acc{};
    unsigned int* d_cov = nullptr;            // integrated coverage (finalize)
    unsigned long long* d_exon_out = nullptr; // canonical (whole | frac) exon sums for the read-back
    unsigned int* d_scan = nullptr;           // scan scratch
    size_t scan_elems = 0;
    unsigned long long frag_cap = 0, pair_cap = 0;
    // contexts that are reset and run again (one sample after another): a second fragment-tracker table is zeroed on a side stream
    // while the hot kernel fills the first one, and rsq_reset swaps the two — the 3.2 GB memset (200 M records) leaves the step
    unsigned long long* frag_spare = nullptr;
    unsigned long long frag_spare_cap = 0;
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_spare_clean = nullptr, ev_reset_mark = nullptr;
    unsigned resets = 0;
    unsigned long long frag_known = 0, frag_pending = 0;   // entries read back from the device / upper bound of those in flight
    unsigned long long exon_depth = 1;                     // deepest stack of overlapping exons (inserts per record <= this)
    unsigned long long* d_export = nullptr;
    // coordinate-partitioned runs (rsq_set_partition / rsq_comm_init)
    uint32_t part_rank = 0, part_world = 1;
    unsigned long long* d_found = nullptr;    // [part_world] per-partition sample counts (phase 1 of the exchange)
    bool found_valid = false;                 // d_found holds the reduced counts of the current run
    ncclComm_t comm = nullptr;
    // rsq_coverage_statistics buffers
    unsigned long long* cs_off = nullptr; uint32_t *cs_eoff = nullptr, *cs_len = nullptr, *cs_slot = nullptr; uint8_t* cs_strand = nullptr;
    double* cs_out = nullptr; unsigned long long* cs_ends = nullptr;   // one allocation: mean | std | cv | exon_cv | five | three
    uint8_t* cs_owned = nullptr;
    double* cs_big = nullptr;                 // window scratch of the wide-window statistics kernel
    size_t cs_big_elems = 0;
    size_t cs_cap_genes = 0, cs_cap_refs = 0;
    size_t export_words = 0;
    unsigned long long* d_hist = nullptr;
    size_t hist_len = 0;
    unsigned int* d_pair_scratch = nullptr;
    size_t pair_scratch_elems = 0;
    size_t sample_cap = 0;                    // capacity of sample_flag / sample_isize / cand (records per sub-batch)
    // staging for host submits (two slots)
    rsq_read* d_reads[2] = {nullptr, nullptr};
    rsq_block* d_blocks[2] = {nullptr, nullptr};
    size_t cap_reads[2] = {0, 0}, cap_blocks[2] = {0, 0};
    // rsq_submit_packed: wire-format staging per slot, scan temporaries shared (compute stream only)
    rsq_read_packed* d_packed[2] = {nullptr, nullptr};
    rsq_block* d_xblocks[2] = {nullptr, nullptr};
    size_t cap_packed[2] = {0, 0}, cap_xblocks[2] = {0, 0};
    unsigned char* d_compact[2] = {nullptr, nullptr};   // rsq_submit_compact: the batch buffer as it came over the wire
    size_t cap_compact[2] = {0, 0};
    unsigned int *d_up_tot = nullptr, *d_up_expl = nullptr, *d_up_scan = nullptr;
    size_t cap_up = 0;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    bool slot_used[2] = {false, false};
    unsigned slot = 0;
    // bookkeeping
    unsigned long long records_submitted = 0;
    unsigned long long mate_launches = 0;     // launches with a mate map on: numbers the rounds of the mate-map stages
    unsigned long long mate_known[2] = {0, 0}, mate_pending[2] = {0, 0};   // candidates read back / records launched since (BED, GC)
    unsigned long long launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timers;
    size_t timers_used = 0;
    double kernel_ms = 0.0;
};

constexpr size_t kSubBatch = 1u << 24;   // records per hot-kernel launch with the BED / GC samplers on (bounds their candidate buffers)
#ifndef RSQ_PLAIN_SUB_BATCH_LOG2
#define RSQ_PLAIN_SUB_BATCH_LOG2 27
#endif
constexpr size_t kSubBatchPlain = (size_t)1 << RSQ_PLAIN_SUB_BATCH_LOG2;   // without them: fewer launches, fewer kernel tails

#define CU_TRY(expr)                                                                                      \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            c->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                                 \
            return RSQ_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

// ---- NCCL, bound at run time ---------------------------------------------------------------------------------------------
// A host process that already carries a libnccl (PyTorch does) must not get a second copy: RTLD_NOLOAD finds the loaded one by
// its soname first; only then is the system library opened.  RSQ_NCCL_LIB overrides both.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};
NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* env = getenv("RSQ_NCCL_LIB");
    if (env && *env) api.handle = dlopen(env, RTLD_NOW | RTLD_LOCAL);
    if (!api.handle) api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!api.handle) api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!api.handle) { const char* e = dlerror(); api.why = std::string("libnccl not found: ") + (e ? e : "?"); return api; }
    auto sym = [&](const char* n) { void* p = dlsym(api.handle, n); if (!p) api.why = std::string("libnccl lacks ") + n; return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!api.why.empty()) api.handle = nullptr;
    return api;
}
static_assert(sizeof(ncclUniqueId) == RSQ_COMM_ID_BYTES, "rsq_comm_unique_id hands out a whole ncclUniqueId");
#define NCCL_TRY(expr)                                                                                    \
    do {                                                                                                  \
        ncclResult_t r__ = (expr);                                                                        \
        if (r__ != ncclSuccess) {                                                                         \
            c->err = std::string(#expr) + ": " + nccl_api().GetErrorString(r__);                          \
            return RSQ_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

unsigned long long sample_quota(const struct Ctx* c);

unsigned long long next_pow2(unsigned long long v) {
    unsigned long long p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <typename T>
int upload(Ctx* c, std::vector<void*>& owner, const T* host, size_t n, const T** out) {
    T* d = nullptr;
    CU_TRY(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
    owner.push_back(d);
    if (n) CU_TRY(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    *out = d;
    return RSQ_OK;
}

unsigned long long sample_quota(const Ctx* c) { return (c->have_bed && c->cfg.fragment_samples) ? c->cfg.fragment_samples : 0ull; }

void free_list(std::vector<void*>& v) {
    for (void* p : v) cudaFree(p);
    v.clear();
}

// The fragment-tracker table of the next run.  First run of a context (and one-shot users, the CLI): zeroed in line.  From the second
// rsq_reset on, a spare table of the same size exists: it was zeroed on the side stream while the previous run's kernels filled the
// other one, so the reset only swaps the two and sends the used one to the side stream — after everything queued on the compute
// stream so far (the kernels that wrote it), and the compute stream waits for the clean one before it goes on.
int fresh_frag_table(Ctx* c, bool from_reset) {
    Accum& a = c->acc;
    const size_t bytes = 8ull * c->frag_cap;
    if (from_reset && c->frag_spare && c->frag_spare_cap == c->frag_cap) {
        CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_spare_clean, 0));
        std::swap(a.frag_set, c->frag_spare);
    } else {
        if (c->frag_spare) {                                           // the table has grown since: the spare no longer fits
            CU_TRY(cudaStreamSynchronize(c->aux_stream));
            cudaFree(c->frag_spare);
            c->frag_spare = nullptr; c->frag_spare_cap = 0;
        }
        CU_TRY(cudaMemsetAsync(a.frag_set, 0, bytes, c->stream));
        if (from_reset && ++c->resets >= 2 && !getenv("RSQ_NO_SPARE_TABLE")) {
            if (!c->aux_stream) {
                CU_TRY(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
                CU_TRY(cudaEventCreateWithFlags(&c->ev_spare_clean, cudaEventDisableTiming));
                CU_TRY(cudaEventCreateWithFlags(&c->ev_reset_mark, cudaEventDisableTiming));
            }
            if (cudaMalloc(&c->frag_spare, bytes) == cudaSuccess) c->frag_spare_cap = c->frag_cap;
            else { c->frag_spare = nullptr; cudaGetLastError(); }      // no room for a second table: resets keep zeroing in line
        }
    }
    if (c->frag_spare) {
        CU_TRY(cudaEventRecord(c->ev_reset_mark, c->stream));
        CU_TRY(cudaStreamWaitEvent(c->aux_stream, c->ev_reset_mark, 0));
        CU_TRY(cudaMemsetAsync(c->frag_spare, 0, bytes, c->aux_stream));
        CU_TRY(cudaEventRecord(c->ev_spare_clean, c->aux_stream));
    }
    return RSQ_OK;
}

int zero_accumulators(Ctx* c, bool from_reset = false) {
    Accum& a = c->acc;
    CU_TRY(cudaMemsetAsync(a.counters, 0, 8 * RSQ_N_COUNTERS, c->stream));
    CU_TRY(cudaMemsetAsync(a.gene_reads, 0, 8ull * c->n_genes, c->stream));
    CU_TRY(cudaMemsetAsync(a.gene_unique, 0, 8ull * c->n_genes, c->stream));
    CU_TRY(cudaMemsetAsync(a.gene_fragments, 0, 8ull * c->n_genes, c->stream));
    CU_TRY(cudaMemsetAsync(a.exon_whole, 0, 8ull * c->n_exons, c->stream));
    CU_TRY(cudaMemsetAsync(a.exon_fhi, 0, 8ull * c->n_exons, c->stream));
    CU_TRY(cudaMemsetAsync(a.exon_flo, 0, 8ull * c->n_exons, c->stream));
    CU_TRY(cudaMemsetAsync(a.cov_diff, 0, 4ull * (c->cov_len + 1), c->stream));
    { const int rc = fresh_frag_table(c, from_reset); if (rc != RSQ_OK) return rc; }
    CU_TRY(cudaMemsetAsync(a.algo_bytes, 0, 8, c->stream));
    CU_TRY(cudaMemsetAsync(a.error_flag, 0, 4, c->stream));
    if (c->have_bed && c->cfg.fragment_samples) {
        // keys and pending mates only: the round scratch (pair_first / pair_second) holds tags of earlier launches, and the launch
        // number keeps growing across resets, so every later tag beats them (run_batch)
        CU_TRY(cudaMemsetAsync(a.pair_keys, 0, 8ull * c->pair_cap, c->stream));
        CU_TRY(cudaMemsetAsync(a.pair_payload, 0xff, 8ull * c->pair_cap, c->stream));      // kMateEmpty
        CU_TRY(cudaMemsetAsync(a.sample_count, 0, 8, c->stream));
    }
    if (c->have_ref) {
        CU_TRY(cudaMemsetAsync(a.gc_keys, 0, 8ull * c->gc_cap, c->stream));
        CU_TRY(cudaMemsetAsync(a.gc_payload, 0xff, 8ull * c->gc_cap, c->stream));
        CU_TRY(cudaMemsetAsync(a.gc_hist, 0, 8 * 100, c->stream));
    }
    c->records_submitted = 0;
    c->mate_known[0] = c->mate_known[1] = c->mate_pending[0] = c->mate_pending[1] = 0;
    if (a.mate_totals) CU_TRY(cudaMemsetAsync(a.mate_totals, 0, 16, c->stream));
    c->found_valid = false;
    c->frag_known = 0; c->frag_pending = 0;
    c->kernel_ms = 0.0;
    c->timers_used = 0;
    return RSQ_OK;
}

int ensure_bed_buffers(Ctx* c, size_t n_reads) {
    if (!(c->have_bed && c->cfg.fragment_samples)) return RSQ_OK;
    const size_t need = std::min(n_reads, kSubBatch);
    if (need > c->sample_cap) {
        CU_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(c->acc.sample_flag); cudaFree(c->acc.sample_isize); cudaFree(c->acc.cand); cudaFree(c->d_pair_scratch);
        c->sample_cap = need;
        CU_TRY(cudaMalloc(&c->acc.sample_flag, need));
        CU_TRY(cudaMalloc(&c->acc.sample_isize, 4 * need));
        CU_TRY(cudaMalloc(&c->acc.cand, sizeof(PairCandidate) * need));
        c->acc.cand_capacity = (unsigned int)need;
        c->pair_scratch_elems = pair_scan_scratch_elems(need);
        CU_TRY(cudaMalloc(&c->d_pair_scratch, 4 * c->pair_scratch_elems));
    }
    return RSQ_OK;
}

// Keeps the fragment-tracker set from filling up.  A record inserts one entry per gene it counts toward, at most `exon_depth`
// (the deepest stack of overlapping exons in the annotation) — a bound that real data stays far below (0.7 entries per record
// on the GENCODE-scale workload, where the depth is 3).  The table is sized 2 x max_records (any number of slots, not the next
// power of two: it is zeroed on every reset; measured: at 1.5 x the longer probe chains cost the hot kernel 10 %, more than the
// smaller memset saves), so typical data leaves it about one-third full.  The guarantee is against OVERFLOW: the host lets a
// launch go while the pessimistic bound (known entries + depth x records in flight) stays below 80 % of the capacity; beyond
// that it synchronises to read the exact entry count (a device counter), and if even that leaves too little room for the whole
// launch, the launch is cut down to what fits (*n shrinks; the caller comes back with the rest) — only when not even 4 M records
// fit is the table grown (rehash).
int reserve_frag_capacity(Ctx* c, size_t* n) {
    const unsigned long long limit = c->frag_cap / 5 * 4;
    auto fits = [&](unsigned long long known, unsigned long long add) { return known + add <= limit; };
    unsigned long long add = (unsigned long long)*n * c->exon_depth;
    if (fits(c->frag_known + c->frag_pending, add)) { c->frag_pending += add; return RSQ_OK; }
    CU_TRY(cudaStreamSynchronize(c->stream));
    unsigned long long exact = 0;
    CU_TRY(cudaMemcpy(&exact, c->acc.counters + RSQ_C_FRAGSET_ENTRIES, 8, cudaMemcpyDeviceToHost));
    c->frag_known = exact;
    c->frag_pending = 0;
    if (!fits(exact, add)) {
        const unsigned long long room = limit > exact ? limit - exact : 0ull;
        const size_t n_fit = (size_t)(room / c->exon_depth);
        if (n_fit >= std::min<size_t>(*n, (size_t)1 << 22)) {
            *n = std::min(*n, n_fit);
            add = (unsigned long long)*n * c->exon_depth;
        } else {
            const unsigned long long new_cap = (2ull * (exact + add) + 1024) & ~3ull;
            unsigned long long* nt = nullptr;
            CU_TRY(cudaMalloc(&nt, 8ull * new_cap));
            CU_TRY(cudaMemsetAsync(nt, 0, 8ull * new_cap, c->stream));
            c->launches += launch_rehash(c->acc.frag_set, c->frag_cap, nt, new_cap, c->acc.error_flag, c->stream);
            CU_TRY(cudaStreamSynchronize(c->stream));
            cudaFree(c->acc.frag_set);
            c->acc.frag_set = nt;
            c->acc.frag_cap = new_cap;
            c->frag_cap = new_cap;
        }
    }
    c->frag_pending = add;
    return RSQ_OK;
}

// The mate maps (BED fragment sizes, --fasta fragment GC) hold one entry per read name that ever was a candidate — a fraction of
// the records.  They are sized by that, not by the records: a launch of n records may go while (candidates known + records
// launched since the last look + n) stays below 3/4 of the capacity; beyond that the host synchronises, reads the exact number of
// candidates so far (a device counter, an upper bound of the names), and grows the table only if even that leaves no room.
int reserve_mate_capacity(Ctx* c, int which, size_t n) {
    const bool on = which == 0 ? (c->have_bed && c->cfg.fragment_samples) : (c->have_ref && !c->cfg.legacy);
    if (!on) return RSQ_OK;
    unsigned long long& cap = which == 0 ? c->pair_cap : c->gc_cap;
    unsigned long long& known = c->mate_known[which];
    unsigned long long& pending = c->mate_pending[which];
    if (known + pending + n <= cap / 4 * 3) { pending += n; return RSQ_OK; }
    CU_TRY(cudaStreamSynchronize(c->stream));
    unsigned long long exact = 0;
    CU_TRY(cudaMemcpy(&exact, c->acc.mate_totals + which, 8, cudaMemcpyDeviceToHost));
    known = exact; pending = n;
    if (exact + 2 * n <= cap / 4 * 3) return RSQ_OK;                   // room for this launch and the next without another look
    const unsigned long long new_cap = next_pow2((exact + 4 * n) / 3 * 4 + 4);   // ... and after a growth for four
    unsigned long long *nk = nullptr, *nf = nullptr, *ns = nullptr, *np = nullptr;
    CU_TRY(cudaMalloc(&nk, 8ull * new_cap));
    CU_TRY(cudaMalloc(&nf, 8ull * new_cap));
    CU_TRY(cudaMalloc(&ns, 8ull * new_cap));
    CU_TRY(cudaMalloc(&np, 8ull * new_cap));
    CU_TRY(cudaMemsetAsync(nk, 0, 8ull * new_cap, c->stream));
    CU_TRY(cudaMemsetAsync(nf, 0, 8ull * new_cap, c->stream));
    CU_TRY(cudaMemsetAsync(ns, 0, 8ull * new_cap, c->stream));
    CU_TRY(cudaMemsetAsync(np, 0xff, 8ull * new_cap, c->stream));
    Accum& a = c->acc;
    unsigned long long*& keys = which == 0 ? a.pair_keys : a.gc_keys;
    unsigned long long*& first = which == 0 ? a.pair_first : a.gc_first;
    unsigned long long*& second = which == 0 ? a.pair_second : a.gc_second;
    unsigned long long*& payload = which == 0 ? a.pair_payload : a.gc_payload;
    c->launches += launch_pair_rehash(keys, payload, cap, nk, np, new_cap - 1, a.error_flag, c->stream);
    CU_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(keys); cudaFree(first); cudaFree(second); cudaFree(payload);
    keys = nk; first = nf; second = ns; payload = np;
    (which == 0 ? a.pair_mask : a.gc_mask) = new_cap - 1;
    cap = new_cap;
    return RSQ_OK;
}

// per-launch candidate list of the GC-content sampler
int ensure_gc_buffers(Ctx* c, size_t n) {
    if (!c->have_ref || c->cfg.legacy) return RSQ_OK;
    const size_t need_cand = std::min(n, kSubBatch);
    if (need_cand > c->gc_cand_cap) {
        CU_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(c->acc.gc_cand);
        c->gc_cand_cap = need_cand;
        CU_TRY(cudaMalloc(&c->acc.gc_cand, sizeof(GcCandidate) * need_cand));
        c->acc.gc_cand_capacity = (unsigned int)need_cand;
    }
    return RSQ_OK;
}

int run_batch(Ctx* c, const rsq_read* d_reads, size_t n_reads, const rsq_block* d_blocks, size_t n_blocks) {
    int rc = ensure_gc_buffers(c, n_reads);
    if (rc != RSQ_OK) return rc;
    rc = ensure_bed_buffers(c, n_reads);
    if (rc != RSQ_OK) return rc;
    Params prm{};
    prm.mapq_threshold = c->cfg.mapq_threshold;
    prm.base_mismatch_threshold = c->cfg.base_mismatch_threshold;
    prm.chimeric_distance = c->cfg.chimeric_distance;
    prm.strand_mode = c->cfg.strand_mode;
    prm.unpaired = c->cfg.unpaired;
    prm.exclude_chimeric = c->cfg.exclude_chimeric;
    prm.n_filter_tags = c->cfg.n_filter_tags;
    prm.collect_coverage = c->cfg.collect_coverage;
    prm.bed_enabled = (c->have_bed && c->cfg.fragment_samples) ? 1u : 0u;
    prm.fragment_samples = c->cfg.fragment_samples;
    prm.legacy = c->cfg.legacy;
    prm.gc_enabled = (c->have_ref && !c->cfg.legacy) ? 1u : 0u;
    prm.ref_present = c->d_ref_present;
    prm.cascade_lut = c->d_cascade_lut;
    prm.cov_len_diag = c->cov_len + 1;
    // timing events around the kernels of this batch
    if (c->timers_used == c->timers.size()) {
        if (c->timers.size() >= 1024) {   // fold what is finished so far
            CU_TRY(cudaStreamSynchronize(c->stream));
            for (size_t k = 0; k < c->timers_used; ++k) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, c->timers[k].first, c->timers[k].second);
                c->kernel_ms += ms;
            }
            c->timers_used = 0;
        } else {
            cudaEvent_t a, b;
            CU_TRY(cudaEventCreate(&a));
            CU_TRY(cudaEventCreate(&b));
            c->timers.emplace_back(a, b);
        }
    }
    auto& tm = c->timers[c->timers_used++];
    CU_TRY(cudaEventRecord(tm.first, c->stream));
    const size_t sub = (prm.bed_enabled || prm.gc_enabled) ? kSubBatch : kSubBatchPlain;
    for (size_t base = 0, n = 0; base < n_reads; base += n) {
        n = std::min(sub, n_reads - base);
        rc = reserve_frag_capacity(c, &n);      // may cut the launch down to the room left in the fragment-tracker set
        if (rc != RSQ_OK) return rc;
        if ((rc = reserve_mate_capacity(c, 0, n)) != RSQ_OK || (rc = reserve_mate_capacity(c, 1, n)) != RSQ_OK) return rc;
        prm.record_base = c->records_submitted + base;
        if (prm.bed_enabled || prm.gc_enabled) {
            if (++c->mate_launches >= (1ull << 23)) {   // the round tags have run out of launch numbers (they survive resets): start over
                Accum& a = c->acc;
                if (a.pair_first) { CU_TRY(cudaMemsetAsync(a.pair_first, 0, 8ull * c->pair_cap, c->stream)); CU_TRY(cudaMemsetAsync(a.pair_second, 0, 8ull * c->pair_cap, c->stream)); }
                if (a.gc_first) { CU_TRY(cudaMemsetAsync(a.gc_first, 0, 8ull * c->gc_cap, c->stream)); CU_TRY(cudaMemsetAsync(a.gc_second, 0, 8ull * c->gc_cap, c->stream)); }
                c->mate_launches = 1;
            }
            prm.mate_round = c->mate_launches << 17;
            CU_TRY(cudaMemsetAsync(c->acc.mate_left_count, 0, 8, c->stream));
        }
        if (prm.bed_enabled) {
            CU_TRY(cudaMemsetAsync(c->acc.cand_count, 0, 4, c->stream));
            CU_TRY(cudaMemsetAsync(c->acc.sample_flag, 0, n, c->stream));
        }
        if (prm.gc_enabled) CU_TRY(cudaMemsetAsync(c->acc.gc_cand_count, 0, 4, c->stream));
        c->launches += launch_hot_path(d_reads + base, d_blocks, n, n_blocks, c->ann, c->bed, c->acc, prm, c->stream);
        if (prm.bed_enabled) c->launches += launch_pair_stages(n, c->acc, prm, c->d_pair_scratch, c->stream);
        if (prm.gc_enabled) c->launches += launch_gc_stages(n, c->ref, c->acc, prm, c->stream);
    }
    CU_TRY(cudaEventRecord(tm.second, c->stream));
    CU_TRY(cudaGetLastError());
    c->records_submitted += n_reads;
    return RSQ_OK;
}

}  // namespace

extern "C" {

const char* rsq_backend(void) { return "cuda-sm_100a"; }
uint32_t rsq_abi_version(void) { return RSQ_ABI_VERSION; }

void* rsq_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void rsq_host_free(void* p) { if (p) cudaFreeHost(p); }

int rsq_create(const rsq_config* cfg, rsq_ctx** out) {
    if (!cfg || !out || cfg->struct_size != sizeof(rsq_config)) { g_create_error = "bad rsq_config"; return RSQ_ERR_INVALID; }
    if (cfg->n_filter_tags > 8) { g_create_error = "at most 8 filter tags"; return RSQ_ERR_INVALID; }
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev <= 0 || cfg->device < 0 || cfg->device >= n_dev) {
        g_create_error = std::string("no usable CUDA device (this build has no CPU fallback): ") + cudaGetErrorString(e);
        cudaGetLastError();
        return RSQ_ERR_NO_DEVICE;
    }
    if (cudaSetDevice(cfg->device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return RSQ_ERR_CUDA; }
    Ctx* c = new Ctx();
    c->cfg = *cfg;
    c->own_stream = cfg->user_stream == 0;
    if (!c->own_stream) c->stream = reinterpret_cast<cudaStream_t>(cfg->user_stream);
    if ((c->own_stream && cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_create_error = "cudaStreamCreate failed";
        delete c;
        return RSQ_ERR_CUDA;
    }
    for (int k = 0; k < 2; ++k) {
        cudaEventCreateWithFlags(&c->ev_copied[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming);
    }
    *out = reinterpret_cast<rsq_ctx*>(c);
    return RSQ_OK;
}

void rsq_destroy(rsq_ctx* ctx) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    cudaDeviceSynchronize();
    if (c->comm && nccl_api().handle) nccl_api().CommDestroy(c->comm);
    free_list(c->ann_allocs);
    free_list(c->bed_allocs);
    free_list(c->ref_allocs);
    Accum& a = c->acc;
    void* ptrs[] = {a.gc_keys, a.gc_first, a.gc_second, a.gc_payload, a.pair_second, a.mate_left, a.mate_left_count, a.mate_totals, a.gc_cand, a.gc_cand_count, a.gc_hist, a.counters, a.gene_reads, a.gene_unique, a.gene_fragments, a.exon_whole, a.exon_fhi, a.exon_flo, c->d_exon_out, a.cov_diff, a.frag_set,
                    a.algo_bytes, a.error_flag, a.pair_keys, a.pair_first, a.pair_payload, a.cand, a.cand_count, a.sample_flag,
                    a.sample_isize, a.sample_vals, a.sample_count, c->d_cov, c->d_scan, c->d_export, c->cs_off, c->cs_eoff, c->cs_len, c->cs_slot, c->cs_strand, c->cs_out, c->cs_owned, c->cs_big, c->d_found, c->d_hist, c->d_pair_scratch,
                    c->d_reads[0], c->d_reads[1], c->d_blocks[0], c->d_blocks[1], c->d_packed[0], c->d_packed[1], c->d_xblocks[0], c->d_xblocks[1],
                    c->d_up_tot, c->d_up_expl, c->d_up_scan, c->d_compact[0], c->d_compact[1]};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto& t : c->timers) { cudaEventDestroy(t.first); cudaEventDestroy(t.second); }
    for (int k = 0; k < 2; ++k) { cudaEventDestroy(c->ev_copied[k]); cudaEventDestroy(c->ev_done[k]); }
    if (c->frag_spare) cudaFree(c->frag_spare);
    if (c->aux_stream) { cudaEventDestroy(c->ev_spare_clean); cudaEventDestroy(c->ev_reset_mark); cudaStreamDestroy(c->aux_stream); }
    if (c->own_stream) cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->copy_stream);
    delete c;
}

const char* rsq_last_error(const rsq_ctx* ctx) {
    return ctx ? reinterpret_cast<const Ctx*>(ctx)->err.c_str() : g_create_error.c_str();
}

int rsq_set_annotation(rsq_ctx* ctx, const rsq_annotation* a) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !a || a->struct_size != sizeof(rsq_annotation)) return RSQ_ERR_INVALID;
    if (c->have_ann) { c->err = "annotation already set"; return RSQ_ERR_INVALID; }
    if (a->n_genes >= (1u << 24)) { c->err = "more than 2^24 genes"; return RSQ_ERR_INVALID; }
    CU_TRY(cudaSetDevice(c->cfg.device));
    c->n_contigs = a->n_contigs; c->n_genes = a->n_genes; c->n_segs = a->n_segs; c->cov_len = a->cov_len;
    c->n_exons = a->n_exon_slots ? a->n_exon_slots : a->n_exons;   // accumulators and read-backs are per output slot
    if (c->n_exons < a->n_exons) { c->err = "n_exon_slots smaller than n_exons"; return RSQ_ERR_INVALID; }
    if (c->cfg.legacy && (!a->gline_off || !a->exon_order || (a->n_gene_lines && (!a->gline_start || !a->gline_end || !a->gline_gene || !a->gline_flags || !a->gline_order)))) {
        c->err = "legacy mode needs the gene-line arrays of rsq_annotation"; return RSQ_ERR_INVALID;
    }
    // RSQ_SETUP_STATS=1: where the time of this call goes (host-side index build against device allocation and upload)
    const bool setup_stats = getenv("RSQ_SETUP_STATS") != nullptr;
    auto t_mark = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!setup_stats) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "rsq_set_annotation: %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_mark).count());
        t_mark = now;
    };
    // repack host-side into the device layout: (start,end) pairs and 16-byte metadata
    std::vector<int2> se((size_t)a->n_exons + 4, make_int2(INT32_MAX, INT32_MAX));   // padded for the rounded-up bulk copies
    std::vector<ExonMeta> meta((size_t)a->n_exons + 4);
    memset(meta.data(), 0, meta.size() * sizeof(ExonMeta));
    for (uint32_t i = 0; i < a->n_exons; ++i) {
        se[i] = make_int2(a->exon_start[i], a->exon_end[i]);
        meta[i].gene_flags = a->exon_gene[i] | ((uint32_t)a->exon_flags[i] << 24);
        meta[i].slot = a->exon_slot[i];
        meta[i].cov_off = a->exon_cov_off[i];
    }
    AnnView& v = c->ann;
    v.n_contigs = a->n_contigs; v.n_genes = a->n_genes; v.n_exons = a->n_exons; v.n_segs = a->n_segs;
    int rc;
    if ((rc = upload(c, c->ann_allocs, a->exon_off, (size_t)a->n_contigs + 1, &v.exon_off))) return rc;
    if ((rc = upload(c, c->ann_allocs, se.data(), se.size(), &v.exon_se))) return rc;
    if ((rc = upload(c, c->ann_allocs, meta.data(), meta.size(), &v.exon_meta))) return rc;
    if ((rc = upload(c, c->ann_allocs, a->gene_flags, (size_t)a->n_genes, &v.gene_flags))) return rc;
    int32_t* pmax = nullptr;
    CU_TRY(cudaMalloc(&pmax, std::max<size_t>(a->n_exons, 1) * 4));
    c->ann_allocs.push_back(pmax);
    c->launches += launch_prefix_max(v.exon_se, v.exon_off, a->n_contigs, pmax, c->stream);
    v.exon_pmax = pmax;
    // gene lines (--legacy rules)
    v.gline_off = nullptr; v.gline_se = nullptr; v.gline_pmax = nullptr; v.gline_gf = nullptr; v.gline_order = nullptr; v.exon_order = nullptr;
    if (c->cfg.legacy) {
        std::vector<int2> gse((size_t)a->n_gene_lines);
        std::vector<uint32_t> ggf((size_t)a->n_gene_lines);
        for (uint32_t i = 0; i < a->n_gene_lines; ++i) {
            gse[i] = make_int2(a->gline_start[i], a->gline_end[i]);
            ggf[i] = (a->gline_gene[i] & 0xffffffu) | ((uint32_t)a->gline_flags[i] << 24);
        }
        if ((rc = upload(c, c->ann_allocs, a->gline_off, (size_t)a->n_contigs + 1, &v.gline_off))) return rc;
        if ((rc = upload(c, c->ann_allocs, gse.data(), gse.size(), &v.gline_se))) return rc;
        if ((rc = upload(c, c->ann_allocs, ggf.data(), ggf.size(), &v.gline_gf))) return rc;
        if ((rc = upload(c, c->ann_allocs, a->gline_order, (size_t)a->n_gene_lines, &v.gline_order))) return rc;
        if ((rc = upload(c, c->ann_allocs, a->exon_order, (size_t)a->n_exons, &v.exon_order))) return rc;
        int32_t* gpmax = nullptr;
        CU_TRY(cudaMalloc(&gpmax, std::max<size_t>(a->n_gene_lines, 1) * 4));
        c->ann_allocs.push_back(gpmax);
        c->launches += launch_prefix_max(v.gline_se, v.gline_off, a->n_contigs, gpmax, c->stream);
        v.gline_pmax = gpmax;
        CU_TRY(cudaStreamSynchronize(c->stream));   // gse / ggf are locals
    }
    lap("exon tables");
    // deepest stack of overlapping exons: bounds the fragment-set inserts of one record
    {
        unsigned long long depth = 1;
        for (uint32_t k = 0; k < a->n_contigs; ++k) {
            std::vector<std::pair<int32_t, int>> ev;
            for (uint32_t i = a->exon_off[k]; i < a->exon_off[k + 1]; ++i) { ev.emplace_back(a->exon_start[i], +1); ev.emplace_back(a->exon_end[i] + 1, -1); }
            std::sort(ev.begin(), ev.end());
            long long cur = 0;
            for (auto& e : ev) { cur += e.second; if (cur > (long long)depth) depth = (unsigned long long)cur; }
        }
        c->exon_depth = depth;
    }
    lap("exon depth");
    // elementary-interval index (rsq_eit.h): what the hot path searches instead of the exon list
    if (!c->cfg.legacy) {
        EiIndex X;
        const std::string why = build_ei_index(a, X);
        lap("elementary-interval index");
        if (!why.empty()) { c->err = "annotation cannot be indexed: " + why; return RSQ_ERR_INVALID; }
        static_assert(sizeof(EiDesc) == sizeof(uint2), "descriptor is read as one 64-bit word");
        if ((rc = upload(c, c->ann_allocs, X.ei_off.data(), X.ei_off.size(), &v.ei_off))) return rc;
        if ((rc = upload(c, c->ann_allocs, X.ei_start.data(), X.ei_start.size(), &v.ei_start))) return rc;
        if ((rc = upload(c, c->ann_allocs, reinterpret_cast<const uint2*>(X.ei_desc.data()), X.ei_desc.size(), &v.ei_desc))) return rc;
        if ((rc = upload(c, c->ann_allocs, X.cover.data(), X.cover.size(), &v.cover))) return rc;
        if ((rc = upload(c, c->ann_allocs, X.eib_off.data(), X.eib_off.size(), &v.eib_off))) return rc;
        if ((rc = upload(c, c->ann_allocs, X.eib.data(), X.eib.size(), &v.eib))) return rc;
        std::vector<unsigned long long> recip;
        build_recip64(recip);
        if ((rc = upload(c, c->ann_allocs, recip.data(), recip.size(), &v.recip64))) return rc;
        CU_TRY(cudaStreamSynchronize(c->stream));   // X and recip are locals
    } else {
        v.recip64 = nullptr;
        v.ei_off = nullptr; v.ei_start = nullptr; v.ei_desc = nullptr; v.cover = nullptr; v.eib_off = nullptr; v.eib = nullptr;
    }
    lap("index upload + reciprocals");
#ifndef RSQ_NO_CASCADE_LUT
    {   // the filter cascade as a pure function of 17 bits of a record: tabulated once per context
        std::vector<unsigned long long> lut(kCascadeEntries);
        for (uint32_t ix = 0; ix < kCascadeEntries; ++ix) lut[ix] = cascade_entry(c->cfg.legacy != 0, c->cfg.exclude_chimeric != 0, c->cfg.unpaired != 0, ix);
        if ((rc = upload(c, c->ann_allocs, lut.data(), lut.size(), &c->d_cascade_lut))) return rc;
        CU_TRY(cudaStreamSynchronize(c->stream));   // lut is a local
    }
#endif
    lap("cascade table");
    // accumulators
    Accum& acc = c->acc;
    CU_TRY(cudaMalloc(&acc.counters, 8 * RSQ_N_COUNTERS));
    CU_TRY(cudaMalloc(&acc.gene_reads, 8ull * std::max(a->n_genes, 1u)));
    CU_TRY(cudaMalloc(&acc.gene_unique, 8ull * std::max(a->n_genes, 1u)));
    CU_TRY(cudaMalloc(&acc.gene_fragments, 8ull * std::max(a->n_genes, 1u)));
    CU_TRY(cudaMalloc(&acc.exon_whole, 8ull * std::max(c->n_exons, 1u)));
    CU_TRY(cudaMalloc(&acc.exon_fhi, 8ull * std::max(c->n_exons, 1u)));
    CU_TRY(cudaMalloc(&acc.exon_flo, 8ull * std::max(c->n_exons, 1u)));
    CU_TRY(cudaMalloc(&c->d_exon_out, 16ull * std::max(c->n_exons, 1u)));
    CU_TRY(cudaMalloc(&acc.cov_diff, 4ull * (a->cov_len + 1)));
    CU_TRY(cudaMalloc(&c->d_cov, 4ull * (a->cov_len + 1)));
    c->scan_elems = coverage_scan_scratch_elems(a->cov_len + 1);
    CU_TRY(cudaMalloc(&c->d_scan, 4ull * c->scan_elems));
    {   // fragment-tracker set: RSQ_FRAG_SLOTS_PER_RECORD_X2 / 2 slots per expected record (default 2.0), any number of slots
#ifndef RSQ_FRAG_SLOTS_PER_RECORD_X2
#define RSQ_FRAG_SLOTS_PER_RECORD_X2 4
#endif
        const unsigned long long recs = c->cfg.max_records ? c->cfg.max_records : (1ull << 24);
        unsigned long long cap = std::max<unsigned long long>(recs * RSQ_FRAG_SLOTS_PER_RECORD_X2 / 2, 1ull << 12);
#ifdef RSQ_FRAG_POW2
        cap = next_pow2(cap);
#endif
        c->frag_cap = (cap + 3) & ~3ull;
    }
    CU_TRY(cudaMalloc(&acc.frag_set, 8ull * c->frag_cap));
    acc.frag_cap = c->frag_cap;
    CU_TRY(cudaMalloc(&acc.algo_bytes, 8));
    CU_TRY(cudaMalloc(&acc.error_flag, 8));   // [0] the flag; [1] spare
    // the export buffer of the end-of-run exchange is sized on first use (its tail depends on rsq_set_bed)
    c->have_ann = true;
    rc = zero_accumulators(c);
    if (rc != RSQ_OK) return rc;
    CU_TRY(cudaStreamSynchronize(c->stream));   // host staging vectors go out of scope
    lap("accumulators (alloc + zero)");
    return RSQ_OK;
}

int rsq_set_reference(rsq_ctx* ctx, const rsq_reference* r) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !r || r->struct_size != sizeof(rsq_reference)) return RSQ_ERR_INVALID;
    if (!c->have_ann) { c->err = "set the annotation first"; return RSQ_ERR_INVALID; }
    if (c->have_ref) { c->err = "reference already set"; return RSQ_ERR_INVALID; }
    if (c->records_submitted) { c->err = "set the reference before the first submit"; return RSQ_ERR_INVALID; }
    if (r->n_contigs != c->n_contigs) { c->err = "reference and annotation contig spaces differ"; return RSQ_ERR_INVALID; }
    CU_TRY(cudaSetDevice(c->cfg.device));
    int rc;
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
    if ((rc = upload(c, c->ref_allocs, r->present, (size_t)r->n_contigs, &c->ref.present))) return rc;
    if ((rc = upload(c, c->ref_allocs, reinterpret_cast<const unsigned long long*>(r->length), (size_t)r->n_contigs, &c->ref.length))) return rc;
    if ((rc = upload(c, c->ref_allocs, reinterpret_cast<const unsigned long long*>(r->word_off), (size_t)r->n_contigs + 1, &c->ref.word_off))) return rc;
    if ((rc = upload(c, c->ref_allocs, r->gc_bits, (size_t)r->word_off[r->n_contigs], &c->ref.bits))) return rc;
    c->d_ref_present = c->ref.present;
    Accum& acc = c->acc;
    c->gc_cap = next_pow2(std::min<unsigned long long>(c->cfg.max_records ? c->cfg.max_records : (1ull << 24), kSubBatch) / 3 * 4 + 4);   // grows with the candidates
    CU_TRY(cudaMalloc(&acc.gc_keys, 8ull * c->gc_cap));
    CU_TRY(cudaMalloc(&acc.gc_first, 8ull * c->gc_cap));
    CU_TRY(cudaMalloc(&acc.gc_second, 8ull * c->gc_cap));
    CU_TRY(cudaMalloc(&acc.gc_payload, 8ull * c->gc_cap));
    acc.gc_mask = c->gc_cap - 1;
    if (!acc.mate_left) {
        CU_TRY(cudaMalloc(&acc.mate_left, 4ull * 2 * kMateLeftCap)); CU_TRY(cudaMalloc(&acc.mate_left_count, 8)); CU_TRY(cudaMalloc(&acc.mate_totals, 16));
        CU_TRY(cudaMemsetAsync(acc.mate_totals, 0, 16, c->stream));
    }
    CU_TRY(cudaMalloc(&acc.gc_cand_count, 4));
    CU_TRY(cudaMalloc(&acc.gc_hist, 8 * 100));
    CU_TRY(cudaMemsetAsync(acc.gc_keys, 0, 8ull * c->gc_cap, c->stream));
    CU_TRY(cudaMemsetAsync(acc.gc_first, 0, 8ull * c->gc_cap, c->stream));
    CU_TRY(cudaMemsetAsync(acc.gc_second, 0, 8ull * c->gc_cap, c->stream));
    CU_TRY(cudaMemsetAsync(acc.gc_payload, 0xff, 8ull * c->gc_cap, c->stream));
    CU_TRY(cudaMemsetAsync(acc.gc_hist, 0, 8 * 100, c->stream));
    CU_TRY(cudaMemsetAsync(acc.gc_cand_count, 0, 4, c->stream));
    c->have_ref = true;
    CU_TRY(cudaStreamSynchronize(c->stream));
    return RSQ_OK;
}

int rsq_set_bed(rsq_ctx* ctx, const rsq_bed* b) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !b || b->struct_size != sizeof(rsq_bed)) return RSQ_ERR_INVALID;
    if (!c->have_ann) { c->err = "set the annotation first"; return RSQ_ERR_INVALID; }
    if (c->have_bed) { c->err = "BED already set"; return RSQ_ERR_INVALID; }
    if (b->n_contigs != c->n_contigs) { c->err = "BED and annotation contig spaces differ"; return RSQ_ERR_INVALID; }
    CU_TRY(cudaSetDevice(c->cfg.device));
    std::vector<int2> se(b->n_intervals);
    for (uint32_t i = 0; i < b->n_intervals; ++i) se[i] = make_int2(b->start[i], b->end[i]);
    BedView& v = c->bed;
    v.n_intervals = b->n_intervals;
    int rc;
    if ((rc = upload(c, c->bed_allocs, b->off, (size_t)b->n_contigs + 1, &v.off))) return rc;
    if ((rc = upload(c, c->bed_allocs, se.data(), se.size(), &v.se))) return rc;
    if ((rc = upload(c, c->bed_allocs, b->name_id, (size_t)b->n_intervals, &v.name_id))) return rc;
    if ((rc = upload(c, c->bed_allocs, b->contig_present, (size_t)b->n_contigs, &v.present))) return rc;
    int32_t* pmax = nullptr;
    CU_TRY(cudaMalloc(&pmax, std::max<size_t>(b->n_intervals, 1) * 4));
    c->bed_allocs.push_back(pmax);
    c->launches += launch_prefix_max(v.se, v.off, b->n_contigs, pmax, c->stream);
    v.pmax = pmax;
    {   // the same elementary-interval index as the annotation's, over the BED intervals (name id in the slot field): a block that
        // lies between two cuts intersects exactly the intervals covering it (fragmentSizeMetrics, src/Expression.cpp:490-507)
        rsq_annotation fake;
        memset(&fake, 0, sizeof fake);
        std::vector<uint32_t> gene(b->n_intervals, 0u), seg_off(b->n_contigs + 1);
        std::vector<uint8_t> flags(b->n_intervals, 0), seg_flags(b->n_contigs, 0), gene_flags(1, 0);
        std::vector<uint64_t> cov(b->n_intervals, 0);
        std::vector<int32_t> seg_start(b->n_contigs, INT32_MIN);
        for (uint32_t k = 0; k <= b->n_contigs; ++k) seg_off[k] = k;
        fake.struct_size = sizeof fake; fake.n_contigs = b->n_contigs; fake.n_genes = 1; fake.n_exons = b->n_intervals; fake.n_segs = b->n_contigs;
        fake.exon_off = b->off; fake.exon_start = b->start; fake.exon_end = b->end; fake.exon_gene = gene.data(); fake.exon_slot = b->name_id;
        fake.exon_flags = flags.data(); fake.exon_cov_off = cov.data(); fake.seg_off = seg_off.data(); fake.seg_start = seg_start.data();
        fake.seg_flags = seg_flags.data(); fake.gene_flags = gene_flags.data();
        EiIndex X;
        const std::string why = build_ei_index(&fake, X);
        if (!why.empty()) { c->err = "BED intervals cannot be indexed: " + why; return RSQ_ERR_INVALID; }
        if ((rc = upload(c, c->bed_allocs, X.ei_off.data(), X.ei_off.size(), &v.ei_off))) return rc;
        if ((rc = upload(c, c->bed_allocs, X.ei_start.data(), X.ei_start.size(), &v.ei_start))) return rc;
        if ((rc = upload(c, c->bed_allocs, reinterpret_cast<const uint2*>(X.ei_desc.data()), X.ei_desc.size(), &v.ei_desc))) return rc;
        if ((rc = upload(c, c->bed_allocs, X.cover.data(), X.cover.size(), &v.cover))) return rc;
        if ((rc = upload(c, c->bed_allocs, X.eib_off.data(), X.eib_off.size(), &v.eib_off))) return rc;
        if ((rc = upload(c, c->bed_allocs, X.eib.data(), X.eib.size(), &v.eib))) return rc;
        CU_TRY(cudaStreamSynchronize(c->stream));   // X is a local
    }
    c->have_bed = true;
    if (c->cfg.fragment_samples) {
        Accum& acc = c->acc;
        c->pair_cap = next_pow2(std::min<unsigned long long>(c->cfg.max_records ? c->cfg.max_records : (1ull << 24), kSubBatch) / 3 * 4 + 4);   // grows with the candidates
        CU_TRY(cudaMalloc(&acc.pair_keys, 8ull * c->pair_cap));
        CU_TRY(cudaMalloc(&acc.pair_first, 8ull * c->pair_cap));
        CU_TRY(cudaMalloc(&acc.pair_second, 8ull * c->pair_cap));
        CU_TRY(cudaMalloc(&acc.pair_payload, 8ull * c->pair_cap));
        acc.pair_mask = c->pair_cap - 1;
        if (!acc.mate_left) {
        CU_TRY(cudaMalloc(&acc.mate_left, 4ull * 2 * kMateLeftCap)); CU_TRY(cudaMalloc(&acc.mate_left_count, 8)); CU_TRY(cudaMalloc(&acc.mate_totals, 16));
        CU_TRY(cudaMemsetAsync(acc.mate_totals, 0, 16, c->stream));
    }
        CU_TRY(cudaMalloc(&acc.cand_count, 4));
        CU_TRY(cudaMalloc(&acc.sample_vals, 4ull * c->cfg.fragment_samples));
        CU_TRY(cudaMalloc(&acc.sample_count, 8));
        CU_TRY(cudaMemsetAsync(acc.pair_keys, 0, 8ull * c->pair_cap, c->stream));
        CU_TRY(cudaMemsetAsync(acc.pair_first, 0, 8ull * c->pair_cap, c->stream));
        CU_TRY(cudaMemsetAsync(acc.pair_second, 0, 8ull * c->pair_cap, c->stream));
        CU_TRY(cudaMemsetAsync(acc.pair_payload, 0xff, 8ull * c->pair_cap, c->stream));
        CU_TRY(cudaMemsetAsync(acc.sample_count, 0, 8, c->stream));
    }
    CU_TRY(cudaStreamSynchronize(c->stream));
    return RSQ_OK;
}

int rsq_submit_device(rsq_ctx* ctx, const rsq_read* d_reads, size_t n_reads, const rsq_block* d_blocks, size_t n_blocks) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !c->have_ann) return RSQ_ERR_INVALID;
    if (!n_reads) return RSQ_OK;
    CU_TRY(cudaSetDevice(c->cfg.device));
    return run_batch(c, d_reads, n_reads, d_blocks, n_blocks);
}

int rsq_submit(rsq_ctx* ctx, const rsq_read* reads, size_t n_reads, const rsq_block* blocks, size_t n_blocks) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !c->have_ann || (n_reads && !reads)) return RSQ_ERR_INVALID;
    if (!n_reads) return RSQ_OK;
    CU_TRY(cudaSetDevice(c->cfg.device));
    const unsigned s = c->slot;
    c->slot ^= 1u;
    // the slot's previous kernels must have drained before its buffers are overwritten (or re-allocated)
    if (c->slot_used[s]) {
        if (n_reads > c->cap_reads[s] || n_blocks > c->cap_blocks[s]) CU_TRY(cudaEventSynchronize(c->ev_done[s]));
        else CU_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_done[s], 0));
    }
    if (n_reads > c->cap_reads[s]) {
        cudaFree(c->d_reads[s]);
        c->cap_reads[s] = n_reads + n_reads / 8;
        CU_TRY(cudaMalloc(&c->d_reads[s], c->cap_reads[s] * sizeof(rsq_read)));
    }
    if (std::max<size_t>(n_blocks, 1) > c->cap_blocks[s]) {
        cudaFree(c->d_blocks[s]);
        c->cap_blocks[s] = std::max<size_t>(n_blocks, 1) + n_blocks / 8;
        CU_TRY(cudaMalloc(&c->d_blocks[s], c->cap_blocks[s] * sizeof(rsq_block)));
    }
    CU_TRY(cudaMemcpyAsync(c->d_reads[s], reads, n_reads * sizeof(rsq_read), cudaMemcpyHostToDevice, c->copy_stream));
    if (n_blocks) CU_TRY(cudaMemcpyAsync(c->d_blocks[s], blocks, n_blocks * sizeof(rsq_block), cudaMemcpyHostToDevice, c->copy_stream));
    CU_TRY(cudaEventRecord(c->ev_copied[s], c->copy_stream));
    CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_copied[s], 0));
    int rc = run_batch(c, c->d_reads[s], n_reads, c->d_blocks[s], n_blocks);
    if (rc != RSQ_OK) return rc;
    CU_TRY(cudaEventRecord(c->ev_done[s], c->stream));
    c->slot_used[s] = true;
    return RSQ_OK;
}

int rsq_submit_packed(rsq_ctx* ctx, const rsq_read_packed* reads, size_t n_reads, const rsq_block* blocks, size_t n_blocks) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !c->have_ann || (n_reads && !reads) || (n_blocks && !blocks)) return RSQ_ERR_INVALID;
    if (!n_reads) return RSQ_OK;
    if (n_reads + n_blocks >= (1ull << 32)) { c->err = "packed batch too large (block offsets are 32-bit)"; return RSQ_ERR_INVALID; }
    CU_TRY(cudaSetDevice(c->cfg.device));
    const unsigned s = c->slot;
    c->slot ^= 1u;
    const size_t total_blocks = n_reads + n_blocks;     // upper bound: every implicit block is materialised on the device
    if (c->slot_used[s]) {
        if (n_reads > c->cap_packed[s] || n_blocks > c->cap_xblocks[s] || n_reads > c->cap_reads[s] || total_blocks > c->cap_blocks[s] || n_reads > c->cap_up)
            CU_TRY(cudaEventSynchronize(c->ev_done[s]));
        else CU_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_done[s], 0));
    }
    if (n_reads > c->cap_packed[s]) {
        cudaFree(c->d_packed[s]);
        c->cap_packed[s] = n_reads + n_reads / 8;
        CU_TRY(cudaMalloc(&c->d_packed[s], c->cap_packed[s] * sizeof(rsq_read_packed)));
    }
    if (std::max<size_t>(n_blocks, 1) > c->cap_xblocks[s]) {
        cudaFree(c->d_xblocks[s]);
        c->cap_xblocks[s] = std::max<size_t>(n_blocks, 1) + n_blocks / 8;
        CU_TRY(cudaMalloc(&c->d_xblocks[s], c->cap_xblocks[s] * sizeof(rsq_block)));
    }
    if (n_reads > c->cap_reads[s]) {
        cudaFree(c->d_reads[s]);
        c->cap_reads[s] = n_reads + n_reads / 8;
        CU_TRY(cudaMalloc(&c->d_reads[s], c->cap_reads[s] * sizeof(rsq_read)));
    }
    if (total_blocks > c->cap_blocks[s]) {
        cudaFree(c->d_blocks[s]);
        c->cap_blocks[s] = total_blocks + total_blocks / 8;
        CU_TRY(cudaMalloc(&c->d_blocks[s], c->cap_blocks[s] * sizeof(rsq_block)));
    }
    if (n_reads > c->cap_up) {
        CU_TRY(cudaStreamSynchronize(c->stream));          // the shared scan temporaries may still be in use by the other slot
        cudaFree(c->d_up_tot); cudaFree(c->d_up_expl); cudaFree(c->d_up_scan);
        c->cap_up = n_reads + n_reads / 8;
        CU_TRY(cudaMalloc(&c->d_up_tot, 4 * c->cap_up));
        CU_TRY(cudaMalloc(&c->d_up_expl, 4 * c->cap_up));
        CU_TRY(cudaMalloc(&c->d_up_scan, 4 * coverage_scan_scratch_elems(c->cap_up)));
    }
    CU_TRY(cudaMemcpyAsync(c->d_packed[s], reads, n_reads * sizeof(rsq_read_packed), cudaMemcpyHostToDevice, c->copy_stream));
    if (n_blocks) CU_TRY(cudaMemcpyAsync(c->d_xblocks[s], blocks, n_blocks * sizeof(rsq_block), cudaMemcpyHostToDevice, c->copy_stream));
    CU_TRY(cudaEventRecord(c->ev_copied[s], c->copy_stream));
    CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_copied[s], 0));
    // expand on the device: per-record block counts -> two inclusive scans -> records + blocks in the hot kernel's layout
    c->launches += launch_unpack_counts_of(c->d_packed[s], n_reads, c->d_up_tot, c->d_up_expl, c->stream);
    c->launches += launch_coverage_scan(c->d_up_tot, c->d_up_tot, n_reads, c->d_up_scan, c->stream);
    c->launches += launch_coverage_scan(c->d_up_expl, c->d_up_expl, n_reads, c->d_up_scan, c->stream);
    c->launches += launch_unpack_records(c->d_packed[s], n_reads, c->d_xblocks[s], n_blocks, c->d_up_tot, c->d_up_expl, c->d_reads[s], c->d_blocks[s],
                                         c->acc.error_flag, c->stream);
    int rc = run_batch(c, c->d_reads[s], n_reads, c->d_blocks[s], total_blocks);
    if (rc != RSQ_OK) return rc;
    CU_TRY(cudaEventRecord(c->ev_done[s], c->stream));
    c->slot_used[s] = true;
    return RSQ_OK;
}

int rsq_submit_compact(rsq_ctx* ctx, const void* batch, size_t bytes) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !c->have_ann || !batch || bytes < sizeof(rsq_batch16)) return RSQ_ERR_INVALID;
    rsq_batch16 H;
    memcpy(&H, batch, sizeof H);
    // the header is input: every section must lie inside the buffer before anything is launched on it
    auto inside = [&](uint64_t off, uint64_t count, uint64_t size) { return (off & 15u) == 0 && off <= bytes && count * size <= bytes - off; };
    if (H.magic != RSQ_BATCH16_MAGIC || H.struct_size != sizeof(rsq_batch16) || H.total_bytes > bytes || H.n_tiles != (H.n_reads + 31u) / 32u ||
        !inside(H.off_tiles, H.n_tiles, sizeof(rsq_tile16)) || !inside(H.off_reads, H.n_reads, sizeof(rsq_read16)) ||
        !inside(H.off_combos, H.n_combos, sizeof(rsq_combo16)) || !inside(H.off_blocks, H.n_blocks, 4) ||
        !inside(H.off_exc, H.n_exc, sizeof(rsq_read)) || !inside(H.off_exc_blocks, H.n_exc_blocks, sizeof(rsq_block))) {
        c->err = "compact batch: bad header";
        return RSQ_ERR_INVALID;
    }
    if (!H.n_reads) return RSQ_OK;
    const size_t n_reads = H.n_reads, total_blocks = std::max<size_t>(H.total_blocks, 1), wire = (size_t)H.total_bytes;
    CU_TRY(cudaSetDevice(c->cfg.device));
    const unsigned s = c->slot;
    c->slot ^= 1u;
    if (c->slot_used[s]) {
        if (wire > c->cap_compact[s] || n_reads > c->cap_reads[s] || total_blocks > c->cap_blocks[s]) CU_TRY(cudaEventSynchronize(c->ev_done[s]));
        else CU_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_done[s], 0));
    }
    if (wire > c->cap_compact[s]) {
        cudaFree(c->d_compact[s]);
        c->cap_compact[s] = wire + wire / 8;
        CU_TRY(cudaMalloc(&c->d_compact[s], c->cap_compact[s]));
    }
    if (n_reads > c->cap_reads[s]) {
        cudaFree(c->d_reads[s]);
        c->cap_reads[s] = n_reads + n_reads / 8;
        CU_TRY(cudaMalloc(&c->d_reads[s], c->cap_reads[s] * sizeof(rsq_read)));
    }
    if (total_blocks > c->cap_blocks[s]) {
        cudaFree(c->d_blocks[s]);
        c->cap_blocks[s] = total_blocks + total_blocks / 8;
        CU_TRY(cudaMalloc(&c->d_blocks[s], c->cap_blocks[s] * sizeof(rsq_block)));
    }
    CU_TRY(cudaMemcpyAsync(c->d_compact[s], batch, wire, cudaMemcpyHostToDevice, c->copy_stream));     // ONE copy per batch
    CU_TRY(cudaEventRecord(c->ev_copied[s], c->copy_stream));
    CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_copied[s], 0));
    c->launches += launch_expand_compact(c->d_compact[s], H, c->d_reads[s], c->d_blocks[s], c->acc.error_flag, c->stream);
    int rc = run_batch(c, c->d_reads[s], n_reads, c->d_blocks[s], H.total_blocks);
    if (rc != RSQ_OK) return rc;
    CU_TRY(cudaEventRecord(c->ev_done[s], c->stream));
    c->slot_used[s] = true;
    return RSQ_OK;
}

int rsq_sync(rsq_ctx* ctx) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RSQ_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->cfg.device));
    CU_TRY(cudaStreamSynchronize(c->copy_stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return RSQ_OK;
}

int rsq_sync_inputs(rsq_ctx* ctx) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RSQ_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->cfg.device));
    CU_TRY(cudaStreamSynchronize(c->copy_stream));
    return RSQ_OK;
}

int rsq_reset(rsq_ctx* ctx) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !c->have_ann) return RSQ_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->cfg.device));
    return zero_accumulators(c, true);
}

int rsq_finalize(rsq_ctx* ctx, rsq_results* out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !out || out->struct_size != sizeof(rsq_results) || !c->have_ann) return RSQ_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->cfg.device));
    CU_TRY(cudaStreamSynchronize(c->copy_stream));
    cudaStream_t st = c->stream;
    const Accum& a = c->acc;
    unsigned int err_flag = 0;
    CU_TRY(cudaMemcpyAsync(&err_flag, a.error_flag, 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(out->counters, a.counters, 8 * RSQ_N_COUNTERS, cudaMemcpyDeviceToHost, st));
    if (out->gene_reads) CU_TRY(cudaMemcpyAsync(out->gene_reads, a.gene_reads, 8ull * c->n_genes, cudaMemcpyDeviceToHost, st));
    if (out->gene_unique) CU_TRY(cudaMemcpyAsync(out->gene_unique, a.gene_unique, 8ull * c->n_genes, cudaMemcpyDeviceToHost, st));
    if (out->gene_fragments) CU_TRY(cudaMemcpyAsync(out->gene_fragments, a.gene_fragments, 8ull * c->n_genes, cudaMemcpyDeviceToHost, st));
    if (out->exon_whole || out->exon_frac) {
        c->launches += launch_combine_exon(a, c->n_exons, c->d_exon_out, c->d_exon_out + c->n_exons, st);
        if (out->exon_whole) CU_TRY(cudaMemcpyAsync(out->exon_whole, c->d_exon_out, 8ull * c->n_exons, cudaMemcpyDeviceToHost, st));
        if (out->exon_frac) CU_TRY(cudaMemcpyAsync(out->exon_frac, c->d_exon_out + c->n_exons, 8ull * c->n_exons, cudaMemcpyDeviceToHost, st));
    }
    if (out->coverage && c->cov_len) {             // raw per-base coverage on request only (rsq_coverage_statistics
        if (c->cfg.collect_coverage) {             // computes the per-gene statistics without it ever leaving the GPU)
            c->launches += launch_coverage_scan(a.cov_diff, c->d_cov, c->cov_len + 1, c->d_scan, st);
            CU_TRY(cudaMemcpyAsync(out->coverage, c->d_cov, 4ull * c->cov_len, cudaMemcpyDeviceToHost, st));
        } else memset(out->coverage, 0, 4ull * c->cov_len);
    }
    out->fragment_samples_taken = 0;
    if (out->gc_hist) {
        memset(out->gc_hist, 0, 8 * 100);
        if (c->have_ref) CU_TRY(cudaMemcpyAsync(out->gc_hist, c->acc.gc_hist, 8 * 100, cudaMemcpyDeviceToHost, st));
    }
    if (out->fragment_hist && out->fragment_hist_len) {
        memset(out->fragment_hist, 0, 8 * out->fragment_hist_len);
        if (c->have_bed && c->cfg.fragment_samples) {
            unsigned long long found = 0;
            CU_TRY(cudaMemcpyAsync(&found, a.sample_count, 8, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
            const unsigned long long take = std::min<unsigned long long>(found, c->cfg.fragment_samples);
            if (c->hist_len < out->fragment_hist_len) {
                cudaFree(c->d_hist);
                c->hist_len = out->fragment_hist_len;
                CU_TRY(cudaMalloc(&c->d_hist, 8ull * c->hist_len));
            }
            CU_TRY(cudaMemsetAsync(c->d_hist, 0, 8ull * out->fragment_hist_len, st));
            c->launches += launch_sample_histogram(a.sample_vals, take, c->d_hist, out->fragment_hist_len, st);
            CU_TRY(cudaMemcpyAsync(out->fragment_hist, c->d_hist, 8ull * out->fragment_hist_len, cudaMemcpyDeviceToHost, st));
            out->fragment_samples_taken = take;
        }
    }
    CU_TRY(cudaStreamSynchronize(st));
    CU_TRY(cudaGetLastError());
    if (err_flag & 8u) { c->err = "packed / compact batch: the records claim more blocks, combinations or exceptions than the batch carries"; return RSQ_ERR_INVALID; }
    if (err_flag) {
        c->err = std::string("device table overflow:") + ((err_flag & 1) ? " fragment-tracker set" : "") +
                 ((err_flag & 2) ? " mate table" : "") + ((err_flag & 4) ? " candidate list" : "") +
                 ((err_flag & 16) ? " more than 65536 records beyond the second of their read name in one launch" : "") + " (raise rsq_config.max_records)";
        return RSQ_ERR_CAPACITY;
    }
    return RSQ_OK;
}

int rsq_coverage_statistics(rsq_ctx* ctx, const rsq_transcripts* q, rsq_coverage_stats* out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !q || !out || !c->have_ann || q->struct_size != sizeof(rsq_transcripts) || out->struct_size != sizeof(rsq_coverage_stats))
        return RSQ_ERR_INVALID;
    if (q->bias_offset < 0 || q->bias_window < 1) { c->err = "bias offset/window out of range (offset >= 0, window >= 1)"; return RSQ_ERR_INVALID; }
    CU_TRY(cudaSetDevice(c->cfg.device));
    CU_TRY(cudaStreamSynchronize(c->copy_stream));
    cudaStream_t st = c->stream;
    // request arrays -> device (a few MB); the buffers persist in the ctx and only grow
    const size_t ng = q->n_genes, nr = q->n_exon_refs, ne = c->n_exons;
    if (std::max<size_t>(ng, 1) != c->cs_cap_genes || nr > c->cs_cap_refs) {   // exact per-gene stride: the all-reduce below spans the block
        CU_TRY(cudaStreamSynchronize(st));
        cudaFree(c->cs_off); cudaFree(c->cs_eoff); cudaFree(c->cs_strand); cudaFree(c->cs_len); cudaFree(c->cs_slot); cudaFree(c->cs_out); cudaFree(c->cs_owned);
        c->cs_cap_genes = std::max<size_t>(ng, 1); c->cs_cap_refs = std::max(c->cs_cap_refs, std::max<size_t>(nr, 1));
        CU_TRY(cudaMalloc(&c->cs_off, 8 * c->cs_cap_genes));
        CU_TRY(cudaMalloc(&c->cs_eoff, 4 * (c->cs_cap_genes + 1)));
        CU_TRY(cudaMalloc(&c->cs_strand, c->cs_cap_genes));
        CU_TRY(cudaMalloc(&c->cs_owned, c->cs_cap_genes));
        CU_TRY(cudaMalloc(&c->cs_len, 4 * c->cs_cap_refs));
        CU_TRY(cudaMalloc(&c->cs_slot, 4 * c->cs_cap_refs));
        CU_TRY(cudaMalloc(&c->cs_out, 8 * (5 * c->cs_cap_genes + std::max<size_t>(ne, 1))));   // mean | std | cv | exon_cv | five | three
    }
    unsigned long long* d_off = c->cs_off; uint32_t *d_eoff = c->cs_eoff, *d_len = c->cs_len, *d_slot = c->cs_slot; uint8_t* d_strand = c->cs_strand;
    double* d_out = c->cs_out;
    double* d_mean = d_out; double* d_std = d_out + c->cs_cap_genes; double* d_cv = d_std + c->cs_cap_genes;
    double* d_ecv = d_cv + c->cs_cap_genes;
    unsigned long long* d_ends = reinterpret_cast<unsigned long long*>(d_ecv + std::max<size_t>(ne, 1));
    c->cs_ends = d_ends;
    CU_TRY(cudaMemcpyAsync(d_off, q->gene_cov_off, 8 * ng, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_eoff, q->gene_exon_off, 4 * (ng + 1), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_strand, q->gene_strand, ng, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_len, q->exon_len, 4 * nr, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_slot, q->exon_slot, 4 * nr, cudaMemcpyHostToDevice, st));
    const uint8_t* d_owned = nullptr;
    if (q->gene_owned) { CU_TRY(cudaMemcpyAsync(c->cs_owned, q->gene_owned, ng, cudaMemcpyHostToDevice, st)); d_owned = c->cs_owned; }
    if (c->cov_len && c->cfg.collect_coverage)
        c->launches += launch_coverage_scan(c->acc.cov_diff, c->d_cov, c->cov_len + 1, c->d_scan, st);
    else
        CU_TRY(cudaMemsetAsync(c->d_cov, 0, 4ull * (c->cov_len + 1), st));
    unsigned big_ctas = 0, big_cap = 0;
    if (coverage_stats_needs_scratch(q->bias_window) && ng) {          // windows beyond shared memory: scratch in global memory
        unsigned long long longest = 0;
        for (size_t g = 0; g < ng; ++g) {
            unsigned long long T = 0;
            for (uint32_t e = q->gene_exon_off[g]; e < q->gene_exon_off[g + 1]; ++e) T += q->exon_len[e];
            longest = std::max(longest, T);
        }
        big_cap = big_window_cap(q->bias_window, longest);
        big_ctas = (unsigned)std::min<size_t>(ng, 296);
        const size_t need = (size_t)big_ctas * 2 * big_cap;
        if (need > c->cs_big_elems) {
            CU_TRY(cudaStreamSynchronize(st));
            cudaFree(c->cs_big);
            c->cs_big = nullptr; c->cs_big_elems = 0;
            CU_TRY(cudaMalloc(&c->cs_big, 8 * need));
            c->cs_big_elems = need;
        }
    }
    c->launches += launch_coverage_stats(c->d_cov, (uint32_t)ng, d_off, d_eoff, d_strand, d_len, d_slot, q->coverage_mask, q->bias_offset,
                                         q->bias_window, q->bias_gene_length, d_mean, d_std, d_cv, d_ends, d_ends + c->cs_cap_genes,
                                         d_ecv, (uint32_t)ne, d_owned, !d_owned || c->part_rank == 0, c->cs_big, big_ctas, big_cap, st);
    if (d_owned && c->comm) {   // owner-gather as a bitwise sum: every value has exactly one non-zero contributor
        if (ng < c->cs_cap_genes) {            // the unused tails of the five per-gene arrays travel too: keep them defined
            for (int k = 0; k < 3; ++k) CU_TRY(cudaMemsetAsync(d_out + k * c->cs_cap_genes + ng, 0, 8 * (c->cs_cap_genes - ng), st));
            for (int k = 0; k < 2; ++k) CU_TRY(cudaMemsetAsync(d_ends + k * c->cs_cap_genes + ng, 0, 8 * (c->cs_cap_genes - ng), st));
        }
        NcclApi& N = nccl_api();
        NCCL_TRY(N.AllReduce(d_out, d_out, 5 * c->cs_cap_genes + std::max<size_t>(ne, 1), ncclUint64, ncclSum, c->comm, st));
    }
    CU_TRY(cudaMemcpyAsync(out->gene_mean, d_mean, 8 * ng, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(out->gene_std, d_std, 8 * ng, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(out->gene_cv, d_cv, 8 * ng, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(out->five_end, d_ends, 8 * ng, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(out->three_end, d_ends + c->cs_cap_genes, 8 * ng, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(out->exon_cv, d_ecv, 8 * ne, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    CU_TRY(cudaGetLastError());
    return RSQ_OK;
}

int rsq_set_partition(rsq_ctx* ctx, uint32_t rank, uint32_t world) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RSQ_ERR_INVALID;
    if (!world || rank >= world || world > 4096) { c->err = "bad partition rank / world"; return RSQ_ERR_INVALID; }
    if (c->comm && (rank != c->part_rank || world != c->part_world)) { c->err = "the partition is fixed by rsq_comm_init"; return RSQ_ERR_INVALID; }
    CU_TRY(cudaSetDevice(c->cfg.device));
    if (world != c->part_world || !c->d_found) {
        CU_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_found);
        c->d_found = nullptr;
        CU_TRY(cudaMalloc(&c->d_found, 8ull * world));
    }
    c->part_rank = rank; c->part_world = world;
    c->found_valid = false;
    return RSQ_OK;
}

static int ensure_export_buffer(Ctx* c) {
    const size_t words = export_count_words(c->n_genes, c->n_exons, sample_quota(c));
    if (words != c->export_words || !c->d_export) {
        CU_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_export);
        c->d_export = nullptr;
        CU_TRY(cudaMalloc(&c->d_export, 8ull * words));
        c->export_words = words;
    }
    return RSQ_OK;
}

int rsq_export_sample_counts(rsq_ctx* ctx, uint64_t** d_words, size_t* n_words) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !d_words || !n_words || !c->have_ann) return RSQ_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->cfg.device));
    if (!c->d_found) { int rc = rsq_set_partition(ctx, c->part_rank, c->part_world); if (rc != RSQ_OK) return rc; }
    Accum a = c->acc;
    if (!sample_quota(c)) a.sample_count = nullptr;
    c->launches += launch_pack_sample_counts(a, c->part_rank, c->part_world, c->d_found, c->stream);
    CU_TRY(cudaStreamSynchronize(c->stream));
    c->found_valid = true;             // ... once the caller has reduced it
    *d_words = reinterpret_cast<uint64_t*>(c->d_found);
    *n_words = c->part_world;
    return RSQ_OK;
}

int rsq_export_counts(rsq_ctx* ctx, uint64_t** d_words, size_t* n_words) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !d_words || !n_words || !c->have_ann) return RSQ_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->cfg.device));
    int rc = ensure_export_buffer(c);
    if (rc != RSQ_OK) return rc;
    const unsigned long long quota = sample_quota(c);
    if (quota && c->part_world > 1 && !c->found_valid) { c->err = "rsq_export_sample_counts (and its all-reduce) must precede rsq_export_counts"; return RSQ_ERR_INVALID; }
    c->launches += launch_pack_counts(c->acc, c->n_genes, c->n_exons, quota, c->part_world > 1 ? c->d_found : nullptr, c->part_rank, c->part_world,
                                      c->d_export, c->stream);
    CU_TRY(cudaStreamSynchronize(c->stream));
    *d_words = reinterpret_cast<uint64_t*>(c->d_export);
    *n_words = c->export_words;
    return RSQ_OK;
}

int rsq_import_counts(rsq_ctx* ctx) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !c->have_ann || !c->d_export) return RSQ_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->cfg.device));
    c->launches += launch_unpack_counts(c->acc, c->n_genes, c->n_exons, sample_quota(c), c->d_export, c->stream);
    CU_TRY(cudaStreamSynchronize(c->stream));
    return RSQ_OK;
}

int rsq_comm_unique_id(uint8_t id[RSQ_COMM_ID_BYTES]) {
    NcclApi& N = nccl_api();
    if (!id) return RSQ_ERR_INVALID;
    if (!N.handle) { g_create_error = N.why; return RSQ_ERR_NO_DEVICE; }
    ncclUniqueId u;
    const ncclResult_t r = N.GetUniqueId(&u);
    if (r != ncclSuccess) { g_create_error = std::string("ncclGetUniqueId: ") + N.GetErrorString(r); return RSQ_ERR_CUDA; }
    memcpy(id, &u, RSQ_COMM_ID_BYTES);
    return RSQ_OK;
}

int rsq_comm_init(rsq_ctx* ctx, const uint8_t id[RSQ_COMM_ID_BYTES], uint32_t rank, uint32_t world) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !id) return RSQ_ERR_INVALID;
    if (c->comm) { c->err = "communicator already attached"; return RSQ_ERR_INVALID; }
    NcclApi& N = nccl_api();
    if (!N.handle) { c->err = N.why; return RSQ_ERR_NO_DEVICE; }
    int rc = rsq_set_partition(ctx, rank, world);
    if (rc != RSQ_OK) return rc;
    CU_TRY(cudaSetDevice(c->cfg.device));
    ncclUniqueId u;
    memcpy(&u, id, RSQ_COMM_ID_BYTES);
    NCCL_TRY(N.CommInitRank(&c->comm, (int)world, u, (int)rank));
    return RSQ_OK;
}

// The whole end-of-run exchange on the ctx stream: sample counts -> all-reduce -> count vector (with the kept samples at their
// global ranks) -> all-reduce -> fold back.  No host synchronisation: rsq_finalize's is the first one.
int rsq_exchange(rsq_ctx* ctx) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !c->have_ann) return RSQ_ERR_INVALID;
    if (!c->comm) { c->err = "rsq_exchange needs rsq_comm_init"; return RSQ_ERR_INVALID; }
    CU_TRY(cudaSetDevice(c->cfg.device));
    CU_TRY(cudaStreamSynchronize(c->copy_stream));      // every submit's copy has been consumed by kernels queued on c->stream
    NcclApi& N = nccl_api();
    int rc = ensure_export_buffer(c);
    if (rc != RSQ_OK) return rc;
    const unsigned long long quota = sample_quota(c);
    if (quota) {
        c->launches += launch_pack_sample_counts(c->acc, c->part_rank, c->part_world, c->d_found, c->stream);
        NCCL_TRY(N.AllReduce(c->d_found, c->d_found, c->part_world, ncclUint64, ncclSum, c->comm, c->stream));
        c->found_valid = true;
    }
    c->launches += launch_pack_counts(c->acc, c->n_genes, c->n_exons, quota, c->d_found, c->part_rank, c->part_world, c->d_export, c->stream);
    NCCL_TRY(N.AllReduce(c->d_export, c->d_export, c->export_words, ncclUint64, ncclSum, c->comm, c->stream));
    c->launches += launch_unpack_counts(c->acc, c->n_genes, c->n_exons, quota, c->d_export, c->stream);
    CU_TRY(cudaGetLastError());
    return RSQ_OK;
}

int rsq_algorithmic_bytes(rsq_ctx* ctx, uint64_t* bytes) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !bytes || !c->have_ann) return RSQ_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->cfg.device));
    CU_TRY(cudaStreamSynchronize(c->stream));
    CU_TRY(cudaMemcpy(bytes, c->acc.algo_bytes, 8, cudaMemcpyDeviceToHost));
    return RSQ_OK;
}

int rsq_launch_count(rsq_ctx* ctx, uint64_t* n) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !n) return RSQ_ERR_INVALID;
    *n = c->launches;
    return RSQ_OK;
}

int rsq_kernel_time_ms(rsq_ctx* ctx, double* ms) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !ms) return RSQ_ERR_INVALID;
    CU_TRY(cudaSetDevice(c->cfg.device));
    CU_TRY(cudaStreamSynchronize(c->stream));
    double total = c->kernel_ms;
    for (size_t k = 0; k < c->timers_used; ++k) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, c->timers[k].first, c->timers[k].second) == cudaSuccess) total += t;
    }
    *ms = total;
    return RSQ_OK;
}

}  // extern "C"
