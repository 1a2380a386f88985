// rsq_device.h — device-side views shared by the kernels (rsq_kernels.cu) and the C-ABI layer (rsq_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "rsq_b200.h"

// Kernel launches go through one macro so that the CPU test harness (tests/cudaemu: the kernels' own sources compiled for the host,
// one OS thread per CUDA thread, barriers and warp shuffles emulated) can run what cannot be tested without a GPU otherwise.
#if defined(RSQ_CUDA_EMU)
#define RSQ_LAUNCH(kernel, grid, block, smem, stream, ...) ::cuda_emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define RSQ_DYNAMIC_SMEM(type, name) type* name = reinterpret_cast<type*>(::cuda_emu::dynamic_smem())
#else
#define RSQ_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define RSQ_DYNAMIC_SMEM(type, name) extern __shared__ type name[]
#endif

namespace rsqd {

// Exon metadata packed for one 128-bit load.
struct __align__(16) ExonMeta {
    uint32_t gene_flags;   // gene index in the low 24 bits, RSQ_EXON_* flags in the high 8
    uint32_t slot;         // output slot (exonList order)
    uint64_t cov_off;      // first base of this exon in the coverage array
};

// ---- elementary-interval index (built by rsq_eit.h, read by rsq_hot_core.cuh) ------------------------------------------
constexpr int kEiShift = 10;                 // direct-address table: one entry per 1 kb of contig
constexpr int kEiPad = 512;                  // INT32_MAX entries behind the last interval: bulk copies may run past the end
// class bits of an elementary interval: (kind << 0) for the forward strand, << 1 reverse, << 2 unknown
enum : uint32_t { EI_GENE = 0x001, EI_GENE_RIBO = 0x008, EI_EXON = 0x040, EI_EXON_RIBO = 0x200,
                  EI_ANY_GENE = 0x007, EI_ANY_GENE_RIBO = 0x038, EI_ANY_EXON = 0x1c0, EI_ANY_EXON_RIBO = 0xe00 };
// CoverEntry::gene_flags, high byte: RSQ_EXON_STRAND_MASK | RSQ_EXON_RIBOSOMAL as in rsq_annotation.exon_flags, plus
constexpr uint32_t kCoverGlobin = 0x8;       // geneNames[gene] is a blacklisted globin (src/Expression.cpp:24,396-399)
// One exon covering an elementary interval, with everything the accumulation needs inlined.
struct __align__(16) CoverEntry {
    int32_t  start, end;       // the exon (closed)
    uint32_t gene_flags;       // gene index in the low 24 bits, flags in the high 8
    uint32_t slot;             // output slot (exonList order)
    uint64_t cov_off;          // first base of this exon in the coverage array
    int32_t  late_after;       // reads that start (0-based) beyond this coordinate find the gene already evicted from the reference's
                               // sliding window (rsq_eit.h, gene_eviction_bounds); INT32_MAX = never
    uint32_t reserved;
};
static_assert(sizeof(CoverEntry) == 32, "two 128-bit loads");

// Annotation resident in HBM (replicated per GPU; a few tens of MB at GENCODE scale, i.e. L2-resident).
struct AnnView {
    uint32_t n_contigs, n_genes, n_exons, n_segs;
    const uint32_t* exon_off;     // [n_contigs+1]
    const int2*     exon_se;      // [n_exons] (start, end), start-sorted per contig
    const int32_t*  exon_pmax;    // [n_exons] running max of end within the contig (monotone)
    const ExonMeta* exon_meta;    // [n_exons]
    const uint8_t*  gene_flags;   // [n_genes]
    // gene lines as intervals in list order (--legacy rules only; null otherwise)
    const uint32_t* gline_off;    // [n_contigs+1]
    const int2*     gline_se;     // [n_gene_lines] (start, end), start-sorted per contig
    const int32_t*  gline_pmax;   // [n_gene_lines] running max of end within the contig
    const uint32_t* gline_gf;     // [n_gene_lines] gene index of the line's gene_id | (strand, ribosomal) << 24
    const uint32_t* gline_order;  // [n_gene_lines] rank among the kept GTF lines (ties of the list's stable sort)
    const uint32_t* exon_order;   // [n_exons] same for the exon lines
    // elementary-interval index (rsq_eit.h): what the non-legacy hot path reads
    const uint32_t*   ei_off;     // [n_contigs+1] first interval of each contig (an INT32_MIN sentinel)
    const int32_t*    ei_start;   // [n_ei + kEiPad] interval starts, ascending inside a contig
    const uint2*      ei_desc;    // [n_ei + 1] x = first covering exon in `cover`, y = class bits | count << 16
    const CoverEntry* cover;      // covering exons, interval-major, list order inside an interval
    const uint32_t*   eib_off;    // [n_contigs+1] first table entry of each contig (n_buckets + 1 entries per contig)
    const uint32_t*   eib;        // interval containing each bucket start; a contig's last entry = its last interval
    const unsigned long long* recip64;   // [65536] floor(2^64 / L): exact exon fractions with one 32-bit divide (frac_q64)
};

struct BedView {
    uint32_t n_intervals;
    const uint32_t* off;          // [n_contigs+1]
    const int2*     se;           // [n_intervals]
    const int32_t*  pmax;         // [n_intervals]
    const uint32_t* name_id;      // [n_intervals]
    const uint8_t*  present;      // [n_contigs]
    // elementary-interval index over the BED intervals (rsq_eit.h; CoverEntry::slot carries the interval's name id)
    const uint32_t*   ei_off;
    const int32_t*    ei_start;
    const uint2*      ei_desc;
    const CoverEntry* cover;
    const uint32_t*   eib_off;
    const uint32_t*   eib;
};

// One read that entered the fragment-size mate map (src/Expression.cpp:508-516), emitted by the hot kernel.
struct __align__(16) PairCandidate {
    unsigned long long name_hash;
    uint32_t local_idx;    // record index inside the batch
    uint32_t name_id;      // BED interval identity
    int32_t  endpos;
    uint32_t abs_isize;
    uint32_t cond_ok;      // !MateReverse && Reverse && pos != mpos (src/Expression.cpp:528, minus the endpos test)
    uint32_t slot;         // table slot, filled by the hot kernel
};

// One read that entered the fragment GC-content mate map (src/Expression.cpp:459-467), emitted by the hot kernel.
struct __align__(16) GcCandidate {
    unsigned long long name_hash;
    uint32_t local_idx;    // record index inside the batch
    uint32_t exon_slot;    // the single exon every block aligned to (alignedExons, :319,348)
    int32_t  endpos;
    int32_t  l_qseq;
    int32_t  contig;
    uint32_t cond_ok;      // Position() != MatePosition() (:471, minus the end test)
    uint32_t slot;         // table slot
    uint32_t pad;
};

// Reference genome as G/C bits (rsq_reference).
struct RefView {
    const uint8_t*  present;     // [n_contigs]
    const unsigned long long* length;
    const unsigned long long* word_off;
    const uint32_t* bits;
};

constexpr unsigned long long kMateEmpty = ~0ull;
constexpr unsigned int kMateLeftCap = 65536;     // candidates beyond the second of their name in one launch, per map
// tag of a candidate in round r of a launch: later rounds beat earlier ones, earlier records beat later ones (atomicMax)
__host__ __device__ inline unsigned long long mate_tag(unsigned long long mate_round, unsigned r, uint32_t local_idx) {
    return ((mate_round + r) << 24) | (unsigned long long)(0xffffffu - (local_idx & 0xffffffu));
}

// Mutable accumulators in HBM.
struct Accum {
    unsigned long long* counters;        // [RSQ_N_COUNTERS]
    unsigned long long* gene_reads;      // [n_genes]
    unsigned long long* gene_unique;
    unsigned long long* gene_fragments;
    // exon counts, exact Q64.64 kept as three carry-free sums (by slot): value = whole + (fhi * 2^32 + flo) * 2^-64
    unsigned long long* exon_whole;      // [n_exons]
    unsigned long long* exon_fhi;        // [n_exons] sum of (frac >> 32)
    unsigned long long* exon_flo;        // [n_exons] sum of (frac & 0xffffffff)
    unsigned int*       cov_diff;        // [cov_len+1] coverage difference array (wrapping u32)
    unsigned long long* frag_set;        // fragment-tracker hash set (open addressing), 0 = empty
    unsigned long long  frag_cap;        // slots (any number, not a power of two)
    unsigned long long* algo_bytes;      // [1]
    unsigned int*       error_flag;      // [2] ([1]: scratch word, see launch_hot_path) bit0: fragment set full, bit1: pair table full, bit2: candidate list full
    // fragment-size sampler (BED mode)
    unsigned long long* pair_keys;       // mate map keys (name hash, 0 = empty)
    // State per read name = what the reference's map holds for it (src/Expression.cpp:511-533): nothing (kMateEmpty) or the
    // (interval, PositionEnd) of the pending mate.  The candidates of one launch are applied to it in stream order: the earliest
    // candidate of a name wins pair_first (atomicMax of a tag that grows with every round and falls with the record index), the
    // earliest of the others pair_second, further ones (three or more records of one name in one launch: malformed input) queue up
    // in mate_left and are replayed round by round (rsq_kernels.cu, mate_* kernels).
    unsigned long long* pair_first;      // scratch: tag of the earliest candidate of the current round
    unsigned long long* pair_second;     // scratch: tag of the earliest candidate among the others
    unsigned long long* pair_payload;    // (name_id << 32) | (uint32)endpos of the pending mate, or kMateEmpty
    unsigned long long  pair_mask;
    unsigned int*       mate_left;       // [2 * kMateLeftCap] candidate indices still to be replayed (BED half, GC half)
    unsigned int*       mate_left_count; // [2]
    unsigned long long* mate_totals;     // [2] candidates since the last reset: what the host sizes the maps by
    PairCandidate*      cand;            // per-batch candidate list
    unsigned int*       cand_count;      // [1]
    unsigned int        cand_capacity;
    unsigned char*      sample_flag;     // [batch capacity] 1 = this record completes a sample
    unsigned int*       sample_isize;    // [batch capacity]
    unsigned int*       sample_vals;     // [fragment_samples] |isize| of the k-th sample in stream order
    unsigned long long* sample_count;    // [1] samples found so far (may exceed fragment_samples)
    // fragment GC-content sampler (--fasta): same mate-map scheme as the BED sampler, its own tables (gcContentFragmentTracker)
    unsigned long long* gc_keys;
    unsigned long long* gc_first;
    unsigned long long* gc_second;
    unsigned long long* gc_payload;      // (exon slot << 32) | (uint32)endpos of the pending mate, or kMateEmpty
    unsigned long long  gc_mask;
    GcCandidate*        gc_cand;
    unsigned int*       gc_cand_count;   // [1]
    unsigned int        gc_cand_capacity;
    unsigned long long* gc_hist;         // [100] gcBins
};

struct Params {
    uint32_t mapq_threshold, base_mismatch_threshold;
    int32_t  chimeric_distance;
    uint32_t strand_mode, unpaired, exclude_chimeric, n_filter_tags, collect_coverage;
    uint32_t bed_enabled;
    uint32_t fragment_samples;
    uint32_t legacy;                     // --legacy counting rules (rsq_legacy.cuh)
    uint32_t gc_enabled;                 // fragment GC-content sampler (rsq_set_reference)
    const uint8_t* ref_present;          // [n_contigs] Fasta::hasContig (gc_enabled only)
    const unsigned long long* cascade_lut;   // [2^17] the filter cascade as a table (cascade_entry, rsq_hot_core.cuh); null = compute
    unsigned long long record_base;      // global index of reads[0] in the stream (ordering of fragment samples)
    unsigned long long mate_round;       // (launch number << 17): base of the round tags of this launch's mate-map stages
    unsigned long long cov_len_diag;     // coverage array length (RSQ_AB_SCATTER diagnostic build only)
};

// launchers (rsq_kernels.cu); every launcher returns the number of kernels it launched
int launch_hot_path(const rsq_read* reads, const rsq_block* blocks, size_t n_reads, size_t n_blocks, const AnnView& ann,
                    const BedView& bed, const Accum& acc, const Params& prm, cudaStream_t stream);
// compact wire format -> rsq_read / rsq_block (rsq_submit_packed): per-record block counts, then (after two scans) the records
int launch_unpack_counts_of(const rsq_read_packed* packed, size_t n, unsigned int* tot, unsigned int* expl, cudaStream_t stream);
int launch_unpack_records(const rsq_read_packed* packed, size_t n, const rsq_block* xblocks, size_t n_xblocks, const unsigned int* incl_tot,
                          const unsigned int* incl_expl, rsq_read* reads, rsq_block* blocks, unsigned int* error_flag, cudaStream_t stream);
// 16-byte wire format (rsq_batch16 in one device buffer) -> rsq_read / rsq_block
int launch_expand_compact(const unsigned char* buf, const rsq_batch16& hdr, rsq_read* reads, rsq_block* blocks, unsigned int* error_flag, cudaStream_t stream);
int launch_gc_stages(size_t n_reads, const RefView& ref, const Accum& acc, const Params& prm, cudaStream_t stream);
int launch_prefix_max(const int2* se, const uint32_t* off, uint32_t n_contigs, int32_t* pmax, cudaStream_t stream);
int launch_pair_rehash(const unsigned long long* old_keys, const unsigned long long* old_payload, unsigned long long old_cap, unsigned long long* nk,
                       unsigned long long* np, unsigned long long new_mask, unsigned int* error_flag, cudaStream_t stream);
int launch_coverage_scan(const unsigned int* diff, unsigned int* cov, unsigned long long n, unsigned int* block_sums,
                         cudaStream_t stream);
size_t coverage_scan_scratch_elems(unsigned long long n);
int launch_rehash(const unsigned long long* old_tab, unsigned long long old_cap, unsigned long long* new_tab,
                  unsigned long long new_cap, unsigned int* error_flag, cudaStream_t stream);
// end-of-run exchange of a coordinate-partitioned run (include/rsq_b200.h): `quota` = fragment_samples when the sampler is on, else 0
size_t export_count_words(uint32_t n_genes, uint32_t n_exons, unsigned long long quota);
int launch_pack_sample_counts(const Accum& acc, uint32_t rank, uint32_t world, unsigned long long* out, cudaStream_t stream);
int launch_pack_counts(const Accum& acc, uint32_t n_genes, uint32_t n_exons, unsigned long long quota, const unsigned long long* found, uint32_t rank,
                       uint32_t world, unsigned long long* out, cudaStream_t stream);
int launch_unpack_counts(const Accum& acc, uint32_t n_genes, uint32_t n_exons, unsigned long long quota, const unsigned long long* in, cudaStream_t stream);
int launch_combine_exon(const Accum& acc, uint32_t n_exons, unsigned long long* whole_out, unsigned long long* frac_out, cudaStream_t stream);
// BED mode, after launch_hot_path of the same batch: owners publish, seconds test, samples are ranked in stream order
int launch_pair_stages(size_t n_reads, const Accum& acc, const Params& prm, unsigned int* scan_scratch, cudaStream_t stream);
size_t pair_scan_scratch_elems(size_t n_reads);
int launch_sample_histogram(const unsigned int* vals, unsigned long long n, unsigned long long* hist, unsigned long long hist_len,
                            cudaStream_t stream);

// rsq_covstats.cu
int launch_coverage_stats(const unsigned int* cov, uint32_t n_genes, const unsigned long long* gene_cov_off, const uint32_t* gene_exon_off,
                          const uint8_t* gene_strand, const uint32_t* exon_len, const uint32_t* exon_slot, uint32_t mask, int bias_offset,
                          int bias_window, uint32_t bias_gene_length, double* gene_mean, double* gene_std, double* gene_cv,
                          unsigned long long* five_end, unsigned long long* three_end, double* exon_cv, uint32_t n_exons,
                          const uint8_t* gene_owned, bool nan_prefill, double* big_scratch, unsigned big_ctas, unsigned big_win_cap,
                          cudaStream_t stream);
// bias windows too wide for shared memory: the caller provides big_ctas x 2 x big_win_cap doubles of scratch, big_win_cap = big_window_cap(...)
bool coverage_stats_needs_scratch(int bias_window);
inline unsigned big_window_cap(int bias_window, unsigned long long longest_transcript) {
    unsigned long long need = std::min<unsigned long long>((unsigned long long)std::max(bias_window, 1), std::max<unsigned long long>(longest_transcript, 1ull));
    unsigned p2 = 1;
    while (p2 < need) p2 <<= 1;
    return p2;
}

}  // namespace rsqd
