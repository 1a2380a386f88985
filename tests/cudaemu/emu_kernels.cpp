// emu_kernels.cpp — TEST INFRASTRUCTURE (oracle/_build/libcudaemu.so): the CUDA sources of the product compiled for the host through
// cuda_emu.h and exported as plain C entry points for tests/test_cudaemu.py.  Nothing here is linked into the product.
#define RSQ_CUDA_EMU 1
#include "cuda_emu.h"

#include "../../rnaseqc_b200/csrc/rsq_kernels.cu"

using namespace rsqd;

extern "C" {

// the three-phase scan of the coverage difference array
int emu_coverage_scan(const unsigned int* diff, unsigned int* cov, unsigned long long n) {
    std::vector<unsigned int> scratch(coverage_scan_scratch_elems(n) + 8);
    return launch_coverage_scan(diff, cov, n, scratch.data(), nullptr);
}

// One launch worth of BED mate-map stages: the hot kernel's part (claiming a table slot per read name, the round-1 bid) is done
// here on the host, then launch_pair_stages runs mate_first / mate_second / mate_rest and the sample ranking.
//   cand_*: the launch's candidates in ANY order (as the device emits them); tables: `cap` slots (a power of two), persistent
//   across calls; sample_vals[quota] / sample_count persistent.  launch_no numbers the launches from 1.
int emu_bed_launch(uint32_t n_reads, uint32_t n_cand, const unsigned long long* name, const uint32_t* local_idx, const uint32_t* name_id,
                   const int32_t* endpos, const uint32_t* abs_isize, const uint32_t* cond_ok, unsigned long long cap, unsigned long long* keys,
                   unsigned long long* first, unsigned long long* second, unsigned long long* payload, unsigned long long launch_no,
                   uint32_t quota, unsigned int* sample_vals, unsigned long long* sample_count, unsigned long long record_base,
                   unsigned int* error_flag, unsigned long long* mate_totals) {
    Accum acc{};
    Params prm{};
    prm.mate_round = launch_no << 17; prm.fragment_samples = quota; prm.record_base = record_base;
    std::vector<PairCandidate> cand(n_cand ? n_cand : 1);
    for (uint32_t k = 0; k < n_cand; ++k) {
        const unsigned long long key = name[k] ? name[k] : 1ull;
        unsigned long long slot = mix64(key) & (cap - 1);
        while (keys[slot] != 0ull && keys[slot] != key) slot = (slot + 1) & (cap - 1);
        keys[slot] = key;
        first[slot] = std::max(first[slot], mate_tag(prm.mate_round, 1u, local_idx[k]));       // atomicMax in the hot kernel
        PairCandidate c{};
        c.name_hash = key; c.local_idx = local_idx[k]; c.name_id = name_id[k]; c.endpos = endpos[k]; c.abs_isize = abs_isize[k];
        c.cond_ok = cond_ok[k]; c.slot = (uint32_t)slot;
        cand[k] = c;
    }
    unsigned int cand_count = n_cand;
    std::vector<unsigned char> flag(n_reads + 1, 0);
    std::vector<unsigned int> isz(n_reads + 1, 0), scratch(pair_scan_scratch_elems(n_reads) + 8), left(2 * kMateLeftCap);
    unsigned int left_count[2] = {0, 0};
    acc.pair_keys = keys; acc.pair_first = first; acc.pair_second = second; acc.pair_payload = payload; acc.pair_mask = cap - 1;
    acc.cand = cand.data(); acc.cand_count = &cand_count; acc.cand_capacity = n_cand ? n_cand : 1;
    acc.sample_flag = flag.data(); acc.sample_isize = isz.data(); acc.sample_vals = sample_vals; acc.sample_count = sample_count;
    acc.mate_left = left.data(); acc.mate_left_count = left_count; acc.mate_totals = mate_totals; acc.error_flag = error_flag;
    return launch_pair_stages(n_reads, acc, prm, scratch.data(), nullptr);
}

// 16-byte wire format -> rsq_read / rsq_block (expand_compact_kernel)
int emu_expand_compact(const unsigned char* buf, rsq_read* reads, rsq_block* blocks, unsigned int* error_flag) {
    rsq_batch16 H;
    memcpy(&H, buf, sizeof H);
    return launch_expand_compact(buf, H, reads, blocks, error_flag, nullptr);
}

}  // extern "C"
