// emu_covstats.cpp — TEST INFRASTRUCTURE: rsq_covstats.cu compiled for the host through cuda_emu.h (see emu_kernels.cpp).
#define RSQ_CUDA_EMU 1
#include "cuda_emu.h"

#include <vector>

#include "../../rnaseqc_b200/csrc/rsq_covstats.cu"

using namespace rsqd;

extern "C" {

// rsq_covstats.cu: per-gene coverage statistics over an integrated coverage array
int emu_coverage_stats(const unsigned int* cov, uint32_t n_genes, const unsigned long long* gene_cov_off, const uint32_t* gene_exon_off,
                       const uint8_t* gene_strand, const uint32_t* exon_len, const uint32_t* exon_slot, uint32_t mask, int bias_offset,
                       int bias_window, uint32_t bias_gene_length, double* gene_mean, double* gene_std, double* gene_cv,
                       unsigned long long* five_end, unsigned long long* three_end, double* exon_cv, uint32_t n_exons,
                       const uint8_t* gene_owned, int nan_prefill) {
    std::vector<double> scratch;
    unsigned ctas = 0, cap = 0;
    if (coverage_stats_needs_scratch(bias_window)) {
        unsigned long long longest = 0;
        for (uint32_t g = 0; g < n_genes; ++g) {
            unsigned long long T = 0;
            for (uint32_t e = gene_exon_off[g]; e < gene_exon_off[g + 1]; ++e) T += exon_len[e];
            longest = std::max(longest, T);
        }
        ctas = std::min<unsigned>(n_genes, 5u); cap = big_window_cap(bias_window, longest);
        scratch.resize((size_t)ctas * 2 * cap);
    }
    return launch_coverage_stats(cov, n_genes, gene_cov_off, gene_exon_off, gene_strand, exon_len, exon_slot, mask, bias_offset, bias_window,
                                 bias_gene_length, gene_mean, gene_std, gene_cv, five_end, three_end, exon_cv, n_exons, gene_owned, nan_prefill != 0,
                                 scratch.empty() ? nullptr : scratch.data(), ctas, cap, nullptr);
}

}  // extern "C"
