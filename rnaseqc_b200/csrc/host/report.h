// report.h — end-of-run statistics and the reference's output files (src/RNASeQC.cpp:397-676,
// src/Metrics.cpp:132-337,342-412), computed from the device results.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "annotation.h"
#include "fasta.h"
#include "rsq_b200.h"

namespace rsqh {

struct RunOptions {
    std::string sample_name;          // SAMPLENAME (src/RNASeQC.cpp:99)
    bool sample_given = false;        // -s present: column headers use it (:433-436,514)
    std::string output_dir;
    bool use_rpkm = false;            // --rpkm
    bool write_coverage = false;      // --coverage (coverage.tsv is /dev/null otherwise, src/Metrics.h:110)
    unsigned coverage_mask = 500;     // --coverage-mask (:91)
    int bias_offset = 0;              // BIAS_OFFSET (:94)
    int bias_window = 100;            // BIAS_WINDOW (:95)
    unsigned long bias_gene_length = 200;  // BIAS_LENGTH (:96)
    unsigned detection_threshold = 5; // DETECTION_THRESHOLD (:100)
    std::vector<std::string> filter_tags;
    bool exact_complexity_scan = false;  // replicate the <=1e9-iteration loop literally (:402-415)
};

// Order-dependent scalars the decoder tracked in stream order.
struct StreamState {
    int read_length = 0;
    uint64_t alignment_count = 0;
    std::vector<uint32_t> dropped_order;  // contigs whose feature lists were dropped, in order
    int64_t final_contig = -1;            // current_chrom at end of stream
    int32_t final_position = 0;           // pos of the last read that reached trimFeatures
};

struct HostResults {
    uint64_t counters[RSQ_N_COUNTERS];
    std::vector<uint64_t> gene_reads, gene_unique, gene_fragments;  // per universe gene
    std::vector<uint64_t> exon_whole, exon_frac;                     // per exon slot, Q64.64
    std::vector<double> exon_seq;         // optional: sequentially summed doubles (oracle only); empty otherwise
    std::vector<uint32_t> coverage;       // per base, laid out by Annotation::exon_cov_off
    std::vector<uint64_t> fragment_hist;  // index = |isize|
    uint64_t fragment_samples_taken = 0;
    uint64_t gc_hist[100] = {0};          // gcBins (src/RNASeQC.cpp:106,368)
    // per-gene coverage statistics computed on the device (rsq_coverage_statistics), indexed by GeneFeature::flat;
    // when present the per-base coverage array is not needed on the host at all
    bool have_cov_stats = false;
    std::vector<double> cs_mean, cs_std, cs_cv, cs_exon_cv;
    std::vector<uint64_t> cs_five, cs_three;
};

// exact Q64.64 -> nearest double
double exon_count_to_double(uint64_t whole, uint64_t frac);

// Per-gene coverage statistics in eviction order (BaseCoverage::compute, src/Metrics.cpp:132-151).
struct CoverageSummary {
    std::vector<std::string> rows;             // coverage.tsv rows (without header)
    std::vector<double> gene_means, gene_stds, gene_cvs;
    std::vector<std::pair<std::string, double>> exon_cvs;  // exonCoverage entries (unordered)
    std::vector<uint64_t> five_end, three_end; // BiasCounter maps, per universe gene
};
void compute_coverage_summary(const Annotation& A, const HostResults& R, const RunOptions& opt,
                              const StreamState& st, CoverageSummary& out);

// Writes every output file the reference writes.  `ref` (may be null) is the --fasta reference: adds the GC column of
// exon_cv.tsv, gc_content.tsv and the four fragment-GC rows of metrics.tsv (src/RNASeQC.cpp:632-640,660-674).  Throws
// HostError with the reference's exit codes.
void write_reports(const Annotation& A, const HostResults& R, const RunOptions& opt, const StreamState& st,
                   const ReferenceModel* ref = nullptr);

}  // namespace rsqh
