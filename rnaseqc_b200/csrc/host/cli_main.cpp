// cli_main.cpp — `rnaseqc gtf bam outdir {OPTIONS}`: drop-in CLI over the C-ABI (src/RNASeQC.cpp:36).
// Linked against rnaseqc_b200/lib/librsq_b200.so this is the product binary (CUDA hot path, no CPU
// fallback).  The SAME source linked against oracle/_build/librsq_oracle.so gives the test harness
// oracle/_build/rnaseqc_oracle used to validate host parsing/reporting without a GPU.
#include <iostream>

#include "pipeline.h"

int main(int argc, char** argv) {
    rsqh::CliOptions opt;
    int rc = rsqh::parse_cli(argc, argv, opt);
    if (rc >= 0) return rc;
    rsqh::Backend be{rsq_backend, rsq_create, rsq_destroy, rsq_last_error, rsq_set_annotation, rsq_set_bed, rsq_set_reference, rsq_submit,
                     rsq_submit_packed, rsq_submit_compact, rsq_sync, rsq_sync_inputs, rsq_finalize, rsq_coverage_statistics, rsq_host_alloc,
                     rsq_host_free, rsq_set_partition, rsq_export_sample_counts, rsq_export_counts, rsq_import_counts, rsq_comm_unique_id,
                     rsq_comm_init, rsq_exchange};
    if (opt.verbosity) std::cout << "hot-path backend: " << be.backend() << std::endl;
    opt.exit_after_reports = true;
    return rsqh::run_sample(be, opt);
}
