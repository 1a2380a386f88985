// pipeline.h — the host driver shared by the CLI and the Python binding: GTF/BED -> annotation,
// SAM/BAM -> record batches -> rsq_submit (whatever library implements include/rsq_b200.h is linked or
// passed in as a function table) -> rsq_finalize -> report files.  Mirrors the phases of main()
// (src/RNASeQC.cpp:36-676); the hot loop itself lives behind the C-ABI.
#pragma once
#include <string>
#include <vector>

#include "alnreader.h"
#include "annotation.h"
#include "report.h"
#include "rsq_b200.h"

namespace rsqh {

// Function table of one backend library (resolved with dlsym by the Python binding, or bound statically
// by the CLI).  Keeps pipeline.cpp independent of WHICH library implements the ABI.
struct Backend {
    decltype(&rsq_backend) backend;
    decltype(&rsq_create) create;
    decltype(&rsq_destroy) destroy;
    decltype(&rsq_last_error) last_error;
    decltype(&rsq_set_annotation) set_annotation;
    decltype(&rsq_set_bed) set_bed;
    decltype(&rsq_set_reference) set_reference;
    decltype(&rsq_submit) submit;
    decltype(&rsq_submit_packed) submit_packed;
    decltype(&rsq_submit_compact) submit_compact;
    decltype(&rsq_sync) sync;
    decltype(&rsq_sync_inputs) sync_inputs;
    decltype(&rsq_finalize) finalize;
    decltype(&rsq_coverage_statistics) coverage_statistics;
    decltype(&rsq_host_alloc) host_alloc;
    decltype(&rsq_host_free) host_free;
    // coordinate-partitioned runs on several GPUs (--gpus N): the exchange protocol of include/rsq_b200.h
    decltype(&rsq_set_partition) set_partition;
    decltype(&rsq_export_sample_counts) export_sample_counts;
    decltype(&rsq_export_counts) export_counts;
    decltype(&rsq_import_counts) import_counts;
    decltype(&rsq_comm_unique_id) comm_unique_id;
    decltype(&rsq_comm_init) comm_init;
    decltype(&rsq_exchange) exchange;
};

struct CliOptions {
    std::string gtf, bam, output_dir, bed;
    RunOptions run;
    DecodeOptions decode;
    rsq_config cfg;
    bool legacy = false;
    std::string fasta;
    size_t batch_records = 1u << 20;
    bool packed_submit = true;        // send batches in the wire formats (rsq_batch16, else rsq_read_packed) when they fit them
    bool compact_submit = true;       // ... the 16-byte one first (--no-compact-submit: A/B)
    std::vector<int> devices;         // --gpus N / --devices a,b,c: one context per GPU, records range-partitioned by coordinate
    size_t records_per_gpu = 0;       // --records-per-gpu: records a GPU takes before the stream moves on to the next (0 = by file size)
    int decode_threads = -1;          // -1: hardware threads (<= 32); 0: sequential reader
    int verbosity = 0;
    bool exit_after_reports = false;  // the CLI binary: once the report files are closed the process ends without tearing down page-locked
                                      // buffers, device tables and ingest threads one by one (the OS does that in one sweep)
};

// Fills `opt` from argv with the reference's flag set (src/RNASeQC.cpp:39-65).  Returns an exit code < 0 to
// continue, or >= 0 to exit immediately (help/version/validation error: 0, 4, 5, 6).
int parse_cli(int argc, char** argv, CliOptions& opt);

// Runs one sample end to end; returns the process exit code (the reference's table, src/RNASeQC.cpp:678-764).
int run_sample(const Backend& be, CliOptions& opt);

// Pulls results out of a finalised ctx into HostResults sized for annotation A.
void fetch_results(const Backend& be, rsq_ctx* ctx, const Annotation& A, bool want_coverage, HostResults& R);
// Per-gene coverage statistics from the backend (device kernel in the product) into R.cs_*.
// `owned` (may be null): rsq_transcripts.gene_owned of a partitioned run.
void fetch_coverage_statistics(const Backend& be, rsq_ctx* ctx, const Annotation& A, const RunOptions& opt, HostResults& R,
                               const std::vector<uint8_t>* owned = nullptr);

}  // namespace rsqh
