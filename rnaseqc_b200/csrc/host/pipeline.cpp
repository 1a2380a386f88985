// pipeline.cpp — host driver (see pipeline.h).
#include "pipeline.h"

#include "bamparallel.h"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <filesystem>
#include <future>
#include <iostream>
#include <thread>

namespace rsqh {

namespace {

const char* kVersion = "RNASeQC 2.4.3";  // the CLI this build is a drop-in for (src/RNASeQC.cpp:25)

void usage(std::ostream& o) {
    o << "  rnaseqc gtf bam output {OPTIONS}\n\n"
         "    " << kVersion << " — B200-native hot path (rnaseqc-b200)\n\n"
         "  OPTIONS:\n"
         "      -h, --help                        Display this message and quit\n"
         "      --version                         Display the version and quit\n"
         "      gtf                               The input GTF file containing features to check the bam against\n"
         "      bam                               The input SAM/BAM file containing reads to process\n"
         "      output                            Output directory\n"
         "      -s[sample], --sample=[sample]     The name of the current sample. Default: The bam's filename\n"
         "      --bed=[BEDFILE]                   BED file with non-overlapping exons used for fragment size calculations\n"
         "      --fasta=[fasta]                   Reference genome FASTA (indexed, .fai): enables GC content statistics\n"
         "      --chimeric-distance=[DISTANCE]    Maximum accepted distance between read mates. Default: 2000000 [bp]\n"
         "      --fragment-samples=[SAMPLES]      Number of fragment size samples. Requires --bed. Default: 1000000\n"
         "      -q[QUALITY], --mapping-quality=[QUALITY]  Lower bound on read quality for exon coverage counting. Default: 255\n"
         "      --base-mismatch=[MISMATCHES]      Maximum number of allowed mismatches. Default: 6\n"
         "      --offset=[OFFSET]                 Offset into the gene for the 3' and 5' windows in bias calculation\n"
         "      --window-size=[SIZE]              Size of the 3' and 5' windows in bias calculation. Default: 100 [bp]\n"
         "      --gene-length=[LENGTH]            Minimum size of a gene for bias calculation\n"
         "      --legacy                          Use legacy counting rules.  Gene and exon counts match output of RNA-SeQC 1.1.9\n"
         "      --stranded=[stranded]             Strand-specific metrics: 'RF', 'rf', 'FR', or 'fr'\n"
         "      -v, --verbose                     Give some feedback; twice for progress updates\n"
         "      -t[TAG...], --tag=[TAG...]        Filter out reads with the specified tag\n"
         "      --chimeric-tag=[TAG]              Reads marked with the specified tag are labeled Chimeric. Default 'ch'\n"
         "      --exclude-chimeric                Exclude chimeric reads from the read counts\n"
         "      -u, --unpaired                    Allow unpaired reads to be quantified\n"
         "      --rpkm                            Output gene RPKM values instead of TPMs\n"
         "      --coverage                        Write per-transcript coverage statistics to a table\n"
         "      --coverage-mask=[SIZE]            Bases masked at both transcript ends for coverage. Default: 500bp\n"
         "      -d[threshold], --detection-threshold=[threshold]  Counts on a gene to consider it detected. Default: 5\n"
         "      --batch-records=[N]               Records per device batch (this build). Default: 1048576\n"
         "      --decode-threads=[N]              BAM inflate/convert worker threads (this build); 0 = sequential reader.\n"
         "      --no-packed-submit                Send 48-byte records instead of the 16- / 32-byte wire formats (this build)\n"
         "      --no-compact-submit               Do not use the 16-byte wire format (this build)\n"
         "      --gpus=[N]                        Range-partition the records by coordinate over N GPUs (this build); results are\n"
         "                                        identical to a single-GPU run.  Needs a coordinate-sorted alignment file\n"
         "      --devices=[a,b,..]                The same with explicit CUDA device ordinals (this build)\n"
         "      --records-per-gpu=[N]             Records a GPU takes before the stream moves on to the next one (this build).\n"
         "                                        Default: the file's estimated record count / GPUs\n"
         "                                        Default: hardware threads, at most 32\n"
         "      --exact-complexity                Literal 1e9-iteration library-complexity scan (this build)\n";
}

struct ArgError { int code; std::string msg; };

}  // namespace

int parse_cli(int argc, char** argv, CliOptions& opt) {
    memset(&opt.cfg, 0, sizeof(opt.cfg));
    opt.cfg.struct_size = sizeof(rsq_config);
    opt.cfg.device = 0;
    opt.cfg.chimeric_distance = 2000000;
    opt.cfg.base_mismatch_threshold = 6;
    opt.cfg.mapq_threshold = 255;
    opt.cfg.collect_coverage = 1;
    unsigned fragment_samples = 1000000u;
    bool mapq_given = false;
    std::vector<std::string> positional;
    try {
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i];
            auto value = [&](const std::string& longname, char shortname) -> std::string {
                // --name=value | --name value | -nvalue | -n value
                if (a.compare(0, 2, "--") == 0) {
                    size_t eq = a.find('=');
                    if (eq != std::string::npos) return a.substr(eq + 1);
                } else if (shortname && a.size() > 2) {
                    return a.substr(2);
                }
                if (i + 1 >= argc) throw ArgError{5, "Flag '" + longname + "' requires an argument but received none"};
                return argv[++i];
            };
            auto is = [&](const char* longname, char shortname = 0) {
                std::string l = std::string("--") + longname;
                if (a == l || a.compare(0, l.size() + 1, l + "=") == 0) return true;
                return shortname && a.size() >= 2 && a[0] == '-' && a[1] == shortname && a[1] != '-';
            };
            if (is("help", 'h')) { usage(std::cout); return 4; }
            else if (a == "--version") { std::cout << kVersion << std::endl; return 0; }
            else if (is("sample", 's')) { opt.run.sample_name = value("sample", 's'); opt.run.sample_given = true; }
            else if (is("bed")) opt.bed = value("bed", 0);
            else if (is("fasta")) opt.fasta = value("fasta", 0);
            else if (is("chimeric-distance")) opt.cfg.chimeric_distance = std::stoi(value("chimeric-distance", 0));
            else if (is("fragment-samples")) fragment_samples = (unsigned)std::stoul(value("fragment-samples", 0));
            else if (is("mapping-quality", 'q')) { opt.cfg.mapq_threshold = (unsigned)std::stoul(value("mapping-quality", 'q')); mapq_given = true; }
            else if (is("base-mismatch")) opt.cfg.base_mismatch_threshold = (unsigned)std::stoul(value("base-mismatch", 0));
            else if (is("offset")) opt.run.bias_offset = std::stoi(value("offset", 0));
            else if (is("window-size")) opt.run.bias_window = std::stoi(value("window-size", 0));
            else if (is("gene-length")) opt.run.bias_gene_length = std::stoul(value("gene-length", 0));
            else if (a == "--legacy") opt.legacy = true;
            else if (is("stranded")) {
                std::string s = value("stranded", 0);
                if (s == "RF" || s == "rf") opt.cfg.strand_mode = 2;
                else if (s == "FR" || s == "fr") opt.cfg.strand_mode = 1;
                else throw ArgError{6, "--stranded argument must be in {'RF', 'rf', 'FR', 'fr'}"};
            }
            else if (a == "--verbose") ++opt.verbosity;
            else if (a.size() >= 2 && a[0] == '-' && a[1] == 'v' && a.find_first_not_of('v', 1) == std::string::npos) opt.verbosity += (int)a.size() - 1;
            else if (is("tag", 't')) opt.decode.filter_tags.push_back(value("tag", 't'));
            else if (is("chimeric-tag")) opt.decode.chimeric_tag = value("chimeric-tag", 0);
            else if (a == "--exclude-chimeric") opt.cfg.exclude_chimeric = 1;
            else if (a == "--unpaired" || a == "-u") opt.cfg.unpaired = 1;
            else if (a == "--rpkm") opt.run.use_rpkm = true;
            else if (a == "--coverage") opt.run.write_coverage = true;
            else if (is("coverage-mask")) opt.run.coverage_mask = (unsigned)std::stoul(value("coverage-mask", 0));
            else if (is("detection-threshold", 'd')) opt.run.detection_threshold = (unsigned)std::stoul(value("detection-threshold", 'd'));
            else if (is("batch-records")) opt.batch_records = std::stoul(value("batch-records", 0));
            else if (is("decode-threads")) opt.decode_threads = std::stoi(value("decode-threads", 0));
            else if (a == "--no-packed-submit") opt.packed_submit = false;
            else if (is("device")) opt.cfg.device = std::stoi(value("device", 0));
            else if (is("gpus")) { const int n = std::stoi(value("gpus", 0)); if (n < 1 || n > 64) throw ArgError{6, "--gpus must be between 1 and 64"}; opt.devices.clear(); for (int k = 0; k < n; ++k) opt.devices.push_back(k); }
            else if (is("devices")) {
                opt.devices.clear();
                std::string v = value("devices", 0), tok;
                for (size_t p0 = 0; p0 <= v.size(); ++p0) {
                    if (p0 == v.size() || v[p0] == ',') { if (!tok.empty()) opt.devices.push_back(std::stoi(tok)); tok.clear(); }
                    else tok += v[p0];
                }
                if (opt.devices.empty()) throw ArgError{6, "--devices needs a comma-separated list of CUDA device ordinals"};
            }
            else if (is("records-per-gpu")) opt.records_per_gpu = std::stoul(value("records-per-gpu", 0));
            else if (a == "--no-compact-submit") opt.compact_submit = false;
            else if (a == "--exact-complexity") opt.run.exact_complexity_scan = true;
            else if (a.size() > 1 && a[0] == '-') throw ArgError{5, "Flag could not be matched: " + a};
            else positional.push_back(a);
        }
        if (positional.size() > 3) throw ArgError{5, "Passed in argument, but no positional arguments were ready to receive it: " + positional[3]};
        if (positional.size() < 1) throw ArgError{6, "No GTF file provided"};
        if (positional.size() < 2) throw ArgError{6, "No BAM file provided"};
        if (positional.size() < 3) throw ArgError{6, "No output directory provided"};
    } catch (ArgError& e) {
        usage(std::cerr);
        std::cerr << (e.code == 5 ? "Argument parsing error: " : "Argument validation error: ") << e.msg << std::endl;
        return e.code;
    } catch (std::exception& e) {  // stoi/stoul failures are ParseErrors in args.hxx
        usage(std::cerr);
        std::cerr << "Argument parsing error: " << e.what() << std::endl;
        return 5;
    }
    opt.gtf = positional[0];
    opt.bam = positional[1];
    opt.output_dir = positional[2];
    opt.run.output_dir = opt.output_dir;
    if (!opt.run.sample_given) opt.run.sample_name = std::filesystem::path(opt.bam).filename().string();  // :99
    if (opt.legacy) {                                                  // src/RNASeQC.cpp:90: the default MAPQ gate drops to 4
        opt.cfg.legacy = 1;
        opt.decode.legacy = true;
        if (!mapq_given) opt.cfg.mapq_threshold = 4;
    }
    if (opt.decode.filter_tags.size() > 8) {
        std::cerr << "At most 8 --tag filters are supported" << std::endl;
        return 6;
    }
    // what rsq_coverage_statistics accepts: refused here, before hours of alignments are processed (with --window-size 0 the
    // reference dies of an empty median as soon as a gene gets that far; with a negative --offset its window loops run over an
    // unsigned wrap-around, read past the end of the coverage vector and end the same way)
    if (opt.run.bias_window < 1 || opt.run.bias_offset < 0) {
        std::cerr << "Argument validation error: --window-size must be at least 1 and --offset must not be negative in this build" << std::endl;
        return 6;
    }
    opt.cfg.n_filter_tags = (uint32_t)opt.decode.filter_tags.size();
    opt.cfg.fragment_samples = opt.bed.empty() ? 0u : fragment_samples;
    if (opt.decode_threads < 0) opt.decode_threads = (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
    opt.decode.exclude_chimeric = opt.cfg.exclude_chimeric;
    opt.decode.chimeric_distance = opt.cfg.chimeric_distance;
    opt.decode.verbosity = opt.verbosity;
    opt.run.filter_tags = opt.decode.filter_tags;
    return -1;
}

void fetch_results(const Backend& be, rsq_ctx* ctx, const Annotation& A, bool want_coverage, HostResults& R) {
    const size_t ng = A.gene_ids.size(), ne = A.exon_ids.size();
    R.gene_reads.assign(ng, 0); R.gene_unique.assign(ng, 0); R.gene_fragments.assign(ng, 0);
    R.exon_whole.assign(ne, 0); R.exon_frac.assign(ne, 0);
    if (want_coverage) R.coverage.assign(A.cov_len + 1, 0); else R.coverage.clear();
    R.fragment_hist.assign(1u << 20, 0);
    rsq_results out;
    memset(&out, 0, sizeof(out));
    out.struct_size = sizeof(out);
    out.gene_reads = R.gene_reads.data(); out.gene_unique = R.gene_unique.data(); out.gene_fragments = R.gene_fragments.data();
    out.exon_whole = R.exon_whole.data(); out.exon_frac = R.exon_frac.data();
    out.coverage = want_coverage ? R.coverage.data() : nullptr;
    out.fragment_hist = R.fragment_hist.data(); out.fragment_hist_len = R.fragment_hist.size();
    out.gc_hist = R.gc_hist;
    int rc = be.finalize(ctx, &out);
    if (rc != RSQ_OK) throw HostError(-1, std::string("rsq_finalize failed: ") + be.last_error(ctx));
    memcpy(R.counters, out.counters, sizeof(R.counters));
    R.fragment_samples_taken = out.fragment_samples_taken;
    // trim the histogram to its used range
    size_t last = 0;
    for (size_t i = 0; i < R.fragment_hist.size(); ++i) if (R.fragment_hist[i]) last = i + 1;
    R.fragment_hist.resize(last);
}

void fetch_coverage_statistics(const Backend& be, rsq_ctx* ctx, const Annotation& A, const RunOptions& opt, HostResults& R,
                               const std::vector<uint8_t>* owned) {
    rsq_transcripts q = A.transcripts(opt.coverage_mask, opt.bias_offset, opt.bias_window, opt.bias_gene_length);
    if (owned) q.gene_owned = owned->data();
    R.cs_mean.assign(q.n_genes, 0); R.cs_std.assign(q.n_genes, 0); R.cs_cv.assign(q.n_genes, 0);
    R.cs_five.assign(q.n_genes, 0); R.cs_three.assign(q.n_genes, 0);
    R.cs_exon_cv.assign(A.exon_ids.size(), 0);
    rsq_coverage_stats out;
    memset(&out, 0, sizeof(out));
    out.struct_size = sizeof(out);
    out.gene_mean = R.cs_mean.data(); out.gene_std = R.cs_std.data(); out.gene_cv = R.cs_cv.data();
    out.five_end = R.cs_five.data(); out.three_end = R.cs_three.data(); out.exon_cv = R.cs_exon_cv.data();
    int rc = be.coverage_statistics(ctx, &q, &out);
    if (rc != RSQ_OK) throw HostError(rc == RSQ_ERR_INVALID ? 6 : -1, std::string("rsq_coverage_statistics failed: ") + be.last_error(ctx));
    R.have_cov_stats = true;
}

namespace {

// One GPU of the run: its context, three page-locked staging buffers (a batch is packed into one while the previous two are on
// the wire or waiting for their copy) and the route its records take.  A single-GPU run is the one-lane case of the same code.
struct Lane {
    const Backend* be = nullptr;
    rsq_ctx* ctx = nullptr;
    int device = 0;
    std::future<int> creating;
    static constexpr int kSlots = 3;
    void* pinned[kSlots] = {nullptr, nullptr, nullptr};
    size_t cap[kSlots] = {0, 0, 0};
    bool in_flight[kSlots] = {false, false, false};
    int next_slot = 0;
    uint64_t records = 0, batches = 0, wire_bytes = 0;
    std::vector<uint8_t> owned;

    // a staging buffer of at least `bytes` bytes whose previous content has left the host
    int acquire(size_t bytes) {
        const int s = next_slot;
        next_slot = (next_slot + 1) % kSlots;
        if (in_flight[s]) {
            const int rc = be->sync_inputs(ctx);
            if (rc != RSQ_OK) throw HostError(-1, std::string("rsq_sync_inputs failed: ") + be->last_error(ctx));
            for (bool& f : in_flight) f = false;
        }
        if (cap[s] < bytes) {
            if (pinned[s]) be->host_free(pinned[s]);
            cap[s] = bytes + bytes / 4;
            pinned[s] = be->host_alloc(cap[s]);
            if (!pinned[s]) { cap[s] = 0; throw HostError(10, "Memory allocation failure. Out of memory"); }
        }
        return s;
    }
    void release() {
        for (int s = 0; s < kSlots; ++s) if (pinned[s]) { be->host_free(pinned[s]); pinned[s] = nullptr; }
        if (ctx) { be->destroy(ctx); ctx = nullptr; }
    }
};

// One piece of a decoded batch on its way to one lane: packed on a helper thread, submitted on the main thread, in order.
struct Staged {
    int lane = 0, slot = 0;
    size_t n_reads = 0, n_blocks = 0, n_packed_blocks = 0, compact_bytes = 0, blocks_off = 0;
    std::future<int> how;        // 2 = rsq_batch16, 1 = rsq_read_packed, 0 = rsq_read
};

}  // namespace

int run_sample(const Backend& be, CliOptions& opt) {
    using clk = std::chrono::steady_clock;
    if (opt.devices.empty()) opt.devices.push_back(opt.cfg.device);
    const size_t n_lanes = opt.devices.size();
    std::vector<Lane> lanes(n_lanes);
    int exit_code = 0;
    std::vector<Staged> staged;
    try {
        // The device contexts come up on helper threads while the GTF is parsed (CUDA start-up costs about as much as the parse
        // on a multi-GPU box); their outcome is looked at where rsq_create used to be called, so errors keep their precedence.
        for (size_t k = 0; k < n_lanes; ++k) {
            Lane& L = lanes[k];
            L.be = &be; L.device = opt.devices[k];
            L.creating = std::async(std::launch::async, [&be, &opt, &L] {
                rsq_config cfg = opt.cfg;
                cfg.device = L.device;
                const int rc = be.create(&cfg, &L.ctx);
                if (rc == RSQ_OK) {
                    // the lane's page-locked staging buffers, sized for a full batch of two-block records: page-locking ~80 MB takes
                    // tens of milliseconds each, which the main thread used to spend inside its first three batches
                    const size_t n = opt.batch_records;
                    const size_t est = std::max(rsq_compact_bound(n, 2 * n), 48 * n + 16 * n + 16);
                    for (int s = 0; s < Lane::kSlots; ++s) {
                        L.pinned[s] = be.host_alloc(est);
                        L.cap[s] = L.pinned[s] ? est : 0;             // no room now: acquire() asks again, and reports it
                    }
                }
                return rc;
            });
        }
        // So does the BAM ingest: inflating, framing and converting records needs the command line and the BAM header only
        // (bamparallel.h: start), so the worker threads run ahead while this thread parses the GTF, builds the index and uploads
        // it.  A file that cannot be opened or is no BAM is reported where the reference opens it (after the GTF, the BED and the
        // output directory), not here.
        std::unique_ptr<ParallelBamDecoder> pbam;
        std::exception_ptr early_bam_error;
        if (opt.decode_threads > 0 && !getenv("RSQ_INGEST_LATE_START")) {
            unsigned char m[2] = {0, 0};
            size_t got = 0;
            if (FILE* fp = std::fopen(opt.bam.c_str(), "rb")) { got = std::fread(m, 1, 2, fp); std::fclose(fp); }
            if (got == 2 && m[0] == 0x1f && m[1] == 0x8b) {
                try {
                    pbam.reset(new ParallelBamDecoder(opt.bam, opt.decode_threads));
                    pbam->start(opt.decode);
                } catch (...) { early_bam_error = std::current_exception(); pbam.reset(); }
            }
        }
        Annotation A;
        ReferenceModel ref;
        if (!opt.fasta.empty()) {                                                  // src/RNASeQC.cpp:117-122: before the GTF
            open_reference(opt.fasta, A.contigs, ref);
            if (opt.verbosity > 1) std::cout << "A FASTA has been provided. This will enable GC-content statistics but adds additional runtime and memory costs" << std::endl;
        }
        auto t0 = clk::now();
        if (opt.verbosity) std::cout << "Reading GTF Features..." << std::endl;
        load_gtf(opt.gtf, A, opt.legacy);
        if (opt.verbosity)
            std::cout << "Finished processing GTF in " << std::chrono::duration<double>(clk::now() - t0).count() << " seconds" << std::endl;
        BedModel bed;
        if (!opt.bed.empty()) {
            if (opt.verbosity) std::cout << "Parsing BED intervals for fragment size computations..." << std::endl;
            load_bed(opt.bed, A.contigs, bed);
        }
        const size_t n_contigs = A.contigs.names.size();
        A.pad_contigs(n_contigs);
        if (!opt.bed.empty()) flatten_bed(bed, n_contigs);
        if (n_contigs > 32767) throw HostError(11, "more than 32767 annotated contigs");
        if (ref.open) load_reference_bits(ref, n_contigs);
        try {
            if (!std::filesystem::exists(opt.output_dir)) std::filesystem::create_directories(opt.output_dir);
        } catch (std::filesystem::filesystem_error& e) {
            throw HostError(8, std::string("Filesystem error:  ") + e.what());
        }

        // BAM goes through the multi-threaded ingest (bamparallel.cpp); SAM text through the sequential reader
        std::unique_ptr<AlnReader> reader;
        if (early_bam_error) std::rethrow_exception(early_bam_error);              // ... with the precedence it always had
        if (!pbam) {
            FILE* fp = std::fopen(opt.bam.c_str(), "rb");
            if (!fp) throw HostError(10, "Unable to open BAM file: " + opt.bam);
            unsigned char m[2] = {0, 0};
            const size_t got = std::fread(m, 1, 2, fp);
            std::fclose(fp);
            if (got == 2 && m[0] == 0x1f && m[1] == 0x8b && opt.decode_threads > 0) pbam.reset(new ParallelBamDecoder(opt.bam, opt.decode_threads));
            else reader = AlnReader::open(opt.bam, opt.fasta, opt.decode_threads);
        }
        const AlnHeader& aln_header = pbam ? pbam->header() : reader->header();
        Decoder decoder(aln_header, A.contigs, n_contigs, opt.decode);
        {
            std::vector<uint8_t> has_features(A.contigs.names.size(), 0);
            for (size_t c = 0; c < A.genes_by_contig.size(); ++c)
                has_features[c] = !A.genes_by_contig[c].empty() || !A.exons_by_contig[c].empty();
            if (!decoder.any_overlap_with(has_features)) {
                throw HostError(11, "BAM file shares no contigs with GTF");       // src/RNASeQC.cpp:234-238
            }
        }
        // An exon that ends before it starts has a "negative" length (1 + end - start in an unsigned type, src/GTF.cpp:108).  The
        // reference meets it when the exon's gene leaves the window — every kept gene line does, at the latest after the loop
        // (src/RNASeQC.cpp:385-386) — and sizes a coverage vector with it (src/Metrics.cpp:138): std::length_error, exit 1
        // (src/RNASeQC.cpp:729-734).  Same outcome here, before any alignment is processed.
        for (const auto& genes : A.genes_by_contig)
            for (const GeneFeature& g : genes)
                for (uint32_t slot : A.exons_for_gene[g.gene])
                    if (A.exon_len[slot] < 0)
                        throw HostError(1, "Unable to parse the GFT lines\ncannot create std::vector larger than max_size()");

        auto t_setup = clk::now();
        for (Lane& L : lanes) {
            int rc = L.creating.get();
            if (rc != RSQ_OK) throw HostError(rc == RSQ_ERR_NO_DEVICE ? 10 : -1, std::string("rsq_create failed: ") + be.last_error(nullptr));
        }
        {   // annotation / BED / reference upload: one thread per GPU
            rsq_annotation av = A.view();
            rsq_bed bv{};
            if (!opt.bed.empty()) bv = bed.view();
            rsq_reference rv{};
            if (ref.open) rv = ref.view();
            std::vector<std::future<std::string>> ups;
            for (Lane& L : lanes)
                ups.push_back(std::async(n_lanes > 1 ? std::launch::async : std::launch::deferred, [&, ctx = L.ctx]() -> std::string {
                    if (be.set_annotation(ctx, &av) != RSQ_OK) return std::string("rsq_set_annotation failed: ") + be.last_error(ctx);
                    if (!opt.bed.empty() && be.set_bed(ctx, &bv) != RSQ_OK) return std::string("rsq_set_bed failed: ") + be.last_error(ctx);
                    if (ref.open && be.set_reference(ctx, &rv) != RSQ_OK) return std::string("rsq_set_reference failed: ") + be.last_error(ctx);
                    return std::string();
                }));
            std::string why;
            for (auto& f : ups) { const std::string w = f.get(); if (why.empty()) why = w; }
            if (!why.empty()) throw HostError(-1, why);
        }
        // Several GPUs: the communicator of the end-of-run exchange.  NCCL inside the library when the backend has it (the CUDA
        // product); the CPU oracle backend of the test harness has none, and the exchange buffers travel through host memory.
        bool nccl = false;
        if (n_lanes > 1) {
            uint8_t id[RSQ_COMM_ID_BYTES];
            // RSQ_HOST_EXCHANGE=1 (test harness only): the emulated CUDA library's "device" buffers are host memory too
            const bool host_exchange = std::string(be.backend()) == "oracle-cpu" || (getenv("RSQ_HOST_EXCHANGE") && getenv("RSQ_HOST_EXCHANGE")[0] == '1');
            if (!host_exchange && be.comm_unique_id(id) == RSQ_OK) {
                std::vector<std::future<int>> inits;
                for (size_t k = 0; k < n_lanes; ++k)
                    inits.push_back(std::async(std::launch::async, [&, k] { return be.comm_init(lanes[k].ctx, id, (uint32_t)k, (uint32_t)n_lanes); }));
                int bad = RSQ_OK;
                for (auto& f : inits) { const int r = f.get(); if (r != RSQ_OK) bad = r; }
                if (bad != RSQ_OK) throw HostError(-1, std::string("rsq_comm_init failed: ") + be.last_error(lanes[0].ctx));
                nccl = true;
            } else if (host_exchange) {
                for (size_t k = 0; k < n_lanes; ++k)
                    if (be.set_partition(lanes[k].ctx, (uint32_t)k, (uint32_t)n_lanes) != RSQ_OK) throw HostError(-1, "rsq_set_partition failed");
            } else throw HostError(-1, std::string("--gpus needs NCCL: ") + be.last_error(nullptr));
        }

        if (opt.verbosity) std::cout << "Device context + annotation upload: " << std::chrono::duration<double>(clk::now() - t_setup).count() << " seconds" << std::endl;
        if (opt.verbosity) std::cout << "Parsing bam..." << std::endl;
        auto t1 = clk::now();

        // ---- routing of the coordinate-sorted stream over the lanes -------------------------------------------
        // Lane g takes records until it has its share, then the stream moves on to lane g + 1 at the next position that lies in no
        // island (gene hull or BED interval): all records that count toward a gene, and both mates of a pair, stay on one GPU.
        Islands islands;
        std::vector<int64_t> contig_rank(n_contigs, -1);          // annotation contig id -> rank in the stream's (header's) order
        if (n_lanes > 1) {
            build_islands(A, opt.bed.empty() ? nullptr : &bed, islands);
            const std::vector<int>& t2c = decoder.tid_to_contig();
            for (size_t t = 0; t < t2c.size(); ++t)
                if (t2c[t] >= 0 && (size_t)t2c[t] < n_contigs && contig_rank[t2c[t]] < 0) contig_rank[t2c[t]] = (int64_t)t;
        }
        std::vector<StreamCut> cuts;          // cuts[g] = first coordinate of lane g + 1
        bool have_pending = false;
        StreamCut pending{0, 0};
        size_t g = 0;
        auto key_of = [&](const rsq_read& r, StreamCut& k) -> bool {     // false: the record carries no usable coordinate
            if (r.contig < 0 || (size_t)r.contig >= n_contigs || contig_rank[r.contig] < 0) return false;
            k.contig_rank = contig_rank[r.contig]; k.pos1 = (int64_t)r.pos + 1;
            return true;
        };
        auto before = [](const StreamCut& a, const StreamCut& b) { return a.contig_rank < b.contig_rank || (a.contig_rank == b.contig_rank && a.pos1 < b.pos1); };
        auto share = [&]() -> uint64_t {                                 // records a lane takes before the stream moves on
            if (opt.records_per_gpu) return opt.records_per_gpu;
            if (pbam && pbam->compressed_consumed() > 0 && pbam->file_bytes() > 0) {
                const double est = (double)decoder.alignment_count() * (double)pbam->file_bytes() / (double)pbam->compressed_consumed();
                return (uint64_t)(est / (double)n_lanes) + 1;
            }
            return (uint64_t)1 << 24;
        };

        // Two batch buffers: while helper threads pack batch i into page-locked slots (16-byte wire format, else the 32-byte
        // one, else the plain copy), this thread already collects batch i+1 from the decoder; submits stay on this thread, in order.
        std::vector<rsq_read> reads_buf[2];
        std::vector<rsq_block> blocks_buf[2];
        reads_buf[0].reserve(opt.batch_records);
        reads_buf[1].reserve(opt.batch_records);
        RawAln aln;
        auto submit_staged = [&]() {            // waits for the packers of the previous batch, then submits its pieces
            std::vector<Staged> todo;
            todo.swap(staged);
            std::string why;
            for (Staged& st : todo) {
                const int how = st.how.get();
                if (!why.empty()) continue;
                Lane& L = lanes[st.lane];
                unsigned char* buf = (unsigned char*)L.pinned[st.slot];
                int r2;
                if (how == 2) { r2 = be.submit_compact(L.ctx, buf, st.compact_bytes); L.wire_bytes += st.compact_bytes; }
                else if (how == 1) { r2 = be.submit_packed(L.ctx, (const rsq_read_packed*)buf, st.n_reads, (const rsq_block*)(buf + st.blocks_off), st.n_packed_blocks); L.wire_bytes += 32 * st.n_reads + 8 * st.n_packed_blocks; }
                else { r2 = be.submit(L.ctx, (const rsq_read*)buf, st.n_reads, (const rsq_block*)(buf + st.blocks_off), st.n_blocks); L.wire_bytes += 48 * st.n_reads + 8 * st.n_blocks; }
                L.in_flight[st.slot] = true;
                L.records += st.n_reads; ++L.batches;
                if (r2 != RSQ_OK) why = std::string("rsq_submit failed: ") + be.last_error(L.ctx);
            }
            if (!why.empty()) throw HostError(-1, why);
        };
        auto stage = [&](size_t lane, const std::vector<rsq_read>& reads, const std::vector<rsq_block>& blocks, size_t i0, size_t i1) {
            if (i1 <= i0) return;
            Lane& L = lanes[lane];
            const size_t n = i1 - i0;
            const size_t b0 = reads[i0].block_off, b1 = (size_t)reads[i1 - 1].block_off + reads[i1 - 1].n_blocks, nb = b1 - b0;
            const size_t blocks_off = (48 * n + 15) & ~(size_t)15;
            const size_t need = std::max(rsq_compact_bound(n, nb), blocks_off + 8 * std::max<size_t>(nb, 1));
            Staged st;
            st.lane = (int)lane; st.slot = L.acquire(need); st.n_reads = n; st.n_blocks = nb; st.blocks_off = blocks_off;
            staged.push_back(std::move(st));
            Staged* sp = &staged.back();
            unsigned char* dst = (unsigned char*)L.pinned[sp->slot];
            const size_t cap = L.cap[sp->slot];
            const rsq_read* src_reads = reads.data() + i0;
            const rsq_block* src_blocks = blocks.data();        // block_off of the records stays relative to the whole batch
            const size_t n_all_blocks = blocks.size();
            const bool try_compact = opt.packed_submit && opt.compact_submit && be.submit_compact;
            const bool try_packed = opt.packed_submit && be.submit_packed;
            size_t* compact_bytes = &sp->compact_bytes;
            size_t* n_packed = &sp->n_packed_blocks;
            sp->how = std::async(std::launch::async, [=]() -> int {
                if (try_compact) {
                    std::vector<uint32_t> scratch(131072);
                    if (rsq_pack_compact(src_reads, n, src_blocks, n_all_blocks, dst, cap, scratch.data(), compact_bytes) == RSQ_OK) return 2;
                }
                if (try_packed && rsq_pack_records(src_reads, n, src_blocks, n_all_blocks, (rsq_read_packed*)dst, (rsq_block*)(dst + blocks_off), n_packed) == RSQ_OK) return 1;
                rsq_read* out = (rsq_read*)dst;
                memcpy(out, src_reads, n * sizeof(rsq_read));
                for (size_t i = 0; i < n; ++i) out[i].block_off -= (uint32_t)b0;
                memcpy(dst + blocks_off, src_blocks + b0, nb * sizeof(rsq_block));
                return 0;
            });
        };
        int cur = 0;
        bool more = true;
        // One GPU, BAM input, 16-byte wire format: the batches are not copied together at all.  The decoder hands every batch over
        // as a list of pieces that stay in the ingest workers' buffers (bamparallel.h: next_batch_segments) and the packers read
        // them there, binding the contig ids on the way — this thread only walks over job digests, and up to three packers (one
        // per page-locked slot) work at a time.  A batch that does not suit the format is copied together by its packer and goes
        // down the usual chain of formats.
        const bool by_segments = pbam && n_lanes == 1 && opt.packed_submit && opt.compact_submit && be.submit_compact && !getenv("RSQ_INGEST_COPY_BATCHES");
        if (by_segments) {
            struct Piece {
                std::shared_ptr<ParallelBamDecoder::Segments> segs;
                Staged st;
            };
            std::deque<Piece> queue;                       // staged batches, oldest first
            size_t by_format[3] = {0, 0, 0};
            const std::vector<int16_t> contig_of_tid = decoder.contig_of_tid();
            Lane& L = lanes[0];
            auto submit_front = [&]() {
                const int how = queue.front().st.how.get();          // the packer writes into this element: wait before it moves
                Piece pc = std::move(queue.front());
                queue.pop_front();
                pbam->recycle(*pc.segs);
                unsigned char* buf = (unsigned char*)L.pinned[pc.st.slot];
                int r2;
                if (how == 2) { r2 = be.submit_compact(L.ctx, buf, pc.st.compact_bytes); L.wire_bytes += pc.st.compact_bytes; }
                else if (how == 1) { r2 = be.submit_packed(L.ctx, (const rsq_read_packed*)buf, pc.st.n_reads, (const rsq_block*)(buf + pc.st.blocks_off), pc.st.n_packed_blocks); L.wire_bytes += 32 * pc.st.n_reads + 8 * pc.st.n_packed_blocks; }
                else { r2 = be.submit(L.ctx, (const rsq_read*)buf, pc.st.n_reads, (const rsq_block*)(buf + pc.st.blocks_off), pc.st.n_blocks); L.wire_bytes += 48 * pc.st.n_reads + 8 * pc.st.n_blocks; }
                L.in_flight[pc.st.slot] = true;
                L.records += pc.st.n_reads; ++L.batches;
                ++by_format[how];
                if (r2 != RSQ_OK) throw HostError(-1, std::string("rsq_submit failed: ") + be.last_error(L.ctx));
            };
            try {
                while (more) {
                    auto segs = std::make_shared<ParallelBamDecoder::Segments>();
                    if (!pbam->next_batch_segments(decoder, opt.batch_records, *segs)) break;
                    more = segs->n_reads == opt.batch_records;
                    while (queue.size() >= (size_t)Lane::kSlots) submit_front();     // its slot comes free
                    const size_t n = segs->n_reads, nb = segs->n_blocks;
                    const size_t blocks_off = (48 * n + 15) & ~(size_t)15;
                    const size_t need = std::max(rsq_compact_bound(n, nb), blocks_off + 8 * std::max<size_t>(nb, 1));
                    queue.emplace_back();
                    Piece& pc = queue.back();
                    pc.segs = segs;
                    pc.st.lane = 0; pc.st.slot = L.acquire(need); pc.st.n_reads = n; pc.st.n_blocks = nb; pc.st.blocks_off = blocks_off;
                    unsigned char* dst = (unsigned char*)L.pinned[pc.st.slot];
                    const size_t cap = L.cap[pc.st.slot];
                    const ParallelBamDecoder::Segments* sg = segs.get();
                    const int16_t* c_of_t = contig_of_tid.data();
                    const size_t n_tids = contig_of_tid.size();
                    const bool try_packed = opt.packed_submit && be.submit_packed;
                    size_t* compact_bytes = &pc.st.compact_bytes;      // (deque: elements stay where they are when others come and go)
                    size_t* n_packed = &pc.st.n_packed_blocks;
                    pc.st.how = std::async(std::launch::async, [=]() -> int {
                        std::vector<uint32_t> scratch(131072);
                        if (rsq_pack_compact_segments(sg->segs.data(), sg->segs.size(), c_of_t, n_tids, dst, cap, scratch.data(), compact_bytes) == RSQ_OK) return 2;
                        // not this batch: copy it together (contig ids bound) and try the other formats
                        std::vector<rsq_read> flat;
                        std::vector<rsq_block> fblocks;
                        flat.reserve(n); fblocks.reserve(nb);
                        for (const rsq_segment& S : sg->segs) {
                            // the blocks a piece's records point at (a piece of a job: all of the job's)
                            const size_t base = fblocks.size();
                            fblocks.insert(fblocks.end(), S.blocks, S.blocks + S.n_blocks);
                            for (size_t k = 0; k < S.n_reads; ++k) {
                                rsq_read r = S.reads[k];
                                if (S.tids) { const int32_t tid = S.tids[k]; r.contig = (tid >= 0 && (size_t)tid < n_tids) ? c_of_t[tid] : (int16_t)-1; }
                                r.block_off += (uint32_t)base;
                                flat.push_back(r);
                            }
                        }
                        if (try_packed && blocks_off + 8 * fblocks.size() <= cap &&
                            rsq_pack_records(flat.data(), n, fblocks.data(), fblocks.size(), (rsq_read_packed*)dst, (rsq_block*)(dst + blocks_off), n_packed) == RSQ_OK) return 1;
                        // plain records: the blocks in record order (what the batch's block count was sized for)
                        rsq_read* out = (rsq_read*)dst;
                        rsq_block* ob = (rsq_block*)(dst + blocks_off);
                        size_t w = 0;
                        for (size_t i = 0; i < n; ++i) {
                            out[i] = flat[i];
                            out[i].block_off = (uint32_t)w;
                            for (unsigned b = 0; b < flat[i].n_blocks; ++b) ob[w++] = fblocks[flat[i].block_off + b];
                        }
                        return 0;
                    });
                    if (opt.verbosity > 1)
                        std::cout << "Time elapsed: " << std::chrono::duration<double>(clk::now() - t1).count()
                                  << "; Alignments processed: " << decoder.alignment_count() << std::endl;
                }
                while (!queue.empty()) submit_front();
                if (opt.verbosity)
                    std::cout << "Batches read in place from the ingest workers' buffers: " << by_format[2] << " in the 16-byte wire format, "
                              << by_format[1] << " in the 32-byte one, " << by_format[0] << " as plain records" << std::endl;
            } catch (...) {
                for (Piece& pc : queue) if (pc.st.how.valid()) pc.st.how.wait();   // the packers still read the jobs and write the slots
                throw;
            }
            more = false;
        }
        try {
        while (more) {
            std::vector<rsq_read>& reads = reads_buf[cur];
            std::vector<rsq_block>& blocks = blocks_buf[cur];
            if (pbam) {
                if (!pbam->next_batch(decoder, opt.batch_records, reads, blocks)) break;
                more = reads.size() == opt.batch_records;
            } else {
                reads.clear();
                blocks.clear();
                while (reads.size() < opt.batch_records) {
                    if (!reader->next(aln)) { more = false; break; }
                    decoder.convert(aln, reads, blocks);
                }
            }
            if (reads.empty()) break;
            submit_staged();
            staged.reserve(n_lanes + 1);           // `stage` keeps pointers into the vector
            if (n_lanes > 1 && decoder.saw_unsorted())
                throw HostError(6, "--gpus needs a coordinate-sorted alignment file: records are range-partitioned by coordinate");
            // split the batch where the stream crosses the pending cut (at most one crossing per lane)
            size_t i = 0;
            const size_t n = reads.size();
            while (i < n) {
                if (g + 1 >= n_lanes || !have_pending) { stage(g, reads, blocks, i, n); break; }
                size_t j = i;
                StreamCut k;
                while (j < n && !(key_of(reads[j], k) && !before(k, pending))) ++j;
                stage(g, reads, blocks, i, j);
                if (j == n) break;
                cuts.push_back(pending);
                have_pending = false;
                ++g;
                i = j;
            }
            if (g + 1 < n_lanes && !have_pending) {
                uint64_t on_lane = lanes[g].records;
                for (const Staged& st : staged) if ((size_t)st.lane == g) on_lane += st.n_reads;
                if (on_lane >= share()) {
                    // the cut: the first island-free coordinate behind the last record of this batch
                    StreamCut k;
                    size_t j = n;
                    while (j > 0 && !key_of(reads[j - 1], k)) --j;
                    if (j > 0) {
                        pending.contig_rank = k.contig_rank;
                        pending.pos1 = islands.next_cut((uint32_t)reads[j - 1].contig, k.pos1 + 1);
                        have_pending = true;
                    }
                }
            }
            cur ^= 1;
            if (opt.verbosity > 1)
                std::cout << "Time elapsed: " << std::chrono::duration<double>(clk::now() - t1).count()
                          << "; Alignments processed: " << decoder.alignment_count() << std::endl;
        }
        submit_staged();
        } catch (...) {
            for (Staged& st : staged) if (st.how.valid()) st.how.wait();   // the packers still read the batch buffers
            throw;
        }
        for (Lane& L : lanes) {
            int rc = be.sync(L.ctx);
            if (rc != RSQ_OK) throw HostError(-1, std::string("rsq_sync failed: ") + be.last_error(L.ctx));
        }
        if (opt.verbosity) {
            const double dt = std::chrono::duration<double>(clk::now() - t1).count();
            std::cout << "Time Elapsed: " << dt << "; Alignments processed: " << decoder.alignment_count() << std::endl;
            if (opt.verbosity > 1) std::cout << "Average Reads/Sec: " << (double)decoder.alignment_count() / dt << std::endl;
            if (n_lanes > 1)
                for (size_t k = 0; k < n_lanes; ++k)
                    std::cout << "GPU " << lanes[k].device << ": " << lanes[k].records << " records in " << lanes[k].batches << " batches, "
                              << lanes[k].wire_bytes << " bytes over the wire" << std::endl;
            std::cout << "Estimating library complexity..." << std::endl;
        }
        auto t_fin = clk::now();
        HostResults R;
        if (n_lanes == 1) {
            fetch_results(be, lanes[0].ctx, A, false, R);            // counts only: the per-base coverage stays on the device
            fetch_coverage_statistics(be, lanes[0].ctx, A, opt.run, R);
        } else {
            // ---- end-of-run exchange (include/rsq_b200.h): afterwards every context holds the single-GPU state -------------
            if (nccl) {
                std::vector<std::future<int>> xs;
                for (Lane& L : lanes) xs.push_back(std::async(std::launch::async, [&be, ctx = L.ctx] { return be.exchange(ctx); }));
                int bad = RSQ_OK;
                for (auto& f : xs) { const int r = f.get(); if (r != RSQ_OK) bad = r; }
                if (bad != RSQ_OK) throw HostError(-1, std::string("rsq_exchange failed: ") + be.last_error(lanes[0].ctx));
            } else {
                auto reduce = [&](decltype(be.export_counts) exporter) {
                    std::vector<uint64_t*> ptr(n_lanes);
                    size_t words = 0;
                    for (size_t k = 0; k < n_lanes; ++k)
                        if (exporter(lanes[k].ctx, &ptr[k], &words) != RSQ_OK) throw HostError(-1, std::string("exchange export failed: ") + be.last_error(lanes[k].ctx));
                    for (size_t w = 0; w < words; ++w) {
                        uint64_t sum = 0;
                        for (size_t k = 0; k < n_lanes; ++k) sum += ptr[k][w];
                        for (size_t k = 0; k < n_lanes; ++k) ptr[k][w] = sum;
                    }
                };
                reduce(be.export_sample_counts);
                reduce(be.export_counts);
                for (Lane& L : lanes) if (be.import_counts(L.ctx) != RSQ_OK) throw HostError(-1, "rsq_import_counts failed");
            }
            fetch_results(be, lanes[0].ctx, A, false, R);
            // per-gene coverage statistics: every GPU evaluates the genes whose reads it holds; gathered as bitwise sums
            for (size_t k = 0; k < n_lanes; ++k) gene_ownership(A, cuts, contig_rank, (uint32_t)k, lanes[k].owned);
            std::vector<HostResults> parts(n_lanes);
            std::vector<std::future<std::string>> cs;
            for (size_t k = 0; k < n_lanes; ++k)
                cs.push_back(std::async(std::launch::async, [&, k]() -> std::string {
                    try { fetch_coverage_statistics(be, lanes[k].ctx, A, opt.run, parts[k], &lanes[k].owned); }
                    catch (HostError& e) { return std::string(e.what()); }
                    return std::string();
                }));
            std::string why;
            for (auto& f : cs) { const std::string w = f.get(); if (why.empty()) why = w; }
            if (!why.empty()) throw HostError(-1, why);
            if (!nccl) {               // host transport: sum the bit patterns
                auto fold = [&](auto member) {
                    auto& dst = parts[0].*member;
                    for (size_t k = 1; k < n_lanes; ++k) {
                        auto& src = parts[k].*member;
                        for (size_t i2 = 0; i2 < dst.size(); ++i2) {
                            uint64_t a2, b2;
                            memcpy(&a2, &dst[i2], 8); memcpy(&b2, &src[i2], 8);
                            a2 += b2;
                            memcpy(&dst[i2], &a2, 8);
                        }
                    }
                };
                fold(&HostResults::cs_mean); fold(&HostResults::cs_std); fold(&HostResults::cs_cv); fold(&HostResults::cs_exon_cv);
                fold(&HostResults::cs_five); fold(&HostResults::cs_three);
            }
            R.cs_mean = parts[0].cs_mean; R.cs_std = parts[0].cs_std; R.cs_cv = parts[0].cs_cv; R.cs_exon_cv = parts[0].cs_exon_cv;
            R.cs_five = parts[0].cs_five; R.cs_three = parts[0].cs_three;
            R.have_cov_stats = true;
        }
        StreamState st;
        st.read_length = decoder.read_length();
        st.alignment_count = decoder.alignment_count();
        st.dropped_order = decoder.dropped_order();
        st.final_contig = decoder.current_contig();
        st.final_position = decoder.trim_position();
        if (opt.verbosity) std::cout << "Finalize + coverage statistics: " << std::chrono::duration<double>(clk::now() - t_fin).count() << " seconds" << std::endl;
        if (opt.verbosity) std::cout << "Generating report" << std::endl;
        auto t_rep = clk::now();
        write_reports(A, R, opt.run, st, ref.open ? &ref : nullptr);
        if (opt.verbosity) std::cout << "Report files: " << std::chrono::duration<double>(clk::now() - t_rep).count() << " seconds" << std::endl;
        if (opt.exit_after_reports && !getenv("RSQ_FULL_TEARDOWN")) {
            for (Staged& st : staged) if (st.how.valid()) st.how.wait();
            if (getenv("RSQ_INGEST_STATS")) pbam.reset();          // its stage times are printed when it goes
            std::cout.flush(); std::cerr.flush(); std::fflush(nullptr);
            std::_Exit(0);
        }
    } catch (HostError& e) {
        std::cerr << e.what() << std::endl;
        exit_code = e.exit_code;
    } catch (std::bad_alloc& e) {
        std::cerr << "Memory allocation failure. Out of memory" << std::endl << e.what() << std::endl;
        exit_code = 10;
    } catch (std::exception& e) {
        std::cerr << "Unknown error" << std::endl << e.what() << std::endl;
        exit_code = -1;
    }
    for (Staged& st : staged) if (st.how.valid()) st.how.wait();
    for (Lane& L : lanes) {
        if (L.creating.valid()) L.creating.wait();
        L.release();
    }
    return exit_code;
}

}  // namespace rsqh
