// pipeline.cpp — host driver (see pipeline.h).
#include "pipeline.h"

#include "bamparallel.h"

#include <chrono>
#include <cstring>
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
         "      --no-packed-submit                Send 48-byte records instead of the 32-byte wire format (this build)\n"
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

void fetch_coverage_statistics(const Backend& be, rsq_ctx* ctx, const Annotation& A, const RunOptions& opt, HostResults& R) {
    rsq_transcripts q = A.transcripts(opt.coverage_mask, opt.bias_offset, opt.bias_window, opt.bias_gene_length);
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

int run_sample(const Backend& be, CliOptions& opt) {
    using clk = std::chrono::steady_clock;
    rsq_ctx* ctx = nullptr;
    void* pinned[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    size_t pinned_cap[2][2] = {{0, 0}, {0, 0}};
    int exit_code = 0;
    // The device context comes up on a helper thread while the GTF is parsed (CUDA start-up costs about as much as the parse on
    // a multi-GPU box); its outcome is looked at where rsq_create used to be called, so errors keep their order of precedence.
    std::future<int> creating;
    try {
        creating = std::async(std::launch::async, [&be, &opt, &ctx] { return be.create(&opt.cfg, &ctx); });
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
        std::unique_ptr<ParallelBamDecoder> pbam;
        {
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
        int rc = creating.get();
        if (rc != RSQ_OK) throw HostError(rc == RSQ_ERR_NO_DEVICE ? 10 : -1, std::string("rsq_create failed: ") + be.last_error(nullptr));
        rsq_annotation av = A.view();
        rc = be.set_annotation(ctx, &av);
        if (rc != RSQ_OK) throw HostError(-1, std::string("rsq_set_annotation failed: ") + be.last_error(ctx));
        if (!opt.bed.empty()) {
            rsq_bed bv = bed.view();
            rc = be.set_bed(ctx, &bv);
            if (rc != RSQ_OK) throw HostError(-1, std::string("rsq_set_bed failed: ") + be.last_error(ctx));
        }
        if (ref.open) {
            rsq_reference rv = ref.view();
            rc = be.set_reference(ctx, &rv);
            if (rc != RSQ_OK) throw HostError(-1, std::string("rsq_set_reference failed: ") + be.last_error(ctx));
        }

        if (opt.verbosity) std::cout << "Device context + annotation upload: " << std::chrono::duration<double>(clk::now() - t_setup).count() << " seconds" << std::endl;
        if (opt.verbosity) std::cout << "Parsing bam..." << std::endl;
        auto t1 = clk::now();
        // Two batch buffers: while a helper thread packs batch i into its page-locked slot (the wire format, or the plain copy
        // when the batch does not fit its rules), this thread already collects batch i+1 from the decoder; submits stay on this
        // thread, in order.  (Packing was a third of this thread's time per record once the decoder's framing stopped being
        // the sequential bottleneck.)
        std::vector<rsq_read> reads_buf[2];
        std::vector<rsq_block> blocks_buf[2];
        reads_buf[0].reserve(opt.batch_records);
        reads_buf[1].reserve(opt.batch_records);
        RawAln aln;
        struct Staged { bool active = false; int slot = 0; size_t n_reads = 0, n_blocks = 0, n_packed_blocks = 0; std::future<int> packed; } staged;
        auto submit_staged = [&]() {            // waits for the packer of the previous batch, then submits it
            if (!staged.active) return;
            staged.active = false;
            const int how = staged.packed.get();
            const int s = staged.slot;
            int r2;
            if (how == 1) r2 = be.submit_packed(ctx, (const rsq_read_packed*)pinned[s][0], staged.n_reads, (const rsq_block*)pinned[s][1], staged.n_packed_blocks);
            else r2 = be.submit(ctx, (const rsq_read*)pinned[s][0], staged.n_reads, (const rsq_block*)pinned[s][1], staged.n_blocks);
            if (r2 != RSQ_OK) throw HostError(-1, std::string("rsq_submit failed: ") + be.last_error(ctx));
        };
        int slot = 0, cur = 0;
        bool more = true;
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
            // stage into page-locked buffers; a slot is reused only after the device drained it
            const size_t need[2] = {reads.size() * sizeof(rsq_read), std::max<size_t>(blocks.size(), 1) * sizeof(rsq_block)};
            if (pinned_cap[slot][0]) { rc = be.sync(ctx); if (rc != RSQ_OK) throw HostError(-1, std::string("rsq_sync failed: ") + be.last_error(ctx)); }
            for (int k = 0; k < 2; ++k)
                if (pinned_cap[slot][k] < need[k]) {
                    if (pinned[slot][k]) be.host_free(pinned[slot][k]);
                    pinned_cap[slot][k] = need[k] + need[k] / 4;
                    pinned[slot][k] = be.host_alloc(pinned_cap[slot][k]);
                    if (!pinned[slot][k]) throw HostError(10, "Memory allocation failure. Out of memory");
                }
            // the compact wire format (32 B per record, no entry for a whole-read block) when the batch fits its rules; the
            // packer writes straight into the pinned buffers, so it replaces the staging memcpy
            staged.active = true; staged.slot = slot; staged.n_reads = reads.size(); staged.n_blocks = blocks.size(); staged.n_packed_blocks = 0;
            const bool try_pack = opt.packed_submit && be.submit_packed;
            void* dst_reads = pinned[slot][0];
            void* dst_blocks = pinned[slot][1];
            size_t* n_packed = &staged.n_packed_blocks;
            const std::vector<rsq_read>* src_reads = &reads;
            const std::vector<rsq_block>* src_blocks = &blocks;
            staged.packed = std::async(std::launch::async, [=]() -> int {
                if (try_pack && rsq_pack_records(src_reads->data(), src_reads->size(), src_blocks->data(), src_blocks->size(),
                                                 (rsq_read_packed*)dst_reads, (rsq_block*)dst_blocks, n_packed) == RSQ_OK) return 1;
                memcpy(dst_reads, src_reads->data(), src_reads->size() * sizeof(rsq_read));
                memcpy(dst_blocks, src_blocks->data(), src_blocks->size() * sizeof(rsq_block));
                return 0;
            });
            slot ^= 1;
            cur ^= 1;
            if (opt.verbosity > 1)
                std::cout << "Time elapsed: " << std::chrono::duration<double>(clk::now() - t1).count()
                          << "; Alignments processed: " << decoder.alignment_count() << std::endl;
        }
        submit_staged();
        } catch (...) {
            if (staged.active && staged.packed.valid()) staged.packed.wait();   // the packer still reads the batch buffers
            throw;
        }
        rc = be.sync(ctx);
        if (rc != RSQ_OK) throw HostError(-1, std::string("rsq_sync failed: ") + be.last_error(ctx));
        if (opt.verbosity) {
            const double dt = std::chrono::duration<double>(clk::now() - t1).count();
            std::cout << "Time Elapsed: " << dt << "; Alignments processed: " << decoder.alignment_count() << std::endl;
            if (opt.verbosity > 1) std::cout << "Average Reads/Sec: " << (double)decoder.alignment_count() / dt << std::endl;
            std::cout << "Estimating library complexity..." << std::endl;
        }
        auto t_fin = clk::now();
        HostResults R;
        fetch_results(be, ctx, A, false, R);                     // counts only: the per-base coverage stays on the device
        fetch_coverage_statistics(be, ctx, A, opt.run, R);
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
    if (creating.valid()) creating.wait();
    for (int s = 0; s < 2; ++s) for (int k = 0; k < 2; ++k) if (pinned[s][k]) be.host_free(pinned[s][k]);
    if (ctx) be.destroy(ctx);
    return exit_code;
}

}  // namespace rsqh
