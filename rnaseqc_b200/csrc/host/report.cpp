// report.cpp — end-of-run statistics and output files (see report.h).
//
// Everything here runs once per sample on a few MB of counts (plus the per-base coverage array for the
// per-gene statistics) and deliberately keeps the reference's sequential fp64 evaluation order so that
// the printed text matches.  It is written from the behaviour of src/RNASeQC.cpp:397-676 and
// src/Metrics.cpp:132-337,342-412; file:line comments give the statement being reproduced.
#include "report.h"

#include <algorithm>
#include <chrono>
#include <charconv>
#include <climits>
#include <cmath>
#include <filesystem>
#include <fstream>
#include <future>
#include <iostream>
#include <map>
#include <sstream>

namespace rsqh {

double exon_count_to_double(uint64_t whole, uint64_t frac) {
    // whole + frac * 2^-64, rounded once
    long double v = (long double)whole + std::ldexp((long double)frac, -64);
    return (double)v;
}

namespace {

const double kMadFactor = 1.4826;  // src/RNASeQC.cpp:26

// computeMedian (src/Metrics.h:147-160): NOTE the inverted parity — an ODD size averages two elements,
// an EVEN size returns element (size-1)/2.
template <typename Get>
double median_quirk(unsigned long size, Get get) {
    if (size == 0) throw HostError(2, "Invalid range\nCannot compute median of an empty list");
    if (size == 1) return (double)get(0);
    const unsigned long mid = (size - 1) / 2;
    if (size % 2) return ((double)get(mid) + (double)get(mid + 1)) / 2.0;
    return (double)get(mid);
}

struct Stats { double avg, med, std, mad; };

// getStatistics (src/Metrics.h:166-186); sorts `data` in place like the reference does.
Stats get_statistics(std::vector<double>& data) {
    if (data.empty()) return Stats{NAN, NAN, NAN, NAN};
    double avg = 0.0, sd = 0.0;
    std::vector<double> deviations;
    std::sort(data.begin(), data.end());
    const double size = (double)data.size();
    const double median = median_quirk((unsigned long)size, [&](unsigned long i) { return data[i]; });
    for (double x : data) {
        avg += x / size;
        deviations.push_back(std::fabs(x - median));
    }
    std::sort(deviations.begin(), deviations.end());
    const double med_dev = median_quirk(deviations.size(), [&](unsigned long i) { return deviations[i]; }) * kMadFactor;
    for (double x : data) sd += std::pow(x - avg, 2.0) / size;
    sd = std::pow(sd, 0.5);
    return Stats{avg, median, sd, med_dev};
}

// `stream << v` with the stream's defaults is printf's %g (precision 6); std::to_chars(general, 6) prints the same digits without
// the stream machinery (the report writer formats about a million numbers); non-finite values keep the stream's spelling
std::string fmt_default(double v) {
    if (std::isfinite(v)) {
        char num[64];
        const auto r = std::to_chars(num, num + sizeof num, v, std::chars_format::general, 6);
        return std::string(num, r.ptr);
    }
    std::ostringstream s;
    s << v;
    return s.str();
}

}  // namespace

void compute_coverage_summary(const Annotation& A, const HostResults& R, const RunOptions& opt,
                              const StreamState& st, CoverageSummary& out) {
    const size_t ng = A.gene_ids.size();
    out.five_end.assign(ng, 0);
    out.three_end.assign(ng, 0);

    // ---- eviction order -------------------------------------------------------------------
    // trimFeatures evicts from the FRONT of the start-sorted list (src/Expression.cpp:80-93) and dropFeatures
    // walks the rest in list order (:96-103), so within a contig the order is always list order.  Contigs come
    // in the order their lists were dropped (chr switches, src/RNASeQC.cpp:346-349); the genes of the final
    // contig that were already trimmed come before the end-of-run sweep over all contigs in id order (:385-386).
    std::vector<const GeneFeature*> order;
    std::vector<uint8_t> done(A.genes_by_contig.size(), 0);
    for (uint32_t c : st.dropped_order) {
        if (c >= A.genes_by_contig.size() || done[c]) continue;
        // the final contig may ALSO appear here if the stream left and came back; its list is then empty
        for (const GeneFeature& g : A.genes_by_contig[c]) order.push_back(&g);
        done[c] = 1;
    }
    std::vector<uint8_t> gene_emitted;
    if (st.final_contig >= 0 && (size_t)st.final_contig < A.genes_by_contig.size() && !done[st.final_contig]) {
        // features popped so far on the final contig: the longest prefix of the combined (gene+exon) start-sorted
        // list whose every element has end < pos of the last read that reached trimFeatures.
        const auto& genes = A.genes_by_contig[st.final_contig];
        const auto& exons = A.exons_by_contig[st.final_contig];
        gene_emitted.assign(genes.size(), 0);
        size_t gi = 0, ei = 0;
        const int64_t P = st.final_position;
        while (gi < genes.size() || ei < exons.size()) {
            // merge by (start, GTF line order) == the stable sort of the interleaved list
            bool take_gene = ei >= exons.size() ||
                             (gi < genes.size() && (genes[gi].start < exons[ei].start ||
                                                    (genes[gi].start == exons[ei].start && genes[gi].line_order < exons[ei].line_order)));
            int64_t end = take_gene ? genes[gi].end : exons[ei].end;
            if (!(end < P)) break;
            if (take_gene) { order.push_back(&genes[gi]); gene_emitted[gi] = 1; ++gi; }
            else ++ei;
        }
    }
    for (size_t c = 0; c < A.genes_by_contig.size(); ++c) {
        if (done[c]) continue;
        const auto& genes = A.genes_by_contig[c];
        for (size_t i = 0; i < genes.size(); ++i) {
            if ((int64_t)c == st.final_contig && !gene_emitted.empty() && gene_emitted[i]) continue;
            order.push_back(&genes[i]);
        }
    }

    // The statistics themselves (BaseCoverage::compute / computeCoverage / computeBias, src/Metrics.cpp:132-337) come from
    // rsq_coverage_statistics — the device kernel rsq_covstats.cu in the product; only the ordering and the text are left.
    if (!R.have_cov_stats) throw HostError(-1, "coverage statistics missing: rsq_coverage_statistics must run before the report writer");
    for (const GeneFeature* gf : order) {
        const uint32_t g = gf->gene, fi = gf->flat;
        out.five_end[g] = (uint64_t)((double)out.five_end[g] + (double)R.cs_five[fi]);
        out.three_end[g] = (uint64_t)((double)out.three_end[g] + (double)R.cs_three[fi]);
        std::string row = A.gene_ids[g] + "\t";
        if (R.cs_mean[fi] != -1.0 || R.cs_std[fi] != -1.0) {
            row += fmt_default(R.cs_mean[fi]) + "\t" + fmt_default(R.cs_std[fi]) + "\t" + fmt_default(R.cs_cv[fi]);
            out.gene_means.push_back(R.cs_mean[fi]);
            out.gene_stds.push_back(R.cs_std[fi]);
            out.gene_cvs.push_back(R.cs_cv[fi]);
        } else row += "0\t0\tnan";
        out.rows.push_back(row);
    }
    for (size_t s = 0; s < R.cs_exon_cv.size(); ++s)
        if (!std::isnan(R.cs_exon_cv[s])) out.exon_cvs.emplace_back(A.exon_ids[s], R.cs_exon_cv[s]);
}

namespace {

// Lander-Waterman scan (src/RNASeQC.cpp:398-415).  The literal loop runs up to 1e9 exp() calls; the estimate
// x*(1-exp(-n/x)) is increasing in x with slope < 1, so the first minimum of the truncated error can be
// bracketed by bisection and then found by replaying the reference loop over a window around it.
unsigned library_complexity(double duplicates, double unique, bool exact) {
    if (!(duplicates > 0)) return 0u;
    const double num_reads = duplicates + unique;
    auto estimate = [&](double x) { return x * (1.0 - std::exp(-1.0 * num_reads / x)); };
    auto scan = [&](double from, double to, unsigned& min_reads, unsigned& min_error) {
        for (double x = from; x < to; ++x) {
            unsigned error = (unsigned)std::fabs(estimate(x) - unique);
            if (error < min_error) { min_error = error; min_reads = (unsigned)x; }
        }
    };
    unsigned min_reads = 0u, min_error = UINT_MAX;
    if (exact || unique < 1.0) {
        scan(unique, 1e9, min_reads, min_error);
        return min_reads;
    }
    // steps k = 0 .. K-1 with x = unique + k < 1e9
    const double K = std::ceil(1e9 - unique);
    if (K <= 0) return 0u;
    const double W = 1 << 20;
    if (K <= 4 * W) { scan(unique, 1e9, min_reads, min_error); return min_reads; }
    // target error value: the smallest truncated error reachable = 0 if the estimate crosses unique-1 before 1e9,
    // else the truncated error at the last step.
    const double x_last = unique + (K - 1);
    double target;  // first x with estimate(x) > unique - target is the answer (monotone part, estimate <= unique)
    if (estimate(x_last) > unique - 1.0) target = 1.0;
    else target = std::floor(unique - estimate(x_last)) + 1.0;
    double lo = 0, hi = K - 1;  // invariant: predicate false at lo-1 (or lo==0), true at hi
    while (lo < hi) {
        double mid = std::floor((lo + hi) / 2);
        if (estimate(unique + mid) > unique - target) hi = mid; else lo = mid + 1;
    }
    // replay the literal loop on [lo-W, lo+W); seed min_error with the error just before the window so that
    // "first strictly smaller" behaves as in the full scan.
    double from = std::max(0.0, lo - W), to = std::min(K, lo + W);
    if (from > 0) {
        // errors before the window are >= the error at from-1 (monotone decreasing up to rounding noise)
        min_error = (unsigned)std::fabs(estimate(unique + from - 1) - unique);
        min_reads = (unsigned)(unique + from - 1);
        // walk back to the first x with that same truncated error is unnecessary: a strictly smaller error
        // exists inside the window by construction, so this seed is overwritten.
    }
    scan(unique + from, unique + to, min_reads, min_error);
    return min_reads;
}

struct CounterName { int idx; const char* name; };
// dump order of operator<<(ofstream&, Metrics&) (src/Metrics.cpp:344-380)
const CounterName kDumpOrder[] = {
    {RSQ_C_END1_ANTISENSE, "End 1 Antisense"}, {RSQ_C_END2_ANTISENSE, "End 2 Antisense"},
    {RSQ_C_END1_BASES, "End 1 Bases"}, {RSQ_C_END2_BASES, "End 2 Bases"},
    {RSQ_C_END1_MAPPED_READS, "End 1 Mapped Reads"}, {RSQ_C_END2_MAPPED_READS, "End 2 Mapped Reads"},
    {RSQ_C_END1_MISMATCHES, "End 1 Mismatches"}, {RSQ_C_END2_MISMATCHES, "End 2 Mismatches"},
    {RSQ_C_END1_SENSE, "End 1 Sense"}, {RSQ_C_END2_SENSE, "End 2 Sense"},
    {RSQ_C_EXONIC_READS, "Exonic Reads"}, {RSQ_C_FAILED_VENDOR_QC, "Failed Vendor QC"},
    {RSQ_C_HIGH_QUALITY_READS, "High Quality Reads"}, {RSQ_C_INTERGENIC_READS, "Intergenic Reads"},
    {RSQ_C_INTRAGENIC_READS, "Intragenic Reads"}, {RSQ_C_AMBIGUOUS_READS, "Ambiguous Reads"},
    {RSQ_C_INTRONIC_READS, "Intronic Reads"}, {RSQ_C_LOW_MAPPING_QUALITY, "Low Mapping Quality"},
    {RSQ_C_LOW_QUALITY_READS, "Low Quality Reads"}, {RSQ_C_MAPPED_DUPLICATE_READS, "Mapped Duplicate Reads"},
    {RSQ_C_MAPPED_READS, "Mapped Reads"}, {RSQ_C_MAPPED_UNIQUE_READS, "Mapped Unique Reads"},
    {RSQ_C_MISMATCHED_BASES, "Mismatched Bases"}, {RSQ_C_NON_GLOBIN_READS, "Non-Globin Reads"},
    {RSQ_C_NON_GLOBIN_DUPLICATE_READS, "Non-Globin Duplicate Reads"},
    {RSQ_C_READS_USED, "Reads used for Intron/Exon counts"}, {RSQ_C_RRNA_READS, "rRNA Reads"},
    {RSQ_C_SPLIT_READS, "Split Reads"}, {RSQ_C_TOTAL_BASES, "Total Bases"},
    {RSQ_C_TOTAL_MAPPED_PAIRS, "Total Mapped Pairs"},
    {RSQ_C_UMVQ, "Unique Mapping, Vendor QC Passed Reads"}, {RSQ_C_UNPAIRED_READS, "Unpaired Reads"},
};

}  // namespace

void write_reports(const Annotation& A, const HostResults& R, const RunOptions& opt, const StreamState& st, const ReferenceModel* ref) {
    namespace fs = std::filesystem;
    try {
        if (!fs::exists(opt.output_dir)) fs::create_directories(opt.output_dir);   // src/RNASeQC.cpp:190-193
    } catch (fs::filesystem_error& e) {
        throw HostError(8, std::string("Filesystem error:  ") + e.what());
    }
    const std::string base = opt.output_dir + "/" + opt.sample_name;
    const bool setup_stats = getenv("RSQ_SETUP_STATS") != nullptr;          // stage times of the report writer on stderr
    auto t_mark = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!setup_stats) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "write_reports: %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_mark).count());
        t_mark = now;
    };
    uint64_t C[RSQ_N_COUNTERS];
    for (int i = 0; i < RSQ_N_COUNTERS; ++i) C[i] = R.counters[i];
    // counter.increment("Total Alignments", alignmentCount): the int parameter truncates (src/RNASeQC.cpp:397,
    // src/Metrics.cpp:33-36)
    const uint64_t total_alignments = (uint64_t)(int64_t)(int)st.alignment_count;
    auto frac = [&](int a, int b) { return (double)C[a] / C[b]; };                   // Metrics::frac, src/Metrics.cpp:43-46
    // ---- exon table (src/RNASeQC.cpp:510-521): the largest file and independent of everything below — written on a helper thread
    std::future<void> exon_table = std::async(std::launch::async, [&A, &R, &opt, base] {
        std::ofstream exon_report(base + ".exon_reads.gct");
        size_t nonzero = 0;
        for (size_t e = 0; e < A.exon_ids.size(); ++e) if (R.exon_whole[e] || R.exon_frac[e]) ++nonzero;
        exon_report << "#1.2" << '\n';
        exon_report << nonzero << "\t1" << '\n';                                  // exonCounts.size(), :513
        exon_report << "Name\tDescription\t" << (opt.sample_given ? opt.sample_name : "Counts") << '\n';
        // `<< std::fixed << v` is printf's %.6f; std::to_chars(fixed, 6) prints the same digits without the stream machinery
        std::string buf;
        buf.reserve(1 << 20);
        char num[400];
        for (size_t e = 0; e < A.exon_ids.size(); ++e) {
            const double v = R.exon_seq.empty() ? exon_count_to_double(R.exon_whole[e], R.exon_frac[e]) : R.exon_seq[e];
            buf += A.exon_ids[e]; buf += '\t'; buf += A.exon_desc[e]; buf += '\t';
            if (std::isfinite(v)) {
                const auto r = std::to_chars(num, num + sizeof num, v, std::chars_format::fixed, 6);
                buf.append(num, r.ptr);
            } else {
                std::ostringstream o; o << std::fixed << v; buf += o.str();
            }
            buf += '\n';
            if (buf.size() > (1 << 20) - 4096) { exon_report.write(buf.data(), (std::streamsize)buf.size()); buf.clear(); }
        }
        exon_report.write(buf.data(), (std::streamsize)buf.size());
    });
    struct Join { std::future<void>& f; ~Join() { if (f.valid()) f.wait(); } } join_exon_table{exon_table};   // also on the error paths
    // ---- coverage statistics (the reference computes them as genes leave the window) -------------------------
    CoverageSummary cov;
    compute_coverage_summary(A, R, opt, st, cov);
    lap("coverage summary");
    {
        std::ofstream w(opt.write_coverage ? base + ".coverage.tsv" : std::string("/dev/null"));
        if (!w.is_open() && opt.write_coverage) throw HostError(-1, "Unable to open BaseCoverage output file");
        w << "gene_id\tcoverage_mean\tcoverage_std\tcoverage_CV" << '\n';     // src/Metrics.h:113
        for (const std::string& r : cov.rows) w << r << '\n';
    }
    lap("coverage.tsv");
    const unsigned min_reads = library_complexity((double)C[RSQ_C_DUPLICATE_PAIRS], (double)C[RSQ_C_UNIQUE_FRAGMENTS],
                                                  opt.exact_complexity_scan);
    lap("library complexity");
    // ---- gene tables (src/RNASeQC.cpp:419-475) -----------------------------------------------------------------
    unsigned genes_detected = 0;
    unsigned bias_genes = 0;
    std::vector<double> ratios;
    {
        std::ofstream gene_report(base + ".gene_reads.gct");
        std::ofstream gene_rpkm(base + ".gene_" + (opt.use_rpkm ? "rpkm" : "tpm") + ".gct");
        std::ofstream fragment_report(base + ".gene_fragments.gct");
        gene_report << "#1.2" << '\n';
        gene_rpkm << "#1.2" << '\n';
        fragment_report << "#1.2" << '\n';
        gene_report << A.gene_list.size() << "\t1" << '\n';
        gene_rpkm << A.gene_list.size() << "\t1" << '\n';
        fragment_report << A.gene_list.size() << "\t1" << '\n';
        gene_report << "Name\tDescription\t" << (opt.sample_given ? opt.sample_name : "Counts") << '\n';
        gene_rpkm << "Name\tDescription\t" << (opt.sample_given ? opt.sample_name : (opt.use_rpkm ? "RPKM" : "TPM")) << '\n';
        gene_rpkm << std::fixed;
        fragment_report << "Name\tDescription\t" << (opt.sample_given ? opt.sample_name : "Fragments") << '\n';
        const double scale_rpkm = (double)C[RSQ_C_EXONIC_READS] / 1000000.0;
        double scale_tpm = 0.0;
        std::vector<double> tpms(A.gene_list.size(), 0.0);
        // rows are formatted into buffers (std::to_chars prints what `<<` and `<< std::fixed <<` print, without the stream machinery)
        std::string b_reads, b_frag, b_rpkm;
        b_reads.reserve(A.gene_list.size() * 48); b_frag.reserve(A.gene_list.size() * 48); b_rpkm.reserve(A.gene_list.size() * 56);
        char num[400];
        auto put_long = [&](std::string& b, long v) { const auto r = std::to_chars(num, num + sizeof num, v); b.append(num, r.ptr); };
        auto put_fixed = [&](std::string& b, double v) {
            if (std::isfinite(v)) { const auto r = std::to_chars(num, num + sizeof num, v, std::chars_format::fixed, 6); b.append(num, r.ptr); }
            else { std::ostringstream o; o << std::fixed << v; b += o.str(); }
        };
        for (size_t i = 0; i < A.gene_list.size(); ++i) {
            const uint32_t g = A.gene_list[i];
            const double count = (double)R.gene_reads[g];
            b_reads += A.gene_ids[g]; b_reads += '\t'; b_reads += A.gene_desc[g]; b_reads += '\t'; put_long(b_reads, (long)R.gene_reads[g]); b_reads += '\n';
            b_frag += A.gene_ids[g]; b_frag += '\t'; b_frag += A.gene_desc[g]; b_frag += '\t'; put_long(b_frag, (long)R.gene_fragments[g]); b_frag += '\n';
            if (opt.use_rpkm) {
                double rpkm = (1000.0 * count / scale_rpkm) / (double)A.gene_coding_len[g];
                b_rpkm += A.gene_ids[g]; b_rpkm += '\t'; b_rpkm += A.gene_desc[g]; b_rpkm += '\t'; put_fixed(b_rpkm, rpkm); b_rpkm += '\n';
            } else {
                double tpm = (1000.0 * count) / (double)A.gene_coding_len[g];
                tpms[i] = tpm;
                scale_tpm += tpm;
            }
            if ((double)R.gene_unique[g] >= opt.detection_threshold) ++genes_detected;   // :461
            // BiasCounter::getBias (src/Metrics.cpp:239-249)
            const double cov5 = (double)cov.five_end[g], cov3 = (double)cov.three_end[g];
            if (cov5 + cov3 > 0.0) { ++bias_genes; ratios.push_back(cov3 / (cov5 + cov3)); }
        }
        if (!opt.use_rpkm) {
            scale_tpm /= 1000000.0;
            for (size_t i = 0; i < A.gene_list.size(); ++i) {
                const uint32_t g = A.gene_list[i];
                b_rpkm += A.gene_ids[g]; b_rpkm += '\t'; b_rpkm += A.gene_desc[g]; b_rpkm += '\t'; put_fixed(b_rpkm, tpms[i] / scale_tpm); b_rpkm += '\n';
            }
        }
        gene_report.write(b_reads.data(), (std::streamsize)b_reads.size());
        fragment_report.write(b_frag.data(), (std::streamsize)b_frag.size());
        gene_rpkm.write(b_rpkm.data(), (std::streamsize)b_rpkm.size());
    }
    lap("gene tables");
    // ---- 3'/5' ratio statistics (src/RNASeQC.cpp:478-508) ------------------------------------------------------
    double ratio_avg = 0.0, ratio_mad = 0.0, ratio_median = 0.0, ratio_std = 0.0, ratio75 = 0.0, ratio25 = 0.0;
    if (ratios.size() > 1) {
        Stats s = get_statistics(ratios);  // sorts ratios in place, as the reference's template does
        ratio_avg = s.avg; ratio_median = s.med; ratio_std = s.std; ratio_mad = s.mad;
        auto at = [&](double idx) {  // ratios[static_cast<int>(idx)]; the reference reads past the end for n in {2,3,4}
            size_t k = (size_t)(int)idx;
            return k < ratios.size() ? ratios[k] : 0.0;
        };
        double index = .25 * ratios.size();
        if (index > std::floor(index)) { index = std::ceil(index); ratio25 = at(index); }
        else { index = std::ceil(index); ratio25 = (at(index) + at(index + 1)) / 2.0; }
        index = .75 * ratios.size();
        if (index > std::floor(index)) { index = std::ceil(index); ratio75 = at(index); }
        else { index = std::ceil(index); ratio75 = (at(index) + at(index + 1)) / 2.0; }
    }
    // ---- metrics.tsv (src/RNASeQC.cpp:523-564, src/Metrics.cpp:381-410) -----------------------------------------
    std::ofstream output(base + ".metrics.tsv");
    output << "Sample\t" << opt.sample_name << '\n';
    output << "Mapping Rate\t" << frac(RSQ_C_MAPPED_READS, RSQ_C_UMVQ) << '\n';
    output << "Unique Rate of Mapped\t" << frac(RSQ_C_MAPPED_UNIQUE_READS, RSQ_C_MAPPED_READS) << '\n';
    output << "Duplicate Rate of Mapped\t" << frac(RSQ_C_MAPPED_DUPLICATE_READS, RSQ_C_MAPPED_READS) << '\n';
    output << "Duplicate Rate of Mapped, excluding Globins\t" << frac(RSQ_C_NON_GLOBIN_DUPLICATE_READS, RSQ_C_NON_GLOBIN_READS) << '\n';
    output << "Base Mismatch\t" << frac(RSQ_C_MISMATCHED_BASES, RSQ_C_TOTAL_BASES) << '\n';
    output << "End 1 Mapping Rate\t" << 2.0 * frac(RSQ_C_END1_MAPPED_READS, RSQ_C_UMVQ) << '\n';
    output << "End 2 Mapping Rate\t" << 2.0 * frac(RSQ_C_END2_MAPPED_READS, RSQ_C_UMVQ) << '\n';
    output << "End 1 Mismatch Rate\t" << frac(RSQ_C_END1_MISMATCHES, RSQ_C_END1_BASES) << '\n';
    output << "End 2 Mismatch Rate\t" << frac(RSQ_C_END2_MISMATCHES, RSQ_C_END2_BASES) << '\n';
    output << "Expression Profiling Efficiency\t" << frac(RSQ_C_EXONIC_READS, RSQ_C_UMVQ) << '\n';
    output << "High Quality Rate\t" << frac(RSQ_C_HIGH_QUALITY_READS, RSQ_C_MAPPED_READS) << '\n';
    output << "Exonic Rate\t" << frac(RSQ_C_EXONIC_READS, RSQ_C_MAPPED_READS) << '\n';
    output << "Intronic Rate\t" << frac(RSQ_C_INTRONIC_READS, RSQ_C_MAPPED_READS) << '\n';
    output << "Intergenic Rate\t" << frac(RSQ_C_INTERGENIC_READS, RSQ_C_MAPPED_READS) << '\n';
    output << "Intragenic Rate\t" << frac(RSQ_C_INTRAGENIC_READS, RSQ_C_MAPPED_READS) << '\n';
    output << "Ambiguous Alignment Rate\t" << frac(RSQ_C_AMBIGUOUS_READS, RSQ_C_MAPPED_READS) << '\n';
    output << "High Quality Exonic Rate\t" << frac(RSQ_C_HQ_EXONIC_READS, RSQ_C_HIGH_QUALITY_READS) << '\n';
    output << "High Quality Intronic Rate\t" << frac(RSQ_C_HQ_INTRONIC_READS, RSQ_C_HIGH_QUALITY_READS) << '\n';
    output << "High Quality Intergenic Rate\t" << frac(RSQ_C_HQ_INTERGENIC_READS, RSQ_C_HIGH_QUALITY_READS) << '\n';
    output << "High Quality Intragenic Rate\t" << frac(RSQ_C_HQ_INTRAGENIC_READS, RSQ_C_HIGH_QUALITY_READS) << '\n';
    output << "High Quality Ambiguous Alignment Rate\t" << frac(RSQ_C_HQ_AMBIGUOUS_READS, RSQ_C_HIGH_QUALITY_READS) << '\n';
    output << "Discard Rate\t" << (double)(C[RSQ_C_MAPPED_READS] - C[RSQ_C_READS_USED]) / C[RSQ_C_MAPPED_READS] << '\n';
    output << "rRNA Rate\t" << frac(RSQ_C_RRNA_READS, RSQ_C_MAPPED_READS) << '\n';
    output << "End 1 Sense Rate\t" << (double)C[RSQ_C_END1_SENSE] / (C[RSQ_C_END1_SENSE] + C[RSQ_C_END1_ANTISENSE]) << '\n';
    output << "End 2 Sense Rate\t" << (double)C[RSQ_C_END2_SENSE] / (C[RSQ_C_END2_SENSE] + C[RSQ_C_END2_ANTISENSE]) << '\n';
    output << "Avg. Splits per Read\t" << frac(RSQ_C_ALIGNMENT_BLOCKS, RSQ_C_MAPPED_READS) - 1.0 << '\n';
    // operator<<(ofstream&, Metrics&)
    output << "Total Alignments\t" << total_alignments << '\n';
    output << "Alternative Alignments\t" << C[RSQ_C_ALTERNATIVE_ALIGNMENTS] << '\n';
    output << "Supplementary Alignments\t" << C[RSQ_C_SUPPLEMENTARY_ALIGNMENTS] << '\n';
    output << "Total Reads\t" << total_alignments - C[RSQ_C_ALTERNATIVE_ALIGNMENTS] - C[RSQ_C_SUPPLEMENTARY_ALIGNMENTS] << '\n';
    output << "Chimeric Fragments\t";
    if (C[RSQ_C_CHIMERIC_TAG]) {
        output << C[RSQ_C_CHIMERIC_TAG] << '\n';
        output << "Chimeric Alignment Rate\t" << frac(RSQ_C_CHIMERIC_TAG, RSQ_C_TOTAL_MAPPED_PAIRS) << '\n';
    } else {
        output << C[RSQ_C_CHIMERIC_AUTO] << '\n';
        output << "Chimeric Alignment Rate\t" << frac(RSQ_C_CHIMERIC_AUTO, RSQ_C_TOTAL_MAPPED_PAIRS) << '\n';
    }
    for (const CounterName& k : kDumpOrder)
        if (k.idx != RSQ_C_SPLIT_READS || C[RSQ_C_SPLIT_READS]) output << k.name << "\t" << C[k.idx] << '\n';
    {
        // "Filtered by tag: X" rows exist only for tags that fired, in std::map (lexicographic) order
        std::map<std::string, uint64_t> fired;
        for (size_t k = 0; k < opt.filter_tags.size() && k < 8; ++k)
            if (C[RSQ_C_FILTERED_BY_TAG0 + k]) fired["Filtered by tag: " + opt.filter_tags[k]] += C[RSQ_C_FILTERED_BY_TAG0 + k];
        for (auto& kv : fired) output << kv.first << "\t" << kv.second << '\n';
    }
    output << "Read Length\t" << st.read_length << '\n';
    output << "Genes Detected\t" << genes_detected << '\n';
    output << "Estimated Library Complexity\t" << min_reads << '\n';
    output << "Genes used in 3' bias\t" << bias_genes << '\n';
    output << "Mean 3' bias\t" << ratio_avg << '\n';
    output << "Median 3' bias\t" << ratio_median << '\n';
    output << "3' bias Std\t" << ratio_std << '\n';
    output << "3' bias MAD_Std\t" << ratio_mad << '\n';
    output << "3' Bias, 25th Percentile\t" << ratio25 << '\n';
    output << "3' Bias, 75th Percentile\t" << ratio75 << '\n';
    // ---- fragment sizes (src/RNASeQC.cpp:570-607) ----------------------------------------------------------------
    {
        std::vector<std::pair<long long, unsigned long>> frag;  // fragmentSizes map contents, ascending
        for (size_t i = 0; i < R.fragment_hist.size(); ++i)
            if (R.fragment_hist[i]) frag.emplace_back((long long)i, (unsigned long)R.fragment_hist[i]);
        if (!frag.empty()) {
            double fragment_avg = 0.0, fragment_std = 0.0, fragment_mad = 0.0;
            unsigned long total = 0;
            for (auto& f : frag) total += f.second;
            const double size = (double)total;
            // median over the expanded sorted list
            auto expanded_at = [&](unsigned long i) {
                unsigned long acc = 0;
                for (auto& f : frag) { acc += f.second; if (i < acc) return (double)f.first; }
                return (double)frag.back().first;
            };
            const double fragment_med = median_quirk((unsigned long)size, expanded_at);
            std::ofstream fragment_list(base + ".fragmentSizes.txt");
            fragment_list << "Fragment Size\tCount" << '\n';
            std::vector<double> deviations;
            for (auto& f : frag) {
                fragment_list << f.first << "\t" << f.second << '\n';
                fragment_avg += (double)(f.first * (long long)f.second) / size;
                double deviation = std::fabs((double)f.first - fragment_med);
                for (unsigned long i = 0; i < f.second; ++i) deviations.push_back(deviation);
            }
            fragment_list.close();
            std::sort(deviations.begin(), deviations.end());
            fragment_mad = median_quirk(deviations.size(), [&](unsigned long i) { return deviations[i]; }) * kMadFactor;
            for (auto& f : frag)
                for (unsigned long i = 0; i < f.second; ++i) fragment_std += std::pow((double)f.first - fragment_avg, 2.0) / size;
            fragment_std = std::pow(fragment_std, 0.5);
            output << "Average Fragment Length\t" << fragment_avg << '\n';
            output << "Fragment Length Median\t" << fragment_med << '\n';
            output << "Fragment Length Std\t" << fragment_std << '\n';
            output << "Fragment Length MAD_Std\t" << fragment_mad << '\n';
        }
    }
    lap("metrics rows, fragment sizes");
    // ---- coverage medians and exon_cv.tsv (src/RNASeQC.cpp:609-659) --------------------------------------------
    {
        std::vector<double> means = cov.gene_means, stds = cov.gene_stds, cvs;
        const unsigned long n_transcripts = means.size();
        std::sort(means.begin(), means.end());
        std::sort(stds.begin(), stds.end());
        for (double v : cov.gene_cvs) if (!(std::isnan(v) || std::isinf(v))) cvs.push_back(v);
        std::sort(cvs.begin(), cvs.end());
        const unsigned long n_cvs = cvs.size();
        output << "Median of Avg Transcript Coverage\t" << median_quirk(n_transcripts, [&](unsigned long i) { return means[i]; }) << '\n';
        output << "Median of Transcript Coverage Std\t" << median_quirk(n_transcripts, [&](unsigned long i) { return stds[i]; }) << '\n';
        output << "Median of Transcript Coverage CV\t" << (n_cvs ? median_quirk(n_cvs, [&](unsigned long i) { return cvs[i]; }) : 0.0) << '\n';
        const bool fasta = ref && ref->open;
        // exonCoverage is a std::map keyed by exon id: rows in key order, later entries of one id overwrite earlier ones — here a
        // stable sort of pointers into the id table and the last entry of every run of equal ids
        std::vector<const std::pair<std::string, double>*> by_id;
        by_id.reserve(cov.exon_cvs.size());
        for (auto& kv : cov.exon_cvs) by_id.push_back(&kv);
        std::stable_sort(by_id.begin(), by_id.end(), [](const std::pair<std::string, double>* a, const std::pair<std::string, double>* b) { return a->first < b->first; });
        std::map<std::string, uint32_t> slot_of;
        if (fasta) for (size_t e = 0; e < A.exon_ids.size(); ++e) slot_of[A.exon_ids[e]] = (uint32_t)e;
        std::ofstream cv_report(base + ".exon_cv.tsv");
        cv_report << "Exon ID\tExon CV";
        if (fasta) cv_report << "\tGC Content";
        cv_report << '\n';
        std::vector<double> total_exon_cv;
        total_exon_cv.reserve(by_id.size());
        std::string buf;
        buf.reserve(1 << 20);
        for (size_t k = 0; k < by_id.size(); ++k) {
            if (k + 1 < by_id.size() && by_id[k + 1]->first == by_id[k]->first) continue;      // overwritten by a later entry
            const std::pair<std::string, double>& kv = *by_id[k];
            buf += kv.first; buf += '\t'; buf += fmt_default(kv.second);
            if (fasta) {
                // gc(getSeq(chromosome, start, start + length)) with the 1-based GTF start used as a 0-based offset (src/Metrics.cpp:299-302)
                const uint32_t s = slot_of[kv.first];
                const double g = ref->present.size() > A.exon_contig[s] && ref->present[A.exon_contig[s]]
                                     ? ref->gc_content(A.exon_contig[s], A.exon_start[s], A.exon_start[s] + A.exon_len[s]) : -1.0;
                buf += '\t'; buf += fmt_default(g);
            }
            buf += '\n';
            if (buf.size() > (1 << 20) - 4096) { cv_report.write(buf.data(), (std::streamsize)buf.size()); buf.clear(); }
            total_exon_cv.push_back(kv.second);
        }
        cv_report.write(buf.data(), (std::streamsize)buf.size());
        Stats cs = get_statistics(total_exon_cv);
        output << "Median Exon CV\t" << cs.med << '\n';
        output << "Exon CV MAD\t" << cs.mad << '\n';
    }
    lap("exon_cv.tsv");
    if (ref && ref->open) {                                                          // src/RNASeQC.cpp:660-674
        std::ofstream gc_report(base + ".gc_content.tsv");
        gc_report << "Content Bin\tCount" << '\n';
        // getAdvancedStatistics over the bins expanded in ascending order (src/Metrics.h:188-206)
        double avg = 0.0, m2 = 0.0, m3 = 0.0, m4 = 0.0, count = 0.0;
        for (unsigned i = 0; i < 100; ++i) {
            gc_report << (double)i / 100.0 << "\t" << R.gc_hist[i] << '\n';
            for (uint64_t j = 0; j < R.gc_hist[i]; ++j) {
                const double prev_count = count++;
                const double delta = (double)i - avg;
                const double delta_n = delta / count;
                const double delta_n2 = delta_n * delta_n;
                const double t = delta * delta_n * prev_count;
                avg += delta_n;
                m4 += t * delta_n2 * (count * count - 3 * count + 3) + 6 * delta_n2 * m2 - 4 * delta_n * m3;
                m3 += t * delta_n * (count - 2) - 3 * delta_n * m2;
                m2 += t;
            }
        }
        double s_avg = NAN, s_skew = NAN, s_std = NAN, s_kurt = NAN;
        if (count > 0) {
            s_std = std::pow(m2 / count, 0.5);
            s_avg = avg; s_skew = m3 / count / std::pow(s_std, 3.0); s_kurt = (count * m4) / (m2 * m2) - 3;
        }
        output << "Fragment GC Content Mean\t" << s_avg / 100.0 << '\n';
        output << "Fragment GC Content Std\t" << s_std / 100.0 << '\n';
        output << "Fragment GC Content Skewness\t" << s_skew << '\n';
        output << "Fragment GC Content Kurtosis\t" << s_kurt << '\n';
    }
    output.close();
    if (exon_table.valid()) exon_table.wait();
    lap("rest + exon_reads.gct (thread)");
}

}  // namespace rsqh
