// decode_bench — BAM ingest throughput (records/s) of the sequential reader vs the multi-threaded decoder.
//   decode_bench file.bam [threads...]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "alnreader.h"
#include "bamparallel.h"

using namespace rsqh;
using clk = std::chrono::steady_clock;

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: decode_bench file.bam [threads...]\n"); return 2; }
    DecodeOptions opt;
    {
        ContigMap contigs;
        auto rd = AlnReader::open(argv[1]);
        Decoder dec(rd->header(), contigs, contigs.names.size(), opt);
        std::vector<rsq_read> R; std::vector<rsq_block> B;
        RawAln a;
        auto t0 = clk::now();
        size_t n = 0;
        while (rd->next(a)) { dec.convert(a, R, B); if (R.size() >= (1u << 20)) { n += R.size(); R.clear(); B.clear(); } }
        n += R.size();
        const double dt = std::chrono::duration<double>(clk::now() - t0).count();
        std::printf("sequential      : %zu records in %.3f s -> %.2f M records/s\n", n, dt, n / dt / 1e6);
    }
    for (int k = 2; k < argc; ++k) {
        const int threads = std::atoi(argv[k]);
        ContigMap contigs;
        ParallelBamDecoder pb(argv[1], threads);
        Decoder dec(pb.header(), contigs, contigs.names.size(), opt);
        std::vector<rsq_read> R; std::vector<rsq_block> B;
        R.resize(1u << 20); B.resize(3u << 20); R.clear(); B.clear();     // like the CLI's batch buffers after their first batch: mapped
        // RSQ_BENCH_PREDECODE=<seconds>: start the workers first and let them run ahead for that long (as the CLI does while it parses
        // the GTF and brings the device up): the loop below then measures the consumer alone
        if (const char* e = std::getenv("RSQ_BENCH_PREDECODE")) { pb.start(opt); std::this_thread::sleep_for(std::chrono::milliseconds((long)(std::atof(e) * 1000))); }
        auto t0 = clk::now();
        size_t n = 0;
        std::vector<rsq_read_packed> P(1u << 20);
        std::vector<rsq_block> PB;
        double t_pack = 0;
        while (pb.next_batch(dec, 1u << 20, R, B)) {
            n += R.size();
            // what the CLI's main thread does next with the batch (pipeline.cpp): the wire-format packer
            PB.resize(B.size() + 1);
            size_t nb = 0;
            auto tp = clk::now();
            rsq_pack_records(R.data(), R.size(), B.data(), B.size(), P.data(), PB.data(), &nb);
            t_pack += std::chrono::duration<double>(clk::now() - tp).count();
        }
        const double dt = std::chrono::duration<double>(clk::now() - t0).count();
        std::printf("                  (of which %.3f s packing the batches into the wire format on the calling thread)\n", t_pack);
        std::printf("parallel x%-2d    : %zu records in %.3f s -> %.2f M records/s (read length %d)\n", threads, n, dt, n / dt / 1e6, dec.read_length());
    }
    return 0;
}
