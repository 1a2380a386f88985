// inflate_bench — BGZF block inflate throughput: zlib against the ingest's own decoder (fastinflate.h), one thread.
//   inflate_bench file.bam [repeats]
#include <zlib.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fastinflate.h"

using clk = std::chrono::steady_clock;

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: inflate_bench file.bam [repeats]\n"); return 2; }
    const int reps = argc > 2 ? std::atoi(argv[2]) : 3;
    FILE* fp = std::fopen(argv[1], "rb");
    if (!fp) return 1;
    std::vector<uint8_t> file;
    { uint8_t buf[1 << 16]; size_t r; while ((r = std::fread(buf, 1, sizeof buf, fp)) > 0) file.insert(file.end(), buf, buf + r); }
    std::fclose(fp);
    struct Blk { size_t off, len; uint32_t isize; };
    std::vector<Blk> blocks;
    size_t total = 0;
    for (size_t p = 0; p + 18 <= file.size();) {
        const unsigned xlen = file[p + 10] | (file[p + 11] << 8);
        const unsigned bsize = file[p + 16] | (file[p + 17] << 8);      // BC subfield first, as every BGZF writer lays it out
        const size_t cdata = (size_t)bsize + 1 - 12 - xlen - 8;
        uint32_t isize; memcpy(&isize, &file[p + 12 + xlen + cdata + 4], 4);
        blocks.push_back({p + 12 + xlen, cdata, isize});
        total += isize;
        p += (size_t)bsize + 1;
    }
    std::vector<uint8_t> out(1 << 16), out2(1 << 16);
    for (int pass = 0; pass < 2; ++pass) {
        double best = 1e30;
        size_t refused = 0;
        for (int r = 0; r < reps; ++r) {
            auto t0 = clk::now();
            for (const Blk& b : blocks) {
                if (!b.isize) continue;
                if (pass == 0) {
                    z_stream zs; memset(&zs, 0, sizeof zs);
                    inflateInit2(&zs, -15);
                    zs.next_in = &file[b.off]; zs.avail_in = (uInt)b.len; zs.next_out = out.data(); zs.avail_out = b.isize;
                    if (inflate(&zs, Z_FINISH) != Z_STREAM_END) return 3;
                    inflateEnd(&zs);
                } else if (!rsqh::fast_inflate(&file[b.off], b.len, out2.data(), b.isize)) ++refused;
            }
            best = std::min(best, std::chrono::duration<double>(clk::now() - t0).count());
        }
        std::printf("%-12s: %zu blocks, %.1f MB -> %.1f MB in %.3f s = %.2f GB/s out%s\n", pass ? "fastinflate" : "zlib", blocks.size(), file.size() / 1e6, total / 1e6, best,
                    total / best / 1e9, pass && refused ? "  (some blocks refused!)" : "");
    }
    return 0;
}
