// fasta.cpp — see fasta.h.
#include "fasta.h"

#include <algorithm>
#include <cstdio>
#include <filesystem>
#include <fstream>

namespace rsqh {

void open_reference(const std::string& fasta_path, ContigMap& contigs, ReferenceModel& ref) {
    {
        std::ifstream probe(fasta_path);
        if (!probe.is_open()) throw HostError(10, "Unable to open reference fasta: " + fasta_path);
    }
    std::string index_path = fasta_path + ".fai";
    std::filesystem::path replaced = std::filesystem::path(fasta_path).replace_extension(".fai");
    if (std::filesystem::exists(replaced)) index_path = replaced.string();
    else if (!std::filesystem::exists(index_path)) throw HostError(10, "Unable to locate fasta index: " + fasta_path);
    std::ifstream in(index_path, std::ios::binary);
    std::string line;
    while (std::getline(in, line)) {
        // name \t length \t offset \t line bases \t line bytes   (bioio.hpp FastaContigIndex)
        FaiEntry e;
        size_t p = line.find('\t');
        e.name = line.substr(0, p);
        unsigned long long v[4] = {0, 0, 0, 0};
        if (p != std::string::npos) std::sscanf(line.c_str() + p + 1, "%llu\t%llu\t%llu\t%llu", &v[0], &v[1], &v[2], &v[3]);
        e.length = v[0]; e.offset = v[1]; e.line_len = v[2]; e.line_bytes = v[3];
        ref.entries.push_back(e);
    }
    if (ref.entries.empty()) throw HostError(10, "No contigs found in fasta index: " + index_path);
    for (const FaiEntry& e : ref.entries) ref.entry_contig.push_back(contigs.get(e.name));
    ref.fasta_path = fasta_path;
    ref.open = true;
}

void load_reference_bits(ReferenceModel& ref, size_t n_contigs) {
    ref.present.assign(n_contigs, 0);
    ref.length.assign(n_contigs, 0);
    ref.word_off.assign(n_contigs + 1, 0);
    // a name listed twice in the index: the later entry wins (unordered_map emplace keeps the FIRST in bioio; chromosomeMap
    // gives both the same id) — keep the first, like read_fasta_index's emplace (bioio.hpp:236-247)
    std::vector<int> entry_of(n_contigs, -1);
    for (size_t k = 0; k < ref.entries.size(); ++k) {
        const uint32_t c = ref.entry_contig[k];
        if (c < n_contigs && entry_of[c] < 0) entry_of[c] = (int)k;
    }
    uint64_t words = 0;
    for (size_t c = 0; c < n_contigs; ++c) {
        ref.word_off[c] = words;
        if (entry_of[c] >= 0) { ref.present[c] = 1; ref.length[c] = ref.entries[entry_of[c]].length; words += (ref.length[c] + 31) / 32; }
    }
    ref.word_off[n_contigs] = words;
    ref.bits.assign(words ? words : 1, 0u);
    FILE* fp = std::fopen(ref.fasta_path.c_str(), "rb");
    if (!fp) throw HostError(10, "Unable to open reference fasta: " + ref.fasta_path);
    std::vector<char> buf;
    for (size_t c = 0; c < n_contigs; ++c) {
        if (entry_of[c] < 0) continue;
        const FaiEntry& e = ref.entries[entry_of[c]];
        if (!e.length || !e.line_len) continue;
        const uint64_t full = e.length / e.line_len, rest = e.length % e.line_len;
        const uint64_t bytes = full * e.line_bytes + rest;
        buf.resize(bytes);
        if (fseeko(fp, (off_t)e.offset, SEEK_SET) != 0) { std::fclose(fp); throw HostError(10, "Unable to read reference fasta: " + ref.fasta_path); }
        const size_t got = std::fread(buf.data(), 1, bytes, fp);
        uint32_t* w = ref.bits.data() + ref.word_off[c];
        const uint64_t pad = e.line_bytes - e.line_len;
        uint64_t base = 0;
        for (uint64_t off = 0; off < got && base < e.length;) {
            const uint64_t n = std::min<uint64_t>(e.line_len, e.length - base);
            for (uint64_t i = 0; i < n && off + i < got; ++i) {
                const char ch = buf[off + i];
                if (ch == 'G' || ch == 'g' || ch == 'C' || ch == 'c') w[(base + i) >> 5] |= 1u << ((base + i) & 31);
            }
            base += n;
            off += n + pad;
        }
    }
    std::fclose(fp);
}

rsq_reference ReferenceModel::view() const {
    rsq_reference v{};
    v.struct_size = sizeof(rsq_reference);
    v.n_contigs = (uint32_t)present.size();
    v.present = present.data();
    v.length = length.data();
    v.word_off = word_off.data();
    v.gc_bits = bits.data();
    return v;
}

double ReferenceModel::gc_content(uint32_t contig, int64_t start, int64_t end) const {
    if (contig >= present.size() || !present[contig]) return -1;
    const int64_t len = (int64_t)length[contig];
    if (start < 0 || start >= len) return -1;
    const int64_t stop = std::min(end, len);
    const double size = (double)(stop - start);
    if (size <= 0) return -1;
    const uint32_t* w = bits.data() + word_off[contig];
    double content = 0.0;
    for (int64_t i = start; i < stop; ++i)
        if ((w[i >> 5] >> (i & 31)) & 1u) content += 1.0 / size;
    return content;
}

}  // namespace rsqh
