// cramreader.cpp — CRAM (2.1 / 3.0) decode into the fields the hot loop touches (RawAln), own implementation.
//
// The reference reads CRAM through SeqLib -> htslib 1.5 (src/BamReader.h:70-101, SeqLib/src/BamReader.cpp:194-204,
// SeqLib/htslib/cram/cram_decode.c).  What the loop sees of a CRAM record is what htslib reconstructs, so this reader follows
// the CRAM specification for the container / block / codec layers and htslib 1.5 for the three places where reconstruction
// is a choice:
//   * CIGAR from the read features, equal neighbouring operations merged (cram_decode.c:1199-1840) — extractBlocks makes one
//     aligned block per M run, so the merge decides "Alignment Blocks";
//   * NM: htslib opens CRAM for reading with CRAM_OPT_DECODE_MD (hts.c:851) and its writer drops NM (cram_encode.c:2113), so
//     the NM the loop reads is REGENERATED: one per substitution ('X'), explicit base ('B', 'b') inside the reference,
//     inserted base ('I', 'i'), deleted base ('D') and per reference 'N' under an implicit match (cram_decode.c:1214-1883).
//     That needs the reference sequence — only where it is 'N' — hence the FASTA (--fasta), as for the reference itself;
//   * mate fields of records that are not "detached": copied from the partner inside the slice, template length from the
//     span of the chain, sign by leftmost / READ1 (cram_decode_slice_xref, cram_decode.c:2020-2190).
// Sequence bases and qualities are consumed (the streams must stay in step) and thrown away.
// Block compression: raw, gzip, rANS 4x8 order 0/1 (CRAM 3.0 specification §13).  bzip2 and lzma blocks are refused, as they
// are by the reference binary this repo builds for its tests (htslib compiled without those libraries).
// Not supported: embedded-reference slices, CRAM 1.0.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <array>
#include <climits>
#include <cstring>
#include <deque>
#include <fstream>
#include <future>
#include <map>
#include <memory>
#include <mutex>

#include "alnreader.h"

namespace rsqh {

namespace {

[[noreturn]] void bad(const std::string& what) { throw HostError(-1, "corrupt or unsupported CRAM: " + what); }

struct Cursor {
    const uint8_t* p;
    const uint8_t* end;
    void need(size_t n) const { if ((size_t)(end - p) < n) bad("truncated"); }
    uint8_t u8() { need(1); return *p++; }
    int32_t i32le() { need(4); int32_t v; memcpy(&v, p, 4); p += 4; return v; }
    const uint8_t* take(size_t n) { need(n); const uint8_t* q = p; p += n; return q; }
    int32_t itf8() {
        const uint8_t b0 = u8();
        if (b0 < 0x80) return b0;
        if (b0 < 0xc0) return ((b0 & 0x3f) << 8) | u8();
        if (b0 < 0xe0) { const uint32_t a = u8(), b = u8(); return ((b0 & 0x1f) << 16) | (a << 8) | b; }
        if (b0 < 0xf0) { const uint32_t a = u8(), b = u8(), c = u8(); return ((b0 & 0x0f) << 24) | (a << 16) | (b << 8) | c; }
        const uint32_t a = u8(), b = u8(), c = u8(), d = u8();
        return (int32_t)(((uint32_t)(b0 & 0x0f) << 28) | (a << 20) | (b << 12) | (c << 4) | (d & 0x0f));
    }
    int64_t ltf8() {
        const uint8_t b0 = u8();
        int extra = 0;
        while (extra < 8 && (b0 & (0x80 >> extra))) ++extra;
        uint64_t v = extra < 8 ? (uint64_t)(b0 & (0xff >> (extra + 1))) : 0;
        for (int k = 0; k < extra; ++k) v = (v << 8) | u8();
        return (int64_t)v;
    }
};

// ---- rANS 4x8 (CRAM 3.0 specification, "rANS codec") ------------------------------------------------------------------
struct RansTable { uint16_t freq[256], start[256]; uint8_t lookup[4096]; };

// frequency table of one context: symbols in ascending order with run-length shortcuts, 12-bit total
void rans_read_table(Cursor& c, RansTable& t, bool zero_means_full) {
    memset(t.freq, 0, sizeof t.freq);
    memset(t.start, 0, sizeof t.start);
    int x = 0, run = 0;
    int sym = c.u8();
    do {
        int f = c.u8();
        if (f >= 128) f = ((f & 127) << 8) | c.u8();
        if (!f && zero_means_full) f = 4096;
        if (x + f > 4096) bad("rANS frequencies");
        t.freq[sym] = (uint16_t)f; t.start[sym] = (uint16_t)x;
        memset(t.lookup + x, sym, (size_t)f);
        x += f;
        if (!run && c.p < c.end && sym + 1 == *c.p) { sym = c.u8(); run = c.u8(); }
        else if (run) { --run; if (++sym > 255) bad("rANS symbols"); }
        else sym = c.u8();
    } while (sym);
    if (x < 4095) bad("rANS frequency total");
    if (x == 4095) t.lookup[4095] = t.lookup[4094];
}

inline void rans_renorm(uint32_t& r, Cursor& c) {
    while (r < (1u << 23) && c.p < c.end) r = (r << 8) | *c.p++;
}

void rans_decode(const uint8_t* in, size_t n, std::vector<uint8_t>& out) {
    if (n < 9) bad("rANS header");
    Cursor c{in, in + n};
    const int order = c.u8();
    const uint32_t in_sz = (uint32_t)c.i32le(), out_sz = (uint32_t)c.i32le();
    if (in_sz != n - 9) bad("rANS size");
    out.assign(out_sz, 0);
    if (!out_sz) return;
    if (order == 0) {
        std::unique_ptr<RansTable> t(new RansTable);
        rans_read_table(c, *t, false);
        uint32_t r[4];
        for (int k = 0; k < 4; ++k) r[k] = (uint32_t)c.i32le();
        const size_t full = out_sz & ~3u;
        for (size_t i = 0; i < full; i += 4)
            for (int k = 0; k < 4; ++k) {
                const uint32_t m = r[k] & 4095u;
                const uint8_t s = t->lookup[m];
                out[i + k] = s;
                r[k] = t->freq[s] * (r[k] >> 12) + m - t->start[s];
                rans_renorm(r[k], c);
            }
        for (size_t k = 0; k < (out_sz & 3u); ++k) out[full + k] = t->lookup[r[k] & 4095u];
    } else if (order == 1) {
        std::vector<std::unique_ptr<RansTable>> t(256);
        int run = 0, ctx = c.u8();
        do {
            t[ctx].reset(new RansTable);
            rans_read_table(c, *t[ctx], true);
            if (!run && c.p < c.end && ctx + 1 == *c.p) { ctx = c.u8(); run = c.u8(); }
            else if (run) { --run; if (++ctx > 255) bad("rANS contexts"); }
            else ctx = c.u8();
        } while (ctx);
        uint32_t r[4];
        for (int k = 0; k < 4; ++k) r[k] = (uint32_t)c.i32le();
        const size_t q = out_sz >> 2;
        size_t at[4] = {0, q, 2 * q, 3 * q};
        uint8_t last[4] = {0, 0, 0, 0};
        auto step = [&](int k) {
            const RansTable* tb = t[last[k]].get();
            if (!tb) bad("rANS context without a table");
            const uint32_t m = r[k] & 4095u;
            const uint8_t s = tb->lookup[m];
            out[at[k]++] = s;
            r[k] = tb->freq[s] * (r[k] >> 12) + m - tb->start[s];
            rans_renorm(r[k], c);
            last[k] = s;
        };
        for (size_t i = 0; i < q; ++i) for (int k = 0; k < 4; ++k) step(k);
        while (at[3] < out_sz) step(3);
    } else bad("rANS order");
}

// ---- blocks --------------------------------------------------------------------------------------------------------------
struct Block {
    int method = 0, content_type = 0, content_id = 0;
    // Blocks of a slice are decompressed on first READ: the record loop only SKIPS over the quality and base series (by far the
    // largest blocks), so those are never inflated at all.
    const uint8_t* comp = nullptr;
    size_t comp_n = 0, raw_n = 0;
    bool loaded = false;
    std::vector<uint8_t> data;
    size_t pos = 0;          // byte cursor (EXTERNAL / BYTE_ARRAY_STOP)
    size_t bit = 0;          // bit cursor (core block), counted from the start of data
    void load();
    void need() { if (!loaded) load(); }
    uint8_t byte() { need(); if (pos >= data.size()) bad("external block exhausted"); return data[pos++]; }
    void skip(size_t n) { if (pos + n > raw_n) bad("external block exhausted"); pos += n; }   // no data needed
    int32_t itf8() { need(); Cursor c{data.data() + pos, data.data() + data.size()}; const int32_t v = c.itf8(); pos = (size_t)(c.p - data.data()); return v; }
    uint32_t bits(int n) {
        need();
        uint32_t v = 0;
        for (int k = 0; k < n; ++k) {
            if ((bit >> 3) >= data.size()) bad("core block exhausted");
            v = (v << 1) | ((data[bit >> 3] >> (7 - (bit & 7))) & 1u);
            ++bit;
        }
        return v;
    }
};

void inflate_gzip(const uint8_t* in, size_t n, size_t raw, std::vector<uint8_t>& out) {
    out.assign(raw, 0);
    z_stream z;
    memset(&z, 0, sizeof z);
    if (inflateInit2(&z, 15 + 32) != Z_OK) bad("zlib init");
    z.next_in = const_cast<Bytef*>(in); z.avail_in = (uInt)n;
    z.next_out = out.data(); z.avail_out = (uInt)raw;
    const int rc = inflate(&z, Z_FINISH);
    inflateEnd(&z);
    if (rc != Z_STREAM_END && !(rc == Z_OK && z.avail_out == 0) && !(rc == Z_BUF_ERROR && z.avail_out == 0)) bad("gzip block");
}

void Block::load() {
    loaded = true;
    switch (method) {
        case 0: data.assign(comp, comp + comp_n); break;
        case 1: inflate_gzip(comp, comp_n, raw_n, data); break;
        case 4: rans_decode(comp, comp_n, data); break;
        default: bad("block compression method " + std::to_string(method));
    }
    if (data.size() != raw_n) bad("block size differs from its header");   // skip() trusted the header
}

Block read_block(Cursor& c, int major, bool lazy = false) {
    Block b;
    b.method = c.u8();
    b.content_type = c.u8();
    b.content_id = c.itf8();
    const int32_t comp = c.itf8(), raw = c.itf8();
    if (comp < 0 || raw < 0) bad("block sizes");
    b.comp = c.take((size_t)comp);
    b.comp_n = (size_t)comp; b.raw_n = (size_t)raw;
    switch (b.method) {
        case 0: case 1: case 4: break;
        case 2: bad("bzip2-compressed block (this build, like the reference's test build, has no bzip2)");
        case 3: bad("lzma-compressed block (this build, like the reference's test build, has no lzma)");
        default: bad("block compression method " + std::to_string(b.method));
    }
    if (major >= 3) c.take(4);   // CRC32
    if (!lazy) b.load();
    return b;
}

// ---- encodings -------------------------------------------------------------------------------------------------------------
enum { E_NULL = 0, E_EXTERNAL = 1, E_GOLOMB = 2, E_HUFFMAN = 3, E_BYTE_ARRAY_LEN = 4, E_BYTE_ARRAY_STOP = 5, E_BETA = 6, E_SUBEXP = 7,
       E_GOLOMB_RICE = 8, E_GAMMA = 9 };

struct Codec {
    int kind = E_NULL;
    int ext_id = 0, offset = 0, nbits = 0, k = 0, m = 0;
    uint8_t stop = 0;
    struct HuffEntry { int len; int32_t sym; uint32_t code; };
    std::vector<HuffEntry> huff;          // canonical, sorted by (len, sym)
    std::unique_ptr<Codec> len_codec, val_codec;
    mutable Block* blk = nullptr;         // the external block of ext_id in the slice being decoded (bind_blocks)
};

std::unique_ptr<Codec> parse_codec(Cursor& c) {
    std::unique_ptr<Codec> e(new Codec);
    e->kind = c.itf8();
    const int32_t plen = c.itf8();
    if (plen < 0) bad("encoding parameters");
    Cursor p{c.take((size_t)plen), nullptr};
    p.end = p.p + plen;
    switch (e->kind) {
        case E_NULL: break;
        case E_EXTERNAL: e->ext_id = p.itf8(); break;
        case E_GOLOMB: case E_GOLOMB_RICE: e->offset = p.itf8(); e->m = p.itf8(); break;
        case E_HUFFMAN: {
            const int32_t ns = p.itf8();
            std::vector<int32_t> syms((size_t)std::max(ns, 0));
            for (auto& s : syms) s = p.itf8();
            const int32_t nl = p.itf8();
            if (nl != ns) bad("huffman table");
            for (int32_t i = 0; i < ns; ++i) e->huff.push_back({p.itf8(), syms[(size_t)i], 0u});
            std::sort(e->huff.begin(), e->huff.end(), [](const Codec::HuffEntry& a, const Codec::HuffEntry& b) {
                return a.len != b.len ? a.len < b.len : a.sym < b.sym; });
            uint32_t code = 0; int prev = e->huff.empty() ? 0 : e->huff[0].len;
            for (auto& h : e->huff) { code <<= (h.len - prev); prev = h.len; h.code = code++; }
            break;
        }
        case E_BYTE_ARRAY_LEN: e->len_codec = parse_codec(p); e->val_codec = parse_codec(p); break;
        case E_BYTE_ARRAY_STOP: e->stop = p.u8(); e->ext_id = p.itf8(); break;
        case E_BETA: e->offset = p.itf8(); e->nbits = p.itf8(); break;
        case E_SUBEXP: e->offset = p.itf8(); e->k = p.itf8(); break;
        case E_GAMMA: e->offset = p.itf8(); break;
        default: bad("encoding " + std::to_string(e->kind));
    }
    return e;
}

struct SliceCtx {
    Block* core = nullptr;
    std::map<int, Block*> ext;
    Block& external(const Codec& e) { if (!e.blk) bad("missing external block " + std::to_string(e.ext_id)); return *e.blk; }
};
void bind_blocks(const Codec& e, const SliceCtx& s) {
    auto it = s.ext.find(e.ext_id);
    e.blk = (e.kind == E_EXTERNAL || e.kind == E_BYTE_ARRAY_STOP) && it != s.ext.end() ? it->second : nullptr;
    if (e.len_codec) bind_blocks(*e.len_codec, s);
    if (e.val_codec) bind_blocks(*e.val_codec, s);
}

int32_t decode_int(const Codec& e, SliceCtx& s) {
    switch (e.kind) {
        case E_EXTERNAL: return s.external(e).itf8();
        case E_HUFFMAN: {
            if (e.huff.empty()) bad("empty huffman code");
            if (e.huff.size() == 1 && e.huff[0].len == 0) return e.huff[0].sym;
            uint32_t code = 0; int len = 0; size_t i = 0;
            for (;;) {
                while (i < e.huff.size() && e.huff[i].len == len) { if (e.huff[i].code == code) return e.huff[i].sym; ++i; }
                if (i >= e.huff.size()) bad("huffman code");
                const int want = e.huff[i].len;
                code = (code << (want - len)) | s.core->bits(want - len);
                len = want;
            }
        }
        case E_BETA: return (int32_t)s.core->bits(e.nbits) - e.offset;
        case E_GAMMA: {
            int z = 0;
            while (s.core->bits(1) == 0) if (++z > 31) bad("gamma code");
            const uint32_t v = (1u << z) | (z ? s.core->bits(z) : 0u);
            return (int32_t)v - e.offset;
        }
        case E_SUBEXP: {
            int u = 0;
            while (s.core->bits(1) == 1) if (++u > 31) bad("subexp code");
            uint32_t v;
            if (u == 0) v = s.core->bits(e.k);
            else { const int n = u + e.k - 1; v = (1u << n) | s.core->bits(n); }
            return (int32_t)v - e.offset;
        }
        case E_GOLOMB_RICE: {
            uint32_t q = 0;
            while (s.core->bits(1) == 1) ++q;
            return (int32_t)((q << e.m) | s.core->bits(e.m)) - e.offset;
        }
        default: bad("integer read through encoding " + std::to_string(e.kind));
    }
}

uint8_t decode_byte(const Codec& e, SliceCtx& s) {
    if (e.kind == E_EXTERNAL) return s.external(e).byte();
    return (uint8_t)decode_int(e, s);
}

// one byte whose value is not needed
void skip_byte(const Codec& e, SliceCtx& s) {
    if (e.kind == E_EXTERNAL) s.external(e).skip(1);
    else (void)decode_int(e, s);
}

// byte array; `n` < 0: the encoding knows the length (BYTE_ARRAY_LEN / BYTE_ARRAY_STOP), else n items of a byte encoding
void decode_bytes(const Codec& e, SliceCtx& s, int32_t n, std::vector<uint8_t>* out) {
    if (out) out->clear();
    if (e.kind == E_BYTE_ARRAY_LEN) {
        const int32_t len = decode_int(*e.len_codec, s);
        if (len < 0) bad("byte array length");
        decode_bytes(*e.val_codec, s, len, out);
    } else if (e.kind == E_BYTE_ARRAY_STOP) {
        Block& b = s.external(e);
        for (;;) { const uint8_t v = b.byte(); if (v == e.stop) break; if (out) out->push_back(v); }
    } else {
        if (n < 0) bad("byte array without a length");
        if (e.kind == E_EXTERNAL) {
            Block& b = s.external(e);
            if (!out) b.skip((size_t)n);
            else {
                b.need();
                if (b.pos + (size_t)n > b.data.size()) bad("external block exhausted");
                out->assign(b.data.begin() + (long)b.pos, b.data.begin() + (long)b.pos + n);
                b.pos += (size_t)n;
            }
        } else for (int32_t i = 0; i < n; ++i) { const uint8_t v = decode_byte(e, s); if (out) out->push_back(v); }
    }
}

// length of a self-delimiting byte array whose content is not needed
int32_t skip_bytes_len(const Codec& e, SliceCtx& s, std::vector<uint8_t>& scratch) {
    if (e.kind == E_BYTE_ARRAY_LEN) {
        const int32_t len = decode_int(*e.len_codec, s);
        if (len < 0) bad("byte array length");
        decode_bytes(*e.val_codec, s, len, nullptr);
        return len;
    }
    decode_bytes(e, s, -1, &scratch);
    return (int32_t)scratch.size();
}

struct CompHeader {
    bool names_included = true, ap_delta = true;
    std::vector<std::vector<std::array<uint8_t, 3>>> tag_lines;     // TD: tag dictionary
    std::map<std::string, std::unique_ptr<Codec>> ds;               // data series by two-letter key
    std::map<int32_t, std::unique_ptr<Codec>> tags;                 // (tag0 << 16 | tag1 << 8 | type) -> value encoding
    const Codec* get(const char* key) const { auto it = ds.find(key); return it == ds.end() ? nullptr : it->second.get(); }
    // the data series the record loop reads, looked up once per container
    const Codec *BF = nullptr, *CF = nullptr, *RI = nullptr, *RL = nullptr, *AP = nullptr, *RG = nullptr, *RN = nullptr, *MF = nullptr, *NS = nullptr,
                *NP = nullptr, *TS = nullptr, *NF = nullptr, *TL = nullptr, *TC = nullptr, *TN = nullptr, *FN = nullptr, *FC = nullptr, *FP = nullptr,
                *DL = nullptr, *BA = nullptr, *BS = nullptr, *IN = nullptr, *SC = nullptr, *HC = nullptr, *PD = nullptr, *RS = nullptr, *BB = nullptr,
                *QQ = nullptr, *MQ = nullptr, *QS = nullptr;
    void resolve() {
        BF = get("BF"); CF = get("CF"); RI = get("RI"); RL = get("RL"); AP = get("AP"); RG = get("RG"); RN = get("RN"); MF = get("MF"); NS = get("NS");
        NP = get("NP"); TS = get("TS"); NF = get("NF"); TL = get("TL"); TC = get("TC"); TN = get("TN"); FN = get("FN"); FC = get("FC"); FP = get("FP");
        DL = get("DL"); BA = get("BA"); BS = get("BS"); IN = get("IN"); SC = get("SC"); HC = get("HC"); PD = get("PD"); RS = get("RS"); BB = get("BB");
        QQ = get("QQ"); MQ = get("MQ"); QS = get("QS");
    }
};

CompHeader parse_comp_header(Block& b) {
    CompHeader h;
    Cursor c{b.data.data(), b.data.data() + b.data.size()};
    {   // preservation map
        c.itf8();
        const int32_t n = c.itf8();
        for (int32_t i = 0; i < n; ++i) {
            const char k0 = (char)c.u8(), k1 = (char)c.u8();
            if (k0 == 'R' && k1 == 'N') h.names_included = c.u8() != 0;
            else if (k0 == 'A' && k1 == 'P') h.ap_delta = c.u8() != 0;
            else if (k0 == 'R' && k1 == 'R') c.u8();
            else if (k0 == 'S' && k1 == 'M') c.take(5);
            else if (k0 == 'T' && k1 == 'D') {
                const int32_t len = c.itf8();
                const uint8_t* d = c.take((size_t)std::max(len, 0));
                std::vector<std::array<uint8_t, 3>> line;
                for (int32_t j = 0; j < len;) {
                    if (d[j] == 0) { h.tag_lines.push_back(line); line.clear(); ++j; continue; }
                    if (j + 3 > len) bad("tag dictionary");
                    line.push_back({d[j], d[j + 1], d[j + 2]});
                    j += 3;
                }
                if (!line.empty()) h.tag_lines.push_back(line);
            } else bad(std::string("preservation key ") + k0 + k1);
        }
    }
    {   // data series encodings
        c.itf8();
        const int32_t n = c.itf8();
        for (int32_t i = 0; i < n; ++i) {
            std::string key; key += (char)c.u8(); key += (char)c.u8();
            h.ds[key] = parse_codec(c);
        }
    }
    {   // tag encodings
        c.itf8();
        const int32_t n = c.itf8();
        for (int32_t i = 0; i < n; ++i) { const int32_t key = c.itf8(); h.tags[key] = parse_codec(c); }
    }
    h.resolve();
    return h;
}

// ---- reference: where it is 'N' (the only thing NM regeneration needs of it) ----------------------------------------------
struct NRuns { std::once_flag once; bool present = false; std::vector<std::pair<int64_t, int64_t>> runs; };   // [begin, end) 0-based

class ReferenceNs {
public:
    explicit ReferenceNs(const std::string& fasta) : path_(fasta) {
        if (fasta.empty()) return;
        std::string fai = fasta + ".fai";
        std::ifstream in(fai);
        if (!in.is_open()) { const size_t dot = fasta.find_last_of('.'); if (dot != std::string::npos) { fai = fasta.substr(0, dot) + ".fai"; in.open(fai); } }
        if (!in.is_open()) throw HostError(10, "Unable to open the FASTA index of " + fasta);
        std::string name; uint64_t len, off, lb, lw;
        while (in >> name >> len >> off >> lb >> lw) index_[name] = {len, off, lb, lw};
    }
    bool given() const { return !path_.empty(); }
    // the contigs that will be asked for (the CRAM header's): after this the map is never modified, so the lazy per-contig
    // loads below are safe from any number of decoding threads
    void declare(const std::vector<std::string>& contigs) { for (const auto& c : contigs) runs_[c]; }
    // number of 'N' / 'n' reference bases in [a, b) of the named contig; -1 when the FASTA does not have it
    int64_t count(const std::string& contig, int64_t a, int64_t b) {
        NRuns& r = slot(contig);
        if (!r.present) return -1;
        int64_t n = 0;
        auto it = std::upper_bound(r.runs.begin(), r.runs.end(), std::make_pair(a, (int64_t)INT64_MAX));
        if (it != r.runs.begin()) --it;
        for (; it != r.runs.end() && it->first < b; ++it) n += std::max<int64_t>(0, std::min(b, it->second) - std::max(a, it->first));
        return n;
    }
    bool has(const std::string& contig) { return slot(contig).present; }

private:
    struct Fai { uint64_t len, off, line_bases, line_width; };
    NRuns& slot(const std::string& contig) {
        auto it = runs_.find(contig);
        if (it == runs_.end()) bad("contig outside the CRAM header: " + contig);
        std::call_once(it->second.once, [&] { load(contig, it->second); });
        return it->second;
    }
    void load(const std::string& contig, NRuns& r) {
        auto it = index_.find(contig);
        if (it == index_.end() || !it->second.line_bases) return;
        std::ifstream in(path_, std::ios::binary);
        if (!in.is_open()) throw HostError(10, "Unable to open FASTA file: " + path_);
        const Fai& f = it->second;
        const uint64_t lines = (f.len + f.line_bases - 1) / f.line_bases;
        std::vector<char> buf((size_t)(f.len + lines * (f.line_width - f.line_bases) + 8));
        in.seekg((std::streamoff)f.off);
        in.read(buf.data(), (std::streamsize)buf.size());
        const size_t got = (size_t)in.gcount();
        int64_t pos = 0, start = -1;
        for (size_t i = 0; i < got && pos < (int64_t)f.len; ++i) {
            const char ch = buf[i];
            if (ch == '\n' || ch == '\r') continue;
            const bool n = ch == 'N' || ch == 'n';
            if (n && start < 0) start = pos;
            if (!n && start >= 0) { r.runs.emplace_back(start, pos); start = -1; }
            ++pos;
        }
        if (start >= 0) r.runs.emplace_back(start, pos);
        r.present = true;
    }
    std::string path_;
    std::map<std::string, Fai> index_;
    std::map<std::string, NRuns> runs_;
};

// ---- the reader ---------------------------------------------------------------------------------------------------------------
class CramReader : public AlnReader {
public:
    CramReader(const std::string& path, const std::string& fasta, int threads) : ref_(fasta), threads_(std::max(threads, 1)) {
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) throw HostError(10, "Unable to open BAM file: " + path);
        struct stat st;
        if (fstat(fd_, &st) != 0 || st.st_size < 26) { ::close(fd_); throw HostError(-1, "not a CRAM file: " + path); }
        size_ = (size_t)st.st_size;
        map_ = (const uint8_t*)mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (map_ == MAP_FAILED) { ::close(fd_); throw HostError(10, "Unable to map " + path); }
        cur_ = Cursor{map_, map_ + size_};
        if (memcmp(cur_.take(4), "CRAM", 4) != 0) bad("magic");
        major_ = cur_.u8(); cur_.u8();
        if (major_ != 2 && major_ != 3) bad("version " + std::to_string(major_) + " (2.x and 3.x are supported)");
        cur_.take(20);
        read_sam_header();
        ref_.declare(hdr_.ref_names);
    }
    ~CramReader() override {
        for (auto& f : inflight_) if (f.valid()) f.wait();   // they read the mapping
        inflight_.clear(); if (map_ && map_ != MAP_FAILED) munmap(const_cast<uint8_t*>(map_), size_); if (fd_ >= 0) ::close(fd_); }
    const AlnHeader& header() const override { return hdr_; }
    bool next(RawAln& out) override {
        while (at_ >= ready_.size()) { ready_.clear(); at_ = 0; if (!next_container()) return false; }
        out = std::move(ready_[at_++]);
        return true;
    }
    // Containers are self-contained (their own compression header, mates resolved within a slice), so with threads_ > 1 the
    // caller's thread only walks the container headers and up to 2 x threads_ bodies are decoded ahead of it, handed back
    // in file order.  An error inside a container surfaces when next() reaches that container, like the sequential reader.

private:
    struct ContainerHeader { int32_t length, ref_id, start, span, n_records, n_blocks; int64_t counter; std::vector<int32_t> landmarks; };
    ContainerHeader read_container_header() {
        ContainerHeader h;
        h.length = cur_.i32le();
        h.ref_id = cur_.itf8(); h.start = cur_.itf8(); h.span = cur_.itf8(); h.n_records = cur_.itf8();
        h.counter = major_ >= 3 ? cur_.ltf8() : cur_.itf8();
        cur_.ltf8();   // bases
        h.n_blocks = cur_.itf8();
        const int32_t nl = cur_.itf8();
        for (int32_t i = 0; i < nl; ++i) h.landmarks.push_back(cur_.itf8());
        if (major_ >= 3) cur_.take(4);
        return h;
    }
    void read_sam_header() {
        const ContainerHeader h = read_container_header();
        const uint8_t* body = cur_.take((size_t)h.length);
        Cursor c{body, body + h.length};
        Block b = read_block(c, major_);
        if (b.data.size() < 4) bad("SAM header block");
        int32_t len; memcpy(&len, b.data.data(), 4);
        const std::string text((const char*)b.data.data() + 4, (size_t)std::min<int64_t>(len, (int64_t)b.data.size() - 4));
        size_t p = 0;
        while (p < text.size()) {
            size_t e = text.find('\n', p);
            if (e == std::string::npos) e = text.size();
            const std::string line = text.substr(p, e - p);
            p = e + 1;
            if (line.compare(0, 3, "@SQ") == 0) {
                std::string name; int64_t ln = 0;
                size_t q = 3;
                while (q < line.size()) {
                    size_t t = line.find('\t', q + 1);
                    if (t == std::string::npos) t = line.size();
                    const std::string f = line.substr(q + 1, t - q - 1);
                    if (f.compare(0, 3, "SN:") == 0) name = f.substr(3);
                    else if (f.compare(0, 3, "LN:") == 0) ln = std::strtoll(f.c_str() + 3, nullptr, 10);
                    q = t;
                }
                hdr_.ref_names.push_back(name); hdr_.ref_lens.push_back(ln);
            } else if (line.compare(0, 3, "@RG") == 0) ++n_read_groups_;
        }
    }

    struct Rec {   // one record of a slice while its mates are being resolved
        RawAln a;
        int32_t apos = 0, aend = 0, mate_line = -1, mate_flags = 0, tlen = INT_MIN, mate_ref = -1, mate_pos = 0;
        bool named = false;
    };

    // the next container with records in it: its body, or {nullptr, 0} at the end of the file
    std::pair<const uint8_t*, size_t> scan_container() {
        while (cur_.p < cur_.end) {
            const ContainerHeader h = read_container_header();
            const uint8_t* body = cur_.take((size_t)std::max(h.length, 0));
            if (h.n_records != 0) return {body, (size_t)h.length};   // else the EOF marker / an empty container
        }
        return {nullptr, 0};
    }
    std::vector<RawAln> decode_container(const uint8_t* body, size_t length) const {
        std::vector<RawAln> out;
        Cursor c{body, body + length};
        Block chb = read_block(c, major_);
        if (chb.content_type != 1) bad("compression header expected");
        const CompHeader ch = parse_comp_header(chb);
        while (c.p < c.end) decode_slice(c, ch, out);
        return out;
    }
    bool next_container() {
        if (threads_ <= 1) {
            const auto b = scan_container();
            if (!b.first) return false;
            ready_ = decode_container(b.first, b.second);
            return true;
        }
        while (!scanned_all_ && inflight_.size() < 2 * (size_t)threads_) {
            const auto b = scan_container();
            if (!b.first) { scanned_all_ = true; break; }
            inflight_.push_back(std::async(std::launch::async, [this, b] { return decode_container(b.first, b.second); }));
        }
        if (inflight_.empty()) return false;
        ready_ = inflight_.front().get();
        inflight_.pop_front();
        return true;
    }

    void decode_slice(Cursor& c, const CompHeader& ch, std::vector<RawAln>& out) const {
        Block shb = read_block(c, major_);
        if (shb.content_type != 2) bad("slice header expected");
        Cursor s{shb.data.data(), shb.data.data() + shb.data.size()};
        const int32_t ref_id = s.itf8();
        const int32_t slice_start = s.itf8();
        s.itf8();                                 // span
        const int32_t n_records = s.itf8();
        const int64_t counter = major_ >= 3 ? s.ltf8() : s.itf8();
        const int32_t n_blocks = s.itf8();
        const int32_t n_ids = s.itf8();
        for (int32_t i = 0; i < n_ids; ++i) s.itf8();
        const int32_t embedded = s.itf8();
        if (embedded >= 0) bad("slice with an embedded reference");
        std::vector<Block> blocks;
        blocks.reserve((size_t)std::max(n_blocks, 0));
        for (int32_t i = 0; i < n_blocks; ++i) blocks.push_back(read_block(c, major_, true));
        SliceCtx ctx;
        for (Block& b : blocks) { if (b.content_type == 5) ctx.core = &b; else if (b.content_type == 4) ctx.ext[b.content_id] = &b; }
        if (!ctx.core) bad("slice without a core block");
        for (const auto& kv : ch.ds) bind_blocks(*kv.second, ctx);
        for (const auto& kv : ch.tags) bind_blocks(*kv.second, ctx);

        std::vector<Rec> recs((size_t)std::max(n_records, 0));
        int32_t last_apos = slice_start;          // AP deltas run from the slice's start (htslib: s->last_apos = hdr->ref_seq_start)
        std::vector<uint8_t> scratch;
        for (int32_t r = 0; r < n_records; ++r) decode_record(ch, ctx, ref_id, recs[(size_t)r], r, last_apos, scratch);
        resolve_mates(recs);
        for (int32_t r = 0; r < n_records; ++r) {
            Rec& x = recs[(size_t)r];
            if (!x.named) {       // names discarded: mates must still share one (htslib cram_to_bam: prefix:counter of the first of the pair)
                const int64_t id = counter + ((x.mate_line >= 0 && x.mate_line < r) ? x.mate_line : r) + 1;
                x.a.qname = "cram:" + std::to_string(id);
            }
            x.a.pos = x.apos - 1;
            x.a.mtid = x.mate_ref; x.a.mpos = x.mate_pos - 1; x.a.isize = x.tlen == INT_MIN ? 0 : x.tlen;
            out.push_back(std::move(x.a));
        }
    }

    static void push_cigar(std::vector<uint32_t>& cig, uint32_t& run_len, int& run_op, int op, uint32_t len) {
        if (run_len && run_op != op) { cig.push_back((run_len << 4) | (uint32_t)run_op); run_len = 0; }
        run_op = op; run_len += len;
    }

    void decode_record(const CompHeader& ch, SliceCtx& ctx, int32_t slice_ref, Rec& x, int32_t /*index*/, int32_t& last_apos,
                       std::vector<uint8_t>& scratch) const {
        auto need = [&](const Codec* e, const char* k) -> const Codec& { if (!e) bad(std::string("data series ") + k + " missing"); return *e; };
        RawAln& a = x.a;
        a = RawAln();
        const int32_t bf = decode_int(need(ch.BF, "BF"), ctx);
        const int32_t cf = decode_int(need(ch.CF, "CF"), ctx);
        a.flag = (uint16_t)bf;
        a.tid = slice_ref == -2 ? decode_int(need(ch.RI, "RI"), ctx) : slice_ref;
        if (a.tid < -1 || a.tid >= (int32_t)hdr_.ref_names.size()) bad("reference id");
        int32_t len = decode_int(need(ch.RL, "RL"), ctx);
        if (len < 0) bad("read length");
        x.apos = decode_int(need(ch.AP, "AP"), ctx);
        if (ch.ap_delta) x.apos += last_apos;
        last_apos = x.apos;
        const int32_t rg = decode_int(need(ch.RG, "RG"), ctx);
        if (ch.names_included) { decode_bytes(need(ch.RN, "RN"), ctx, -1, &scratch); a.qname.assign(scratch.begin(), scratch.end()); x.named = !scratch.empty(); }
        x.mate_line = -1; x.mate_ref = -1; x.mate_pos = 0; x.mate_flags = 0; x.tlen = INT_MIN;
        if (cf & 2) {                                   // detached: the mate's fields are stored
            x.mate_flags = decode_int(need(ch.MF, "MF"), ctx);
            if (!ch.names_included) { decode_bytes(need(ch.RN, "RN"), ctx, -1, &scratch); a.qname.assign(scratch.begin(), scratch.end()); x.named = !scratch.empty(); }
            x.mate_ref = decode_int(need(ch.NS, "NS"), ctx);
            x.mate_pos = decode_int(need(ch.NP, "NP"), ctx);
            x.tlen = decode_int(need(ch.TS, "TS"), ctx);
        } else if (cf & 4) {                            // the mate follows in this slice
            x.mate_line = decode_int(need(ch.NF, "NF"), ctx);   // distance; made absolute by the caller's index below
            x.mate_line = -2 - x.mate_line;              // marker: resolved in resolve_mates (needs the record index)
        }
        // auxiliary tags
        bool has_nm = false, has_md = false;
        if (const Codec* tl = ch.TL) {
            const int32_t line = decode_int(*tl, ctx);
            if (line < 0 || line >= (int32_t)ch.tag_lines.size()) bad("tag line");
            for (const auto& t : ch.tag_lines[(size_t)line]) {
                const int32_t key = (t[0] << 16) | (t[1] << 8) | t[2];
                auto it = ch.tags.find(key);
                if (it == ch.tags.end()) bad("tag encoding");
                decode_bytes(*it->second, ctx, -1, &scratch);
                add_aux(a, t[0], t[1], t[2], scratch);
                if (t[0] == 'N' && t[1] == 'M') has_nm = true;
                if (t[0] == 'M' && t[1] == 'D') has_md = true;
            }
        } else if (const Codec* tc = ch.TC) {    // CRAM 2.x: count + ids
            const int32_t n = decode_int(*tc, ctx);
            for (int32_t i = 0; i < n; ++i) {
                const int32_t key = decode_int(need(ch.TN, "TN"), ctx);
                auto it = ch.tags.find(key);
                if (it == ch.tags.end()) bad("tag encoding");
                decode_bytes(*it->second, ctx, -1, &scratch);
                add_aux(a, (uint8_t)(key >> 16), (uint8_t)(key >> 8), (uint8_t)key, scratch);
                if (((key >> 16) & 0xff) == 'N' && ((key >> 8) & 0xff) == 'M') has_nm = true;
                if (((key >> 16) & 0xff) == 'M' && ((key >> 8) & 0xff) == 'D') has_md = true;
            }
        }
        x.aend = x.apos;
        if (!(bf & 0x4)) {                              // mapped: features -> CIGAR, NM
            if (x.apos <= 0) bad("mapped read without a position");
            const std::string& contig = hdr_.ref_names[(size_t)std::max(a.tid, 0)];
            const int64_t ref_len = a.tid >= 0 ? hdr_.ref_lens[(size_t)a.tid] : 0;
            const bool have_ref = a.tid >= 0 && ref_.given() && ref_.has(contig);
            if (a.tid >= 0 && !have_ref)
                // the reference: referenceHTSMismatch -> exit 11 when --fasta was given (src/BamReader.cpp:22, src/RNASeQC.cpp:725),
                // the bare htslib runtime_error -> "Unknown error", exit -1, when it was not
                throw HostError(ref_.given() ? 11 : -1, "HTSLib was unable to find a suitable reference while decoding a cram: no sequence for " +
                                    contig + (ref_.given() ? " in the FASTA" : " (CRAM input needs --fasta)"));
            const bool count_nm = have_ref && !has_nm && !(cf & 8);
            uint32_t nm = 0, run_len = 0;
            int run_op = 0;                             // BAM op codes: M0 I1 D2 N3 S4 H5 P6
            int64_t ref_pos = (int64_t)x.apos - 1;      // 0-based
            int32_t seq_pos = 1, prev = 0;
            auto implicit_match = [&](int32_t n) {      // n read bases that equal the reference
                if (n <= 0) return;
                if (have_ref && ref_pos + n <= ref_len) { if (count_nm) nm += (uint32_t)ref_.count(contig, ref_pos, ref_pos + n); }
                push_cigar(a.cigar, run_len, run_op, 0, (uint32_t)n);
                ref_pos += n; seq_pos += n;
            };
            const int32_t fn = decode_int(need(ch.FN, "FN"), ctx);
            for (int32_t f = 0; f < fn; ++f) {
                const uint8_t op = decode_byte(need(ch.FC, "FC"), ctx);
                const int32_t pos = decode_int(need(ch.FP, "FP"), ctx) + prev;
                if (pos <= 0 || pos > len + 1) bad("feature position");
                if (pos > seq_pos) implicit_match(pos - seq_pos);
                prev = pos;
                switch (op) {
                    case 'S': { const int32_t n = skip_bytes_len(need(ch.SC, "SC"), ctx, scratch);
                                if (run_len) { a.cigar.push_back((run_len << 4) | (uint32_t)run_op); run_len = 0; }
                                a.cigar.push_back(((uint32_t)n << 4) | 4u); run_op = 4;
                                seq_pos += n; break; }
                    case 'X': { skip_byte(need(ch.BS, "BS"), ctx);
                                push_cigar(a.cigar, run_len, run_op, 0, 1);
                                if (count_nm && ref_pos < ref_len) ++nm;
                                ++seq_pos; ++ref_pos; break; }
                    case 'D': { const int32_t n = decode_int(need(ch.DL, "DL"), ctx);
                                push_cigar(a.cigar, run_len, run_op, 2, (uint32_t)n);
                                if (count_nm) nm += (uint32_t)std::max<int64_t>(0, std::min<int64_t>(n, ref_len - ref_pos));
                                ref_pos += n; break; }
                    case 'I': { const int32_t n = skip_bytes_len(need(ch.IN, "IN"), ctx, scratch);
                                push_cigar(a.cigar, run_len, run_op, 1, (uint32_t)n);
                                seq_pos += n; if (count_nm) nm += (uint32_t)n; break; }
                    case 'i': { skip_byte(need(ch.BA, "BA"), ctx);
                                push_cigar(a.cigar, run_len, run_op, 1, 1);
                                ++seq_pos; if (count_nm) ++nm; break; }
                    case 'b': { const int32_t n = skip_bytes_len(need(ch.BB, "BB"), ctx, scratch);
                                push_cigar(a.cigar, run_len, run_op, 0, (uint32_t)n);
                                if (count_nm) nm += (uint32_t)std::max<int64_t>(0, std::min<int64_t>((int64_t)n, ref_len - ref_pos));
                                seq_pos += n; ref_pos += (int64_t)n; break; }
                    case 'q': { const int32_t n = skip_bytes_len(need(ch.QQ, "QQ"), ctx, scratch);
                                push_cigar(a.cigar, run_len, run_op, 0, (uint32_t)n);
                                seq_pos += n; ref_pos += (int64_t)n; break; }
                    case 'B': { skip_byte(need(ch.BA, "BA"), ctx); skip_byte(need(ch.QS, "QS"), ctx);
                                push_cigar(a.cigar, run_len, run_op, 0, 1);
                                if (count_nm && ref_pos < ref_len) ++nm;
                                ++seq_pos; ++ref_pos; break; }
                    case 'Q': skip_byte(need(ch.QS, "QS"), ctx); break;
                    case 'H': { const int32_t n = decode_int(need(ch.HC, "HC"), ctx); push_cigar(a.cigar, run_len, run_op, 5, (uint32_t)n); break; }
                    case 'P': { const int32_t n = decode_int(need(ch.PD, "PD"), ctx); push_cigar(a.cigar, run_len, run_op, 6, (uint32_t)n); break; }
                    case 'N': { const int32_t n = decode_int(need(ch.RS, "RS"), ctx); push_cigar(a.cigar, run_len, run_op, 3, (uint32_t)n); ref_pos += n; break; }
                    default: bad(std::string("feature code ") + (char)op);
                }
            }
            if (len >= seq_pos) implicit_match(len - seq_pos + 1);
            if (run_len) a.cigar.push_back((run_len << 4) | (uint32_t)run_op);
            x.aend = (int32_t)ref_pos;
            a.mapq = (uint8_t)decode_int(need(ch.MQ, "MQ"), ctx);
            if (cf & 1) decode_bytes(need(ch.QS, "QS"), ctx, len, nullptr);
            if (have_ref && !has_md && !(cf & 8)) { AuxField m; m.tag[0] = 'M'; m.tag[1] = 'D'; m.type = 'Z'; m.ival = 0; a.aux.push_back(m); }   // regenerated too
            if (count_nm) { AuxField f; f.tag[0] = 'N'; f.tag[1] = 'M'; f.type = 'I'; f.ival = nm; a.aux.push_back(f); }
            if (cf & 8) len = 0;
        } else {
            if (len) decode_bytes(need(ch.BA, "BA"), ctx, len, nullptr);
            if (cf & 1) decode_bytes(need(ch.QS, "QS"), ctx, len, nullptr);
        }
        a.l_qseq = len;
        if (rg >= 0 && rg < n_read_groups_) { AuxField f; f.tag[0] = 'R'; f.tag[1] = 'G'; f.type = 'Z'; f.ival = 0; a.aux.push_back(f); }
    }

    static void add_aux(RawAln& a, uint8_t t0, uint8_t t1, uint8_t type, const std::vector<uint8_t>& v) {
        AuxField f;
        f.tag[0] = (char)t0; f.tag[1] = (char)t1; f.type = (char)type; f.ival = 0;
        auto le = [&](size_t n) { uint64_t x = 0; for (size_t i = 0; i < n && i < v.size(); ++i) x |= (uint64_t)v[i] << (8 * i); return x; };
        switch (type) {
            case 'A': case 'C': f.ival = v.empty() ? 0 : v[0]; break;
            case 'c': f.ival = v.empty() ? 0 : (int8_t)v[0]; break;
            case 's': f.ival = (int16_t)le(2); break;
            case 'S': f.ival = (uint16_t)le(2); break;
            case 'i': f.ival = (int32_t)le(4); break;
            case 'I': f.ival = (uint32_t)le(4); break;
            default: break;
        }
        a.aux.push_back(f);
    }

    // mate fields of the records that were not stored "detached" (htslib cram_decode_slice_xref)
    static void resolve_mates(std::vector<Rec>& recs) {
        const int32_t n = (int32_t)recs.size();
        for (int32_t r = 0; r < n; ++r) if (recs[(size_t)r].mate_line <= -2) recs[(size_t)r].mate_line = r + 1 + (-2 - recs[(size_t)r].mate_line);
        for (int32_t r = 0; r < n; ++r) {
            Rec& cr = recs[(size_t)r];
            if (cr.mate_line >= 0) {
                if (cr.mate_line < n) {
                    if (cr.tlen == INT_MIN) {
                        int32_t id2 = r, aleft = cr.apos, aright = cr.aend, ref = cr.a.tid, left_cnt = 0;
                        for (;;) {
                            Rec& y = recs[(size_t)id2];
                            if (aleft > y.apos) { aleft = y.apos; left_cnt = 1; } else if (aleft == y.apos) ++left_cnt;
                            if (aright < y.aend) aright = y.aend;
                            if (y.mate_line == -1) { y.mate_line = r; break; }
                            if (y.mate_line <= id2 || y.mate_line >= n) bad("mate chain");
                            id2 = y.mate_line;
                            if (recs[(size_t)id2].a.tid != ref) ref = -1;
                            if (id2 == r) break;
                        }
                        const int32_t tlen = aright - aleft + 1;
                        id2 = r;
                        do {
                            Rec& y = recs[(size_t)id2];
                            if (ref == -1) y.tlen = 0;
                            else if (y.apos == aleft) y.tlen = (left_cnt == 1 || (y.a.flag & 0x40)) ? tlen : -tlen;
                            else y.tlen = -tlen;
                            id2 = y.mate_line;
                        } while (id2 != r);
                    }
                    const Rec& m = recs[(size_t)cr.mate_line];
                    cr.mate_pos = m.apos; cr.mate_ref = m.a.tid;
                    cr.a.flag |= 0x1;
                    if (m.a.flag & 0x4) { cr.a.flag |= 0x8; cr.tlen = 0; }
                    if (cr.a.flag & 0x4) cr.tlen = 0;
                    if (m.a.flag & 0x10) cr.a.flag |= 0x20;
                }
            } else {
                if (cr.mate_flags & 1) cr.a.flag |= 0x1 | 0x20;
                if (cr.mate_flags & 2) cr.a.flag |= 0x8;
                if (!(cr.a.flag & 0x1)) cr.mate_ref = -1;
            }
            if (cr.tlen == INT_MIN) cr.tlen = 0;
        }
    }

    int fd_ = -1;
    const uint8_t* map_ = nullptr;
    size_t size_ = 0;
    Cursor cur_{nullptr, nullptr};
    int major_ = 3;
    int32_t n_read_groups_ = 0;
    AlnHeader hdr_;
    mutable ReferenceNs ref_;   // lazily loaded per contig, under a once-flag
    int threads_ = 1;
    bool scanned_all_ = false;
    std::deque<std::future<std::vector<RawAln>>> inflight_;
    std::vector<RawAln> ready_;
    size_t at_ = 0;
};

}  // namespace

std::unique_ptr<AlnReader> open_cram(const std::string& path, const std::string& fasta, int threads) {
    return std::unique_ptr<AlnReader>(new CramReader(path, fasta, threads));
}

}  // namespace rsqh
