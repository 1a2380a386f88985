// fastinflate.cpp — see fastinflate.h.  Written from RFC 1951; the table layout and the refill are this file's own.
#include "fastinflate.h"

#include <cstring>

namespace rsqh {

namespace {

// ---- decode-table entries ------------------------------------------------------------------------------------------------------
//   bits  0..7   bits to consume: code length (sub-table entries: the WHOLE code, root bits included) + extra bits — one shift
//                takes both off the bit buffer, the extra bits are read from a copy (off the critical path)
//   bits  8..11  number of extra bits (lengths, distances) / index bits of the sub-table (pointers)
//   bits 12..15  kind flags
//   bits 16..31  literal byte / base length / base distance / offset of the sub-table
constexpr uint32_t kLit = 0x8000u, kEob = 0x4000u, kSub = 0x2000u, kBad = 0x1000u;
constexpr int kLitRoot = 11, kDistRoot = 8, kPreRoot = 7;
constexpr int kLitCap = (1 << kLitRoot) + 1024, kDistCap = (1 << kDistRoot) + 512, kPreCap = 1 << kPreRoot;

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kPreOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

enum Kind { kPrecode, kLitLen, kDist };

inline uint32_t make_symbol_entry(Kind kind, int sym) {
    if (kind == kPrecode) return (uint32_t)sym << 16;
    if (kind == kLitLen) {
        if (sym < 256) return ((uint32_t)sym << 16) | kLit;
        if (sym == 256) return kEob;
        if (sym < 286) return ((uint32_t)kLenBase[sym - 257] << 16) | ((uint32_t)kLenExtra[sym - 257] << 8);
        return kBad;
    }
    if (sym < 30) return ((uint32_t)kDistBase[sym] << 16) | ((uint32_t)kDistExtra[sym] << 8);
    return kBad;
}

// entry templates per alphabet and the byte-reversal table, built once
struct Constants {
    uint32_t entry[3][288];
    uint8_t rev8[256];
    Constants() {
        for (int k = 0; k < 3; ++k)
            for (int s = 0; s < 288; ++s) entry[k][s] = make_symbol_entry((Kind)k, s);
        for (int v = 0; v < 256; ++v) {
            int r = 0;
            for (int b = 0; b < 8; ++b) r |= ((v >> b) & 1) << (7 - b);
            rev8[v] = (uint8_t)r;
        }
    }
};
const Constants kConst;

inline uint32_t symbol_entry(Kind kind, int sym) { return kConst.entry[kind][sym]; }

inline uint32_t reverse_bits(uint32_t code, int len) {                 // len <= 15
    return (((uint32_t)kConst.rev8[code & 0xffu] << 8) | kConst.rev8[code >> 8]) >> (16 - len);
}

// Canonical Huffman code of `n` symbols -> a root table of 2^root entries indexed by the next stream bits (LSB first) and, for
// codes longer than the root, sub-tables sized by the longest code under each root prefix.  False unless the code is complete
// (the one exception RFC 1951 / zlib allow is handled: a distance alphabet with no code, or with a single 1-bit code).
bool build_table(const uint8_t* lens, int n, Kind kind, int root, uint32_t* table, int cap) {
    int count[16] = {0};
    for (int i = 0; i < n; ++i) ++count[lens[i]];
    const int used = n - count[0];
    int left = 1;
    for (int l = 1; l <= 15; ++l) {
        left = (left << 1) - count[l];
        if (left < 0) return false;                                  // over-subscribed
    }
    const int n_root = 1 << root;
    if (left > 0) {                                                  // incomplete
        if (kind != kDist) return false;
        if (used == 0) { for (int i = 0; i < n_root; ++i) table[i] = kBad | 1u; return true; }
        if (used != 1 || count[1] != 1) return false;
        int sym = 0;
        while (lens[sym] != 1) ++sym;
        const uint32_t t = symbol_entry(kind, sym);
        for (int i = 0; i < n_root; i += 2) { table[i] = t | (1u + ((t >> 8) & 15u)); table[i + 1] = kBad | 1u; }
        return true;
    }
    uint32_t next[16];
    {
        uint32_t code = 0;
        count[0] = 0;
        for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
    }
    bool any_long = false;
    for (int l = root + 1; l <= 15; ++l) any_long |= count[l] != 0;
    uint8_t longest[1 << kLitRoot];                                   // longest code under each root prefix (long codes only)
    uint32_t rev_of[288];
    if (any_long) memset(longest, 0, (size_t)n_root);
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t rev = reverse_bits(next[l]++, l);
        rev_of[s] = rev;
        if (l <= root) {
            const uint32_t t = symbol_entry(kind, s), e = t | ((uint32_t)l + ((t >> 8) & 15u));
            for (int i = (int)rev; i < n_root; i += 1 << l) table[i] = e;
        } else {
            uint8_t& m = longest[rev & (uint32_t)(n_root - 1)];
            if (l > m) m = (uint8_t)l;
        }
    }
    if (!any_long) return true;
    int free_at = n_root;
    for (int p = 0; p < n_root; ++p) {
        if (!longest[p]) continue;
        const int sub_bits = longest[p] - root, size = 1 << sub_bits;
        if (free_at + size > cap) return false;
        table[p] = ((uint32_t)free_at << 16) | kSub | ((uint32_t)sub_bits << 8) | (uint32_t)root;
        for (int i = 0; i < size; ++i) table[free_at + i] = kBad | (uint32_t)longest[p];
        free_at += size;
    }
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (l <= root) continue;
        const uint32_t rev = rev_of[s], ptr = table[rev & (uint32_t)(n_root - 1)];
        const int size = 1 << ((ptr >> 8) & 15);
        uint32_t* sub = table + (ptr >> 16);
        const uint32_t t = symbol_entry(kind, s), e = t | ((uint32_t)l + ((t >> 8) & 15u));
        for (int i = (int)(rev >> root); i < size; i += 1 << (l - root)) sub[i] = e;
    }
    return true;
}

inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }   // little-endian hosts only (x86-64, aarch64)

struct Bits {
    const uint8_t* in;
    const uint8_t* in_end;
    uint64_t buf = 0;
    int cnt = 0;                       // valid bits in buf; they end exactly at `in`
    // after refill: at least 56 valid bits, or everything that is left of the input
    inline void refill() {
        if (in_end - in >= 8) {
            buf |= load64(in) << cnt;
            in += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt < 56 && in < in_end) { buf |= (uint64_t)*in++ << cnt; cnt += 8; }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)buf & ((1u << n) - 1u); }
    inline void drop(int n) { buf >>= n; cnt -= n; }
    inline bool take(int n, uint32_t& v) {          // n <= 16
        if (cnt < n) return false;
        v = peek(n); drop(n);
        return true;
    }
};

inline uint32_t lookup(const uint32_t* table, int root, uint64_t buf) {
    uint32_t e = table[(uint32_t)buf & ((1u << root) - 1u)];
    if (__builtin_expect(e & kSub, 0)) e = table[(e >> 16) + (((uint32_t)(buf >> root)) & ((1u << ((e >> 8) & 15)) - 1u))];
    return e;
}

// One block of Huffman-coded data.  False on any violation; true after its end-of-block code.
bool huffman_block(Bits& b, const uint32_t* lt, const uint32_t* dt, uint8_t* const out_begin, uint8_t*& out_io, uint8_t* const out_end) {
    uint8_t* out = out_io;
    // ---- fast loop: one unconditional 8-byte refill per iteration; room for 3 literals or one longest match + copy slack.  The bit
    // buffer lives in locals: byte stores through `out` may alias anything the compiler cannot prove private ----------------------------
    {
        uint64_t buf = b.buf;
        int cnt = b.cnt;
        const uint8_t* in = b.in;
        const uint8_t* const in_fast_end = b.in_end - 8;
        uint8_t* const out_fast_end = out_end - 320;
        bool ok = true, done = false;
#define RSQ_DROP(n) do { buf >>= (n); cnt -= (n); } while (0)
        while (in <= in_fast_end && out <= out_fast_end) {
            buf |= load64(in) << cnt;
            in += (63 - cnt) >> 3;
            cnt |= 56;
            uint32_t e = lookup(lt, kLitRoot, buf);
            if (e & kLit) {                                          // up to three literals on one refill (3 x 15 bits)
                RSQ_DROP((int)(e & 0xffu)); *out++ = (uint8_t)(e >> 16);
                e = lookup(lt, kLitRoot, buf);
                if (e & kLit) {
                    RSQ_DROP((int)(e & 0xffu)); *out++ = (uint8_t)(e >> 16);
                    e = lookup(lt, kLitRoot, buf);
                    if (e & kLit) { RSQ_DROP((int)(e & 0xffu)); *out++ = (uint8_t)(e >> 16); }
                }
                continue;                                            // whatever follows is looked up again after the refill
            }
            if (__builtin_expect(e & (kEob | kBad), 0)) {
                if (e & kBad) ok = false;
                else { RSQ_DROP((int)(e & 0xffu)); done = true; }
                break;
            }
            // a match: length code (<= 15 + 5 bits) and distance code (<= 15 + 13 bits) fit the 56 bits of one refill
            const uint64_t at_len = buf;
            RSQ_DROP((int)(e & 0xffu));                                   // code + extra bits in one shift
            const uint32_t d = lookup(dt, kDistRoot, buf);
            const int lx = (int)((e >> 8) & 15u);
            const uint32_t length = (e >> 16) + ((uint32_t)(at_len >> ((e & 0xffu) - lx)) & ((1u << lx) - 1u));
            if (__builtin_expect(d & kBad, 0)) { ok = false; break; }
            const uint64_t at_dist = buf;
            RSQ_DROP((int)(d & 0xffu));
            const int dx = (int)((d >> 8) & 15u);
            const uint32_t dist = (d >> 16) + ((uint32_t)(at_dist >> ((d & 0xffu) - dx)) & ((1u << dx) - 1u));
            if (__builtin_expect((size_t)dist > (size_t)(out - out_begin), 0)) { ok = false; break; }
            const uint8_t* src = out - dist;
            uint8_t* dst = out;
            out += length;
            if (dist >= 8) {
                // words in ascending order: each source word was complete before its target is written; three of them without a
                // test (most matches are short: no loop, no mispredicted exit), the rest in pairs
                memcpy(dst, src, 8); memcpy(dst + 8, src + 8, 8); memcpy(dst + 16, src + 16, 8);
                if (length > 24) {
                    dst += 24; src += 24;
                    do { memcpy(dst, src, 8); memcpy(dst + 8, src + 8, 8); dst += 16; src += 16; } while (dst < out);
                }
            } else if (dist == 1) {
                const uint64_t v = 0x0101010101010101ull * (uint64_t)*src;
                do { memcpy(dst, &v, 8); dst += 8; } while (dst < out);
            } else {
                do { *dst++ = *src++; } while (dst < out);
            }
        }
#undef RSQ_DROP
        b.buf = buf; b.cnt = cnt; b.in = in;
        if (!ok) return false;
        if (done) { out_io = out; return true; }
    }
    // ---- careful loop: the last bytes of the input or the output ---------------------------------------------------------------
    for (;;) {
        b.refill();
        const uint32_t e = lookup(lt, kLitRoot, b.buf);
        const int len = (int)(e & 0xffu);                              // code and extra bits together
        if ((e & kBad) || b.cnt < len) return false;
        if (e & kLit) {
            if (out == out_end) return false;
            b.drop(len);
            *out++ = (uint8_t)(e >> 16);
            continue;
        }
        if (e & kEob) { b.drop(len); out_io = out; return true; }
        const int lx = (int)((e >> 8) & 15u);
        const uint32_t length = (e >> 16) + ((uint32_t)(b.buf >> (len - lx)) & ((1u << lx) - 1u));
        b.drop(len);
        const uint32_t d = lookup(dt, kDistRoot, b.buf);
        const int dl = (int)(d & 0xffu);
        if ((d & kBad) || b.cnt < dl) return false;
        const int dx = (int)((d >> 8) & 15u);
        const uint32_t dist = (d >> 16) + ((uint32_t)(b.buf >> (dl - dx)) & ((1u << dx) - 1u));
        b.drop(dl);
        if ((size_t)dist > (size_t)(out - out_begin) || (size_t)length > (size_t)(out_end - out)) return false;
        const uint8_t* src = out - dist;
        for (uint32_t k = 0; k < length; ++k) out[k] = src[k];
        out += length;
    }
}

}  // namespace

bool fast_inflate(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
    Bits b;
    b.in = in; b.in_end = in + in_len;
    uint8_t* const out_begin = out;
    uint8_t* const out_end = out + out_len;
    uint32_t lt[kLitCap], dt[kDistCap], pt[kPreCap];
    uint8_t lens[320];
    for (;;) {
        b.refill();
        uint32_t final_block, type;
        if (!b.take(1, final_block) || !b.take(2, type)) return false;
        if (type == 0) {                                              // stored: back to the byte boundary, LEN, ~LEN, bytes
            b.drop(b.cnt & 7);
            b.in -= b.cnt >> 3;
            b.buf = 0; b.cnt = 0;
            if (b.in_end - b.in < 4) return false;
            const uint32_t len = (uint32_t)b.in[0] | ((uint32_t)b.in[1] << 8), nlen = (uint32_t)b.in[2] | ((uint32_t)b.in[3] << 8);
            if ((len ^ nlen) != 0xffffu) return false;
            b.in += 4;
            if ((size_t)(b.in_end - b.in) < len || (size_t)(out_end - out) < len) return false;
            memcpy(out, b.in, len);
            b.in += len; out += len;
        } else if (type == 1) {                                       // fixed code, RFC 1951 3.2.6
            int i = 0;
            for (; i < 144; ++i) lens[i] = 8;
            for (; i < 256; ++i) lens[i] = 9;
            for (; i < 280; ++i) lens[i] = 7;
            for (; i < 288; ++i) lens[i] = 8;
            if (!build_table(lens, 288, kLitLen, kLitRoot, lt, kLitCap)) return false;
            for (i = 0; i < 32; ++i) lens[i] = 5;
            if (!build_table(lens, 32, kDist, kDistRoot, dt, kDistCap)) return false;
            if (!huffman_block(b, lt, dt, out_begin, out, out_end)) return false;
        } else if (type == 2) {                                       // dynamic code, RFC 1951 3.2.7
            uint32_t hlit, hdist, hclen;
            if (!b.take(5, hlit) || !b.take(5, hdist) || !b.take(4, hclen)) return false;
            const int n_lit = (int)hlit + 257, n_dist = (int)hdist + 1, n_pre = (int)hclen + 4;
            if (n_lit > 286 || n_dist > 30) return false;
            uint8_t pre[19] = {0};
            b.refill();                                               // 19 x 3 bits = 57 > 56: refill once more on the way
            for (int k = 0; k < n_pre; ++k) {
                if (k == 12) b.refill();
                uint32_t v;
                if (!b.take(3, v)) return false;
                pre[kPreOrder[k]] = (uint8_t)v;
            }
            if (!build_table(pre, 19, kPrecode, kPreRoot, pt, kPreCap)) return false;
            const int total = n_lit + n_dist;
            int at = 0;
            while (at < total) {
                b.refill();
                const uint32_t e = pt[b.peek(kPreRoot)];
                const int len = (int)(e & 0xffu);
                if (b.cnt < len) return false;
                b.drop(len);
                const uint32_t sym = e >> 16;
                if (sym < 16) { lens[at++] = (uint8_t)sym; continue; }
                uint32_t rep, v;
                uint8_t what = 0;
                if (sym == 16) {
                    if (at == 0 || !b.take(2, v)) return false;
                    rep = 3 + v; what = lens[at - 1];
                } else if (sym == 17) {
                    if (!b.take(3, v)) return false;
                    rep = 3 + v;
                } else {
                    if (!b.take(7, v)) return false;
                    rep = 11 + v;
                }
                if (at + (int)rep > total) return false;
                memset(lens + at, what, rep);
                at += (int)rep;
            }
            if (lens[256] == 0) return false;                         // no end-of-block code
            if (!build_table(lens, n_lit, kLitLen, kLitRoot, lt, kLitCap)) return false;
            if (!build_table(lens + n_lit, n_dist, kDist, kDistRoot, dt, kDistCap)) return false;
            if (!huffman_block(b, lt, dt, out_begin, out, out_end)) return false;
        } else {
            return false;
        }
        if (final_block) return out == out_end;
    }
}

}  // namespace rsqh
