// alnreader.cpp — SAM text and BAM (BGZF) readers + record conversion (see alnreader.h).
#include "alnreader.h"
#include "fastinflate.h"

#include <zlib.h>

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <unordered_map>

namespace rsqh {

uint64_t hash_qname(const char* s, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; ++i) {
        h ^= (unsigned char)s[i];
        h *= 0x100000001b3ull;
    }
    h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ull;
    h ^= h >> 27; h *= 0x94d049bb133111ebull;
    h ^= h >> 31;
    return h;
}

namespace {

const char kCigarOps[] = "MIDNSHP=XB";

// -------------------------------------------------------------------------------------------------------------
// SAM text
// -------------------------------------------------------------------------------------------------------------
class SamReader : public AlnReader {
public:
    explicit SamReader(const std::string& path) : in_(path) {
        if (!in_.is_open()) throw HostError(10, "Unable to open BAM file: " + path);
        // header
        while (in_.peek() == '@') {
            std::string line;
            std::getline(in_, line);
            if (line.compare(0, 3, "@SQ") == 0) {
                std::string name;
                int64_t len = 0;
                size_t p = 3;
                while (p < line.size()) {
                    size_t q = line.find('\t', p + 1);
                    std::string f = line.substr(p + 1, q == std::string::npos ? std::string::npos : q - p - 1);
                    if (f.compare(0, 3, "SN:") == 0) name = f.substr(3);
                    else if (f.compare(0, 3, "LN:") == 0) len = std::atoll(f.c_str() + 3);
                    if (q == std::string::npos) break;
                    p = q;
                }
                ref_index_.emplace(name, (int)hdr_.ref_names.size());
                hdr_.ref_names.push_back(name);
                hdr_.ref_lens.push_back(len);
            }
        }
    }
    const AlnHeader& header() const override { return hdr_; }

    bool next(RawAln& a) override {
        while (std::getline(in_, line_)) {
            if (line_.empty()) continue;
            parse(a);
            return true;
        }
        return false;
    }

private:
    void parse(RawAln& a) {
        const char* s = line_.c_str();
        const char* fields[11];
        size_t lens[11];
        const char* p = s;
        int nf = 0;
        const char* end = s + line_.size();
        const char* rest = nullptr;
        while (nf < 11) {
            const char* t = (const char*)memchr(p, '\t', end - p);
            fields[nf] = p;
            lens[nf] = (t ? t : end) - p;
            ++nf;
            if (!t) { rest = end; break; }
            p = t + 1;
            rest = p;
        }
        if (nf < 11) throw HostError(-1, "truncated SAM record: " + line_);
        a.qname.assign(fields[0], lens[0]);
        a.flag = (uint16_t)std::strtol(fields[1], nullptr, 10);
        std::string rname(fields[2], lens[2]);
        a.tid = lookup(rname);
        a.pos = (int32_t)std::strtol(fields[3], nullptr, 10) - 1;
        a.mapq = (uint8_t)std::strtol(fields[4], nullptr, 10);
        a.cigar.clear();
        if (!(lens[5] == 1 && fields[5][0] == '*')) {
            const char* c = fields[5];
            const char* ce = c + lens[5];
            while (c < ce) {
                char* e2;
                unsigned long n = std::strtoul(c, &e2, 10);
                if (e2 == c || e2 >= ce) throw HostError(-1, "malformed CIGAR: " + line_);
                const char* op = strchr(kCigarOps, *e2);
                if (!op || !*e2) throw HostError(-1, "unrecognized CIGAR operator: " + line_);
                a.cigar.push_back((uint32_t)(n << 4) | (uint32_t)(op - kCigarOps));
                c = e2 + 1;
            }
        }
        if (lens[6] == 1 && fields[6][0] == '=') a.mtid = a.tid;
        else if (lens[6] == 1 && fields[6][0] == '*') a.mtid = -1;
        else a.mtid = lookup(std::string(fields[6], lens[6]));
        a.mpos = (int32_t)std::strtol(fields[7], nullptr, 10) - 1;
        a.isize = (int32_t)std::strtol(fields[8], nullptr, 10);
        a.l_qseq = (lens[9] == 1 && fields[9][0] == '*') ? 0 : (int32_t)lens[9];
        a.aux.clear();
        // optional fields TAG:TYPE:VALUE
        const char* q = (nf == 11 && rest && rest <= end) ? rest : end;
        // `rest` points just past the tab following QUAL when one exists
        if (fields[10] + lens[10] >= end) q = end;
        while (q < end) {
            const char* t = (const char*)memchr(q, '\t', end - q);
            const char* fe = t ? t : end;
            if (fe - q >= 5 && q[2] == ':' && q[4] == ':') {
                AuxField f;
                f.tag[0] = q[0]; f.tag[1] = q[1];
                f.type = q[3];
                f.ival = 0;
                if (f.type == 'i') f.ival = std::strtoll(q + 5, nullptr, 10);
                else if (f.type == 'A') f.ival = (unsigned char)q[5];
                a.aux.push_back(f);
            }
            if (!t) break;
            q = t + 1;
        }
    }
    int lookup(const std::string& n) {
        if (n == "*") return -1;
        auto it = ref_index_.find(n);
        return it == ref_index_.end() ? -1 : it->second;
    }
    std::ifstream in_;
    AlnHeader hdr_;
    std::unordered_map<std::string, int> ref_index_;
    std::string line_;
};

// -------------------------------------------------------------------------------------------------------------
// BAM over BGZF
// -------------------------------------------------------------------------------------------------------------
class BgzfStream {
public:
    explicit BgzfStream(const std::string& path) {
        fp_ = std::fopen(path.c_str(), "rb");
        if (!fp_) throw HostError(10, "Unable to open BAM file: " + path);
    }
    ~BgzfStream() { if (fp_) std::fclose(fp_); }
    // reads exactly n bytes unless EOF; returns bytes read
    size_t read(void* dst, size_t n) {
        size_t got = 0;
        uint8_t* d = (uint8_t*)dst;
        while (got < n) {
            if (off_ == out_.size()) {
                if (!fill()) break;
                continue;
            }
            size_t k = std::min(n - got, out_.size() - off_);
            memcpy(d + got, out_.data() + off_, k);
            off_ += k;
            got += k;
        }
        return got;
    }

private:
    bool fill() {
        uint8_t h[18];
        size_t r = std::fread(h, 1, 18, fp_);
        if (r == 0) return false;
        if (r < 18 || h[0] != 0x1f || h[1] != 0x8b) throw HostError(-1, "corrupt BGZF block header");
        unsigned xlen = h[10] | (h[11] << 8);
        // fixed-layout BC subfield is the norm; general case: scan extra field
        std::vector<uint8_t> extra(xlen);
        memcpy(extra.data(), h + 12, std::min<size_t>(6, xlen));
        if (xlen > 6 && std::fread(extra.data() + 6, 1, xlen - 6, fp_) != xlen - 6)
            throw HostError(-1, "truncated BGZF block");
        int bsize = -1;
        for (size_t i = 0; i + 4 <= xlen;) {
            unsigned slen = extra[i + 2] | (extra[i + 3] << 8);
            if (extra[i] == 'B' && extra[i + 1] == 'C' && slen == 2) bsize = extra[i + 4] | (extra[i + 5] << 8);
            i += 4 + slen;
        }
        if (bsize < 0) throw HostError(-1, "BGZF block without BC subfield");
        size_t cdata = (size_t)bsize + 1 - 12 - xlen - 8;
        in_.resize(cdata + 8);
        if (std::fread(in_.data(), 1, cdata + 8, fp_) != cdata + 8) throw HostError(-1, "truncated BGZF block");
        uint32_t isize;
        memcpy(&isize, in_.data() + cdata + 4, 4);
        out_.resize(isize);
        off_ = 0;
        if (isize == 0) return true;  // empty block (EOF marker): keep going
        if (!std::getenv("RSQ_INGEST_ZLIB") && fast_inflate(in_.data(), cdata, out_.data(), isize)) return true;   // fastinflate.h; zlib decides the rest
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) throw HostError(-1, "zlib init failed");
        zs.next_in = in_.data();
        zs.avail_in = (uInt)cdata;
        zs.next_out = out_.data();
        zs.avail_out = isize;
        int rc = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        if (rc != Z_STREAM_END) throw HostError(-1, "BGZF inflate failed");
        return true;
    }
    FILE* fp_ = nullptr;
    std::vector<uint8_t> in_, out_;
    size_t off_ = 0;
};

class BamReader : public AlnReader {
public:
    explicit BamReader(const std::string& path) : z_(path) {
        char magic[4];
        if (z_.read(magic, 4) != 4 || memcmp(magic, "BAM\1", 4) != 0) throw HostError(-1, "not a BAM file: " + path);
        int32_t l_text = rd32();
        std::string text(l_text, '\0');
        if (l_text && z_.read(&text[0], l_text) != (size_t)l_text) throw HostError(-1, "truncated BAM header");
        int32_t n_ref = rd32();
        for (int i = 0; i < n_ref; ++i) {
            int32_t l_name = rd32();
            std::string name(l_name, '\0');
            z_.read(&name[0], l_name);
            if (!name.empty() && name.back() == '\0') name.pop_back();
            int32_t l_ref = rd32();
            hdr_.ref_names.push_back(name);
            hdr_.ref_lens.push_back(l_ref);
        }
    }
    const AlnHeader& header() const override { return hdr_; }
    bool next(RawAln& a) override {
        int32_t block_size;
        size_t r = z_.read(&block_size, 4);
        if (r == 0) return false;
        if (r != 4 || block_size < 32) throw HostError(-1, "truncated BAM record");
        buf_.resize(block_size);
        if (z_.read(buf_.data(), block_size) != (size_t)block_size) throw HostError(-1, "truncated BAM record");
        const uint8_t* b = buf_.data();
        auto i32 = [&](size_t o) { int32_t v; memcpy(&v, b + o, 4); return v; };
        auto u16 = [&](size_t o) { uint16_t v; memcpy(&v, b + o, 2); return v; };
        a.tid = i32(0);
        a.pos = i32(4);
        uint8_t l_read_name = b[8];
        a.mapq = b[9];
        uint16_t n_cigar = u16(12);
        a.flag = u16(14);
        a.l_qseq = i32(16);
        a.mtid = i32(20);
        a.mpos = i32(24);
        a.isize = i32(28);
        size_t o = 32;
        a.qname.assign((const char*)b + o, l_read_name ? l_read_name - 1 : 0);
        o += l_read_name;
        a.cigar.resize(n_cigar);
        if (n_cigar) memcpy(a.cigar.data(), b + o, 4 * (size_t)n_cigar);
        o += 4 * (size_t)n_cigar;
        o += ((size_t)a.l_qseq + 1) / 2 + (size_t)a.l_qseq;
        a.aux.clear();
        const size_t n = buf_.size();
        while (o + 3 <= n) {
            AuxField f;
            f.tag[0] = (char)b[o]; f.tag[1] = (char)b[o + 1];
            f.type = (char)b[o + 2];
            f.ival = 0;
            o += 3;
            switch (f.type) {
                case 'A': f.ival = b[o]; o += 1; break;
                case 'c': f.ival = (int8_t)b[o]; o += 1; break;
                case 'C': f.ival = b[o]; o += 1; break;
                case 's': { int16_t v; memcpy(&v, b + o, 2); f.ival = v; o += 2; break; }
                case 'S': { uint16_t v; memcpy(&v, b + o, 2); f.ival = v; o += 2; break; }
                case 'i': { int32_t v; memcpy(&v, b + o, 4); f.ival = v; o += 4; break; }
                case 'I': { uint32_t v; memcpy(&v, b + o, 4); f.ival = v; o += 4; break; }
                case 'f': o += 4; break;
                case 'd': o += 8; break;
                case 'Z': case 'H': while (o < n && b[o]) ++o; ++o; break;
                case 'B': {
                    char sub = (char)b[o];
                    int32_t cnt; memcpy(&cnt, b + o + 1, 4);
                    size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                    o += 5 + es * (size_t)cnt;
                    break;
                }
                default: throw HostError(-1, "corrupt BAM aux field");
            }
            a.aux.push_back(f);
        }
        return true;
    }

private:
    int32_t rd32() {
        int32_t v;
        if (z_.read(&v, 4) != 4) throw HostError(-1, "truncated BAM header");
        return v;
    }
    BgzfStream z_;
    AlnHeader hdr_;
    std::vector<uint8_t> buf_;
};

}  // namespace

std::unique_ptr<AlnReader> AlnReader::open(const std::string& path, const std::string& fasta, int threads) {
    FILE* fp = std::fopen(path.c_str(), "rb");
    if (!fp) throw HostError(10, "Unable to open BAM file: " + path);
    unsigned char m[4] = {0, 0, 0, 0};
    size_t r = std::fread(m, 1, 4, fp);
    std::fclose(fp);
    if (r >= 2 && m[0] == 0x1f && m[1] == 0x8b) return std::unique_ptr<AlnReader>(new BamReader(path));
    if (r >= 4 && memcmp(m, "CRAM", 4) == 0) return open_cram(path, fasta, threads);
    return std::unique_ptr<AlnReader>(new SamReader(path));
}

// -------------------------------------------------------------------------------------------------------------
// Decoder
// -------------------------------------------------------------------------------------------------------------
Decoder::Decoder(const AlnHeader& hdr, ContigMap& contigs, size_t n_annotation_contigs, const DecodeOptions& opt)
    : opt_(opt), n_annotation_contigs_(n_annotation_contigs) {
    // chromosomeMap(sequence->Name) for every header sequence, in order (src/RNASeQC.cpp:224-227)
    for (const std::string& n : hdr.ref_names) tid_to_contig_.push_back((int)contigs.get(n));
    dropped_.assign(contigs.names.size() + 1, 0);
}

bool Decoder::any_overlap_with(const std::vector<uint8_t>& contig_has_features) const {
    for (int c : tid_to_contig_)
        if (c >= 0 && (size_t)c < contig_has_features.size() && contig_has_features[c]) return true;
    return false;
}

void Decoder::convert(const RawAln& a, std::vector<rsq_read>& reads, std::vector<rsq_block>& blocks) {
    rsq_read r;
    memset(&r, 0, sizeof(r));
    r.pos = a.pos;
    r.mpos = a.mpos;
    r.isize = a.isize;
    r.l_qseq = a.l_qseq;
    r.flag = a.flag;
    r.mapq = a.mapq;
    r.name_hash = hash_qname(a.qname.data(), a.qname.size());

    // aux: first occurrence wins (bam_aux_get)
    bool nm_seen = false, ch_seen = false;
    std::vector<char> filt_seen(opt_.filter_tags.size(), 0);
    for (const AuxField& f : a.aux) {
        const bool is_int = f.type == 'i' || f.type == 'C' || f.type == 'S' || f.type == 's' || f.type == 'I' || f.type == 'c';
        if (!nm_seen && f.tag[0] == 'N' && f.tag[1] == 'M') {
            nm_seen = true;
            if (is_int) { r.aux |= RSQ_AUX_HAS_NM; r.nm = (int32_t)f.ival; }  // GetIntTag, BamRecord.h:740-750
        }
        if (!ch_seen && opt_.chimeric_tag.size() == 2 && f.tag[0] == opt_.chimeric_tag[0] && f.tag[1] == opt_.chimeric_tag[1]) {
            ch_seen = true;
            // readStringTag (src/RNASeQC.cpp:780-800): a Z tag, or an A tag whose character is not NUL
            if (f.type == 'Z' || (f.type == 'A' && f.ival != 0)) r.aux |= RSQ_AUX_CHIMERIC_TAG;
        }
        for (size_t k = 0; k < opt_.filter_tags.size(); ++k) {
            const std::string& t = opt_.filter_tags[k];
            if (!filt_seen[k] && t.size() == 2 && f.tag[0] == t[0] && f.tag[1] == t[1]) {
                filt_seen[k] = 1;
                // GetTag (SeqLib/src/BamRecord.cpp:474-497): Z, any int type, or float
                if (f.type == 'Z' || is_int || f.type == 'f' || f.type == 'd') r.filter_mask |= (uint8_t)(1u << k);
            }
        }
    }
    if (a.tid != a.mtid) r.aux |= RSQ_AUX_MATE_OTHER_CONTIG;
    const int n_refs = (int)tid_to_contig_.size();
    const bool tid_ok = a.tid >= 0 && a.tid < n_refs;
    if (!tid_ok) r.aux |= RSQ_AUX_TID_OUT_OF_RANGE;
    if (a.tid > 127) r.aux |= RSQ_AUX_TID_GT_127;
    const int chr = tid_ok ? tid_to_contig_[a.tid] : -1;
    r.contig = (chr >= 0 && (size_t)chr < n_annotation_contigs_ && chr <= INT16_MAX) ? (int16_t)chr : (int16_t)-1;

    // extractBlocks (src/Expression.cpp:26-67) and bam_cigar2rlen / NumQueryConsumed for PositionEnd
    r.block_off = (uint32_t)blocks.size();
    int64_t start = (int64_t)a.pos + 1;
    int64_t rlen = 0, qlen = 0;
    unsigned nb = 0;
    for (uint32_t c : a.cigar) {
        const unsigned op = c & 0xf;
        const int64_t len = c >> 4;
        switch (op) {
            case 0: case 7: case 8:  // M = X
                blocks.push_back(rsq_block{(int32_t)start, (int32_t)len});
                ++nb;
                start += len; rlen += len; qlen += len;
                break;
            case 2: case 3:          // D N
                start += len; rlen += len;
                break;
            case 1: case 4:          // I S
                qlen += len;
                break;
            case 5: case 6:          // H P
                break;
            default:
                std::cerr << "Bad cigar operation: " << (op < 10 ? kCigarOps[op] : '?') << std::endl;
                throw HostError(7, "Invalid argument type provided: Unrecognized Cigar Op ");
        }
    }
    if (nb > 0xffff) throw HostError(-1, "alignment with more than 65535 aligned blocks");
    r.n_blocks = (uint16_t)nb;
    // PositionEnd (SeqLib/src/BamRecord.cpp:51-53) over bam_endpos (htslib/sam.c:347-353)
    if (a.l_qseq > 0) r.endpos = (!(a.flag & RSQ_F_UNMAP) && !a.cigar.empty()) ? (int32_t)(a.pos + rlen) : a.pos + 1;
    else r.endpos = (int32_t)(a.pos + qlen);

    observe(r, a.tid);
    reads.push_back(r);
}

// Pure conversion of one raw BAM record (no RawAln intermediate): same field semantics as convert(RawAln).
void Decoder::convert_bam_bytes(const uint8_t* b, size_t n, rsq_read& r, std::vector<rsq_block>& blocks, int32_t& tid_out) const {
    convert_bam_record(opt_, (int)tid_to_contig_.size(), b, n, r, blocks, tid_out);
    bind_contig(r, tid_out);
}

void Decoder::convert_bam_record(const DecodeOptions& opt_, int n_refs, const uint8_t* b, size_t n, rsq_read& r, std::vector<rsq_block>& blocks, int32_t& tid_out) {
    if (n < 32) throw HostError(-1, "truncated BAM record");
    auto i32 = [&](size_t o) { int32_t v; memcpy(&v, b + o, 4); return v; };
    auto u16 = [&](size_t o) { uint16_t v; memcpy(&v, b + o, 2); return v; };
    memset(&r, 0, sizeof(r));
    const int32_t tid = i32(0), mtid = i32(20);
    tid_out = tid;
    r.pos = i32(4);
    const uint8_t l_read_name = b[8];
    r.mapq = b[9];
    const uint16_t n_cigar = u16(12);
    r.flag = u16(14);
    r.l_qseq = i32(16);
    r.mpos = i32(24);
    r.isize = i32(28);
    size_t o = 32;
    if (o + l_read_name + 4 * (size_t)n_cigar > n) throw HostError(-1, "truncated BAM record");
    r.name_hash = hash_qname((const char*)b + o, l_read_name ? l_read_name - 1 : 0);
    o += l_read_name;
    // CIGAR -> blocks
    r.block_off = (uint32_t)blocks.size();
    int64_t start = (int64_t)r.pos + 1, rlen = 0, qlen = 0;
    unsigned nb = 0;
    for (unsigned k = 0; k < n_cigar; ++k) {
        uint32_t c; memcpy(&c, b + o + 4 * k, 4);
        const unsigned op = c & 0xf;
        const int64_t len = c >> 4;
        switch (op) {
            case 0: case 7: case 8: blocks.push_back(rsq_block{(int32_t)start, (int32_t)len}); ++nb; start += len; rlen += len; qlen += len; break;
            case 2: case 3: start += len; rlen += len; break;
            case 1: case 4: qlen += len; break;
            case 5: case 6: break;
            default:
                std::cerr << "Bad cigar operation: " << (op < 10 ? kCigarOps[op] : '?') << std::endl;
                throw HostError(7, "Invalid argument type provided: Unrecognized Cigar Op ");
        }
    }
    if (nb > 0xffff) throw HostError(-1, "alignment with more than 65535 aligned blocks");
    r.n_blocks = (uint16_t)nb;
    if (r.l_qseq > 0) r.endpos = (!(r.flag & RSQ_F_UNMAP) && n_cigar) ? (int32_t)(r.pos + rlen) : r.pos + 1;
    else r.endpos = (int32_t)(r.pos + qlen);
    o += 4 * (size_t)n_cigar;
    o += ((size_t)r.l_qseq + 1) / 2 + (size_t)r.l_qseq;
    // aux: first occurrence of each tag wins (bam_aux_get)
    bool nm_seen = false, ch_seen = false;
    unsigned filt_seen = 0;
    const char c0 = opt_.chimeric_tag.size() == 2 ? opt_.chimeric_tag[0] : 0, c1 = opt_.chimeric_tag.size() == 2 ? opt_.chimeric_tag[1] : 0;
    while (o + 3 <= n) {
        const char t0 = (char)b[o], t1 = (char)b[o + 1], type = (char)b[o + 2];
        o += 3;
        int64_t ival = 0;
        switch (type) {
            case 'A': ival = b[o]; o += 1; break;
            case 'c': ival = (int8_t)b[o]; o += 1; break;
            case 'C': ival = b[o]; o += 1; break;
            case 's': { int16_t v; memcpy(&v, b + o, 2); ival = v; o += 2; break; }
            case 'S': { uint16_t v; memcpy(&v, b + o, 2); ival = v; o += 2; break; }
            case 'i': { int32_t v; memcpy(&v, b + o, 4); ival = v; o += 4; break; }
            case 'I': { uint32_t v; memcpy(&v, b + o, 4); ival = v; o += 4; break; }
            case 'f': o += 4; break;
            case 'd': o += 8; break;
            case 'Z': case 'H': while (o < n && b[o]) ++o; ++o; break;
            case 'B': {
                if (o + 5 > n) throw HostError(-1, "corrupt BAM aux field");
                const char sub = (char)b[o];
                int32_t cnt; memcpy(&cnt, b + o + 1, 4);
                const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                o += 5 + es * (size_t)cnt;
                break;
            }
            default: throw HostError(-1, "corrupt BAM aux field");
        }
        const bool is_int = type == 'i' || type == 'C' || type == 'S' || type == 's' || type == 'I' || type == 'c';
        if (!nm_seen && t0 == 'N' && t1 == 'M') {
            nm_seen = true;
            if (is_int) { r.aux |= RSQ_AUX_HAS_NM; r.nm = (int32_t)ival; }
        }
        if (!ch_seen && c0 && t0 == c0 && t1 == c1) {
            ch_seen = true;
            if (type == 'Z' || (type == 'A' && ival != 0)) r.aux |= RSQ_AUX_CHIMERIC_TAG;
        }
        for (size_t k = 0; k < opt_.filter_tags.size(); ++k) {
            const std::string& t = opt_.filter_tags[k];
            if (!(filt_seen & (1u << k)) && t.size() == 2 && t0 == t[0] && t1 == t[1]) {
                filt_seen |= 1u << k;
                if (type == 'Z' || is_int || type == 'f' || type == 'd') r.filter_mask |= (uint8_t)(1u << k);
            }
        }
    }
    if (tid != mtid) r.aux |= RSQ_AUX_MATE_OTHER_CONTIG;
    const bool tid_ok = tid >= 0 && tid < n_refs;
    if (!tid_ok) r.aux |= RSQ_AUX_TID_OUT_OF_RANGE;
    if (tid > 127) r.aux |= RSQ_AUX_TID_GT_127;
    r.contig = (int16_t)-1;                                                      // Decoder::bind_contig
}

}  // namespace rsqh
