// annotation.cpp — GTF/BED -> flattened annotation (see annotation.h).
#include "annotation.h"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <set>
#include <unordered_set>

namespace rsqh {

namespace {

// The 12 globin gene names the reference excludes (src/Expression.cpp:24).
const char* const kGlobins[] = {"HBA1", "HBA2", "HBB",  "HBD", "HBG1",  "HBG2",
                                "HBE1", "HBM",  "HBQ1", "HBZ", "HBBP1", "HBZP1"};

struct Span {      // a piece of the GTF text held in memory
    const char* p; size_t n;
    std::string str() const { return std::string(p, n); }
    bool is(const char* lit, size_t ln) const { return n == ln && !memcmp(p, lit, ln); }
};

// std::stoull semantics as used at src/GTF.cpp:54,57: skip leading whitespace, optional sign, digits;
// no digits -> invalid_argument (-> gtfException, exit 11).
bool parse_u64(Span s, unsigned long long& v) {
    char tmp[48];
    std::string big;
    const char* p;
    if (s.n < sizeof tmp) { memcpy(tmp, s.p, s.n); tmp[s.n] = 0; p = tmp; }   // bounded copy: strtoull must not run into the next field
    else { big = s.str(); p = big.c_str(); }
    char* endp = nullptr;
    errno = 0;
    v = std::strtoull(p, &endp, 10);
    return endp != p;
}

// parseAttributes (src/GTF.cpp:146-161): split on every ';' (quotes are NOT honoured), key = text before
// the first '"' minus its last character, left-trimmed of blanks/tabs; value = text up to the next '"'.
// A piece with no quote at all leaves the value equal to the whole piece (std::getline does not clear its
// target once eofbit is set).  Later duplicates overwrite earlier ones (std::map assignment).  Only the five keys the
// reference ever looks up are kept (no per-line std::map: this is the hot loop of GTF loading).
struct Attrs {
    bool has_gene_id = false, has_transcript_id = false, has_exon_id = false, has_transcript_type = false, has_gene_name = false;
    std::string gene_id, transcript_id, exon_id, transcript_type, gene_name;
    void clear() { has_gene_id = has_transcript_id = has_exon_id = has_transcript_type = has_gene_name = false; }
};

void parse_attributes(Span text, Attrs& out) {
    size_t pos = 0;
    const size_t n = text.n;
    const char* t = text.p;
    auto find = [&](char ch, size_t from) -> size_t {
        if (from >= n) return std::string::npos;
        const void* q = memchr(t + from, ch, n - from);
        return q ? (size_t)((const char*)q - t) : std::string::npos;
    };
    size_t q1 = find('"', 0);                                            // the next quote at or after pos (kept across pieces)
    while (pos < n) {  // getline(tokenizer, buffer, ';') stops yielding once the input is exhausted
        const size_t semi = find(';', pos);
        const size_t pe = semi == std::string::npos ? n : semi;          // piece = [pos, pe)
        if (q1 != std::string::npos && q1 < pos) q1 = find('"', pos);
        const bool has_q = q1 != std::string::npos && q1 < pe;
        size_t ks = pos, ke = has_q ? q1 : pe;                            // head = [ks, ke)
        if (ke > ks) --ke;                                                // drop the last character of the head
        while (ks < ke && (t[ks] == ' ' || t[ks] == '\t')) ++ks;
        const size_t kl = ke - ks;
        std::string* dst = nullptr;
        bool* flag = nullptr;
        switch (kl) {
            case 7:
                if (!memcmp(t + ks, "gene_id", 7)) { dst = &out.gene_id; flag = &out.has_gene_id; }
                else if (!memcmp(t + ks, "exon_id", 7)) { dst = &out.exon_id; flag = &out.has_exon_id; }
                break;
            case 9: if (!memcmp(t + ks, "gene_name", 9)) { dst = &out.gene_name; flag = &out.has_gene_name; } break;
            case 13: if (!memcmp(t + ks, "transcript_id", 13)) { dst = &out.transcript_id; flag = &out.has_transcript_id; } break;
            case 15: if (!memcmp(t + ks, "transcript_type", 15)) { dst = &out.transcript_type; flag = &out.has_transcript_type; } break;
            default: break;
        }
        if (dst) {
            *flag = true;
            if (!has_q) dst->assign(t + pos, pe - pos);                   // stale `current` == the piece just read
            else {
                const size_t q2 = find('"', q1 + 1);
                const size_t ve = (q2 == std::string::npos || q2 > pe) ? pe : q2;
                dst->assign(t + q1 + 1, ve - q1 - 1);
            }
        }
        pos = (semi == std::string::npos) ? n : semi + 1;
    }
}

bool parse_u64(const std::string& s, unsigned long long& v) { return parse_u64(Span{s.data(), s.size()}, v); }

bool split_tab(Span line, size_t& cursor, Span& field) {
    // getline(tokenizer, buffer, '\t'): fails only when no characters remain at all
    if (cursor >= line.n) return false;   // stream exhausted: a previous read consumed the last char; getline extracts nothing
    const void* q = memchr(line.p + cursor, '\t', line.n - cursor);
    if (!q) {
        field = Span{line.p + cursor, line.n - cursor};
        cursor = line.n + 1;
    } else {
        const size_t tab = (size_t)((const char*)q - line.p);
        field = Span{line.p + cursor, tab - cursor};
        cursor = tab + 1;                 // a trailing tab leaves cursor == size: the next getline yields nothing
    }
    return true;
}

}  // namespace

uint32_t Annotation::gene_universe(const std::string& id) {
    auto it = gene_index.find(id);
    if (it != gene_index.end()) return it->second;
    uint32_t g = (uint32_t)gene_ids.size();
    gene_ids.push_back(id);
    gene_index.emplace(id, g);
    gene_coding_len.push_back(0);
    exons_for_gene.emplace_back();
    return g;
}

void load_gtf(const std::string& path, Annotation& A, bool legacy) {
    // RSQ_SETUP_STATS=1: where the time of the start-up goes (like rsq_set_annotation in the library)
    const bool setup_stats = getenv("RSQ_SETUP_STATS") != nullptr;
    auto t_mark = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!setup_stats) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "load_gtf: %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_mark).count());
        t_mark = now;
    };
    // the whole file in memory, lines and fields as views into it (the loop below runs ~0.5 M times for a GENCODE-scale GTF)
    std::string text;
    {
        std::ifstream in(path, std::ios::binary);
        if (!in.is_open()) throw HostError(10, "Unable to open GTF file: " + path);
        in.seekg(0, std::ios::end);
        const std::streamoff sz = in.tellg();
        if (sz > 0) {
            text.resize((size_t)sz);
            in.seekg(0, std::ios::beg);
            in.read(&text[0], sz);
            text.resize((size_t)std::max<std::streamsize>(in.gcount(), 0));
        } else {                                                     // not seekable (a pipe): stream it
            in.clear();
            in.seekg(0, std::ios::beg);
            in.clear();
            text.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
        }
    }

    lap("file read");
    // State of the single re-used Feature object (src/RNASeQC.cpp:109): these three survive from line to line
    // because operator>> only assigns them when the attribute exists (src/GTF.cpp:81,89,90,95,110).
    std::string cur_feature_id, cur_gene_id, cur_transcript_type;
    // One table over feature ids: geneNames (src/GTF.cpp:22, keyed by feature_id: gene, transcript and exon ids share it) and
    // the set of exon ids seen so far (the uniqueness check, src/GTF.cpp:101-104).
    struct IdInfo { std::string name; bool named = false, exon_seen = false; };
    std::unordered_map<std::string, IdInfo> ids;
    ids.reserve(text.size() / 200 + 16);
    std::vector<const IdInfo*> exon_info;                             // per exon slot: its entry (looked up again for the names below otherwise)
    std::unordered_set<std::string> seen_gene_ids;
    std::map<std::string, unsigned> unnamed_exon_counter;            // exon_names (src/GTF.cpp:27)
    std::vector<GeneFeature> kept_genes;
    std::vector<ExonFeature> kept_exons;
    size_t kept_counter = 0;
    // gene_universe(cur_gene_id) is asked up to three times per line, and neighbouring lines share the gene
    std::string memo_gene_id; uint32_t memo_gene = 0; bool memo_ok = false;
    auto universe = [&](const std::string& id) -> uint32_t {
        if (memo_ok && id == memo_gene_id) return memo_gene;
        memo_gene = A.gene_universe(id); memo_gene_id = id; memo_ok = true;
        return memo_gene;
    };

    std::string memo_chr; uint32_t memo_contig = 0; bool memo_chr_ok = false;
    Attrs at;
    size_t lp = 0;
    while (lp < text.size()) {
        const void* nl = memchr(text.data() + lp, '\n', text.size() - lp);
        const size_t le = nl ? (size_t)((const char*)nl - text.data()) : text.size();
        const Span line{text.data() + lp, le - lp};
        lp = le + 1;
        if (line.n && line.p[0] == '#') continue;
        size_t cur = 0;
        Span f_chr, f_track, f_type, f_start, f_end, f_score, f_strand, f_frame;
        if (!split_tab(line, cur, f_chr)) throw HostError(11, "Failed to parse the GTF: Unable to parse chromosome. Invalid GTF line: " + line.str());
        if (!(memo_chr_ok && f_chr.n == memo_chr.size() && !memcmp(f_chr.p, memo_chr.data(), f_chr.n))) {   // neighbouring lines share the contig
            memo_chr.assign(f_chr.p, f_chr.n);
            memo_contig = A.contigs.get(memo_chr);
            memo_chr_ok = true;
        }
        const uint32_t contig = memo_contig;
        if (!split_tab(line, cur, f_track)) throw HostError(11, "Failed to parse the GTF: Unable to parse track. Invalid GTF line: " + line.str());
        if (!split_tab(line, cur, f_type)) throw HostError(11, "Failed to parse the GTF: Unable to parse feature type. Invalid GTF line: " + line.str());
        enum { T_GENE, T_TRANSCRIPT, T_EXON, T_OTHER } type =
            f_type.is("exon", 4) ? T_EXON : f_type.is("gene", 4) ? T_GENE : f_type.is("transcript", 10) ? T_TRANSCRIPT : T_OTHER;
        if (!split_tab(line, cur, f_start)) throw HostError(11, "Failed to parse the GTF: Unable to parse start. Invalid GTF line: " + line.str());
        unsigned long long ustart = 0, uend = 0;
        if (!parse_u64(f_start, ustart)) throw HostError(11, "Failed to parse the GTF: GTF is in an invalid format: stoull");
        if (!split_tab(line, cur, f_end)) throw HostError(11, "Failed to parse the GTF: Unable to parse end. Invalid GTF line: " + line.str());
        if (!parse_u64(f_end, uend)) throw HostError(11, "Failed to parse the GTF: GTF is in an invalid format: stoull");
        if (!split_tab(line, cur, f_score)) throw HostError(11, "Failed to parse the GTF: Unable to parse score. Invalid GTF line: " + line.str());
        if (!split_tab(line, cur, f_strand)) throw HostError(11, "Failed to parse the GTF: Unable to parse strand. Invalid GTF line: " + line.str());
        uint8_t strand = !f_strand.n ? RSQ_STRAND_UNKNOWN
                         : f_strand.p[0] == '+' ? RSQ_STRAND_FWD
                         : f_strand.p[0] == '-' ? RSQ_STRAND_REV
                                                : RSQ_STRAND_UNKNOWN;
        if (!split_tab(line, cur, f_frame)) throw HostError(11, "Failed to parse the GTF: Unable to parse frame. Invalid GTF line: " + line.str());
        // attributes = the rest of the line, tabs included (getline with '\n')
        if (cur >= line.n) throw HostError(11, "Failed to parse the GTF: Unable to parse attributes. Invalid GTF line: " + line.str());
        at.clear();                                                                    // the strings keep their storage; only flagged ones are read
        parse_attributes(Span{line.p + cur, line.n - cur}, at);
        const int64_t start = (int64_t)ustart, end = (int64_t)uend;
        if (end < start) std::cerr << "Bad feature range:" << start << " - " << end << std::endl;

        if (type == T_GENE && at.has_gene_id) {
            cur_feature_id = at.gene_id;
            if (!seen_gene_ids.insert(cur_feature_id).second)
                throw HostError(11, "Failed to parse the GTF: Detected non-unique Gene ID: " + cur_feature_id);
            A.gene_list.push_back(universe(cur_feature_id));
        }
        if (type == T_TRANSCRIPT && at.has_transcript_id) cur_feature_id = at.transcript_id;
        if (at.has_gene_id) cur_gene_id = at.gene_id;
        uint32_t exon_slot = 0;
        IdInfo* info = nullptr;                                                        // the table entry of cur_feature_id, once looked up
        if (type == T_EXON) {
            if (at.has_exon_id) {
                cur_feature_id = at.exon_id;
            } else if (at.has_gene_id) {
                const std::string& gid = at.gene_id;
                cur_feature_id = gid + "_" + std::to_string(++unnamed_exon_counter[gid]);
                std::cerr << "Unnamed exon: Gene: " << gid << " Position: [" << start << ", " << end
                          << "] Inferred Exon Name: " << cur_feature_id << std::endl;
            } else {
                throw HostError(11, "Failed to parse the GTF: Exon missing exon_id and gene_id fields: " + line.str());
            }
            info = &ids[cur_feature_id];
            if (info->exon_seen)
                throw HostError(11, "Failed to parse the GTF: Detected non-unique Exon ID: " + cur_feature_id);
            info->exon_seen = true;
            exon_info.push_back(info);                                                 // elements of an unordered_map stay where they are
            exon_slot = (uint32_t)A.exon_ids.size();
            A.exon_ids.push_back(cur_feature_id);
            A.exon_len.push_back(1 + (end - start));
            A.exon_start.push_back(start);
            A.exon_contig.push_back(contig);
            uint32_t g = universe(cur_gene_id);
            A.gene_coding_len[g] += 1 + (end - start);
        }
        if (at.has_transcript_type) cur_transcript_type = at.transcript_type;
        if (at.has_gene_name || at.has_gene_id) {
            if (!info) info = &ids[cur_feature_id];
            info->name = at.has_gene_name ? at.gene_name : at.gene_id;
            info->named = true;
        }
        const bool ribosomal = cur_transcript_type.find("rRNA") != std::string::npos;  // regex "rRNA", src/GTF.cpp:21,113

        if (legacy && end == start) {                                                  // src/RNASeQC.cpp:129-135
            if (type == T_EXON) A.gene_coding_len[universe(cur_gene_id)] -= 1;
            continue;
        }
        // main(): keep genes and exons only (src/RNASeQC.cpp:137-139)
        if (type == T_GENE) {
            GeneFeature gf{contig, start, end, strand, ribosomal, universe(cur_feature_id), (uint32_t)kept_counter++};
            gf.gid = universe(cur_gene_id);
            kept_genes.push_back(gf);
        } else if (type == T_EXON) {
            ExonFeature ef{contig, start, end, strand, ribosomal, universe(cur_gene_id), exon_slot, (uint32_t)kept_counter++};
            kept_exons.push_back(ef);
        }
    }

    lap("line loop");
    if (A.gene_list.empty() || A.exon_ids.empty()) {
        std::cerr << "There were either no genes or no exons in the GTF" << std::endl;
        std::cerr << A.gene_list.size() << " genes parsed" << std::endl;
        std::cerr << A.exon_ids.size() << " exons parsed" << std::endl;
        throw HostError(11, "There were either no genes or no exons in the GTF");
    }

    // geneNames lookups happen after parsing completes, so the final map state is what counts.
    A.gene_desc.assign(A.gene_ids.size(), std::string());
    A.gene_flags.assign(A.gene_ids.size(), 0);
    for (size_t g = 0; g < A.gene_ids.size(); ++g) {
        auto it = ids.find(A.gene_ids[g]);
        if (it != ids.end() && it->second.named) A.gene_desc[g] = it->second.name;
        for (const char* glob : kGlobins)
            if (A.gene_desc[g] == glob) A.gene_flags[g] |= RSQ_GENE_GLOBIN;
    }
    A.exon_desc.assign(A.exon_ids.size(), std::string());
    for (size_t e = 0; e < A.exon_ids.size(); ++e)
        if (exon_info[e]->named) A.exon_desc[e] = exon_info[e]->name;

    lap("names");
    // Per contig, list.sort(compIntervalStart) — a stable sort by start over genes and exons interleaved
    // (src/RNASeQC.cpp:150-152).  Genes and exons are separated here; relative order inside each kind is
    // what the stable sort would give, which is all that exonsForGene and the eviction order depend on.
    const size_t nc = A.contigs.names.size();
    A.genes_by_contig.assign(nc, {});
    A.exons_by_contig.assign(nc, {});
    for (auto& g : kept_genes) A.genes_by_contig[g.contig].push_back(g);
    for (auto& e : kept_exons) A.exons_by_contig[e.contig].push_back(e);
    for (size_t c = 0; c < nc; ++c) {
        std::stable_sort(A.genes_by_contig[c].begin(), A.genes_by_contig[c].end(),
                         [](const GeneFeature& a, const GeneFeature& b) { return a.start < b.start; });
        std::stable_sort(A.exons_by_contig[c].begin(), A.exons_by_contig[c].end(),
                         [](const ExonFeature& a, const ExonFeature& b) { return a.start < b.start; });
        // exonsForGene, filled in chromosome-id order then list order (src/RNASeQC.cpp:150-155)
        for (auto& e : A.exons_by_contig[c]) A.exons_for_gene[e.gene].push_back(e.slot);
    }

    // Coverage layout: one contiguous run per gene in stitched (exonsForGene) order, so a gene's transcript
    // coverage (src/Metrics.cpp:307-308) is a single contiguous slice of the device array.
    A.exon_cov_off.assign(A.exon_ids.size(), 0);
    uint64_t off = 0;
    for (size_t g = 0; g < A.gene_ids.size(); ++g)
        for (uint32_t slot : A.exons_for_gene[g]) {
            A.exon_cov_off[slot] = off;
            off += (uint64_t)std::max<int64_t>(A.exon_len[slot], 0);
        }
    A.cov_len = off;

    // ---- flatten ------------------------------------------------------------------------------------------
    A.f_exon_off.assign(nc + 1, 0);
    A.f_seg_off.assign(nc + 1, 0);
    A.f_gline_off.assign(nc + 1, 0);
    for (size_t c = 0; c < nc; ++c) {
        A.f_gline_off[c] = (uint32_t)A.f_gline_start.size();
        for (auto& g : A.genes_by_contig[c]) {
            if (g.start > INT32_MAX - 2 || g.end > INT32_MAX - 2)
                throw HostError(11, "Failed to parse the GTF: coordinate beyond 2^31");
            A.f_gline_start.push_back((int32_t)g.start);
            A.f_gline_end.push_back((int32_t)g.end);
            A.f_gline_gene.push_back(g.gid);
            A.f_gline_fid.push_back(g.gene);
            A.f_gline_flags.push_back((uint8_t)(g.strand | (g.ribosomal ? RSQ_EXON_RIBOSOMAL : 0)));
            A.f_gline_order.push_back(g.line_order);
        }
        A.f_exon_off[c] = (uint32_t)A.f_exon_start.size();
        for (auto& e : A.exons_by_contig[c]) {
            if (e.start > INT32_MAX - 2 || e.end > INT32_MAX - 2)
                throw HostError(11, "Failed to parse the GTF: coordinate beyond 2^31");
            A.f_exon_start.push_back((int32_t)e.start);
            A.f_exon_end.push_back((int32_t)e.end);
            A.f_exon_gene.push_back(e.gene);
            A.f_exon_slot.push_back(e.slot);
            A.f_exon_flags.push_back((uint8_t)(e.strand | (e.ribosomal ? RSQ_EXON_RIBOSOMAL : 0)));
            A.f_exon_cov_off.push_back(A.exon_cov_off[e.slot]);
            A.f_exon_order.push_back(e.line_order);
        }
        // Elementary gene segments: sweep over [start, end+1) of every Gene feature with one counter per
        // (strand, ribosomal) class; a segment's flags say which classes cover it.
        A.f_seg_off[c] = (uint32_t)A.f_seg_start.size();
        struct Ev { int64_t pos; int cls; int delta; };
        std::vector<Ev> evs;
        for (auto& g : A.genes_by_contig[c]) {
            if (g.end < g.start) continue;
            if (g.end > INT32_MAX - 2) throw HostError(11, "Failed to parse the GTF: coordinate beyond 2^31");
            evs.push_back({g.start, g.strand, +1});
            evs.push_back({g.end + 1, g.strand, -1});
            if (g.ribosomal) {
                evs.push_back({g.start, 3 + g.strand, +1});
                evs.push_back({g.end + 1, 3 + g.strand, -1});
            }
        }
        std::sort(evs.begin(), evs.end(), [](const Ev& a, const Ev& b) { return a.pos < b.pos; });
        int counts[6] = {0, 0, 0, 0, 0, 0};
        A.f_seg_start.push_back(INT32_MIN);  // sentinel segment: nothing before the first gene
        A.f_seg_flags.push_back(0);
        size_t i = 0;
        while (i < evs.size()) {
            int64_t p = evs[i].pos;
            while (i < evs.size() && evs[i].pos == p) { counts[evs[i].cls] += evs[i].delta; ++i; }
            uint8_t fl = 0;
            for (int k = 0; k < 6; ++k) if (counts[k] > 0) fl |= (uint8_t)(1u << k);
            if (fl == A.f_seg_flags.back()) continue;  // merge equal neighbours
            A.f_seg_start.push_back((int32_t)p);
            A.f_seg_flags.push_back(fl);
        }
    }
    A.f_exon_off[nc] = (uint32_t)A.f_exon_start.size();
    A.f_seg_off[nc] = (uint32_t)A.f_seg_start.size();
    A.f_gline_off[nc] = (uint32_t)A.f_gline_start.size();
    lap("sort + flatten");
    A.build_transcripts();
    lap("transcripts");
}

void Annotation::pad_contigs(size_t n_contigs) {
    while (genes_by_contig.size() < n_contigs) {
        genes_by_contig.emplace_back();
        exons_by_contig.emplace_back();
        f_exon_off.push_back(f_exon_off.back());
        f_gline_off.push_back(f_gline_off.back());
        f_seg_start.push_back(INT32_MIN);
        f_seg_flags.push_back(0);
        f_seg_off.push_back((uint32_t)f_seg_start.size());
    }
}

void Annotation::build_transcripts() {
    t_gene_cov_off.clear(); t_gene_exon_off.assign(1, 0); t_exon_len.clear(); t_exon_slot.clear(); t_gene_strand.clear();
    uint32_t flat = 0;
    for (auto& genes : genes_by_contig)
        for (GeneFeature& g : genes) {
            g.flat = flat++;
            const auto& slots = exons_for_gene[g.gene];
            t_gene_cov_off.push_back(slots.empty() ? 0 : exon_cov_off[slots[0]]);   // a gene's exons are contiguous in the coverage array
            for (uint32_t s : slots) { t_exon_len.push_back((uint32_t)std::max<int64_t>(exon_len[s], 0)); t_exon_slot.push_back(s); }
            t_gene_exon_off.push_back((uint32_t)t_exon_len.size());
            t_gene_strand.push_back(g.strand);
        }
}

rsq_transcripts Annotation::transcripts(unsigned coverage_mask, int bias_offset, int bias_window, unsigned long bias_gene_length) const {
    rsq_transcripts q{};
    q.struct_size = sizeof(rsq_transcripts);
    q.n_genes = (uint32_t)t_gene_strand.size();
    q.n_exon_refs = (uint32_t)t_exon_len.size();
    q.coverage_mask = coverage_mask;
    q.bias_offset = bias_offset; q.bias_window = bias_window; q.bias_gene_length = (uint32_t)bias_gene_length;
    q.gene_cov_off = t_gene_cov_off.data(); q.gene_exon_off = t_gene_exon_off.data(); q.gene_strand = t_gene_strand.data();
    q.exon_len = t_exon_len.data(); q.exon_slot = t_exon_slot.data();
    return q;
}

rsq_annotation Annotation::view() const {
    rsq_annotation v{};
    v.struct_size = sizeof(rsq_annotation);
    v.n_contigs = (uint32_t)genes_by_contig.size();
    v.n_genes = (uint32_t)gene_ids.size();
    v.n_exons = (uint32_t)f_exon_start.size();
    v.n_segs = (uint32_t)f_seg_start.size();
    v.n_exon_slots = (uint32_t)exon_ids.size();
    v.exon_off = f_exon_off.data();
    v.exon_start = f_exon_start.data();
    v.exon_end = f_exon_end.data();
    v.exon_gene = f_exon_gene.data();
    v.exon_slot = f_exon_slot.data();
    v.exon_flags = f_exon_flags.data();
    v.exon_cov_off = f_exon_cov_off.data();
    v.seg_off = f_seg_off.data();
    v.seg_start = f_seg_start.data();
    v.seg_flags = f_seg_flags.data();
    v.gene_flags = gene_flags.data();
    v.cov_len = cov_len;
    v.n_gene_lines = (uint32_t)f_gline_start.size();
    v.gline_off = f_gline_off.data();
    v.gline_start = f_gline_start.data();
    v.gline_end = f_gline_end.data();
    v.gline_gene = f_gline_gene.data();
    v.gline_flags = f_gline_flags.data();
    v.gline_order = f_gline_order.data();
    v.exon_order = f_exon_order.data();
    v.gline_fid = f_gline_fid.data();
    return v;
}

void load_bed(const std::string& path, ContigMap& contigs, BedModel& bed) {
    std::ifstream in(path);
    if (!in.is_open()) throw HostError(10, "Unable to open BED file: " + path);
    std::unordered_map<std::string, uint32_t> name_index;
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line[0] == '#') continue;
        // `tokenizer >> buffer` x3: whitespace-separated chromosome, start, end (src/BED.cpp:27-33)
        size_t p = 0;
        auto next_tok = [&](std::string& tok) {
            while (p < line.size() && isspace((unsigned char)line[p])) ++p;
            if (p >= line.size()) return false;
            size_t q = p;
            while (q < line.size() && !isspace((unsigned char)line[q])) ++q;
            tok = line.substr(p, q - p);
            p = q;
            return true;
        };
        // a failed `>>` leaves the buffer as it was: a missing end repeats the start, a missing start repeats the chromosome
        // name (a number, or stoull throws), a blank line leaves the fresh buffer empty (stoull throws)
        std::string t_chr, t_start, t_end;
        next_tok(t_chr);
        t_start = t_chr;
        next_tok(t_start);
        t_end = t_start;
        next_tok(t_end);
        unsigned long long s = 0, e = 0;
        if (!parse_u64(t_start, s) || !parse_u64(t_end, e))
            throw HostError(11, "Failed to parse the BED: Encountered an unknown error while parsing the BED: stoull");
        uint32_t c = contigs.get(t_chr);
        if (bed.starts.size() <= c) {
            bed.starts.resize(c + 1);
            bed.ends.resize(c + 1);
            bed.name_ids.resize(c + 1);
            bed.present.resize(c + 1, 0);
        }
        auto it = name_index.find(line);
        uint32_t nid;
        if (it == name_index.end()) {
            nid = (uint32_t)bed.names.size();
            bed.names.push_back(line);
            name_index.emplace(line, nid);
        } else nid = it->second;
        if (!bed.starts[c].empty() && (int64_t)(s + 1) < bed.starts[c].back()) bed.was_sorted = false;
        bed.starts[c].push_back((int64_t)s + 1);
        bed.ends[c].push_back((int64_t)e + 1);
        bed.name_ids[c].push_back(nid);
        bed.present[c] = 1;
    }
}

void flatten_bed(BedModel& bed, size_t n_contigs) {
    bed.starts.resize(n_contigs);
    bed.ends.resize(n_contigs);
    bed.name_ids.resize(n_contigs);
    bed.present.resize(n_contigs, 0);
    bed.f_off.assign(n_contigs + 1, 0);
    bed.f_start.clear(); bed.f_end.clear(); bed.f_name.clear();
    if (!bed.was_sorted)
        std::cerr << "Warning: BED intervals are not sorted by start; they are sorted here (the reference walks "
                     "them in file order, src/Expression.cpp:69-78,111)" << std::endl;
    for (size_t c = 0; c < n_contigs; ++c) {
        bed.f_off[c] = (uint32_t)bed.f_start.size();
        std::vector<size_t> order(bed.starts[c].size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return bed.starts[c][a] < bed.starts[c][b]; });
        for (size_t i : order) {
            bed.f_start.push_back((int32_t)bed.starts[c][i]);
            bed.f_end.push_back((int32_t)bed.ends[c][i]);
            bed.f_name.push_back(bed.name_ids[c][i]);
        }
    }
    bed.f_off[n_contigs] = (uint32_t)bed.f_start.size();
}

rsq_bed BedModel::view() const {
    rsq_bed v{};
    v.struct_size = sizeof(rsq_bed);
    v.n_contigs = (uint32_t)(f_off.size() - 1);
    v.n_intervals = (uint32_t)f_start.size();
    v.off = f_off.data();
    v.start = f_start.data();
    v.end = f_end.data();
    v.name_id = f_name.data();
    v.contig_present = present.data();
    return v;
}

void build_islands(const Annotation& A, const BedModel* bed, Islands& out) {
    const size_t nc = A.genes_by_contig.size();
    out.by_contig.assign(nc, {});
    // hull of every gene-universe entry per contig (a gene id may, in odd annotations, have lines on several contigs)
    std::vector<std::unordered_map<uint32_t, std::pair<int64_t, int64_t>>> hull(nc);
    auto grow = [&](uint32_t contig, uint32_t gene, int64_t s, int64_t e) {
        if (e < s) std::swap(s, e);
        auto it = hull[contig].find(gene);
        if (it == hull[contig].end()) hull[contig].emplace(gene, std::make_pair(s, e));
        else { it->second.first = std::min(it->second.first, s); it->second.second = std::max(it->second.second, e); }
    };
    for (size_t c = 0; c < nc; ++c) {
        for (const GeneFeature& g : A.genes_by_contig[c]) { grow((uint32_t)c, g.gene, g.start, g.end); if (g.gid != g.gene) grow((uint32_t)c, g.gid, g.start, g.end); }
        for (const ExonFeature& e : A.exons_by_contig[c]) grow((uint32_t)c, e.gene, e.start, e.end);
    }
    for (size_t c = 0; c < nc; ++c) {
        std::vector<std::pair<int64_t, int64_t>> iv;
        iv.reserve(hull[c].size());
        for (auto& kv : hull[c]) iv.push_back(kv.second);
        if (bed && c < bed->starts.size())
            for (size_t i = 0; i < bed->starts[c].size(); ++i) iv.emplace_back(std::min(bed->starts[c][i], bed->ends[c][i]), std::max(bed->starts[c][i], bed->ends[c][i]));
        std::sort(iv.begin(), iv.end());
        auto& m = out.by_contig[c];
        for (auto& x : iv) {
            // the reference's closed block end is one past the last base (src/Expression.cpp:45): keep one base of slack
            if (!m.empty() && x.first <= m.back().second + 1) m.back().second = std::max(m.back().second, x.second);
            else m.push_back(x);
        }
    }
}

int64_t Islands::next_cut(uint32_t contig, int64_t pos1) const {
    if (contig >= by_contig.size()) return pos1;
    const auto& m = by_contig[contig];
    // last island starting at or before pos1
    auto it = std::upper_bound(m.begin(), m.end(), std::make_pair(pos1, INT64_MAX));
    if (it == m.begin()) return pos1;
    --it;
    return pos1 <= it->second + 1 ? it->second + 2 : pos1;
}

void gene_ownership(const Annotation& A, const std::vector<StreamCut>& cuts, const std::vector<int64_t>& contig_rank, uint32_t rank,
                    std::vector<uint8_t>& owned) {
    owned.assign(A.t_gene_strand.size(), 0);
    for (size_t c = 0; c < A.genes_by_contig.size(); ++c) {
        const int64_t cr = c < contig_rank.size() ? contig_rank[c] : -1;
        for (const GeneFeature& g : A.genes_by_contig[c]) {
            uint32_t owner = 0;
            if (cr >= 0)
                for (const StreamCut& k : cuts) if (cr > k.contig_rank || (cr == k.contig_rank && g.start >= k.pos1)) ++owner;
            owned[g.flat] = owner == rank;
        }
    }
}

}  // namespace rsqh
