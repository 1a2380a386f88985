// rsq_legacy.cuh — device restatement of the --legacy counting rules (legacyExonAlignmentMetrics,
// src/Expression.cpp:129-304).  Included by rsq_kernels.cu only; called once per record by the <*, kLegacy=true>
// instantiations of hot_path_kernel in place of the exonAlignmentMetrics section.  The rules are sequential per read and
// order-dependent (list order of gene and exon lines, float dosage sums in block order), so every lane walks its own
// read; this mode exists for parity with RNA-SeQC 1.1.9 counts, not for throughput, and does not use the staged window.
#pragma once

namespace rsqd {
namespace {

struct LegacyRead {
    int contig, pos, endpos, l_qseq, rstrand;
    uint32_t flag, nb;
    bool cleared, hq;
    unsigned long long name_hash;
    const rsq_block* bp;
};

// exact Q64.64 image of a finite positive float (24-bit mantissa: nothing is lost for v >= 2^-40)
__device__ __forceinline__ void float_q64(float v, unsigned long long& whole, unsigned long long& lo) {
    const uint32_t bits = __float_as_uint(v);
    const int exp = (int)((bits >> 23) & 0xffu);
    const unsigned long long man = (unsigned long long)((bits & 0x7fffffu) | (exp ? 0x800000u : 0u));
    const int s = (exp ? exp : 1) - 150 + 64;          // value = man * 2^(s-64)
    whole = 0ull; lo = 0ull;
    if (s >= 64) whole = man << (s - 64);
    else if (s > 0) { lo = man << s; whole = s > 40 ? (man >> (64 - s)) : 0ull; }
    else if (s > -24) lo = man >> (-s);
}

// The exon the reference's inner loop settles on for (gene id, block): the first exon line in list order, among the lines
// the WHOLE read intersects ([cs, ce], :148-150), that belongs to the gene and fully contains the block (:177-195).  Exon
// lines of the gene met before it that merely touch the block set `ribosomal` (when the gene line is rRNA) and
// `partial` (partialIntersect > 0 but not the block length).
__device__ __forceinline__ int legacy_find_exon(const AnnView& ann, int e_lo, int e_hi, int cs, uint32_t gid, int bs, int be,
                                                bool& touched, bool& partial) {
    for (int e = e_lo; e < e_hi; ++e) {
        const int2 x = __ldg(&ann.exon_se[e]);
        if (x.y < cs) continue;                                                        // not in the read's result list
        if ((__ldg(&ann.exon_meta[e].gene_flags) & 0xffffffu) != gid) continue;
        if (x.y < bs || x.x > be) continue;                                            // intersectInterval, GTF.cpp:176-179
        touched = true;
        if (x.x <= bs && be - 1 <= x.y) return e;                                      // partialIntersect == block length
        const int pi = 1 + min(x.y, be - 1) - max(x.x, bs);
        if (pi > 0) partial = true;
    }
    return -1;
}

__device__ __noinline__ void legacy_metrics(const AnnView& ann, const Accum& acc, const Params& prm, const LegacyRead r,
                                            unsigned long long& cm, int& n_fresh, unsigned int& my_bytes) {
    const bool hq = r.hq;
    const uint32_t nb = r.nb;
    // split detection (:135-141, LEGACY_SPLIT_DISTANCE = 100)
    bool split = false;
    {
        long long last_end = -1;
        for (uint32_t b = 0; b < nb; ++b) {
            const int2 blk = __ldg(reinterpret_cast<const int2*>(r.bp + b));
            if (last_end > 0 && !split) split = ((long long)blk.x - last_end) > 99ll;
            last_end = (long long)blk.x + blk.y;
        }
    }
    const int cs = r.pos + 1, ce = r.endpos;                                           // the read as one closed interval, :144-147
    int g_lo = 0, g_hi = 0, e_lo = 0, e_hi = 0;
    if (!r.cleared) {
        const int gl0 = (int)__ldg(&ann.gline_off[r.contig]), gl1 = (int)__ldg(&ann.gline_off[r.contig + 1]);
        g_hi = upper_bound_start(ann.gline_se, gl0, gl1, ce);
        g_lo = lower_bound_pmax(ann.gline_pmax, gl0, g_hi, cs);
        const int c0 = (int)__ldg(&ann.exon_off[r.contig]), c1 = (int)__ldg(&ann.exon_off[r.contig + 1]);
        e_hi = upper_bound_start(ann.exon_se, c0, c1, ce);
        e_lo = lower_bound_pmax(ann.exon_pmax, c0, e_hi, cs);
    }
    // the last line of the result list decides what legacyNotSplit holds after the loop (:163 resets it for EVERY line)
    int last_gene = -1;
    for (int g = g_hi - 1; g >= g_lo; --g) if (__ldg(&ann.gline_se[g].y) >= cs) { last_gene = g; break; }
    if (last_gene >= 0) {
        for (int e = e_hi - 1; e >= e_lo; --e) {
            const int2 x = __ldg(&ann.exon_se[e]);
            if (x.y < cs) continue;
            const int gs = __ldg(&ann.gline_se[last_gene].x);
            if (x.x > gs || (x.x == gs && __ldg(&ann.exon_order[e]) > __ldg(&ann.gline_order[last_gene]))) last_gene = -1;
            break;                                                                     // exons are in list order: the first hit from the back is the last exon line
        }
    }
    bool intragenic = false, plus = false, minus = false, ribosomal = false, do_exon = false, exonic = false, junction = false,
         not_exonic = false, final_not_split = false;
    constexpr int kCache = 16;
    int found_cache[kCache];
    for (int g = g_lo; g < g_hi; ++g) {                                                // :154
        const int2 gx = __ldg(&ann.gline_se[g]);
        if (gx.y < cs) continue;
        const uint32_t ggf = __ldg(&ann.gline_gf[g]);
        const uint32_t gid = ggf & 0xffffffu;
        const int gstrand = (int)((ggf >> 24) & RSQ_EXON_STRAND_MASK);
        const bool g_ribo = (ggf >> 24) & RSQ_EXON_RIBOSOMAL;
        if (gstrand == RSQ_STRAND_FWD) plus = true; else if (gstrand == RSQ_STRAND_REV) minus = true;
        bool not_split = false, found_last = false, t_exon = false, t_intron = false;
        int last_exon = -1;
        const bool strand_ok = r.rstrand == RSQ_STRAND_UNKNOWN || r.rstrand == gstrand;    // :167
        if (strand_ok) {
            for (uint32_t b = 0; b < nb; ++b) {
                const int2 blk = __ldg(reinterpret_cast<const int2*>(r.bp + b));
                const int bs = blk.x, be = blk.x + blk.y;
                intragenic = true;
                if (bs > gx.y) not_exonic = true;                                      // :170
                bool touched = false, partial = false;
                const int e = legacy_find_exon(ann, e_lo, e_hi, cs, gid, bs, be, touched, partial);
                if (touched && g_ribo) ribosomal = true;                               // :179
                if (partial) t_intron = true;
                found_last = e >= 0;
                if (found_last) { t_exon = true; last_exon = e; }
                if (b < (uint32_t)kCache) found_cache[b] = e;
                if (split && !not_split && !found_last) not_split = true;             // :196-203
            }
        }
        if (found_last) {                                                              // :209
            if (hq) {
                if (split && !not_split) {
                    // legacySplitDosage: one float per exon, summed in block order (:200); every block found an exon here
                    for (uint32_t b = 0; b < nb; ++b) {
                        bool t0 = false, t1 = false;
                        int2 blk = __ldg(reinterpret_cast<const int2*>(r.bp + b));
                        const int eb = b < (uint32_t)kCache ? found_cache[b] : legacy_find_exon(ann, e_lo, e_hi, cs, gid, blk.x, blk.x + blk.y, t0, t1);
                        bool first = true;
                        for (uint32_t p = 0; p < b && first; ++p) {
                            const int2 pb = __ldg(reinterpret_cast<const int2*>(r.bp + p));
                            const int ep = p < (uint32_t)kCache ? found_cache[p] : legacy_find_exon(ann, e_lo, e_hi, cs, gid, pb.x, pb.x + pb.y, t0, t1);
                            if (ep == eb) first = false;
                        }
                        if (!first) continue;
                        float dose = __fdiv_rn((float)blk.y, (float)r.l_qseq);
                        for (uint32_t q = b + 1; q < nb; ++q) {
                            const int2 qb = __ldg(reinterpret_cast<const int2*>(r.bp + q));
                            const int eq = q < (uint32_t)kCache ? found_cache[q] : legacy_find_exon(ann, e_lo, e_hi, cs, gid, qb.x, qb.x + qb.y, t0, t1);
                            if (eq == eb) dose = __fadd_rn(dose, __fdiv_rn((float)qb.y, (float)r.l_qseq));
                        }
                        if (dose > 0.f && dose <= 3.0e38f) {                           // finite (Length() == 0 gives inf: skipped, as in the oracle)
                            unsigned long long whole, lo;
                            float_q64(dose, whole, lo);
                            const uint32_t slot = __ldg(&ann.exon_meta[eb].slot);
                            if (whole) atomicAdd(&acc.exon_whole[slot], whole);
                            if (lo >> 32) atomicAdd(&acc.exon_fhi[slot], lo >> 32);
                            if (lo & 0xffffffffull) atomicAdd(&acc.exon_flo[slot], lo & 0xffffffffull);
                        }
                        my_bytes += 8u;
                    }
                } else {
                    atomicAdd(&acc.exon_whole[__ldg(&ann.exon_meta[last_exon].slot)], 1ull);   // :223: only the last exon, whole
                    my_bytes += 8u;
                }
                atomicAdd(&acc.gene_reads[gid], 1ull);                                 // :226
                const unsigned long long key0 = mix64(r.name_hash ^ ((unsigned long long)(gid + 1u) * 0x9E3779B97F4A7C15ull));
                if (set_insert(acc.frag_set, acc.frag_cap, key0 ? key0 : 1ull, acc.error_flag, 1u)) {
                    atomicAdd(&acc.gene_fragments[gid], 1ull);
                    ++n_fresh;
                }
                if (!(r.flag & RSQ_F_DUP)) atomicAdd(&acc.gene_unique[gid], 1ull);
                my_bytes += 8u;
                if (prm.collect_coverage) {
                    // baseCoverage.commit(gene id) (:231, Metrics.cpp:106-124) replays everything cached under the id in this
                    // read: one entry per (non-skipped gene line with this id so far, block with an exon)
                    unsigned int mult = 0;
                    for (int p = g_lo; p <= g; ++p) {
                        if (__ldg(&ann.gline_se[p].y) < cs) continue;
                        const uint32_t pgf = __ldg(&ann.gline_gf[p]);
                        if ((pgf & 0xffffffu) != gid) continue;
                        const int ps = (int)((pgf >> 24) & RSQ_EXON_STRAND_MASK);
                        if (r.rstrand == RSQ_STRAND_UNKNOWN || r.rstrand == ps) ++mult;
                    }
                    for (uint32_t b = 0; b < nb; ++b) {
                        bool t0 = false, t1 = false;
                        const int2 blk = __ldg(reinterpret_cast<const int2*>(r.bp + b));
                        const int eb = b < (uint32_t)kCache ? found_cache[b] : legacy_find_exon(ann, e_lo, e_hi, cs, gid, blk.x, blk.x + blk.y, t0, t1);
                        if (eb < 0) continue;
                        const unsigned long long o = __ldg(&ann.exon_meta[eb].cov_off) + (unsigned long long)(blk.x - __ldg(&ann.exon_se[eb].x));
                        atomicAdd(&acc.cov_diff[o], mult);
                        atomicAdd(&acc.cov_diff[o + (unsigned long long)blk.y], 0u - mult);
                        my_bytes += 8u;
                    }
                }
            }
            do_exon = true;
        }
        if (t_intron && t_exon) junction = true;                                       // :236
        if (t_exon) exonic = true;
        if (g == last_gene) final_not_split = not_split;
    }
#define CNT(k) (cm |= 1ull << (k))
    if (not_exonic || junction || !exonic) {                                           // :242-257
        if (intragenic) {
            CNT(RSQ_C_INTRONIC_READS); CNT(RSQ_C_INTRAGENIC_READS);
            if (hq) { CNT(RSQ_C_HQ_INTRONIC_READS); CNT(RSQ_C_HQ_INTRAGENIC_READS); }
        } else {
            CNT(RSQ_C_INTERGENIC_READS);
            if (hq) CNT(RSQ_C_HQ_INTERGENIC_READS);
        }
    } else if (do_exon) {                                                              // :258-268 (junction and not_exonic are false here)
        CNT(RSQ_C_EXONIC_READS); CNT(RSQ_C_INTRAGENIC_READS);
        if (hq) { CNT(RSQ_C_HQ_EXONIC_READS); CNT(RSQ_C_HQ_INTRAGENIC_READS); }
        if (split && !final_not_split) CNT(RSQ_C_SPLIT_READS);
    } else if (intragenic) {                                                           // :269-280
        CNT(RSQ_C_EXONIC_READS); CNT(RSQ_C_INTRAGENIC_READS);
        if (hq) { CNT(RSQ_C_HQ_EXONIC_READS); CNT(RSQ_C_HQ_INTRAGENIC_READS); }
    }
    if (ribosomal) CNT(RSQ_C_RRNA_READS);                                              // :281
    if ((minus != plus) && (prm.unpaired || (r.flag & RSQ_F_PAIRED))) {                // :283-296
        const bool sense = (r.flag & RSQ_F_REVERSE) ? minus : plus;
        if (prm.unpaired || (r.flag & RSQ_F_READ1)) { if (sense) CNT(RSQ_C_END1_SENSE); else CNT(RSQ_C_END1_ANTISENSE); }
        else { if (sense) CNT(RSQ_C_END2_SENSE); else CNT(RSQ_C_END2_ANTISENSE); }
    }
#undef CNT
}

}  // namespace
}  // namespace rsqd
