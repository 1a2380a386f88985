/*
 * rsq_b200.h — C-ABI of the B200-native RNA-SeQC per-read hot path.
 *
 * The reference (getzlab/rnaseqc 2.4.3) exposes NO plugin/FFI interface for this path: it is inline in
 * main() (src/RNASeQC.cpp:242-382) plus free functions that mutate globals (src/Expression.h:19-37,
 * src/Metrics.h:208-209).  The boundary is therefore cut at the natural seam — after BAM decode / GTF+BED
 * parsing and before report generation.  Every entry point below names the reference
 * interface it replaces.  Plain pointers and sizes only; no C++/torch types cross this boundary; no
 * exceptions cross it (int status, rsq_last_error() for text).
 *
 * Two shared libraries export this identical ABI:
 *   rnaseqc_b200/lib/librsq_b200.so   — the product: hand-written sm_100a CUDA kernels (no CPU fallback;
 *                                       rsq_create fails with RSQ_ERR_NO_DEVICE when no GPU is present)
 *   oracle/_build/librsq_oracle.so    — TEST INFRASTRUCTURE: CPU restatement of the reference algorithm
 */
#ifndef RSQ_B200_H
#define RSQ_B200_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSQ_ABI_VERSION 3u

/* ---- status codes -------------------------------------------------------------------------------------- */
enum {
    RSQ_OK = 0,
    RSQ_ERR_INVALID = 1,    /* bad argument / struct size / sequence of calls          */
    RSQ_ERR_NO_DEVICE = 2,  /* CUDA device missing (the product never falls back to CPU) */
    RSQ_ERR_CUDA = 3,       /* CUDA runtime error (text in rsq_last_error)              */
    RSQ_ERR_NOMEM = 4,
    RSQ_ERR_CAPACITY = 5    /* name-hash set or pair table full (raise max_records)      */
};

/* ---- SAM flag bits used on the path (SeqLib/SeqLib/BamRecord.h:245-302) --------------------------------- */
enum {
    RSQ_F_PAIRED = 0x1, RSQ_F_PROPER = 0x2, RSQ_F_UNMAP = 0x4, RSQ_F_MUNMAP = 0x8, RSQ_F_REVERSE = 0x10,
    RSQ_F_MREVERSE = 0x20, RSQ_F_READ1 = 0x40, RSQ_F_READ2 = 0x80, RSQ_F_SECONDARY = 0x100,
    RSQ_F_QCFAIL = 0x200, RSQ_F_DUP = 0x400, RSQ_F_SUPPLEMENTARY = 0x800
};

/* ---- rsq_read.aux bits (sufficient statistics the decoder extracts) -------------- */
enum {
    RSQ_AUX_HAS_NM = 0x01,          /* int-typed NM aux present: GetIntTag, SeqLib/SeqLib/BamRecord.h:740-750  */
    RSQ_AUX_CHIMERIC_TAG = 0x02,    /* readStringTag(chimeric_tag): src/RNASeQC.cpp:780-800                     */
    RSQ_AUX_MATE_OTHER_CONTIG = 0x04, /* ChrID() != MateChrID(): src/RNASeQC.cpp:287                           */
    RSQ_AUX_TID_OUT_OF_RANGE = 0x08,  /* ChrID()<0 || >= nChrs: src/RNASeQC.cpp:333                              */
    RSQ_AUX_CONTIG_DROPPED = 0x10,  /* features[chr] already cleared by dropFeatures (src/RNASeQC.cpp:346-349,
                                       src/Expression.cpp:96-103): a BAM that RETURNS to a contig finds nothing */
    RSQ_AUX_TID_GT_127 = 0x20,      /* ChrID() > 127: chimeric under --legacy only (src/RNASeQC.cpp:287)       */
    RSQ_AUX_IMPLICIT_BLOCK = 0x40   /* rsq_read_packed only: the record's single block is (pos+1, l_qseq) and is
                                       not in the blocks stream                                                  */
};

/* One decoded alignment record: 48 bytes, 16-byte aligned.  Replaces the SeqLib::BamRecord accessors the
 * loop touches (src/RNASeQC.cpp:254-376, src/Expression.cpp:26-67,119-125,308-479,482-540). */
typedef struct rsq_read {
    int32_t  pos;         /* core.pos, 0-based                                                     */
    int32_t  mpos;        /* core.mpos                                                             */
    int32_t  endpos;      /* BamRecord::PositionEnd(), SeqLib/src/BamRecord.cpp:51-53               */
    int32_t  isize;       /* core.isize                                                            */
    int32_t  l_qseq;      /* BamRecord::Length()                                                   */
    int32_t  nm;          /* value of NM when RSQ_AUX_HAS_NM, else 0                                */
    uint16_t flag;        /* core.flag                                                             */
    uint8_t  mapq;        /* core.qual                                                             */
    uint8_t  aux;         /* RSQ_AUX_* bits                                                        */
    int16_t  contig;      /* annotation contig index of the read's reference name, -1 if the name
                             carries no GTF feature (chromosomeMap, src/Fasta.cpp:17-25)            */
    uint8_t  filter_mask; /* bit k set: --tag k present (GetTag, SeqLib/src/BamRecord.cpp:474-497)  */
    uint8_t  reserved0;
    uint64_t name_hash;   /* 64-bit hash of Qname(): identity for fragmentTracker (src/Metrics.cpp:22)
                             and the fragment-size mate map (src/Expression.cpp:511-531).  Must be a
                             MIXED hash (all 64 bits depend on the whole name; the host library's
                             hash_qname ends with a splitmix64 finaliser): the device tables take their
                             home slots from it after one multiply only                            */
    uint32_t block_off;   /* index of this read's first aligned block in the blocks array of the SAME
                             submit call                                                           */
    uint16_t n_blocks;    /* number of M/=/X blocks (extractBlocks, src/Expression.cpp:26-67)       */
    uint16_t reserved1;
} rsq_read;

/* One aligned block (extractBlocks, src/Expression.cpp:39-48): 1-based start and M/=/X length; the
 * reference's quirky closed end is start+len (src/Expression.cpp:45). */
typedef struct rsq_block {
    int32_t start;
    int32_t len;
} rsq_block;

/* Compact wire format of one record for the host->device hop: 32 bytes, and no block entry at all for the common
 * "one M run as long as the read" case, instead of 48 + 8 per block.  The end-to-end rate of this path is bounded by
 * PCIe, so bytes per record are what matters there; the device expands packed records into
 * rsq_read/rsq_block before the hot kernel.  Field meaning as in rsq_read.  Derived on the device:
 *   endpos    = pos + 1                      when n_blocks == 0   (htslib bam_endpos of a record without reference span)
 *             = pos + l_qseq                 when RSQ_AUX_IMPLICIT_BLOCK
 *             = last block start - 1 + len   otherwise
 *   block_off = running sum of n_blocks in record order
 * Records that do not fit these rules (l_qseq or NM above 65535, more than 255 blocks, an end position that is not the
 * end of the last block) cannot be packed: rsq_pack_records refuses the batch and the caller uses rsq_submit. */
typedef struct rsq_read_packed {
    int32_t  pos;
    int32_t  mpos;
    int32_t  isize;
    uint16_t flag;
    uint8_t  mapq;
    uint8_t  aux;          /* RSQ_AUX_* incl. RSQ_AUX_IMPLICIT_BLOCK                                  */
    uint64_t name_hash;
    int16_t  contig;
    uint16_t l_qseq;
    uint16_t nm;
    uint8_t  n_blocks;
    uint8_t  filter_mask;
} rsq_read_packed;

/* ---- 16-byte wire format ("compact batch") --------------------------------------------------------------------------------------
 * The end-to-end rate of this path is PCIe-bound, so bytes per record set it.  A compact batch is ONE contiguous
 * buffer (one host->device copy per submit):  rsq_batch16 header | tiles | reads | combos | blocks | exception reads | exception blocks.
 *   - records travel in tiles of 32 consecutive records; a tile header carries what its records share or what a data-parallel
 *     expansion needs to know where a tile's variable-length parts start;
 *   - a record is 16 bytes: the 64-bit read-name hash, its position as a 16-bit offset from the tile's lowest position, the mate
 *     position and the insert size as signed 16-bit offsets, and a 16-bit index into the batch's DICTIONARY of the distinct
 *     (flag, MAPQ, aux bits, read length, NM, block count, filter mask) combinations (a few thousand per batch in practice);
 *   - aligned blocks of records that are not one whole-read run travel as one 32-bit word each (offset from the read start in the
 *     high 21 bits, length in the low 11);
 *   - whatever these fields cannot express (mate more than 32 kb away, an intron above 2 Mb, a block of 2048 bases or more,
 *     another contig inside the tile, a 65,536th combination ...) travels as an EXCEPTION: the record's 16-byte slot carries
 *     combo 0xFFFF and the full rsq_read (+ rsq_block entries) sits in the exception lists.  Results never depend on the format:
 *     rsq_pack_compact refuses a batch with more than 1/8 exceptions and the caller falls back to rsq_submit_packed / rsq_submit.
 * Workload of BASELINE.json: 21 bytes per record on the wire instead of 40 (rsq_read_packed) or 60 (rsq_read + rsq_block). */
#define RSQ_BATCH16_MAGIC 0x36314252u   /* "RB16" */
#define RSQ_COMBO_EXCEPTION 0xFFFFu
typedef struct rsq_tile16 {
    int32_t  base_pos;        /* lowest pos among the tile's non-exception records                                  */
    uint32_t block_base;      /* first entry of the tile in `blocks` (explicit blocks of non-exception records)       */
    uint32_t out_block_base;  /* first block of the tile in the expanded rsq_block array (implicit and exception
                                 blocks included): lets every tile be expanded independently                          */
    uint32_t exc_base;        /* first entry of the tile in the exception reads                                       */
    int16_t  contig;          /* contig of every non-exception record of the tile                                     */
    uint16_t reserved;
    uint32_t reserved2;
} rsq_tile16;                 /* 24 bytes per 32 records */
typedef struct rsq_read16 {
    uint64_t name_hash;
    uint16_t dpos;            /* pos - tile.base_pos                                                                  */
    int16_t  dmate;           /* mpos - pos                                                                           */
    int16_t  isize;
    uint16_t combo;           /* dictionary index, or RSQ_COMBO_EXCEPTION                                             */
} rsq_read16;
typedef struct rsq_combo16 {  /* one dictionary entry: field meaning as in rsq_read / rsq_read_packed                 */
    uint16_t flag;
    uint8_t  mapq;
    uint8_t  aux;             /* RSQ_AUX_* incl. RSQ_AUX_IMPLICIT_BLOCK                                               */
    uint16_t l_qseq;
    uint16_t nm;
    uint8_t  n_blocks;
    uint8_t  filter_mask;
    uint16_t reserved;
    uint32_t reserved2;
} rsq_combo16;                /* 16 bytes */
typedef struct rsq_batch16 {  /* header of the buffer; every offset is in bytes from the start of the buffer, 16-byte aligned */
    uint32_t magic, struct_size;
    uint32_t n_reads, n_tiles, n_combos, n_blocks, n_exc, n_exc_blocks;
    uint32_t total_blocks;    /* blocks of the expanded batch                                                         */
    uint32_t reserved;
    uint64_t off_tiles, off_reads, off_combos, off_blocks, off_exc, off_exc_blocks, total_bytes;
} rsq_batch16;
#define RSQ_BLOCK16_LEN_BITS 11u

/* ---- run constants: the locals of main() the loop captures (src/RNASeQC.cpp:78-100) --------------------- */
typedef struct rsq_config {
    uint32_t struct_size;             /* = sizeof(rsq_config)                                       */
    int32_t  device;                  /* CUDA device ordinal                                        */
    uint32_t mapq_threshold;          /* MAPPING_QUALITY_THRESHOLD, default 255 (:90)               */
    uint32_t base_mismatch_threshold; /* BASE_MISMATCH_THRESHOLD, default 6 (:89)                   */
    int32_t  chimeric_distance;       /* CHIMERIC_DISTANCE, default 2000000 (:87)                   */
    uint32_t strand_mode;             /* 0 = unstranded, 1 = FR (Strand::Forward), 2 = RF (:78-85)  */
    uint32_t unpaired;                /* -u (:61)                                                   */
    uint32_t exclude_chimeric;        /* --exclude-chimeric (:60)                                   */
    uint32_t n_filter_tags;           /* number of --tag filters, <= 8 (:58)                        */
    uint32_t fragment_samples;        /* FRAGMENT_SIZE_SAMPLES when --bed given, else 0 (:88,176)   */
    uint64_t max_records;             /* upper bound on records this ctx will see (sizes the
                                         fragment-tracker hash set); 0 = 1<<24                     */
    uint32_t collect_coverage;        /* 1: accumulate per-base exon coverage (always on in the
                                         reference; 0 only for ablation benchmarks)               */
    uint32_t legacy;                  /* --legacy counting rules (src/RNASeQC.cpp:55,90,258,276,279,287,364;
                                         src/Expression.cpp:129-304); needs the gene-line arrays of
                                         rsq_annotation                                             */
    uint64_t user_stream;             /* cudaStream_t to run the kernels on (0 = the ctx creates its
                                         own); lets the caller time the path with its own events   */
    uint32_t reserved[4];
} rsq_config;

/* ---- exon flags / gene-segment flags ------------------------------------------------------------------- */
enum { RSQ_STRAND_FWD = 0, RSQ_STRAND_REV = 1, RSQ_STRAND_UNKNOWN = 2 }; /* src/Fasta.h:40 */
enum { RSQ_EXON_STRAND_MASK = 0x3, RSQ_EXON_RIBOSOMAL = 0x4 };
enum { /* OR over the Gene features covering one elementary segment, per gene strand */
    RSQ_SEG_GENE_FWD = 0x01, RSQ_SEG_GENE_REV = 0x02, RSQ_SEG_GENE_UNK = 0x04,
    RSQ_SEG_RIBO_FWD = 0x08, RSQ_SEG_RIBO_REV = 0x10, RSQ_SEG_RIBO_UNK = 0x20
};
enum { RSQ_GENE_GLOBIN = 0x1 }; /* geneNames[id] in blacklistedGlobins, src/Expression.cpp:24,396-399 */

/* Flattened annotation: replaces map<chrom, list<Feature>> features (src/RNASeQC.cpp:104,150-156) and the
 * string-keyed globals of src/GTF.cpp:22-27.  All arrays are caller-owned and copied by the call. */
typedef struct rsq_annotation {
    uint32_t struct_size;
    uint32_t n_contigs;
    uint32_t n_genes;           /* gene universe: every distinct gene_id on a Gene or Exon line          */
    uint32_t n_exons;           /* exon lines kept as intervals                                           */
    uint32_t n_segs;
    uint32_t n_exon_slots;      /* output slots = size of exonList (src/GTF.cpp:106); 0 means n_exons.  Larger
                                   than n_exons under --legacy, which drops single-base exons from the
                                   interval list but not from the outputs (src/RNASeQC.cpp:129-135)       */
    const uint32_t* exon_off;   /* [n_contigs+1] exons of contig c are [exon_off[c], exon_off[c+1])       */
    const int32_t*  exon_start; /* [n_exons] 1-based closed, stable-sorted by start within a contig       */
    const int32_t*  exon_end;   /* [n_exons]                                                              */
    const uint32_t* exon_gene;  /* [n_exons] gene index                                                   */
    const uint32_t* exon_slot;  /* [n_exons] output slot (index into exonList, src/GTF.cpp:106)           */
    const uint8_t*  exon_flags; /* [n_exons] RSQ_EXON_*                                                   */
    const uint64_t* exon_cov_off; /* [n_exons] first base of this exon in the coverage array              */
    const uint32_t* seg_off;    /* [n_contigs+1]                                                          */
    const int32_t*  seg_start;  /* [n_segs] elementary gene segments: segment k of a contig covers
                                   [seg_start[k], seg_start[k+1]) ; strictly increasing within a contig   */
    const uint8_t*  seg_flags;  /* [n_segs] RSQ_SEG_*                                                     */
    const uint8_t*  gene_flags; /* [n_genes] RSQ_GENE_*                                                   */
    uint64_t cov_len;           /* total exon bases (sum of exonLengths, src/GTF.cpp:108)                 */
    /* Gene lines as intervals — read by the --legacy rules only, which walk the Gene features a whole read
     * intersects in list order (src/Expression.cpp:143-159).  May be NULL / 0 when rsq_config.legacy == 0.   */
    uint32_t n_gene_lines;
    uint32_t reserved1;
    const uint32_t* gline_off;   /* [n_contigs+1]                                                          */
    const int32_t*  gline_start; /* [n_gene_lines] stable-sorted by start within a contig                  */
    const int32_t*  gline_end;   /* [n_gene_lines]                                                         */
    const uint32_t* gline_gene;  /* [n_gene_lines] gene index of the line's gene_id (compared with
                                    exon_gene, src/Expression.cpp:175)                                     */
    const uint8_t*  gline_flags; /* [n_gene_lines] RSQ_EXON_STRAND_MASK | RSQ_EXON_RIBOSOMAL                */
    const uint32_t* gline_order; /* [n_gene_lines] position among the kept GTF lines: ties of the list's
                                    stable sort by start (src/RNASeQC.cpp:150-152)                         */
    const uint32_t* exon_order;  /* [n_exons] same, for the exon lines                                     */
    /* Gene index of each gene line's feature_id (it differs from gline_gene only when the line's gene_id was carried over
     * from another line, src/GTF.cpp:90).  With the gene-line arrays and this one given, the path reproduces what happens once
     * a gene line has left the reference's sliding window while one of the gene's exons reaches past it (trimFeatures,
     * src/Expression.cpp:80-93: BaseCoverage::compute marks the gene "seen", fragmentTracker.erase): later reads of the gene
     * start a new fragment set and their coverage is dropped (src/Metrics.cpp:108-112).  NULL: no gene is ever treated so. */
    const uint32_t* gline_fid;   /* [n_gene_lines]                                                         */
} rsq_annotation;

/* BED intervals for the fragment-size sampler (src/BED.cpp:18-43, src/RNASeQC.cpp:172-187); intervals of a
 * contig sorted by start.  name_id: identical BED lines share an id (feature_id = whole line). */
typedef struct rsq_bed {
    uint32_t struct_size;
    uint32_t n_contigs;         /* same contig index space as rsq_annotation                               */
    uint32_t n_intervals;
    uint32_t reserved;
    const uint32_t* off;        /* [n_contigs+1]                                                           */
    const int32_t*  start;      /* [n_intervals] = BED start + 1                                          */
    const int32_t*  end;        /* [n_intervals] = BED end + 1                                            */
    const uint32_t* name_id;    /* [n_intervals]                                                          */
    const uint8_t*  contig_present; /* [n_contigs] bedFeatures->find(chr) != end (src/RNASeQC.cpp:372)     */
} rsq_bed;

/* Reference genome as a G/C bit mask, for the fragment GC-content sampler of --fasta runs: replaces Fasta::hasContig /
 * Fasta::getSeq + gc() on the path (src/Fasta.cpp:66-75,103-150, src/Expression.cpp:459-476).  Bit b of word w of a contig
 * is set when base 32*w+b (0-based) is one of G, C, g, c. */
typedef struct rsq_reference {
    uint32_t struct_size;
    uint32_t n_contigs;          /* same contig index space as rsq_annotation                                */
    const uint8_t*  present;     /* [n_contigs] Fasta::hasContig                                              */
    const uint64_t* length;      /* [n_contigs] bases (0 when absent)                                         */
    const uint64_t* word_off;    /* [n_contigs+1] first 32-bit word of each contig in gc_bits                 */
    const uint32_t* gc_bits;     /* [word_off[n_contigs]]                                                     */
} rsq_reference;

/* ---- counters: fixed indices replacing the string-keyed Metrics map (src/Metrics.h:31-41) --------------- */
enum {
    RSQ_C_ALTERNATIVE_ALIGNMENTS = 0, RSQ_C_SUPPLEMENTARY_ALIGNMENTS, RSQ_C_FAILED_VENDOR_QC,
    RSQ_C_LOW_MAPPING_QUALITY, RSQ_C_CHIMERIC_AUTO, RSQ_C_CHIMERIC_TAG, RSQ_C_UMVQ, RSQ_C_UNPAIRED_READS,
    RSQ_C_MAPPED_READS, RSQ_C_MAPPED_DUPLICATE_READS, RSQ_C_MAPPED_UNIQUE_READS, RSQ_C_TOTAL_MAPPED_PAIRS,
    RSQ_C_END1_MAPPED_READS, RSQ_C_END1_MISMATCHES, RSQ_C_END1_BASES, RSQ_C_DUPLICATE_PAIRS,
    RSQ_C_UNIQUE_FRAGMENTS, RSQ_C_END2_MAPPED_READS, RSQ_C_END2_MISMATCHES, RSQ_C_END2_BASES,
    RSQ_C_MISMATCHED_BASES, RSQ_C_TOTAL_BASES, RSQ_C_HIGH_QUALITY_READS, RSQ_C_LOW_QUALITY_READS,
    RSQ_C_READS_USED, RSQ_C_ALIGNMENT_BLOCKS, RSQ_C_NON_GLOBIN_READS, RSQ_C_NON_GLOBIN_DUPLICATE_READS,
    RSQ_C_INTRONIC_READS, RSQ_C_INTRAGENIC_READS, RSQ_C_HQ_INTRONIC_READS, RSQ_C_HQ_INTRAGENIC_READS,
    RSQ_C_INTERGENIC_READS, RSQ_C_HQ_INTERGENIC_READS, RSQ_C_EXONIC_READS, RSQ_C_HQ_EXONIC_READS,
    RSQ_C_AMBIGUOUS_READS, RSQ_C_HQ_AMBIGUOUS_READS, RSQ_C_RRNA_READS, RSQ_C_END1_SENSE,
    RSQ_C_END1_ANTISENSE, RSQ_C_END2_SENSE, RSQ_C_END2_ANTISENSE,
    RSQ_C_TOTAL_ALIGNMENTS,        /* alignmentCount, src/RNASeQC.cpp:245 (the int truncation of :397 is the writer's) */
    RSQ_C_FILTERED_BY_TAG0,        /* .. +7: "Filtered by tag: X", src/RNASeQC.cpp:325 */
    RSQ_C_SPLIT_READS = RSQ_C_FILTERED_BY_TAG0 + 8, /* --legacy only (src/Expression.cpp:279) */
    RSQ_C_FRAGMENT_CANDIDATES,     /* diagnostics: reads that entered the fragment-size mate map */
    RSQ_C_FRAGSET_ENTRIES,         /* diagnostics: distinct (gene, read name) pairs in the fragment tracker = sum of gene_fragments */
    RSQ_N_COUNTERS = 64
};

/* ---- results ------------------------------------------------------------------------------------------ */
typedef struct rsq_results {
    uint32_t struct_size;
    uint32_t reserved;
    uint64_t counters[RSQ_N_COUNTERS];
    /* caller-provided arrays (NULL = skip); replace the global count tables of src/Metrics.cpp:20 */
    uint64_t* gene_reads;      /* [n_genes]  geneCounts                                                  */
    uint64_t* gene_unique;     /* [n_genes]  uniqueGeneCounts                                            */
    uint64_t* gene_fragments;  /* [n_genes]  geneFragmentCounts                                          */
    uint64_t* exon_whole;      /* [n_exons by slot] integer part of exonCounts                           */
    uint64_t* exon_frac;       /* [n_exons by slot] fractional part in units of 2^-64 (exact Q64.64 sum
                                  of isz/length, src/Expression.cpp:345, src/Metrics.cpp:59-66)          */
    uint32_t* coverage;        /* [cov_len] per-base exon coverage (BaseCoverage::coverage after all
                                  commits, src/Metrics.cpp:106-124)                                     */
    uint64_t* fragment_hist;   /* [fragment_hist_len] fragmentSizes[|isize|] (src/Expression.cpp:530);
                                  the last bin collects |isize| >= fragment_hist_len-1                   */
    uint64_t  fragment_hist_len;
    uint64_t  fragment_samples_taken;
    uint64_t* gc_hist;         /* [100] gcBins (src/RNASeQC.cpp:106,368); all zero without rsq_set_reference */
} rsq_results;

typedef struct rsq_ctx rsq_ctx;

/* Library identity: "cuda-sm_100a" for the product, "oracle-cpu" for the test oracle. */
const char* rsq_backend(void);
uint32_t    rsq_abi_version(void);

/* Page-locked host memory for submit buffers (plain malloc in the oracle). */
void* rsq_host_alloc(size_t bytes);
void  rsq_host_free(void* p);

/* Replaces the loop-constant locals of main() (src/RNASeQC.cpp:78-100,204-211). */
int rsq_create(const rsq_config* cfg, rsq_ctx** out);
void rsq_destroy(rsq_ctx* ctx);
const char* rsq_last_error(const rsq_ctx* ctx); /* ctx may be NULL: error of the failed rsq_create */

/* Replaces map<chrom,list<Feature>> features + GTF globals (src/RNASeQC.cpp:104-156, src/GTF.cpp:22-27). */
int rsq_set_annotation(rsq_ctx* ctx, const rsq_annotation* ann);
/* Replaces bedFeatures (src/RNASeQC.cpp:169-187). */
int rsq_set_bed(rsq_ctx* ctx, const rsq_bed* bed);
/* Enables the fragment GC-content sampler of --fasta runs (src/Expression.cpp:459-476; gcContentFragmentTracker,
 * src/RNASeQC.cpp:170): after the annotation, before the first submit.  Ignored by the --legacy rules, which never
 * sample (src/RNASeQC.cpp:364). */
int rsq_set_reference(rsq_ctx* ctx, const rsq_reference* ref);

/* Replaces one or more iterations of the loop body src/RNASeQC.cpp:242-382 (and everything it calls:
 * extractBlocks' output is the blocks array; trimFeatures/intersectBlock/exonAlignmentMetrics/
 * fragmentSizeMetrics, src/Expression.cpp:80-117,308-540; Metrics/Collector/BaseCoverage add/commit,
 * src/Metrics.cpp:28-129).  Records must arrive in stream (BAM) order across calls.  Host buffers
 * (pinned or pageable); asynchronous: the buffers are owned by the caller until rsq_sync returns. */
int rsq_submit(rsq_ctx* ctx, const rsq_read* reads, size_t n_reads, const rsq_block* blocks, size_t n_blocks);
/* Same, for batches that already live in device memory of cfg.device (device pointers). */
int rsq_submit_device(rsq_ctx* ctx, const rsq_read* d_reads, size_t n_reads, const rsq_block* d_blocks,
                      size_t n_blocks);
/* Same as rsq_submit for a batch in the compact wire format: `blocks` holds, in record order, the blocks of every
 * record that does not carry RSQ_AUX_IMPLICIT_BLOCK. */
int rsq_submit_packed(rsq_ctx* ctx, const rsq_read_packed* reads, size_t n_reads, const rsq_block* blocks, size_t n_blocks);
/* Same for a compact batch (rsq_batch16 and what follows it in one buffer of `bytes` bytes, written by rsq_pack_compact). */
int rsq_submit_compact(rsq_ctx* ctx, const void* batch, size_t bytes);
int rsq_sync(rsq_ctx* ctx);
/* Returns once the HOST buffers of every submit so far may be reused (their copies have left the host); the kernels that
 * consume them may still be running.  A pipelined caller waits on this, not on rsq_sync, before it refills a buffer. */
int rsq_sync_inputs(rsq_ctx* ctx);

/* Clears all accumulators (annotation and BED stay): a new sample / a new timed step. */
int rsq_reset(rsq_ctx* ctx);

/* End of stream: resolves the fragment-size first-N cut, integrates the coverage difference array, and
 * copies results to the caller's arrays (replaces the reads of the globals at src/RNASeQC.cpp:397-521). */
int rsq_finalize(rsq_ctx* ctx, rsq_results* out);

/* ---- per-gene coverage statistics on the device ------------------------------------------------------------------
 * Replaces BaseCoverage::compute / computeCoverage (src/Metrics.cpp:132-151,265-337) and BiasCounter::computeBias
 * (src/Metrics.cpp:160-235) for every kept "gene" line: per-exon masked mean/std/CV, stitched-transcript 3'/5' bias
 * windows (with the reference's peak-window / 5th-percentile flank trimming) and masked transcript mean/std/CV, computed
 * from the integrated per-base coverage that never leaves the GPU.  The caller (host report writer) keeps the text
 * formatting, the eviction order of coverage.tsv and the medians over genes. */
typedef struct rsq_transcripts {
    uint32_t struct_size;
    uint32_t n_genes;              /* Gene features to evaluate                                              */
    uint32_t n_exon_refs;          /* exons over all listed genes, stitched (exonsForGene) order             */
    uint32_t coverage_mask;        /* COVERAGE_MASK, src/RNASeQC.cpp:91                                       */
    int32_t  bias_offset;          /* BIAS_OFFSET :94 (>= 0)                                                  */
    int32_t  bias_window;          /* BIAS_WINDOW :95 (1..8192)                                               */
    uint32_t bias_gene_length;     /* BIAS_LENGTH :96                                                         */
    uint32_t reserved;
    const uint64_t* gene_cov_off;  /* [n_genes] first base of the gene's stitched transcript in the coverage array */
    const uint32_t* gene_exon_off; /* [n_genes+1] into exon_len / exon_slot                                   */
    const uint8_t*  gene_strand;   /* [n_genes] RSQ_STRAND_*                                                  */
    const uint32_t* exon_len;      /* [n_exon_refs]                                                           */
    const uint32_t* exon_slot;     /* [n_exon_refs]                                                           */
    const uint8_t*  gene_owned;    /* [n_genes] or NULL (= all).  Coordinate-partitioned runs (rsq_set_partition): 1 when the
                                      gene's reads are in THIS partition.  Genes that are not owned are skipped and every
                                      output of theirs is written as all-zero BITS, so the bitwise u64 sum of the outputs
                                      over the partitions is the owner's value (the owner-gather of SURVEY.md §8e); exon
                                      slots that no listed gene references hold NaN on partition 0 and zero bits elsewhere */
} rsq_transcripts;

typedef struct rsq_coverage_stats {
    uint32_t struct_size;
    uint32_t reserved;
    double*   gene_mean;           /* [n_genes] -1 when nothing is left after masking (row "0 0 nan", not counted) */
    double*   gene_std;            /* [n_genes]                                                                */
    double*   gene_cv;             /* [n_genes]                                                                */
    uint64_t* five_end;            /* [n_genes] BiasCounter::fiveEnd (truncated median, src/Metrics.h:76)      */
    uint64_t* three_end;           /* [n_genes]                                                                */
    double*   exon_cv;             /* [n_exons by slot] NaN = no exonCoverage entry (src/Metrics.cpp:298-304)  */
} rsq_coverage_stats;

/* Runs after all records are submitted (integrates the coverage difference array itself).  With a communicator attached
 * (rsq_comm_init) and req->gene_owned given, the call is COLLECTIVE: the per-partition results are summed over NVLink before
 * they are copied out, so every rank receives the statistics of all genes. */
int rsq_coverage_statistics(rsq_ctx* ctx, const rsq_transcripts* req, rsq_coverage_stats* out);

/* ---- coordinate-partitioned runs ------------------------------------------------------------------------
 * The sorted record stream is cut into `world` contiguous coordinate ranges, one ctx (one GPU) per range, cut only where no gene,
 * exon or BED interval lies: the fragment tracker (src/Metrics.cpp:22), both mate maps (src/Expression.cpp:461-474,511-533) and
 * every gene's coverage then have one owner.  What is left to exchange at the end of the run:
 *   - the additive tables: counters, the three gene tables, exon sums, GC-content bins;
 *   - the fragment-size samples: the reference keeps the FIRST fragment_samples pairs of the stream (src/RNASeQC.cpp:372,
 *     src/Expression.cpp:532).  Partition r ranks its own samples in stream order; with the per-partition sample counts known,
 *     it keeps those whose GLOBAL rank (samples of partitions 0..r-1 before its own) is below the quota;
 *   - the per-gene coverage statistics (rsq_transcripts.gene_owned).
 * Two ways to run the exchange: (a) the caller moves the buffers (any transport; the CPU tests use gloo):
 *       rsq_set_partition -> ... submits ... -> rsq_export_sample_counts -> all-reduce(sum) -> rsq_export_counts ->
 *       all-reduce(sum) -> rsq_import_counts -> rsq_finalize
 *   or (b) the library does, over NCCL (product only): rsq_comm_init -> ... submits ... -> rsq_exchange -> rsq_finalize.
 * After either, every partition holds the state of the equivalent single-GPU run (per-base coverage excepted: it stays with
 * its owner). */
int rsq_set_partition(rsq_ctx* ctx, uint32_t rank, uint32_t world);
/* Phase 1 (only needed when the fragment-size sampler is on): `world` uint64 words, zero except [rank] = samples this
 * partition found.  DEVICE pointer in the product, host pointer in the oracle; the caller sum-reduces it in place. */
int rsq_export_sample_counts(rsq_ctx* ctx, uint64_t** d_words, size_t* n_words);
/* Phase 2: count vector for the end-of-run all-reduce (counters ‖ gene_reads ‖ gene_unique ‖ gene_fragments ‖ exon_whole ‖
 * exon_frac_hi32 ‖ exon_frac_lo32 ‖ gc_hist[100] ‖ total samples found ‖ the kept fragment-size samples at their global
 * ranks, two 32-bit values per word), all uint64, additive across partitions without carries.  Returns a DEVICE pointer and
 * the number of uint64 words; valid until rsq_destroy.  rsq_finalize must not have been called between rsq_export_counts and
 * rsq_import_counts. */
int rsq_export_counts(rsq_ctx* ctx, uint64_t** d_words, size_t* n_words);
/* After the caller reduced the exported buffer in place (sum), fold it back into the ctx state. */
int rsq_import_counts(rsq_ctx* ctx);

/* NCCL inside the product (the north_star's "final NCCL allreduce of the per-gene/per-exon count vectors over NVLink").  The
 * library binds libnccl at run time (dlopen; the copy a host process has already loaded — e.g. PyTorch's — is reused).
 * rsq_comm_unique_id: one caller creates the id and hands it to every rank (any transport).  rsq_comm_init: collective over
 * all `world` contexts (one per GPU; threads of one process or one process each); implies rsq_set_partition(rank, world).
 * rsq_exchange: collective; runs phases 1 and 2 above on the ctx stream without a host round trip. */
#define RSQ_COMM_ID_BYTES 128
int rsq_comm_unique_id(uint8_t id[RSQ_COMM_ID_BYTES]);
int rsq_comm_init(rsq_ctx* ctx, const uint8_t id[RSQ_COMM_ID_BYTES], uint32_t rank, uint32_t world);
int rsq_exchange(rsq_ctx* ctx);

/* Algorithmic HBM bytes of everything submitted since the last reset, by the definition of
 * 40 + 8*nb per record, + (8 + 8*nb_a) for HQ assigned reads, + 8*nb_a coverage. */
int rsq_algorithmic_bytes(rsq_ctx* ctx, uint64_t* bytes);
/* Number of kernel launches issued by this ctx since creation (bench.py's gpu_launches). */
int rsq_launch_count(rsq_ctx* ctx, uint64_t* launches);
/* Device time of all hot-path kernels since the last reset, from CUDA events on the ctx stream (ms). */
int rsq_kernel_time_ms(rsq_ctx* ctx, double* ms);

#ifdef __cplusplus
}
#endif

/* Packs one rsq_submit batch (block_off relative to `blocks`) into the compact wire format.  out_reads: n_reads entries;
 * out_blocks: room for n_blocks entries; *n_out_blocks receives the number written.  Returns RSQ_OK, or RSQ_ERR_INVALID
 * when some record cannot be represented (nothing useful is left in the outputs then). */
static inline int rsq_pack_records(const rsq_read* reads, size_t n_reads, const rsq_block* blocks, size_t n_blocks,
                                   rsq_read_packed* out_reads, rsq_block* out_blocks, size_t* n_out_blocks) {
    size_t w = 0, i, b;
    for (i = 0; i < n_reads; ++i) {
        const rsq_read* r = &reads[i];
        rsq_read_packed* p = &out_reads[i];
        const rsq_block* rb = blocks + r->block_off;
        if ((size_t)r->block_off + r->n_blocks > n_blocks) return RSQ_ERR_INVALID;
        if (r->l_qseq < 0 || r->l_qseq > 65535 || r->nm < 0 || r->nm > 65535 || r->n_blocks > 255 ||
            (r->aux & RSQ_AUX_IMPLICIT_BLOCK)) return RSQ_ERR_INVALID;
        p->pos = r->pos; p->mpos = r->mpos; p->isize = r->isize; p->flag = r->flag; p->mapq = r->mapq; p->aux = r->aux;
        p->name_hash = r->name_hash; p->contig = r->contig; p->l_qseq = (uint16_t)r->l_qseq; p->nm = (uint16_t)r->nm;
        p->n_blocks = (uint8_t)r->n_blocks; p->filter_mask = r->filter_mask;
        if (r->n_blocks == 0) {
            if (r->endpos != r->pos + 1) return RSQ_ERR_INVALID;
        } else if (r->n_blocks == 1 && rb[0].start == r->pos + 1 && rb[0].len == r->l_qseq && r->endpos == r->pos + r->l_qseq) {
            p->aux |= RSQ_AUX_IMPLICIT_BLOCK;
        } else {
            if (r->endpos != rb[r->n_blocks - 1].start - 1 + rb[r->n_blocks - 1].len) return RSQ_ERR_INVALID;
            for (b = 0; b < r->n_blocks; ++b) out_blocks[w++] = rb[b];
        }
    }
    *n_out_blocks = w;
    return RSQ_OK;
}

/* Upper bound of the buffer rsq_pack_compact may need for a batch (exceptions are capped at n_reads / 8 + 32). */
static inline size_t rsq_compact_bound(size_t n_reads, size_t n_blocks) {
    const size_t n_tiles = (n_reads + 31) / 32, max_exc = n_reads / 8 + 32;
    return 256 + 24 * n_tiles + 16 * n_reads + 16 * 65536 + 4 * n_blocks + 48 * max_exc + 8 * n_blocks + 7 * 16;
}

/* A batch handed over in pieces: the ingest workers convert ~4 MB jobs of the BAM stream on their own, and copying their records
 * into one array was the serial ceiling of the whole ingest — rsq_pack_compact_segments reads them where they lie.  Each piece has
 * its own block array (block_off relative to it).  `tids` non-NULL: the records' contig field is not final yet and is taken from
 * contig_of_tid[tids[i]] (-1 outside [0, n_tids)) — the annotation's contig ids did not exist when the workers converted them. */
typedef struct rsq_segment {
    const rsq_read* reads;
    const int32_t* tids;
    size_t n_reads;
    const rsq_block* blocks;
    size_t n_blocks;
} rsq_segment;

/* Packs one batch (the concatenation of the segments) into a compact batch at `out` (capacity `cap` bytes, 16-byte aligned;
 * rsq_compact_bound over the totals says how much is enough).  `scratch` must hold 131072 uint32_t (the dictionary's hash table).
 * Returns RSQ_OK and the bytes written, or RSQ_ERR_INVALID when the batch does not suit the format (too many exceptions, too
 * many blocks for 32-bit offsets, capacity): the caller then uses rsq_submit_packed / rsq_submit.  The bytes do not depend on
 * how the batch is cut into segments. */
static inline int rsq_pack_compact_segments(const rsq_segment* segs, size_t n_segs, const int16_t* contig_of_tid, size_t n_tids, void* out,
                                            size_t cap, uint32_t* scratch, size_t* bytes_out) {
    size_t n_reads = 0, n_blocks = 0, s_at = 0, k_at = 0;
    size_t n_tiles, max_exc;
    unsigned char* base = (unsigned char*)out;
    rsq_batch16* H = (rsq_batch16*)out;
    size_t off = 0, t, i, b, n_combos = 0, n_xb = 0, n_exc = 0, n_eb = 0, out_blocks = 0;
    rsq_tile16* tiles; rsq_read16* R; rsq_combo16* D; uint32_t* XB; rsq_read* E; rsq_block* EB;
    rsq_read tile_reads[32];                 /* the tile's records, gathered from the segments (contig bound) */
    const rsq_block* tile_blocks[32];        /* ... and where each one's blocks start */
    for (i = 0; i < n_segs; ++i) { n_reads += segs[i].n_reads; n_blocks += segs[i].n_blocks; }
    n_tiles = (n_reads + 31) / 32; max_exc = n_reads / 8 + 32;
    if (n_reads >= 0xffffffffu || n_blocks >= 0xffffffffu || cap < rsq_compact_bound(n_reads, n_blocks) || ((size_t)out & 15u)) return RSQ_ERR_INVALID;
#define RSQ_P16_ALIGN(x) (((x) + 15u) & ~(size_t)15u)
    off = RSQ_P16_ALIGN(sizeof(rsq_batch16));
    tiles = (rsq_tile16*)(base + off); H->off_tiles = off; off = RSQ_P16_ALIGN(off + 24 * n_tiles);
    R = (rsq_read16*)(base + off); H->off_reads = off; off = RSQ_P16_ALIGN(off + 16 * n_reads);
    D = (rsq_combo16*)(base + off); H->off_combos = off;
    /* the dictionary grows while packing: blocks and exceptions are staged behind its maximum size and moved up at the end */
    XB = (uint32_t*)(base + RSQ_P16_ALIGN(off + 16 * 65536));
    E = (rsq_read*)((unsigned char*)XB + RSQ_P16_ALIGN(4 * n_blocks));
    EB = (rsq_block*)((unsigned char*)E + RSQ_P16_ALIGN(48 * max_exc));
    for (i = 0; i < 131072; ++i) scratch[i] = 0xffffffffu;
    for (t = 0; t < n_tiles; ++t) {
        const size_t lo = 32 * t, hi = lo + 32 < n_reads ? lo + 32 : n_reads;
        rsq_tile16* T = &tiles[t];
        int32_t bp = 0x7fffffff;
        int16_t contig;
        for (i = lo; i < hi; ++i) {          /* gather */
            const rsq_segment* S;
            rsq_read* q = &tile_reads[i - lo];
            while (k_at == segs[s_at].n_reads) { ++s_at; k_at = 0; }
            S = &segs[s_at];
            *q = S->reads[k_at];
            if (S->tids) { const int32_t tid = S->tids[k_at]; q->contig = (tid >= 0 && (size_t)tid < n_tids) ? contig_of_tid[tid] : (int16_t)-1; }
            if ((size_t)q->block_off + q->n_blocks > S->n_blocks) return RSQ_ERR_INVALID;
            tile_blocks[i - lo] = S->blocks + q->block_off;
            ++k_at;
        }
        contig = tile_reads[0].contig;
        for (i = lo; i < hi; ++i) if (tile_reads[i - lo].contig == contig && tile_reads[i - lo].pos < bp) bp = tile_reads[i - lo].pos;
        T->base_pos = bp; T->block_base = (uint32_t)n_xb; T->out_block_base = (uint32_t)out_blocks; T->exc_base = (uint32_t)n_exc;
        T->contig = contig; T->reserved = 0; T->reserved2 = 0;
        for (i = lo; i < hi; ++i) {
            const rsq_read* r = &tile_reads[i - lo];
            const rsq_block* rb = tile_blocks[i - lo];
            rsq_read16* p = &R[i];
            int ok, implicit = 0;
            const long long dpos = (long long)r->pos - bp, dm = (long long)r->mpos - r->pos;
            ok = r->contig == contig && dpos >= 0 && dpos <= 65535 && dm >= -32768 && dm <= 32767 && r->isize >= -32768 && r->isize <= 32767 &&
                 r->l_qseq >= 0 && r->l_qseq <= 65535 && r->nm >= 0 && r->nm <= 65535 && r->n_blocks <= 255 && !(r->aux & RSQ_AUX_IMPLICIT_BLOCK);
            if (ok) {
                if (r->n_blocks == 0) ok = r->endpos == r->pos + 1;
                else if (r->n_blocks == 1 && rb[0].start == r->pos + 1 && rb[0].len == r->l_qseq && r->endpos == r->pos + r->l_qseq) implicit = 1;
                else {
                    ok = r->endpos == rb[r->n_blocks - 1].start - 1 + rb[r->n_blocks - 1].len;
                    for (b = 0; ok && b < r->n_blocks; ++b) {
                        const long long o = (long long)rb[b].start - ((long long)r->pos + 1);
                        ok = o >= 0 && o < (1ll << 21) && rb[b].len >= 0 && rb[b].len < (1 << RSQ_BLOCK16_LEN_BITS);
                    }
                }
            }
            p->name_hash = r->name_hash;
            if (ok) {      /* dictionary look-up (open addressing over `scratch`) */
                rsq_combo16 c;
                uint32_t h, slot;
                c.flag = r->flag; c.mapq = r->mapq; c.aux = (uint8_t)(r->aux | (implicit ? RSQ_AUX_IMPLICIT_BLOCK : 0)); c.l_qseq = (uint16_t)r->l_qseq;
                c.nm = (uint16_t)r->nm; c.n_blocks = (uint8_t)r->n_blocks; c.filter_mask = r->filter_mask; c.reserved = 0; c.reserved2 = 0;
                h = (uint32_t)c.flag * 0x9E3779B1u ^ ((uint32_t)c.mapq << 8 | c.aux) * 0x85EBCA77u ^ ((uint32_t)c.l_qseq << 16 | c.nm) * 0xC2B2AE3Du ^
                    ((uint32_t)c.n_blocks << 8 | c.filter_mask) * 0x27D4EB2Fu;
                h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
                for (slot = h & 131071u;; slot = (slot + 1) & 131071u) {
                    const uint32_t k = scratch[slot];
                    if (k == 0xffffffffu) {
                        if (n_combos >= 0xFFFF) { ok = 0; break; }
                        scratch[slot] = (uint32_t)n_combos; D[n_combos] = c; p->combo = (uint16_t)n_combos++;
                        break;
                    }
                    if (D[k].flag == c.flag && D[k].mapq == c.mapq && D[k].aux == c.aux && D[k].l_qseq == c.l_qseq && D[k].nm == c.nm &&
                        D[k].n_blocks == c.n_blocks && D[k].filter_mask == c.filter_mask) { p->combo = (uint16_t)k; break; }
                }
            }
            if (ok) {
                p->dpos = (uint16_t)dpos; p->dmate = (int16_t)dm; p->isize = (int16_t)r->isize;
                if (!implicit)
                    for (b = 0; b < r->n_blocks; ++b)
                        XB[n_xb++] = ((uint32_t)(rb[b].start - (r->pos + 1)) << RSQ_BLOCK16_LEN_BITS) | (uint32_t)rb[b].len;
            } else {
                if (n_exc >= max_exc) return RSQ_ERR_INVALID;
                p->dpos = 0; p->dmate = 0; p->isize = 0; p->combo = RSQ_COMBO_EXCEPTION;
                E[n_exc] = *r; E[n_exc].block_off = (uint32_t)n_eb; ++n_exc;
                for (b = 0; b < r->n_blocks; ++b) EB[n_eb++] = rb[b];
            }
            out_blocks += r->n_blocks;
        }
    }
    if (n_exc * 8 > n_reads + 256) return RSQ_ERR_INVALID;
    /* close the gaps behind the dictionary */
    off = RSQ_P16_ALIGN(H->off_combos + 16 * n_combos);
    H->off_blocks = off; memmove(base + off, XB, 4 * n_xb); off = RSQ_P16_ALIGN(off + 4 * n_xb);
    H->off_exc = off; memmove(base + off, E, 48 * n_exc); off = RSQ_P16_ALIGN(off + 48 * n_exc);
    H->off_exc_blocks = off; memmove(base + off, EB, 8 * n_eb); off = RSQ_P16_ALIGN(off + 8 * n_eb);
#undef RSQ_P16_ALIGN
    H->magic = RSQ_BATCH16_MAGIC; H->struct_size = (uint32_t)sizeof(rsq_batch16);
    H->n_reads = (uint32_t)n_reads; H->n_tiles = (uint32_t)n_tiles; H->n_combos = (uint32_t)n_combos; H->n_blocks = (uint32_t)n_xb;
    H->n_exc = (uint32_t)n_exc; H->n_exc_blocks = (uint32_t)n_eb; H->total_blocks = (uint32_t)out_blocks; H->reserved = 0;
    H->total_bytes = off;
    *bytes_out = off;
    return RSQ_OK;
}

/* The same for a batch that lies in one piece (block_off relative to `blocks`). */
static inline int rsq_pack_compact(const rsq_read* reads, size_t n_reads, const rsq_block* blocks, size_t n_blocks, void* out, size_t cap,
                                   uint32_t* scratch, size_t* bytes_out) {
    rsq_segment one;
    one.reads = reads; one.tids = NULL; one.n_reads = n_reads; one.blocks = blocks; one.n_blocks = n_blocks;
    return rsq_pack_compact_segments(&one, 1, NULL, 0, out, cap, scratch, bytes_out);
}
#endif /* RSQ_B200_H */
