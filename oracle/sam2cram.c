/* sam2cram.c — TEST INFRASTRUCTURE.  Writes a CRAM from a SAM/BAM with the reference's own vendored htslib 1.5 (linked
 * from the objects oracle/Makefile.ref builds into oracle/_ref/), so that the tests can feed the same CRAM to the unmodified
 * reference binary and to this repo's own CRAM reader (rnaseqc_b200/csrc/host/cramreader.cpp).  Never part of the product.
 * An output name that ends in ".bam" gives a BAM instead (the reference FASTA is not read then): odd records for the BAM readers.
 *   usage: sam2cram in.sam ref.fa out.cram [records_per_slice [slices_per_container [version [lossy_names [multi_ref [no_rans]]]]]] */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "htslib/hts.h"
#include "htslib/sam.h"

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: sam2cram in.sam ref.fa out.cram [records_per_slice [slices_per_container]]\n"); return 2; }
    htsFile* in = hts_open(argv[1], "r");
    if (!in) { perror(argv[1]); return 1; }
    bam_hdr_t* h = sam_hdr_read(in);
    if (!h) { fprintf(stderr, "no header\n"); return 1; }
    const size_t ln = strlen(argv[3]);
    const int bam = ln > 4 && strcmp(argv[3] + ln - 4, ".bam") == 0;
    htsFile* out = hts_open(argv[3], bam ? "wb" : "wc");
    if (!out) { perror(argv[3]); return 1; }
    if (!bam && hts_set_fai_filename(out, argv[2]) != 0) { fprintf(stderr, "cannot use reference %s\n", argv[2]); return 1; }
    if (argc > 4) { int n = atoi(argv[4]); hts_set_opt(out, CRAM_OPT_SEQS_PER_SLICE, n); }
    if (argc > 5) { int n = atoi(argv[5]); hts_set_opt(out, CRAM_OPT_SLICES_PER_CONTAINER, n); }
    if (argc > 6 && argv[6][0] != '-') hts_set_opt(out, CRAM_OPT_VERSION, argv[6]);              /* "2.1" or "3.0"; must precede the header */
    if (argc > 7) hts_set_opt(out, CRAM_OPT_LOSSY_NAMES, atoi(argv[7]));                          /* 1: read names are not stored */
    if (argc > 8) hts_set_opt(out, CRAM_OPT_MULTI_SEQ_PER_SLICE, atoi(argv[8]));                  /* 1: slices may span contigs */
    if (argc > 9 && atoi(argv[9])) hts_set_opt(out, CRAM_OPT_USE_RANS, 0);                        /* 1: gzip blocks only */
    if (sam_hdr_write(out, h) != 0) { fprintf(stderr, "header write failed\n"); return 1; }
    bam1_t* b = bam_init1();
    long n = 0;
    int r;
    while ((r = sam_read1(in, h, b)) >= 0) {
        if (sam_write1(out, h, b) < 0) { fprintf(stderr, "write failed at record %ld\n", n); return 1; }
        ++n;
    }
    if (r < -1) { fprintf(stderr, "read error after %ld records\n", n); return 1; }
    bam_destroy1(b);
    bam_hdr_destroy(h);
    if (hts_close(out) != 0) { fprintf(stderr, "close failed\n"); return 1; }
    hts_close(in);
    fprintf(stderr, "sam2cram: %ld records\n", n);
    return 0;
}
