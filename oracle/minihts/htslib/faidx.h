/* minihts -- faidx subset.  The index object's layout is the one src/read_reference.c:18-33 re-declares for itself. */
#ifndef MINIHTS_FAIDX_H
#define MINIHTS_FAIDX_H
enum fai_format_options { FAI_NONE, FAI_FASTA, FAI_FASTQ };
typedef struct __faidx_t faidx_t;
faidx_t *fai_load(const char *fn);
void fai_destroy(faidx_t *fai);
int faidx_nseq(const faidx_t *fai);
const char *faidx_iseq(const faidx_t *fai, int i);
int faidx_seq_len(const faidx_t *fai, const char *seq);
#endif
