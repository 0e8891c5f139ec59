/* Minimal stand-in for <htslib/sam.h>: TEST INFRASTRUCTURE ONLY.
 * Declares just the public API surface that the reference's src/input_sam.c uses (record struct, accessor macros,
 * flag constants, the two read calls), so that file compiles UNMODIFIED where htslib is absent.  The other hot-path
 * files only need the type names.  The harness (oracle/ref_harness.c) implements sam_read1() over a memory buffer of
 * raw BAM alignment records; the in-memory data[] layout is the on-disk one (qname, cigar, seq, qual, aux). */
#ifndef BSGPU_STUB_HTS_SAM_H
#define BSGPU_STUB_HTS_SAM_H
#include <stdint.h>

typedef struct htsFile htsFile;
typedef struct hts_idx_t hts_idx_t;
/* the fields of the header the hot path reads (src/get_template_vector.c:123: target_name) */
typedef struct bam_hdr_t {
	int32_t n_targets;
	uint32_t *target_len;
	char **target_name;
	char *text;
} bam_hdr_t;
typedef struct hts_itr_t hts_itr_t;
typedef int64_t hts_pos_t;

typedef struct {
	hts_pos_t pos;
	int32_t tid;
	uint16_t bin;
	uint8_t qual;
	uint8_t l_extranul;
	uint16_t flag;
	uint16_t l_qname;
	uint32_t n_cigar;
	int32_t l_qseq;
	int32_t mtid;
	hts_pos_t mpos;
	hts_pos_t isize;
} bam1_core_t;

typedef struct {
	bam1_core_t core;
	uint64_t id;
	uint8_t *data;
	int l_data;
	uint32_t m_data;
} bam1_t;

#define BAM_FPAIRED        1
#define BAM_FPROPER_PAIR   2
#define BAM_FUNMAP         4
#define BAM_FMUNMAP        8
#define BAM_FREVERSE      16
#define BAM_FMREVERSE     32
#define BAM_FREAD1        64
#define BAM_FREAD2       128
#define BAM_FSECONDARY   256
#define BAM_FQCFAIL      512
#define BAM_FDUP        1024
#define BAM_FSUPPLEMENTARY 2048

#define BAM_CIGAR_STR   "MIDNSHP=XB"
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK  0xf
#define bam_cigar_op(c) ((c) & BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c) >> BAM_CIGAR_SHIFT)
#define bam_cigar_opchr(c) (BAM_CIGAR_STR "??????"[bam_cigar_op(c)])

#define bam_get_qname(b) ((char *)(b)->data)
#define bam_get_cigar(b) ((uint32_t *)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b)   ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b)  ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b)   ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)

int sam_read1(htsFile *fp, bam_hdr_t *h, bam1_t *b);
int sam_itr_next(htsFile *fp, hts_itr_t *itr, bam1_t *b);
hts_itr_t *sam_itr_queryi(const hts_idx_t *idx, int tid, hts_pos_t beg, hts_pos_t end);
void hts_itr_destroy(hts_itr_t *itr);
bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
#endif
