/* minihts -- sam.h subset: the alignment record, its accessor macros and flag constants (SAM specification section 4.2) and
 * the read calls the reference uses.  Records keep the on-disk variable-length layout in data[] (qname, cigar, seq, qual,
 * aux), no padding of the read name. */
#ifndef MINIHTS_SAM_H
#define MINIHTS_SAM_H
#include <stdint.h>
#include "hts.h"
typedef struct sam_hdr_t {
	int32_t n_targets, ignore_sam_err;
	size_t l_text;
	uint32_t *target_len;
	const int8_t *cigar_tab;
	char **target_name;
	char *text;
	void *sdict;
	void *hrecs;
	uint32_t ref_count;
} sam_hdr_t;
typedef sam_hdr_t bam_hdr_t;
typedef struct bam1_core_t {
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
typedef struct bam1_t {
	bam1_core_t core;
	uint64_t id;
	uint8_t *data;
	int l_data;
	uint32_t m_data;
	uint32_t mempolicy:2, :30;
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
sam_hdr_t *sam_hdr_read(htsFile *fp);
void sam_hdr_destroy(sam_hdr_t *h);
int sam_read1(htsFile *fp, sam_hdr_t *h, bam1_t *b);
hts_idx_t *sam_index_load(htsFile *fp, const char *fn);
hts_itr_t *sam_itr_queryi(const hts_idx_t *idx, int tid, hts_pos_t beg, hts_pos_t end);
int sam_itr_next(htsFile *fp, hts_itr_t *itr, bam1_t *b);
int bam_name2id(sam_hdr_t *h, const char *ref);
bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
#endif
