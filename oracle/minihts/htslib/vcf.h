/* minihts -- vcf.h subset: the BCF2 header object as far as src/print_vcf.c reads it, the record object, the typed-value
 * encoders (VCF/BCF specification v4.3 section 6.3) and bcf_write for BCF and VCF text output. */
#ifndef MINIHTS_VCF_H
#define MINIHTS_VCF_H
#include <stdint.h>
#include "hts.h"
#define FT_UNKN 0
#define FT_GZ 1
#define FT_VCF 2
#define FT_VCF_GZ 3
#define FT_BCF 4
#define FT_BCF_GZ 5
#define BCF_BT_NULL   0
#define BCF_BT_INT8   1
#define BCF_BT_INT16  2
#define BCF_BT_INT32  3
#define BCF_BT_FLOAT  5
#define BCF_BT_CHAR   7
#define bcf_int8_vector_end  (-127)
#define bcf_int16_vector_end (-32767)
#define bcf_int32_vector_end (-2147483647)
#define bcf_int8_missing     (-128)
#define bcf_int16_missing    (-32767 - 1)
#define bcf_int32_missing    (-2147483647 - 1)
#define BCF_MAX_BT_INT8  (0x7f)
#define BCF_MAX_BT_INT16 (0x7fff)
#define BCF_MIN_BT_INT8  (-120)
#define BCF_MIN_BT_INT16 (-32760)
#define BCF_DT_ID  0
#define BCF_DT_CTG 1
#define BCF_DT_SAMPLE 2
typedef struct bcf_idinfo_t { uint64_t info[3]; void *hrec[3]; int id; } bcf_idinfo_t;
typedef struct bcf_idpair_t { const char *key; const bcf_idinfo_t *val; } bcf_idpair_t;
typedef struct bcf_hdr_t {
	int32_t n[3];
	bcf_idpair_t *id[3];
	void *dict[3];
	char **samples;
	void *priv;                       /* minihts: the header lines in order */
} bcf_hdr_t;
typedef struct bcf1_t {
	hts_pos_t pos;
	hts_pos_t rlen;
	int32_t rid;
	float qual;
	uint32_t n_info:16, n_allele:16;
	uint32_t n_fmt:8, n_sample:24;
	kstring_t shared, indiv;
} bcf1_t;
void bcf_enc_size(kstring_t *s, int size, int type);
void bcf_enc_int1(kstring_t *s, int32_t x);
void bcf_enc_vint(kstring_t *s, int n, int32_t *a, int wsize);
void bcf_enc_vfloat(kstring_t *s, int n, float *a);
void bcf_enc_vchar(kstring_t *s, int l, const char *a);
bcf1_t *bcf_init(void);
void bcf_destroy(bcf1_t *v);
void bcf_clear(bcf1_t *v);
int bcf_write(htsFile *fp, bcf_hdr_t *h, bcf1_t *v);
bcf_hdr_t *bcf_hdr_init(const char *mode);
void bcf_hdr_destroy(bcf_hdr_t *h);
int bcf_hdr_append(bcf_hdr_t *h, const char *line);
int bcf_hdr_printf(bcf_hdr_t *h, const char *format, ...);
const char *bcf_hdr_get_version(const bcf_hdr_t *hdr);
int bcf_hdr_add_sample(bcf_hdr_t *hdr, const char *sample);
int bcf_hdr_write(htsFile *fp, bcf_hdr_t *h);
#endif
