/* Minimal stand-in for <htslib/vcf.h> (+ the kstring surface it pulls in): TEST INFRASTRUCTURE ONLY.
 * htslib is a system dependency of bs_call that is not vendored under /root/reference and is absent from this image.
 * The hot-path files only need the type names and the FT_* constants (src/init_param.c); src/print_vcf.c builds BCF
 * records by hand with the typed-value encoders of htslib's public API, so those are declared here with the API's
 * names and signatures and RESTATED from the published BCF2 encoding (VCF/BCF specification v4.3 section 6.3:
 * a type byte = length << 4 | type, lengths >= 15 spilled into a following typed integer; integers little endian
 * in the smallest of int8 / int16 / int32 that holds every value, with the reserved end-of-vector / missing values)
 * in oracle/ref_harness.c.  bcf_write() in the harness captures the record instead of writing a file. */
#ifndef BSGPU_STUB_HTS_VCF_H
#define BSGPU_STUB_HTS_VCF_H
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifndef BSGPU_STUB_HTS_SAM_H
typedef struct htsFile htsFile;
typedef int64_t hts_pos_t;
#endif

#define FT_UNKN 0
#define FT_GZ 1
#define FT_VCF 2
#define FT_VCF_GZ 3
#define FT_BCF 4
#define FT_BCF_GZ 5

/* ---- kstring ---- */
typedef struct { size_t l, m; char *s; } kstring_t;
static inline int bsstub_ks_room(kstring_t *s, size_t extra) {
	if (s->l + extra + 1 > s->m) {
		size_t m = s->m ? s->m : 64;
		while (m < s->l + extra + 1) m <<= 1;
		char *p = (char *)realloc(s->s, m);
		if (!p) return -1;
		s->s = p; s->m = m;
	}
	return 0;
}
static inline int kputsn_(const void *p, size_t l, kstring_t *s) { if (bsstub_ks_room(s, l)) return -1; memcpy(s->s + s->l, p, l); s->l += l; return (int)l; }
static inline int kputsn(const char *p, size_t l, kstring_t *s) { if (kputsn_(p, l, s) < 0) return -1; s->s[s->l] = 0; return (int)l; }
static inline int kputc_(int c, kstring_t *s) { if (bsstub_ks_room(s, 1)) return -1; s->s[s->l++] = (char)c; return 1; }
static inline int kputc(int c, kstring_t *s) { if (kputc_(c, s) < 0) return -1; s->s[s->l] = 0; return (unsigned char)c; }

/* ---- BCF value types and reserved values ---- */
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

typedef struct { uint64_t info[3]; void *hrec[3]; int id; } bcf_idinfo_t;
typedef struct { const char *key; const bcf_idinfo_t *val; } bcf_idpair_t;
typedef struct bcf_hdr_t {
	int32_t n[3];
	bcf_idpair_t *id[3];
	void *dict[3];
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

/* typed-value encoders (restated in oracle/ref_harness.c) */
void bcf_enc_size(kstring_t *s, int size, int type);
void bcf_enc_int1(kstring_t *s, int32_t x);
void bcf_enc_vint(kstring_t *s, int n, int32_t *a, int wsize);
void bcf_enc_vfloat(kstring_t *s, int n, float *a);
void bcf_enc_vchar(kstring_t *s, int l, const char *a);

bcf1_t *bcf_init(void);
void bcf_destroy(bcf1_t *v);
void bcf_clear(bcf1_t *v);
int bcf_write(htsFile *fp, bcf_hdr_t *h, bcf1_t *v);

/* header side: only print_vcf_header() uses these, which the harness never calls */
bcf_hdr_t *bcf_hdr_init(const char *mode);
int bcf_hdr_append(bcf_hdr_t *h, const char *line);
int bcf_hdr_printf(bcf_hdr_t *h, const char *format, ...);
const char *bcf_hdr_get_version(const bcf_hdr_t *hdr);
int bcf_hdr_add_sample(bcf_hdr_t *hdr, const char *sample);
int bcf_hdr_write(htsFile *fp, bcf_hdr_t *h);
htsFile *hts_open(const char *fn, const char *mode);
int hts_set_threads(htsFile *fp, int n);
#endif
