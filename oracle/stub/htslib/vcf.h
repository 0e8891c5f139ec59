/* Opaque stand-ins for htslib VCF types + the FT_* constants the reference's init_param.c uses. */
#ifndef BSGPU_STUB_HTS_VCF_H
#define BSGPU_STUB_HTS_VCF_H
typedef struct bcf_hdr_t bcf_hdr_t;
typedef struct bcf1_t bcf1_t;
#define FT_UNKN 0
#define FT_GZ 1
#define FT_VCF 2
#define FT_VCF_GZ 3
#define FT_BCF 4
#define FT_BCF_GZ 5
#endif
