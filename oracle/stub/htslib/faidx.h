/* Opaque stand-in for htslib's faidx_t. */
#ifndef BSGPU_STUB_HTS_FAIDX_H
#define BSGPU_STUB_HTS_FAIDX_H
typedef struct __faidx_t faidx_t;
#endif
