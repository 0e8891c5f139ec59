/* Minimal stand-in for <htslib/khash.h>: TEST INFRASTRUCTURE ONLY.
 * src/print_vcf.c instantiates one string-keyed map type (the BCF header dictionary) and looks sixteen names up in
 * it once, inside print_vcf_header(), which the harness never calls: the macros only have to compile and link. */
#ifndef BSGPU_STUB_HTS_KHASH_H
#define BSGPU_STUB_HTS_KHASH_H
#include <stdint.h>
typedef uint32_t khint_t;
typedef khint_t khiter_t;
khint_t bsstub_kh_get(const void *h, const char *key);
khint_t bsstub_kh_end(const void *h);
#define KHASH_MAP_INIT_STR(name, val_t) typedef struct { khint_t n; val_t *vals; } kh_##name##_t;
#define khash_t(name) kh_##name##_t
#define kh_get(name, h, key) bsstub_kh_get((h), (key))
#define kh_end(h) bsstub_kh_end((h))
#define kh_val(h, k) ((h)->vals[(k)])
#endif
