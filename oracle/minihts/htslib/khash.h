/* minihts -- khash subset: string-keyed open-addressing maps behind htslib's macro names (KHASH_MAP_INIT_STR, khash_t,
 * kh_init / kh_get / kh_put / kh_val / kh_value / kh_key / kh_end / kh_exist / kh_size / kh_destroy).  Written for this
 * repository (linear probing over a power-of-two table, no deletion), not htslib's implementation: only the API is the same.
 * src/read_reference.c and src/print_vcf.c instantiate their own map types with these macros and read maps that
 * minihts.c filled through the same macros, so both sides agree on the layout. */
#ifndef MINIHTS_KHASH_H
#define MINIHTS_KHASH_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
typedef uint32_t khint_t;
typedef khint_t khiter_t;
typedef const char *kh_cstr_t;
static inline khint_t minihts_strhash(const char *s) { khint_t h = 2166136261u; for (; *s; s++) h = (h ^ (uint8_t)*s) * 16777619u; return h; }
#define KHASH_MAP_INIT_STR(name, val_t) \
	typedef struct kh_##name##_s { khint_t n_buckets, size; uint8_t *used; kh_cstr_t *keys; val_t *vals; } kh_##name##_t; \
	static inline __attribute__((unused)) kh_##name##_t *kh_init_##name(void) { return (kh_##name##_t *)calloc(1, sizeof(kh_##name##_t)); } \
	static inline __attribute__((unused)) void kh_destroy_##name(kh_##name##_t *h) { if (h) { free(h->used); free((void *)h->keys); free(h->vals); free(h); } } \
	static inline __attribute__((unused)) khint_t kh_get_##name(const kh_##name##_t *h, kh_cstr_t key) { \
		if (!h->n_buckets) return 0; \
		const khint_t mask = h->n_buckets - 1; \
		for (khint_t i = minihts_strhash(key) & mask;; i = (i + 1) & mask) { \
			if (!h->used[i]) return h->n_buckets; \
			if (!strcmp(h->keys[i], key)) return i; \
		} \
	} \
	static inline __attribute__((unused)) khint_t kh_put_##name(kh_##name##_t *h, kh_cstr_t key, int *ret) { \
		if ((h->size + 1) * 2 > h->n_buckets) { \
			const khint_t nb = h->n_buckets ? h->n_buckets * 2 : 16; \
			uint8_t *u = (uint8_t *)calloc(nb, 1); kh_cstr_t *k = (kh_cstr_t *)calloc(nb, sizeof(kh_cstr_t)); val_t *v = (val_t *)calloc(nb, sizeof(val_t)); \
			for (khint_t i = 0; i < h->n_buckets; i++) if (h->used[i]) { \
				khint_t j = minihts_strhash(h->keys[i]) & (nb - 1); \
				while (u[j]) j = (j + 1) & (nb - 1); \
				u[j] = 1; k[j] = h->keys[i]; v[j] = h->vals[i]; \
			} \
			free(h->used); free((void *)h->keys); free(h->vals); \
			h->used = u; h->keys = k; h->vals = v; h->n_buckets = nb; \
		} \
		const khint_t mask = h->n_buckets - 1; \
		khint_t i = minihts_strhash(key) & mask; \
		for (; h->used[i]; i = (i + 1) & mask) if (!strcmp(h->keys[i], key)) { *ret = 0; return i; } \
		h->used[i] = 1; h->keys[i] = key; h->size++; *ret = 1; \
		return i; \
	}
#define khash_t(name) kh_##name##_t
#define kh_init(name) kh_init_##name()
#define kh_destroy(name, h) kh_destroy_##name(h)
#define kh_get(name, h, k) kh_get_##name(h, k)
#define kh_put(name, h, k, r) kh_put_##name(h, k, r)
#define kh_exist(h, x) ((h)->used[x])
#define kh_key(h, x) ((h)->keys[x])
#define kh_val(h, x) ((h)->vals[x])
#define kh_value(h, x) ((h)->vals[x])
#define kh_begin(h) ((khint_t)0)
#define kh_end(h) ((h)->n_buckets)
#define kh_size(h) ((h)->size)
#endif
