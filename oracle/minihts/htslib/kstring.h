/* minihts -- kstring subset.  Part of the htslib stand-in described in ../minihts.c (TEST / MEASUREMENT INFRASTRUCTURE). */
#ifndef MINIHTS_KSTRING_H
#define MINIHTS_KSTRING_H
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
typedef struct kstring_t { size_t l, m; char *s; } kstring_t;
static inline int ks_resize(kstring_t *s, size_t size) {
	if (s->m < size) {
		size_t m = s->m ? s->m : 64;
		while (m < size) m <<= 1;
		char *p = (char *)realloc(s->s, m);
		if (!p) return -1;
		s->s = p; s->m = m;
	}
	return 0;
}
static inline int kputsn_(const void *p, size_t l, kstring_t *s) { if (ks_resize(s, s->l + l + 1)) return -1; memcpy(s->s + s->l, p, l); s->l += l; return (int)l; }
static inline int kputsn(const char *p, size_t l, kstring_t *s) { if (kputsn_(p, l, s) < 0) return -1; s->s[s->l] = 0; return (int)l; }
static inline int kputs(const char *p, kstring_t *s) { return kputsn(p, strlen(p), s); }
static inline int kputc_(int c, kstring_t *s) { if (ks_resize(s, s->l + 2)) return -1; s->s[s->l++] = (char)c; return 1; }
static inline int kputc(int c, kstring_t *s) { if (kputc_(c, s) < 0) return -1; s->s[s->l] = 0; return (unsigned char)c; }
#endif
