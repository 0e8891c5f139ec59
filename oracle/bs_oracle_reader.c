/*
 * bs_oracle_reader.c -- TEST INFRASTRUCTURE ONLY (see bs_oracle.h).
 *
 * CPU restatement of the reader side of the bs_call 2.1.7 hot path:
 *   bso_decode_records   BAM alignment record -> filter verdict, positions, packed read, CIGAR events, strand
 *                        src/input_sam.c:42-88 (get_seq_and_qual), 90-136 (get_bam_misms), 144-220 (get_bs_strand),
 *                        222-312 (get_next_align_details)
 *   bso_read_input       mate pairing, positional duplicate removal, block cutting
 *                        src/get_template_vector.c:49-389 (read_input), 18-45 (handle_end_of_block)
 * Input is the byte stream that follows the header in an uncompressed BAM file.  The restatement keeps templates as
 * indices into the decoded records instead of the reference's pointer-swapped align_details objects; behaviour
 * (including the quirks noted inline) is pinned against the reference's own compiled sources by
 * tests/test_oracle_vs_reference.py.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bs_oracle.h"

enum { F_PAIRED = 1, F_PROPER = 2, F_UNMAP = 4, F_MUNMAP = 8, F_REVERSE = 16, F_READ2 = 128, F_SECONDARY = 256,
       F_QCFAIL = 512, F_DUP = 1024, F_SUPP = 2048 };
/* gt_filter_reason, include/bs_call.h:50 */
enum { FLT_NONE = 0, FLT_UNMAPPED, FLT_QC, FLT_SECONDARY, FLT_MATE_UNMAPPED, FLT_DUPLICATE, FLT_NOPOS, FLT_NOMATEPOS,
       FLT_MISMATCH_CHR, FLT_ORIENTATION, FLT_INSERT_SIZE, FLT_NOSEQ, FLT_MAPQ, FLT_NOT_CORRECTLY_ALIGNED };

static int32_t rd_i32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }
static uint32_t rd_u32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint16_t rd_u16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

typedef struct {
	int32_t tid, pos, l_qseq, mtid, mpos, isize;
	uint32_t l_qname, mapq, n_cigar, flag;
	const uint8_t *qname, *cigar, *seq, *qual, *aux, *end;
} bam_view;

/* 0 ok, -1 end of stream, -2 truncated */
static int next_record(const uint8_t *bam, size_t nbytes, size_t *at, bam_view *v) {
	if (*at == nbytes) return -1;
	if (*at + 4 > nbytes) return -2;
	const int32_t bs = rd_i32(bam + *at);
	if (bs < 32 || *at + 4 + (size_t)bs > nbytes) return -2;
	const uint8_t *p = bam + *at + 4;
	v->tid = rd_i32(p); v->pos = rd_i32(p + 4); v->l_qname = p[8]; v->mapq = p[9];
	v->n_cigar = rd_u16(p + 12); v->flag = rd_u16(p + 14); v->l_qseq = rd_i32(p + 16);
	v->mtid = rd_i32(p + 20); v->mpos = rd_i32(p + 24); v->isize = rd_i32(p + 28);
	v->qname = p + 32;
	v->cigar = v->qname + v->l_qname;
	v->seq = v->cigar + 4 * (size_t)v->n_cigar;
	v->qual = v->seq + ((size_t)v->l_qseq + 1) / 2;
	v->aux = v->qual + v->l_qseq;
	v->end = p + bs;
	*at += 4 + (size_t)bs;
	return 0;
}

/* src/input_sam.c:42-59: nibble -> code (A 1, C 2, G 3, T 4, anything else 0) */
static inline uint8_t nib_code(uint8_t n) { return n == 1 ? 1 : n == 2 ? 2 : n == 4 ? 3 : n == 8 ? 4 : 0; }

/* src/input_sam.c:61-88 */
static void decode_seq(const bam_view *v, uint8_t *out) {
	for (int32_t k = 0; k < v->l_qseq; k++) {
		const uint8_t byte = v->seq[k >> 1];
		const uint8_t c = nib_code((k & 1) ? (byte & 15) : (byte >> 4));
		uint8_t q = v->qual[k];
		if (q > 43) q = 43;                                  /* MAX_QUAL */
		out[k] = c ? (uint8_t)((c - 1) | (q << 2)) : 0;      /* N: the whole byte is 0 */
	}
}

/* src/input_sam.c:90-136.  Returns the read length the CIGAR implies; *span = reference span. */
static uint32_t decode_cigar(const bam_view *v, bso_misms *mm, uint32_t *nmm, uint32_t *span) {
	uint32_t position = 0, reference_span = 0, n = 0;
	for (uint32_t i = 0; i < v->n_cigar; i++) {
		const uint32_t c = rd_u32(v->cigar + 4 * (size_t)i), len = c >> 4, op = c & 15;
		switch (op) {
		case 0: case 7: case 8:            /* M = X */
			position += len; reference_span += len; break;
		case 6: case 4:                    /* P and S: soft clip */
			mm[n].type = BSO_SOFT; mm[n].position = position; mm[n].size = len; n++; position += len; break;
		case 1:                            /* I: bases the reference lacks -> DEL */
			mm[n].type = BSO_DEL; mm[n].position = position; mm[n].size = len; n++; position += len; break;
		case 2:                            /* D: reference bases the read lacks -> INS */
			mm[n].type = BSO_INS; mm[n].position = position; mm[n].size = len; n++; reference_span += len; break;
		default: break;                    /* H, N, B and undefined codes: ignored */
		}
	}
	*nmm = n; *span = reference_span;
	return position;
}

/* src/input_sam.c:144-220 */
static uint8_t decode_strand(const bam_view *v) {
	static const int sub_size[256] = { ['A'] = 1, ['C'] = 1, ['c'] = 1, ['s'] = 2, ['S'] = 2, ['i'] = 4, ['I'] = 4,
		['f'] = 4, ['d'] = 8, ['Z'] = 'Z', ['H'] = 'H', ['B'] = 'B' };
	enum { UNK, GEM, BOWTIE, NOVO, BSMAP, BWAMETH };
	uint8_t strand = 0;
	const uint8_t *s = v->aux, *end = v->end;
	int ok = 1;
	while (ok && s + 4 <= end) {
		int al = UNK;
		if (s[0] == 'Z') { if (s[1] == 'B') al = NOVO; else if (s[1] == 'S') al = BSMAP; }
		else if (s[0] == 'X') { if (s[1] == 'G') al = BOWTIE; else if (s[1] == 'B') al = GEM; }
		else if (s[0] == 'Y' && s[1] == 'D') al = BWAMETH;
		s += 2;
		const uint8_t type = *s++;
		switch (type) {
		case 'A':
			if (al == GEM) { if (*s == 'C') strand = 1; else if (*s == 'G') strand = 2; }
			s++; break;
		case 'C': case 'c': s++; break;
		case 'S': case 's': if (s + 2 <= end) s += 2; else ok = 0; break;
		case 'I': case 'i': case 'f': if (s + 4 <= end) s += 4; else ok = 0; break;
		case 'd': if (s + 8 <= end) s += 8; else ok = 0; break;
		case 'Z':
			if (al == BOWTIE || al == NOVO) { if (*s == 'C') strand = 1; else if (*s == 'G') strand = 2; }
			else if (al == BSMAP) { if (*s == '+') strand = 1; else if (*s == '-') strand = 2; }
			else if (al == BWAMETH) { if (*s == 'f') strand = 1; else if (*s == 'r') strand = 2; }
			/* fall through */
		case 'H':
			while (s < end && *s) s++;
			if (s < end) s++; else ok = 0;
			break;
		case 'B': {
			const int sz = sub_size[*s++];
			if (s + 4 <= end && sz != 0) {
				const uint32_t n = rd_u32(s);
				s += 4;
				const uint32_t bytes = n * (uint32_t)sz;         /* 32-bit product, like the reference */
				if (s + bytes <= end) s += bytes; else ok = 0;
			} else ok = 0;
			break; }
		default: break;                                          /* unknown type: nothing consumed */
		}
	}
	return strand;
}

/* the record header part of get_next_align_details, src/input_sam.c:231-300 */
static void classify(const bam_view *v, uint32_t thresh, uint32_t max_tlen, int keep_unmatched, int ignore_dup, bso_record *o) {
	const uint32_t flag = v->flag;
	uint32_t flt = FLT_NONE;
	if ((flag & F_PAIRED) && !keep_unmatched) {
		if ((flag & (F_PROPER | F_UNMAP | F_MUNMAP | F_QCFAIL | F_SECONDARY | F_SUPP | F_DUP)) != F_PROPER) {
			if (flag & (F_SECONDARY | F_SUPP)) flt = FLT_SECONDARY;
			else if (flag & F_UNMAP) flt = FLT_UNMAPPED;
			else if (flag & F_MUNMAP) flt = FLT_MATE_UNMAPPED;
			else if (flag & F_QCFAIL) flt = FLT_QC;
			else if (flag & F_DUP) { if (!ignore_dup) flt = FLT_DUPLICATE; }
			else flt = FLT_NOT_CORRECTLY_ALIGNED;
		}
	} else if (flag & (F_UNMAP | F_QCFAIL | F_SECONDARY | F_SUPP | F_DUP)) {
		if (flag & (F_SECONDARY | F_SUPP)) flt = FLT_SECONDARY;
		else if (flag & F_UNMAP) flt = FLT_UNMAPPED;
		else if (flag & F_QCFAIL) flt = FLT_QC;
		else if (flag & F_DUP) flt = FLT_DUPLICATE;
	}
	int mis_matched = (flag & (F_MUNMAP | F_PROPER)) != F_PROPER;
	const int reverse = (flag & F_REVERSE) != 0, second = (flag & F_READ2) != 0;
	o->reverse = (uint8_t)reverse;
	o->orientation = ((second && reverse) || !(second || reverse)) ? 0 : 1;
	const int mult_seg = (flag & (F_PAIRED | F_MUNMAP)) == F_PAIRED;
	if (reverse) { o->forward_position = (uint32_t)(v->mpos + 1); o->reverse_position = (uint32_t)(v->pos + 1); }
	else { o->forward_position = (uint32_t)(v->pos + 1); o->reverse_position = (uint32_t)(v->mpos + 1); }
	o->mapq = (uint8_t)v->mapq;
	if (v->mapq < thresh && !flt) flt = FLT_MAPQ;
	uint32_t aflag = flag;
	if (mult_seg) {
		if (v->tid != v->mtid) { if (!flt) flt = FLT_MISMATCH_CHR; if (keep_unmatched) mis_matched = 1; }
		if (!flt) {
			const int64_t is = v->isize < 0 ? -(int64_t)v->isize : (int64_t)v->isize;
			if ((uint64_t)is > (uint64_t)max_tlen) { flt = FLT_INSERT_SIZE; if (keep_unmatched) mis_matched = 1; }
		}
		if (reverse) {
			if (v->pos < v->mpos) { if (!flt) flt = FLT_ORIENTATION; if (keep_unmatched) mis_matched = 1; }
			if (mis_matched) o->forward_position = 0;
		} else {
			if (v->pos > v->mpos) { if (!flt) flt = FLT_ORIENTATION; if (keep_unmatched) mis_matched = 1; }
			if (mis_matched) o->reverse_position = 0;
		}
	}
	if (!mult_seg || mis_matched) aflag &= ~(uint32_t)F_PAIRED;
	o->alignment_flag = aflag;
	o->filtered = flt;
	o->ret = 0;
	if (flt && !(keep_unmatched && (flt == FLT_INSERT_SIZE || flt == FLT_MISMATCH_CHR || flt == FLT_ORIENTATION))) o->ret = 1;
}

int bso_decode_records(const uint8_t *bam, size_t nbytes, int mapq_thresh, uint32_t max_template_len, int keep_unmatched,
		int ignore_dup, bso_record *out, size_t cap, size_t *nrec, uint8_t *bases_out, size_t bases_cap, size_t *nbases,
		bso_misms *misms_out, size_t misms_cap, size_t *nmisms) {
	size_t at = 0, n = 0, nb = 0, nm = 0;
	int rc = 0;
	for (;;) {
		bam_view v;
		const int r = next_record(bam, nbytes, &at, &v);
		if (r) { if (r != -1) rc = -2; break; }
		if (n >= cap) { rc = -3; break; }
		bso_record *o = out + n++;
		memset(o, 0, sizeof(*o));
		classify(&v, (uint32_t)mapq_thresh, max_template_len, keep_unmatched, ignore_dup, o);
		if (o->ret == 0) {                 /* sequence, CIGAR and tags are only decoded for records that are kept */
			if (nb + (size_t)v.l_qseq > bases_cap || nm + v.n_cigar > misms_cap) { rc = -3; break; }
			o->align_length = decode_cigar(&v, misms_out + nm, &o->mm_n, &o->reference_span);
			o->mm_off = (uint32_t)nm; nm += o->mm_n;
			decode_seq(&v, bases_out + nb);
			o->read_off = (uint32_t)nb; o->read_len = (uint32_t)v.l_qseq; nb += (size_t)v.l_qseq;
			o->bs_strand = decode_strand(&v);
		}
	}
	*nrec = n; *nbases = nb; *nmisms = nm;
	return rc;
}

/* ---------------------------------------------------------------------------------------------------------------
 * read_input.  A template is {positions, per-mate record index}.  `hash` maps a read name to the list slot of the
 * template that waits for its mate (uthash keyed on QNAME in the reference, :227-241, 328-332).
 * --------------------------------------------------------------------------------------------------------------- */
typedef struct {
	uint32_t fwd, rev, span[2];
	int64_t rec[2];                /* decoded record of each mate, -1 = none */
	uint8_t mapq[2], orientation, bs_strand;
} tmpl_s;

typedef struct hent { const uint8_t *name; uint32_t len, flag, ix; int live; struct hent *next; } hent;

#define NBUCKET 4096
typedef struct { hent *bucket[NBUCKET]; } htab;

static uint32_t name_hash(const uint8_t *s, uint32_t n) { uint32_t h = 2166136261u; for (uint32_t i = 0; i < n; i++) h = (h ^ s[i]) * 16777619u; return h; }
static hent *h_find(htab *t, const uint8_t *s, uint32_t n) {
	for (hent *e = t->bucket[name_hash(s, n) % NBUCKET]; e; e = e->next) if (e->live && e->len == n && !memcmp(e->name, s, n)) return e;
	return NULL;
}
static hent *h_add(htab *t, const uint8_t *s, uint32_t n, uint32_t flag, uint32_t ix) {
	hent *e = calloc(1, sizeof(hent));
	const uint32_t b = name_hash(s, n) % NBUCKET;
	e->name = s; e->len = n; e->flag = flag; e->ix = ix; e->live = 1; e->next = t->bucket[b]; t->bucket[b] = e;
	return e;
}
static void h_clear(htab *t) {
	for (int b = 0; b < NBUCKET; b++) { hent *e = t->bucket[b]; while (e) { hent *nx = e->next; free(e); e = nx; } t->bucket[b] = NULL; }
}

typedef struct {
	const bso_record *rec; const uint8_t *rbases;
	tmpl_s *list; hent **list_h; size_t used, cap;
} rd_state;

/* get_al_qual, src/al_utils.c:19-35: the reference indexes sq[k] with the MATE index, so the "mean quality" of a
 * template is the quality of byte k of mate k, weighted by read length */
static uint32_t al_qual(const rd_state *st, const tmpl_s *t) {
	uint32_t qual = 0, n = 0;
	for (int k = 0; k < 2; k++) {
		if (t->rec[k] < 0) continue;
		const bso_record *r = st->rec + t->rec[k];
		const uint8_t q = st->rbases[r->read_off + (uint32_t)k] >> 2;
		if (q != 63) { qual += q * r->read_len; n += r->read_len; }
	}
	return n > 0 ? qual / n : 0;
}

static void list_put(rd_state *st, size_t ix, const tmpl_s *t, hent *h) {
	if (ix >= st->cap) {
		st->cap = st->cap ? 2 * st->cap : 256;
		st->list = realloc(st->list, st->cap * sizeof(tmpl_s));
		st->list_h = realloc(st->list_h, st->cap * sizeof(hent *));
	}
	st->list[ix] = *t; st->list_h[ix] = h;
	if (st->used <= ix) st->used = ix + 1;
}

typedef struct {
	bso_block *blocks; size_t block_cap, nblocks;
	bso_template *tmpl; size_t tmpl_cap, ntmpl;
	int err;
} rd_out;

static void publish(rd_state *st, rd_out *o, uint32_t tid, uint32_t y) {
	if (!st->used) return;
	if (o->nblocks >= o->block_cap || o->ntmpl + st->used > o->tmpl_cap) { o->err = -3; return; }
	bso_block *bk = o->blocks + o->nblocks++;
	memset(bk, 0, sizeof(*bk));
	const uint32_t first = st->list[0].fwd ? st->list[0].fwd : st->list[0].rev;
	bk->tid = tid; bk->y = y; bk->x = first > 2 ? first - 2 : 1;          /* src/process_template.c:24-28 */
	bk->first_template = (uint32_t)o->ntmpl; bk->n_templates = (uint32_t)st->used;
	for (size_t i = 0; i < st->used; i++) {
		const tmpl_s *t = st->list + i;
		bso_template *d = o->tmpl + o->ntmpl++;
		memset(d, 0, sizeof(*d));
		d->forward_position = t->fwd; d->reverse_position = t->rev;
		d->orientation = t->orientation; d->bs_strand = t->bs_strand;
		for (int k = 0; k < 2; k++) {
			d->mapq[k] = t->mapq[k];
			if (t->rec[k] < 0) continue;
			const bso_record *r = st->rec + t->rec[k];
			d->present[k] = 1; d->reference_span[k] = t->span[k];
			d->read_off[k] = r->read_off; d->read_len[k] = r->read_len; d->mm_off[k] = r->mm_off; d->mm_n[k] = r->mm_n;
		}
	}
	st->used = 0;
}

/* Templates point into the arrays bso_decode_records produced (rec / rbases); names come from the BAM stream. */
static int build_blocks(const uint8_t *bam, size_t nbytes, const bso_record *rec, size_t nrec, const uint8_t *rbases,
		int keep_unmatched, int keep_duplicates, rd_out *o) {
	rd_state st;
	memset(&st, 0, sizeof(st));
	st.rec = rec; st.rbases = rbases;
	htab *hash = calloc(1, sizeof(htab));
	int curr_tid = -1, old_tid = -1;
	uint32_t max_pos = 0, start_pos = 0, read_idx = 0, curr_pos = 0, start_idx = 0;
	size_t at = 0;
	for (size_t ri = 0; ri < nrec && !o->err; ri++) {
		bam_view v;
		if (next_record(bam, nbytes, &at, &v)) { o->err = -2; break; }
		const bso_record *r = rec + ri;
		if (r->ret > 0) {                                             /* filtered (:100-106) */
			bso_profile_tally((int)r->filtered, 1, (int)r->filtered, (uint64_t)(uint32_t)v.l_qseq);
			continue;
		}
		const int reverse = r->reverse, ix = reverse ? 1 : 0;
		tmpl_s al;                                                     /* the incoming alignment as a one-mate template */
		memset(&al, 0, sizeof(al));
		al.fwd = r->forward_position; al.rev = r->reverse_position; al.orientation = r->orientation; al.bs_strand = r->bs_strand;
		al.rec[0] = al.rec[1] = -1; al.rec[ix] = (int64_t)ri; al.mapq[ix] = r->mapq; al.span[ix] = r->reference_span;
		int new_block = 0, new_contig = 0;
		if (curr_tid < 0 || curr_tid != v.tid) { new_contig = new_block = 1; old_tid = curr_tid; curr_tid = v.tid; }      /* :111-124 */
		int insert = 1;
		if (!new_contig) {                                             /* :131-149 */
			if ((r->alignment_flag & F_PAIRED) && al.fwd > 0 && al.rev > 0) {
				if (al.fwd == al.rev) insert = h_find(hash, v.qname, v.l_qname) == NULL;
				else if (reverse) insert = al.fwd > al.rev;
				else insert = al.fwd < al.rev;
			}
			if (insert && start_pos > 0) {
				if (al.fwd > 0) {
					if (al.fwd > max_pos && (al.rev > max_pos || al.rev == 0)) { if (al.fwd - max_pos > 1) new_block = 1; }
				} else if (al.rev > max_pos && al.rev - max_pos > 1) new_block = 1;
			}
		}
		if (new_block) {                                               /* :151-207 */
			h_clear(hash);
			read_idx = 0; start_idx = 0; curr_pos = 0;
			publish(&st, o, (uint32_t)(new_contig ? old_tid : curr_tid), max_pos);
			max_pos = start_pos = 0;
		}
		{                                                              /* :209-220 */
			const uint32_t s0 = reverse ? al.rev : al.fwd, ml = s0 + r->reference_span;
			if (ml > max_pos) max_pos = ml;
			if (start_pos == 0 || start_pos > s0) start_pos = s0;
		}
		if (r->alignment_flag & F_PAIRED) {
			if (!insert) {                                             /* the mate should be waiting (:224-274) */
				hent *h = h_find(hash, v.qname, v.l_qname);
				if (h) {
					tmpl_s *t = st.list + h->ix;
					if (al.fwd != t->fwd || al.rev != t->rev) { o->err = -5; break; }      /* assert() in the reference (:239) */
					t->rec[ix] = (int64_t)ri; t->mapq[ix] = r->mapq; t->span[ix] = r->reference_span;
					st.list_h[h->ix] = NULL;
					h->live = 0;
				} else {
					bso_profile_tally(14, 1, 14, r->read_len);           /* :243-246 */
					int skip = 0;
					if (!keep_duplicates) { const uint32_t xx = reverse ? al.rev : al.fwd; if (xx >= start_pos) skip = 1; }
					if (!skip && keep_unmatched) {
						const uint32_t xx = (al.fwd > 0 ? al.fwd : al.rev) + r->align_length;
						if (xx > max_pos) max_pos = xx;
						list_put(&st, read_idx++, &al, NULL);
					}
				}
			} else {                                                   /* forward-facing mate: duplicate check, then store (:275-341) */
				int skip = 0;
				const uint8_t *nm = v.qname;                           /* name travelling with `al` through replacements */
				uint32_t nml = v.l_qname;
				if (!keep_duplicates) {
					const uint32_t pos = al.fwd > 0 ? al.fwd : al.rev;
					if (pos == curr_pos) {
						for (uint32_t i = start_idx; i < read_idx; i++) {
							tmpl_s *a1 = st.list + i;
							if (al.fwd != a1->fwd || al.rev != a1->rev || al.bs_strand != a1->bs_strand) continue;
							int maxq = 0, maxq1 = 0, kn = 0, kn1 = 0;
							for (int k = 0; k < 2; k++) {
								if (al.rec[k] >= 0 && rec[al.rec[k]].read_len > 0) { maxq += al.mapq[k]; kn++; }
								if (a1->rec[k] >= 0 && rec[a1->rec[k]].read_len > 0) { maxq1 += a1->mapq[k]; kn1++; }
							}
							maxq /= kn; maxq1 /= kn1;
							if (maxq1 < maxq || (maxq == maxq1 && al_qual(&st, a1) < al_qual(&st, &al))) {
								/* the newcomer takes the slot (and the hash entry of the slot); the old one is carried on */
								hent *h = h_find(hash, nm, nml);
								if (h && st.list_h[i]) { o->err = -4; break; }      /* duplicate read name: fatal in the reference */
								const int from_slot = !h && st.list_h[i];
								if (!h) h = st.list_h[i];
								const tmpl_s old = *a1;
								*a1 = al;
								if (h) h->live = 0;                                 /* the entry object is re-keyed to the newcomer's name */
								hent *nh = h_add(hash, nm, nml, r->alignment_flag, i);
								if (from_slot) st.list_h[i] = nh;                   /* same object in the reference; otherwise the slot keeps NULL */
								al = old;
								/* `al` is now the displaced template; its name is not needed again: a second match would
								 * look it up under the NEWCOMER's tag, exactly as the reference does (tag is not swapped) */
							}
							if (bso_profile_is_on()) {                              /* the template that lost (:314-319) */
								const uint32_t len1 = al.rec[0] >= 0 ? rec[al.rec[0]].read_len : 0, len2 = al.rec[1] >= 0 ? rec[al.rec[1]].read_len : 0;
								bso_profile_tally(5, len1 && len2 ? 2 : 1, 5, (uint64_t)len1 + len2);
							}
							skip = 1;
						}
					} else { curr_pos = pos; start_idx = read_idx; }
				}
				if (!skip && !o->err) {
					if (h_find(hash, nm, nml)) { o->err = -4; break; }
					hent *h = h_add(hash, nm, nml, r->alignment_flag, read_idx);
					list_put(&st, read_idx, &al, h);
					read_idx++;
				}
			}
		} else {                                                       /* single reads (:342-383) */
			int skip = 0;
			if (!keep_duplicates) {
				const uint32_t pos = al.fwd > 0 ? al.fwd : al.rev;
				if (pos == curr_pos) {
					for (uint32_t i = start_idx; i < read_idx; i++) {
						tmpl_s *a1 = st.list + i;
						const hent *h = st.list_h[i];
						if (al.fwd == a1->fwd && al.rev == a1->rev && al.bs_strand == a1->bs_strand &&
								(h == NULL || (h->flag & 9) == 9 || (h->flag & 9) == 0)) {
							/* mapq[0] on both sides, whichever strand the reads are on (reference quirk) */
							if (a1->mapq[0] < al.mapq[0] || (a1->mapq[0] == al.mapq[0] && al_qual(&st, a1) < al_qual(&st, &al))) {
								const tmpl_s old = *a1; *a1 = al; al = old;
							}
							/* counted as a duplicate, its bases under gt_flt_none (reference behaviour, :361-364) */
							bso_profile_tally(5, 1, 0, al.rec[ix] >= 0 ? rec[al.rec[ix]].read_len : 0);
							skip = 1;
						}
					}
				} else { curr_pos = pos; start_idx = read_idx; }
			}
			if (!skip) list_put(&st, read_idx++, &al, NULL);
		}
	}
	if (!o->err && curr_tid >= 0) publish(&st, o, (uint32_t)curr_tid, max_pos);      /* handle_end_of_block (:18-45) */
	h_clear(hash);
	free(hash); free(st.list); free(st.list_h);
	return o->err;
}

static bso_params chain_params;
static int chain_params_set = 0;
void bso_set_params(const bso_params *p) { chain_params = *p; chain_params_set = 1; }

int bso_read_input(const uint8_t *bam, size_t nbytes, int n_targets, const uint32_t *target_len, const uint8_t *const *ctg_codes,
		int mapq_thresh, uint32_t max_template_len, int keep_unmatched, int ignore_duplicates, int keep_duplicates, int run_chain,
		bso_block *blocks, size_t block_cap, size_t *nblocks, bso_template *tmpl, size_t tmpl_cap, size_t *ntmpl,
		uint8_t *bases, size_t bases_cap, size_t *nbases, bso_misms *misms, size_t misms_cap, size_t *nmisms,
		bso_gt_vcf *vcf, size_t vcf_cap, size_t *nvcf) {
	/* decode every record first; templates index into these arrays */
	size_t cap = nbytes / 36 + 8, nrec = 0, nb = 0, nm = 0;
	bso_record *rec = calloc(cap, sizeof(bso_record));
	uint8_t *rb = malloc(nbytes + 64);
	bso_misms *rm = malloc((nbytes / 4 + 64) * sizeof(bso_misms));
	int rc = bso_decode_records(bam, nbytes, mapq_thresh, max_template_len, keep_unmatched, ignore_duplicates, rec, cap, &nrec,
			rb, nbytes + 64, &nb, rm, nbytes / 4 + 64, &nm);
	rd_out o;
	memset(&o, 0, sizeof(o));
	o.blocks = blocks; o.block_cap = block_cap; o.tmpl = tmpl; o.tmpl_cap = tmpl_cap;
	if (!rc) rc = build_blocks(bam, nbytes, rec, nrec, rb, keep_unmatched, keep_duplicates, &o);
	/* compact: copy the bytes / events of the templates that survived, in template order (what a snapshot of the
	 * reference's align_list looks like) */
	size_t ob = 0, om = 0, ov = 0;
	for (size_t i = 0; !rc && i < o.ntmpl; i++) for (int k = 0; k < 2; k++) {
		bso_template *t = tmpl + i;
		if (!t->present[k]) { t->read_off[k] = (uint32_t)ob; t->mm_off[k] = (uint32_t)om; continue; }
		if (ob + t->read_len[k] > bases_cap || om + t->mm_n[k] > misms_cap) { rc = -3; break; }
		memcpy(bases + ob, rb + t->read_off[k], t->read_len[k]);
		memcpy(misms + om, rm + t->mm_off[k], t->mm_n[k] * sizeof(bso_misms));
		t->read_off[k] = (uint32_t)ob; t->mm_off[k] = (uint32_t)om;
		ob += t->read_len[k]; om += t->mm_n[k];
	}
	if (!rc && run_chain) {
		bso_params p;
		if (chain_params_set) p = chain_params; else bso_default_params(&p);
		for (size_t b = 0; b < o.nblocks; b++) {
			bso_block *bk = blocks + b;
			const uint32_t sz = bk->y - bk->x + 1;
			bk->vcf_off = ov;
			if (ov + sz > vcf_cap || (int)bk->tid >= n_targets) { rc = -3; break; }
			/* reference window [x, y]: codes of the contig, N beyond its end (src/get_sequence.c:20-55) */
			uint8_t *rc_ = calloc((size_t)sz + 4, 1);
			for (uint32_t i = 0; i <= sz; i++) { const uint32_t pos = bk->x + i; rc_[i] = pos < target_len[bk->tid] ? ctg_codes[bk->tid][pos - 1] : 0; }      /* one past y: the profile looks ahead */
			uint32_t xo = 0;
			bso_pileup *pile = malloc(sizeof(bso_pileup) * sz);
			const int e = bso_process_block(tmpl + bk->first_template, bk->n_templates, bases, misms, rc_, bk->y, &p, &xo, pile, vcf + ov);
			free(pile); free(rc_);
			if (e) { rc = -2; break; }
			ov += sz;
		}
	}
	free(rec); free(rb); free(rm);
	*nblocks = o.nblocks; *ntmpl = o.ntmpl; *nbases = ob; *nmisms = om; *nvcf = ov;
	return rc;
}
