/*
 * minihts.c -- a small stand-in for the part of htslib that bs_call uses.  TEST / MEASUREMENT INFRASTRUCTURE ONLY.
 *
 * htslib is an external system dependency of bs_call (configure.ac:9-19, README.md: "htslib 1.10 / 1.11"); it is not
 * vendored under /root/reference and it is absent from this image and from the GPU boxes.  SURVEY.md section 8(d) asks for
 * the unmodified reference binary as the whole-program CPU baseline, so the API surface the reference calls is restated
 * here from the published formats, under htslib's public names and signatures:
 *
 *   BGZF          RFC 1952 gzip members with the 'BC' extra field, 64 KiB blocks, empty EOF block (SAM spec 4.1); zlib does
 *                 the inflate / deflate.  Reading also accepts a plain (uncompressed) file.
 *   BAM           magic, header text, reference table, alignment records (SAM spec 4.2); sam_read1() hands out records in
 *                 htslib's bam1_t shape (fixed fields in core, the variable part as it is on disk in data[]).
 *   FASTA + .fai  fai_load() reads <file>.fai (name, length, offset, line bases, line width) or builds the same table by
 *                 scanning the file; the index object has the layout src/read_reference.c:18-33 re-declares for itself.
 *   BCF2 / VCF    header object with the ID / contig / sample dictionaries src/print_vcf.c:745-765 reads; typed-value
 *                 encoders and the record layout of the VCF/BCF specification v4.3 section 6.3; bcf_write() writes BCF
 *                 records as they are, or formats the same record as a VCF text line (-O v / z).
 *
 * Not provided: SAM text and CRAM input, .bai files (regions are served by scanning, see sam_itr_queryi), multi-threaded (de)compression (hts_set_threads is accepted and ignored).  Number
 * formatting of VCF text (QUAL, GL) is C's "%g", which is not guaranteed to be htslib's digit for digit; BCF output is
 * byte-exact by construction (the record bytes are the encoders' output).
 *
 * Used by oracle/Makefile for _ref/bs_call (every reference source unmodified) and _ref/bs_call_gpu (the same with the
 * product's seam files in place of get_template_vector.c / process_template.c / call_genotypes.c).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <stdint.h>
#include <unistd.h>
#include <zlib.h>

#include <htslib/kstring.h>
#include <htslib/khash.h>
#include <htslib/hts.h>
#include <htslib/sam.h>
#include <htslib/vcf.h>
#include <htslib/faidx.h>

/* ------------------------------------------------------------------------------------------------------------------
 * BGZF
 * ------------------------------------------------------------------------------------------------------------------ */
#define BLOCK_MAX 65536
#define WRITE_CHUNK 0xff00           /* uncompressed bytes per written block */

struct hFILE { int fd; };
hFILE *hdopen(int fd, const char *mode) { (void)mode; hFILE *h = calloc(1, sizeof(hFILE)); if (h) h->fd = fd; return h; }

struct BGZF {
	FILE *f;
	int is_write, compressed, level, eof;
	uint8_t *ubuf;                   /* uncompressed block (read: current block; write: pending bytes) */
	int ulen, upos;
	uint8_t *cbuf;
	uint64_t upos_total;             /* uncompressed offset of ubuf[0] (plain files) */
	off_t block_off;                 /* file offset at which the current block (ubuf) begins: bgzf_rewind_to() */
};

static BGZF *bgzf_from_file(FILE *f, int is_write, int level) {
	BGZF *b = calloc(1, sizeof(BGZF));
	if (!b) return NULL;
	b->f = f; b->is_write = is_write; b->level = level;
	b->ubuf = malloc(BLOCK_MAX); b->cbuf = malloc(BLOCK_MAX + 1024);
	setvbuf(f, NULL, _IOFBF, 1 << 20);
	if (!is_write) {
		const int c0 = getc(f), c1 = getc(f);
		b->compressed = c0 == 0x1f && c1 == 0x8b;
		if (c1 != EOF) ungetc(c1, f);
		if (c0 != EOF) ungetc(c0, f);
	} else b->compressed = level >= 0;
	return b;
}

/* read: next block into ubuf; 0 = ok, 1 = end of file, -1 = error */
static int bgzf_fill(BGZF *b) {
	b->upos_total += (uint64_t)b->ulen;
	b->upos = b->ulen = 0;
	b->block_off = ftello(b->f);
	if (!b->compressed) {
		const size_t n = fread(b->ubuf, 1, BLOCK_MAX, b->f);
		if (n == 0) { b->eof = 1; return ferror(b->f) ? -1 : 1; }
		b->ulen = (int)n;
		return 0;
	}
	for (;;) {
		uint8_t h[12];
		const size_t n = fread(h, 1, 12, b->f);
		if (n == 0) { b->eof = 1; return 1; }
		if (n != 12 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return -1;
		const int xlen = h[10] | h[11] << 8;
		uint8_t extra[256];
		if (xlen > 256 || fread(extra, 1, (size_t)xlen, b->f) != (size_t)xlen) return -1;
		int bsize = -1;
		for (int i = 0; i + 4 <= xlen;) {
			const int sl = extra[i + 2] | extra[i + 3] << 8;
			if (extra[i] == 'B' && extra[i + 1] == 'C' && sl == 2) bsize = extra[i + 4] | extra[i + 5] << 8;
			i += 4 + sl;
		}
		if (bsize < 0) return -1;
		const int clen = bsize + 1 - 12 - xlen - 8;
		if (clen < 0 || clen > BLOCK_MAX + 1024) return -1;
		uint8_t tail[8];
		if (fread(b->cbuf, 1, (size_t)clen, b->f) != (size_t)clen || fread(tail, 1, 8, b->f) != 8) return -1;
		const uint32_t isize = (uint32_t)tail[4] | (uint32_t)tail[5] << 8 | (uint32_t)tail[6] << 16 | (uint32_t)tail[7] << 24;
		if (isize > BLOCK_MAX) return -1;
		if (isize == 0) { b->block_off = ftello(b->f); continue; }          /* empty block (the EOF marker, or padding) */
		z_stream zs;
		memset(&zs, 0, sizeof zs);
		if (inflateInit2(&zs, -15) != Z_OK) return -1;
		zs.next_in = b->cbuf; zs.avail_in = (uInt)clen; zs.next_out = b->ubuf; zs.avail_out = BLOCK_MAX;
		const int r = inflate(&zs, Z_FINISH);
		inflateEnd(&zs);
		if (r != Z_STREAM_END || zs.total_out != isize) return -1;
		b->ulen = (int)isize;
		return 0;
	}
}

static long bgzf_read_bytes(BGZF *b, void *dst, size_t n) {
	uint8_t *d = dst;
	size_t got = 0;
	while (got < n) {
		if (b->upos == b->ulen) { const int r = bgzf_fill(b); if (r < 0) return -1; if (r > 0) break; }
		size_t k = (size_t)(b->ulen - b->upos);
		if (k > n - got) k = n - got;
		memcpy(d + got, b->ubuf + b->upos, k);
		b->upos += (int)k; got += k;
	}
	return (long)got;
}

ssize_t bgzf_read(BGZF *b, void *data, size_t length) { return (ssize_t)bgzf_read_bytes(b, data, length); }

int bgzf_getc(BGZF *b) {
	if (b->upos == b->ulen) { if (bgzf_fill(b) != 0) return -1; }
	return b->ubuf[b->upos++];
}

int bgzf_useek(BGZF *b, off_t uoffset, int where) {
	if (b->is_write || b->compressed || where != SEEK_SET) return -1;      /* FASTA files are read uncompressed here */
	if (fseeko(b->f, uoffset, SEEK_SET)) return -1;
	b->upos_total = (uint64_t)uoffset; b->ulen = b->upos = 0; b->eof = 0;
	return 0;
}

static int bgzf_flush_block(BGZF *b) {
	if (!b->compressed) {
		if (b->ulen && fwrite(b->ubuf, 1, (size_t)b->ulen, b->f) != (size_t)b->ulen) return -1;
		b->ulen = 0;
		return 0;
	}
	z_stream zs;
	memset(&zs, 0, sizeof zs);
	if (deflateInit2(&zs, b->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
	zs.next_in = b->ubuf; zs.avail_in = (uInt)b->ulen; zs.next_out = b->cbuf + 18; zs.avail_out = BLOCK_MAX + 1024 - 26;
	const int r = deflate(&zs, Z_FINISH);
	const uint32_t clen = (uint32_t)zs.total_out;
	deflateEnd(&zs);
	if (r != Z_STREAM_END) return -1;
	static const uint8_t head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
	memcpy(b->cbuf, head, 16);
	const uint32_t bsize = clen + 25;
	b->cbuf[16] = (uint8_t)bsize; b->cbuf[17] = (uint8_t)(bsize >> 8);
	const uint32_t crc = (uint32_t)crc32(crc32(0L, NULL, 0), b->ubuf, (uInt)b->ulen), isz = (uint32_t)b->ulen;
	uint8_t *t = b->cbuf + 18 + clen;
	for (int i = 0; i < 4; i++) { t[i] = (uint8_t)(crc >> (8 * i)); t[4 + i] = (uint8_t)(isz >> (8 * i)); }
	if (fwrite(b->cbuf, 1, clen + 26, b->f) != clen + 26) return -1;
	b->ulen = 0;
	return 0;
}

static int bgzf_write_bytes(BGZF *b, const void *src, size_t n) {
	const uint8_t *s = src;
	while (n) {
		size_t k = (size_t)(WRITE_CHUNK - b->ulen);
		if (k > n) k = n;
		memcpy(b->ubuf + b->ulen, s, k);
		b->ulen += (int)k; s += k; n -= k;
		if (b->ulen == WRITE_CHUNK && bgzf_flush_block(b)) return -1;
	}
	return 0;
}

static int bgzf_close_file(BGZF *b) {
	int r = 0;
	if (b->is_write) {
		if (b->ulen) r |= bgzf_flush_block(b);
		if (b->compressed) {                               /* the canonical 28-byte empty block that marks the end of a BGZF file */
			static const uint8_t eof_block[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
			if (fwrite(eof_block, 1, 28, b->f) != 28) r = -1;
		}
		r |= fflush(b->f);
	}
	if (b->f != stdout && b->f != stdin) fclose(b->f);
	free(b->ubuf); free(b->cbuf); free(b);
	return r;
}

/* ------------------------------------------------------------------------------------------------------------------
 * htsFile
 * ------------------------------------------------------------------------------------------------------------------ */
static void scan_state_free(void *state);          /* the region scanner's state hangs off htsFile.state (below) */

static htsFile *hts_wrap(FILE *f, const char *fn, const char *mode) {
	htsFile *fp = calloc(1, sizeof(htsFile));
	if (!fp) return NULL;
	fp->fn = strdup(fn);
	if (mode[0] == 'w') {
		fp->is_write = 1;
		const int is_b = strchr(mode, 'b') != NULL, is_u = strchr(mode, 'u') != NULL, is_z = strchr(mode, 'z') != NULL;
		fp->is_bin = (unsigned)is_b;
		fp->format.category = variant_data;
		fp->format.format = is_b ? bcf : vcf;
		/* BCF is always BGZF-framed (level 0 for "wbu", like htslib); VCF text is plain unless 'z' */
		const int framed = is_b || is_z;
		fp->format.compression = framed ? bgzf : no_compression;
		fp->fp.bgzf = bgzf_from_file(f, 1, framed ? (is_b && is_u ? 0 : 6) : -1);
	} else {
		BGZF *b = bgzf_from_file(f, 0, 0);
		fp->fp.bgzf = b;
		fp->format.category = sequence_data;
		fp->format.compression = b->compressed ? bgzf : no_compression;
		/* look at the first four payload bytes: "BAM\1" */
		if (bgzf_fill(b) == 0 && b->ulen >= 4 && !memcmp(b->ubuf, "BAM\1", 4)) { fp->format.format = bam; fp->is_bin = 1; }
		else fp->format.format = unknown_format;
	}
	return fp;
}

htsFile *hts_open(const char *fn, const char *mode) {
	FILE *f;
	if (!strcmp(fn, "-")) f = mode[0] == 'w' ? stdout : stdin;
	else f = fopen(fn, mode[0] == 'w' ? "wb" : "rb");
	if (!f) return NULL;
	return hts_wrap(f, fn, mode);
}

htsFile *hts_hopen(hFILE *h, const char *fn, const char *mode) {
	FILE *f = fdopen(h->fd, mode[0] == 'w' ? "wb" : "rb");
	free(h);
	return f ? hts_wrap(f, fn, mode) : NULL;
}

int hts_close(htsFile *fp) {
	if (!fp) return 0;
	const int r = bgzf_close_file(fp->fp.bgzf);
	if (fp->state) scan_state_free(fp->state);
	free(fp->fn); free(fp->fn_aux); free(fp->line.s); free(fp);
	return r;
}

int hts_set_threads(htsFile *fp, int n) { (void)fp; (void)n; return 0; }
int hts_set_fai_filename(htsFile *fp, const char *fn_aux) { free(fp->fn_aux); fp->fn_aux = strdup(fn_aux); return 0; }
/* ---- regions without a .bai: a scanning stand-in for the index.  sam_index_load() hands out a handle for any seekable BAM
 * file; sam_itr_queryi() / sam_itr_next() then return, in file order, the records of contig tid that overlap [beg, end) --
 * htslib's contract -- by reading on from where the file stands when the region lies ahead (regions of a coordinate-sorted
 * file asked for in ascending order: one pass over the file), and from the first record again otherwise.  Records already read
 * that reach beyond the end of a region are kept, because they also overlap a region that begins there. ---- */
typedef struct {
	bam1_t *pend;                    /* a record read from the file that lies beyond the region that was being served */
	int has_pend;
	bam1_t **keep;                   /* records served or skipped that end beyond the last region's end */
	int nkeep, mkeep;
	int last_tid;                    /* the last region served: the next one may go forward iff it begins at or beyond its end */
	hts_pos_t last_end;
	off_t rec0_off;                  /* where the first alignment record lies: block start in the file, offset inside the block */
	int rec0_upos;
} scan_state;
struct hts_idx_t { htsFile *fp; };
struct hts_itr_t { htsFile *fp; int tid; hts_pos_t beg, end; bam1_t **serve; int nserve, iserve; };

static hts_pos_t rec_endpos(const bam1_t *b) {
	const uint32_t *cig = bam_get_cigar(b);
	hts_pos_t len = 0;
	for (uint32_t i = 0; i < b->core.n_cigar; i++) {
		const int op = bam_cigar_op(cig[i]);
		if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) len += bam_cigar_oplen(cig[i]);       /* M D N = X consume the reference */
	}
	return b->core.pos + (len ? len : 1);
}
static bam1_t *rec_dup(const bam1_t *b) {
	bam1_t *c = calloc(1, sizeof(bam1_t));
	*c = *b;
	c->data = malloc((size_t)b->l_data + 1);
	memcpy(c->data, b->data, (size_t)b->l_data);
	c->m_data = (uint32_t)b->l_data + 1;
	return c;
}
static void rec_copy(bam1_t *dst, const bam1_t *src) {
	if ((uint32_t)src->l_data > dst->m_data) { dst->data = realloc(dst->data, (size_t)src->l_data + 1); dst->m_data = (uint32_t)src->l_data + 1; }
	uint8_t *d = dst->data;
	const uint32_t m = dst->m_data;
	*dst = *src;
	dst->data = d; dst->m_data = m;
	memcpy(dst->data, src->data, (size_t)src->l_data);
}
static void keep_push(scan_state *st, bam1_t *r) {
	if (st->nkeep == st->mkeep) { st->mkeep = st->mkeep ? 2 * st->mkeep : 64; st->keep = realloc(st->keep, sizeof(bam1_t *) * (size_t)st->mkeep); }
	st->keep[st->nkeep++] = r;
}

static void scan_state_free(void *state) {
	scan_state *st = state;
	for (int i = 0; i < st->nkeep; i++) bam_destroy1(st->keep[i]);
	free(st->keep);
	if (st->pend) bam_destroy1(st->pend);
	free(st);
}

hts_idx_t *sam_index_load(htsFile *fp, const char *fn) {
	(void)fn;
	if (!fp || fp->format.format != bam || fp->fp.bgzf->f == stdin) return NULL;      /* (the scanner's state is made by sam_hdr_read) */
	hts_idx_t *idx = calloc(1, sizeof(hts_idx_t));
	idx->fp = fp;
	return idx;
}
void hts_idx_destroy(hts_idx_t *idx) { free(idx); }

hts_itr_t *sam_itr_queryi(const hts_idx_t *idx, int tid, hts_pos_t beg, hts_pos_t end) {
	if (!idx || !idx->fp->state) return NULL;
	htsFile *fp = idx->fp;
	scan_state *st = fp->state;
	hts_itr_t *it = calloc(1, sizeof(hts_itr_t));
	it->fp = fp; it->tid = tid; it->beg = beg; it->end = end;
	const int forward = st->last_tid < 0 || tid > st->last_tid || (tid == st->last_tid && beg >= st->last_end);
	if (!forward) {
		/* from the first record again */
		BGZF *b = fp->fp.bgzf;
		for (int i = 0; i < st->nkeep; i++) bam_destroy1(st->keep[i]);
		st->nkeep = 0; st->has_pend = 0;
		fseeko(b->f, st->rec0_off, SEEK_SET);
		b->ulen = b->upos = 0; b->eof = 0;
		if (bgzf_fill(b) == 0) b->upos = st->rec0_upos;
	}
	/* kept records that overlap this region are served first; those that end beyond it stay for the next one */
	it->serve = calloc((size_t)st->nkeep + 1, sizeof(bam1_t *));
	int nk = 0;
	for (int i = 0; i < st->nkeep; i++) {
		bam1_t *r = st->keep[i];
		const int mine = r->core.tid == tid && r->core.pos < end && rec_endpos(r) > beg;
		const int later = r->core.tid > tid || (r->core.tid == tid && rec_endpos(r) > end);
		if (mine) it->serve[it->nserve++] = later ? rec_dup(r) : r;
		if (later) st->keep[nk++] = r;
		else if (!mine) bam_destroy1(r);
	}
	st->nkeep = nk;
	st->last_tid = tid; st->last_end = end;
	return it;
}

int sam_itr_next(htsFile *fp, hts_itr_t *it, bam1_t *b) {
	if (!it) return -1;
	scan_state *st = fp->state;
	if (it->iserve < it->nserve) {
		bam1_t *r = it->serve[it->iserve++];
		rec_copy(b, r);
		bam_destroy1(r);
		return b->l_data + 36;
	}
	for (;;) {
		int ret;
		if (st->has_pend) { rec_copy(b, st->pend); st->has_pend = 0; ret = b->l_data + 36; }
		else ret = sam_read1(fp, NULL, b);
		if (ret < 0) return ret;
		const bam1_core_t *c = &b->core;
		if (c->tid < 0 || c->tid > it->tid || (c->tid == it->tid && c->pos >= it->end)) {
			/* beyond the region (unmapped records sort last): keep it for the region that may follow */
			if (!st->pend) st->pend = bam_init1();
			rec_copy(st->pend, b);
			st->has_pend = 1;
			return -1;
		}
		if (c->tid < it->tid) continue;
		const hts_pos_t e = rec_endpos(b);
		if (e > it->end) keep_push(st, rec_dup(b));
		if (e <= it->beg) continue;
		return ret;
	}
}

void hts_itr_destroy(hts_itr_t *it) {
	if (!it) return;
	for (int i = it->iserve; i < it->nserve; i++) bam_destroy1(it->serve[i]);
	free(it->serve);
	free(it);
}

/* ------------------------------------------------------------------------------------------------------------------
 * BAM input
 * ------------------------------------------------------------------------------------------------------------------ */
static int rd_i32(BGZF *b, int32_t *v) { uint8_t x[4]; if (bgzf_read_bytes(b, x, 4) != 4) return -1; *v = (int32_t)((uint32_t)x[0] | (uint32_t)x[1] << 8 | (uint32_t)x[2] << 16 | (uint32_t)x[3] << 24); return 0; }

sam_hdr_t *sam_hdr_read(htsFile *fp) {
	if (fp->format.format != bam) { fprintf(stderr, "minihts: input is not a BAM file (SAM text and CRAM are not supported by this stand-in)\n"); return NULL; }
	BGZF *b = fp->fp.bgzf;
	uint8_t magic[4];
	int32_t l_text, n_ref;
	if (bgzf_read_bytes(b, magic, 4) != 4 || memcmp(magic, "BAM\1", 4) || rd_i32(b, &l_text) || l_text < 0) return NULL;
	sam_hdr_t *h = calloc(1, sizeof(sam_hdr_t));
	h->text = calloc((size_t)l_text + 1, 1);
	h->l_text = (size_t)l_text;
	if (bgzf_read_bytes(b, h->text, (size_t)l_text) != l_text || rd_i32(b, &n_ref) || n_ref < 0) return NULL;
	h->l_text = strlen(h->text);
	h->n_targets = n_ref;
	h->target_name = calloc((size_t)n_ref + 1, sizeof(char *));
	h->target_len = calloc((size_t)n_ref + 1, sizeof(uint32_t));
	for (int i = 0; i < n_ref; i++) {
		int32_t l_name, l_ref;
		if (rd_i32(b, &l_name) || l_name <= 0) return NULL;
		h->target_name[i] = calloc((size_t)l_name + 1, 1);
		if (bgzf_read_bytes(b, h->target_name[i], (size_t)l_name) != l_name || rd_i32(b, &l_ref)) return NULL;
		h->target_len[i] = (uint32_t)l_ref;
	}
	{   /* where the first alignment record lies (regions without an index read from here) */
		scan_state *st = calloc(1, sizeof(scan_state));
		if (b->upos == b->ulen && bgzf_fill(b) < 0) { free(st); return NULL; }
		st->rec0_off = b->block_off; st->rec0_upos = b->upos; st->last_tid = -1;
		fp->state = st;
	}
	return h;
}

void sam_hdr_destroy(sam_hdr_t *h) {
	if (!h) return;
	for (int i = 0; i < h->n_targets; i++) free(h->target_name[i]);
	free(h->target_name); free(h->target_len); free(h->text); free(h);
}

int bam_name2id(sam_hdr_t *h, const char *ref) {
	for (int i = 0; i < h->n_targets; i++) if (!strcmp(h->target_name[i], ref)) return i;
	return -1;
}

bam1_t *bam_init1(void) { return calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t *b) { if (b) { free(b->data); free(b); } }

/* >= 0 record read, -1 end of file, < -1 error (htslib's convention; src/get_template_vector.c:88-91) */
int sam_read1(htsFile *fp, sam_hdr_t *h, bam1_t *b) {
	(void)h;
	BGZF *z = fp->fp.bgzf;
	uint8_t x[36];
	const long n = bgzf_read_bytes(z, x, 4);
	if (n == 0) return -1;
	if (n != 4) return -2;
	const uint32_t block_size = (uint32_t)x[0] | (uint32_t)x[1] << 8 | (uint32_t)x[2] << 16 | (uint32_t)x[3] << 24;
	if (block_size < 32 || bgzf_read_bytes(z, x + 4, 32) != 32) return -3;
#define U32(o) ((uint32_t)x[o] | (uint32_t)x[(o) + 1] << 8 | (uint32_t)x[(o) + 2] << 16 | (uint32_t)x[(o) + 3] << 24)
	bam1_core_t *c = &b->core;
	c->tid = (int32_t)U32(4);
	c->pos = (int32_t)U32(8);
	c->l_qname = x[12];
	c->qual = x[13];
	c->bin = (uint16_t)(x[14] | x[15] << 8);
	c->n_cigar = (uint32_t)(x[16] | x[17] << 8);
	c->flag = (uint16_t)(x[18] | x[19] << 8);
	c->l_qseq = (int32_t)U32(20);
	c->mtid = (int32_t)U32(24);
	c->mpos = (int32_t)U32(28);
	c->isize = (int32_t)U32(32);
	c->l_extranul = 0;
#undef U32
	const uint32_t l_data = block_size - 32;
	if (l_data > b->m_data) {
		uint32_t m = b->m_data ? b->m_data : 256;
		while (m < l_data) m <<= 1;
		uint8_t *p = realloc(b->data, m);
		if (!p) return -4;
		b->data = p; b->m_data = m;
	}
	if (bgzf_read_bytes(z, b->data, l_data) != (long)l_data) return -3;
	b->l_data = (int)l_data;
	if ((size_t)c->l_qname + 4ul * c->n_cigar + (size_t)((c->l_qseq + 1) >> 1) + (size_t)c->l_qseq > l_data) return -5;
	return (int)block_size + 4;
}

/* ------------------------------------------------------------------------------------------------------------------
 * FASTA index
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
	int id;
	uint32_t line_len, line_blen;
	uint64_t len;
	uint64_t seq_offset;
	uint64_t qual_offset;
} faidx1_t;
KHASH_MAP_INIT_STR(s, faidx1_t)

struct __faidx_t {
	BGZF *bgzf;
	int n, m;
	char **name;
	khash_t(s) *hash;
	enum fai_format_options format;
};

static void fai_insert(faidx_t *fai, const char *name, faidx1_t v) {
	if (fai->n == fai->m) { fai->m = fai->m ? fai->m * 2 : 16; fai->name = realloc(fai->name, sizeof(char *) * (size_t)fai->m); }
	char *nm = strdup(name);
	int ret;
	const khint_t k = kh_put(s, fai->hash, nm, &ret);
	if (!ret) { free(nm); return; }                 /* duplicate name: first one wins */
	v.id = fai->n;
	kh_val(fai->hash, k) = v;
	fai->name[fai->n++] = nm;
}

static int fai_scan(faidx_t *fai, FILE *f) {       /* what `samtools faidx` writes, computed in memory */
	char *line = NULL;
	size_t cap = 0;
	ssize_t l;
	uint64_t off = 0;
	char name[1024];
	faidx1_t cur;
	int have = 0;
	memset(&cur, 0, sizeof cur);
	while ((l = getline(&line, &cap, f)) >= 0) {
		if (line[0] == '>') {
			if (have) fai_insert(fai, name, cur);
			size_t k = 0;
			while (1 + k < (size_t)l && k < sizeof name - 1 && line[1 + k] > ' ') { name[k] = line[1 + k]; k++; }
			name[k] = 0;
			memset(&cur, 0, sizeof cur);
			cur.seq_offset = off + (uint64_t)l;
			have = 1;
		} else if (have) {
			size_t bl = (size_t)l;
			while (bl && (line[bl - 1] == '\n' || line[bl - 1] == '\r')) bl--;
			if (bl && !cur.line_len) { cur.line_len = (uint32_t)l; cur.line_blen = (uint32_t)bl; }
			cur.len += bl;
		}
		off += (uint64_t)l;
	}
	if (have) fai_insert(fai, name, cur);
	free(line);
	return 0;
}

faidx_t *fai_load(const char *fn) {
	FILE *f = fopen(fn, "rb");
	if (!f) return NULL;
	faidx_t *fai = calloc(1, sizeof(faidx_t));
	fai->hash = kh_init(s);
	fai->format = FAI_FASTA;
	char *idxname = malloc(strlen(fn) + 5);
	sprintf(idxname, "%s.fai", fn);
	FILE *fi = fopen(idxname, "r");
	free(idxname);
	if (fi) {
		char name[1024];
		unsigned long long len, offs;
		unsigned bl, ll;
		while (fscanf(fi, "%1023s %llu %llu %u %u%*[^\n]", name, &len, &offs, &bl, &ll) == 5 ||
				0) {
			faidx1_t v;
			memset(&v, 0, sizeof v);
			v.len = len; v.seq_offset = offs; v.line_blen = bl; v.line_len = ll;
			fai_insert(fai, name, v);
		}
		fclose(fi);
	} else {
		fai_scan(fai, f);
		rewind(f);
	}
	fai->bgzf = bgzf_from_file(f, 0, 0);
	if (fai->bgzf->compressed) { fprintf(stderr, "minihts: compressed reference FASTA is not supported by this stand-in\n"); return NULL; }
	return fai;
}

void fai_destroy(faidx_t *fai) {
	if (!fai) return;
	for (int i = 0; i < fai->n; i++) free(fai->name[i]);
	free(fai->name);
	kh_destroy(s, fai->hash);
	if (fai->bgzf) bgzf_close_file(fai->bgzf);
	free(fai);
}

int faidx_nseq(const faidx_t *fai) { return fai->n; }
const char *faidx_iseq(const faidx_t *fai, int i) { return fai->name[i]; }
int faidx_seq_len(const faidx_t *fai, const char *seq) {
	const khint_t k = kh_get(s, fai->hash, seq);
	return k == kh_end(fai->hash) ? -1 : (int)kh_val(fai->hash, k).len;
}

/* ------------------------------------------------------------------------------------------------------------------
 * BCF2 typed values and records (VCF/BCF specification v4.3, section 6.3)
 * ------------------------------------------------------------------------------------------------------------------ */
void bcf_enc_size(kstring_t *s, int size, int type) {
	if (size >= 15) {
		kputc(15 << 4 | type, s);
		if (size >= 128) {
			if (size >= 32768) { int32_t x = size; kputc(1 << 4 | BCF_BT_INT32, s); kputsn((char *)&x, 4, s); }
			else { int16_t x = (int16_t)size; kputc(1 << 4 | BCF_BT_INT16, s); kputsn((char *)&x, 2, s); }
		} else { kputc(1 << 4 | BCF_BT_INT8, s); kputc(size, s); }
	} else kputc(size << 4 | type, s);
}
void bcf_enc_int1(kstring_t *s, int32_t x) {
	if (x == bcf_int32_vector_end) { bcf_enc_size(s, 1, BCF_BT_INT8); kputc(bcf_int8_vector_end, s); }
	else if (x == bcf_int32_missing) { bcf_enc_size(s, 1, BCF_BT_INT8); kputc(bcf_int8_missing, s); }
	else if (x <= BCF_MAX_BT_INT8 && x >= BCF_MIN_BT_INT8) { bcf_enc_size(s, 1, BCF_BT_INT8); kputc(x, s); }
	else if (x <= BCF_MAX_BT_INT16 && x >= BCF_MIN_BT_INT16) { int16_t z = (int16_t)x; bcf_enc_size(s, 1, BCF_BT_INT16); kputsn((char *)&z, 2, s); }
	else { int32_t z = x; bcf_enc_size(s, 1, BCF_BT_INT32); kputsn((char *)&z, 4, s); }
}
void bcf_enc_vint(kstring_t *s, int n, int32_t *a, int wsize) {
	if (n <= 0) { bcf_enc_size(s, 0, BCF_BT_NULL); return; }
	if (n == 1) { bcf_enc_int1(s, a[0]); return; }
	int32_t max = INT32_MIN + 1, min = INT32_MAX;
	if (wsize <= 0) wsize = n;
	for (int i = 0; i < n; i++) {
		if (a[i] == bcf_int32_missing || a[i] == bcf_int32_vector_end) continue;
		if (max < a[i]) max = a[i];
		if (min > a[i]) min = a[i];
	}
	if (max <= BCF_MAX_BT_INT8 && min >= BCF_MIN_BT_INT8) {
		bcf_enc_size(s, wsize, BCF_BT_INT8);
		for (int i = 0; i < n; i++) kputc(a[i] == bcf_int32_vector_end ? bcf_int8_vector_end : a[i] == bcf_int32_missing ? bcf_int8_missing : a[i], s);
	} else if (max <= BCF_MAX_BT_INT16 && min >= BCF_MIN_BT_INT16) {
		bcf_enc_size(s, wsize, BCF_BT_INT16);
		for (int i = 0; i < n; i++) {
			int16_t z = a[i] == bcf_int32_vector_end ? bcf_int16_vector_end : a[i] == bcf_int32_missing ? bcf_int16_missing : (int16_t)a[i];
			kputsn((char *)&z, 2, s);
		}
	} else {
		bcf_enc_size(s, wsize, BCF_BT_INT32);
		for (int i = 0; i < n; i++) { int32_t z = a[i]; kputsn((char *)&z, 4, s); }
	}
}
void bcf_enc_vfloat(kstring_t *s, int n, float *a) {
	bcf_enc_size(s, n, BCF_BT_FLOAT);
	kputsn((char *)a, (size_t)n << 2, s);          /* little-endian host */
}
void bcf_enc_vchar(kstring_t *s, int l, const char *a) {
	bcf_enc_size(s, l, BCF_BT_CHAR);
	kputsn(a, (size_t)l, s);
}
bcf1_t *bcf_init(void) { bcf1_t *v = calloc(1, sizeof(bcf1_t)); if (v) bcf_clear(v); return v; }
void bcf_destroy(bcf1_t *v) { if (v) { free(v->shared.s); free(v->indiv.s); free(v); } }
void bcf_clear(bcf1_t *v) {
	v->rid = 0; v->pos = 0; v->rlen = 0;
	{ uint32_t miss = 0x7F800001u; memcpy(&v->qual, &miss, 4); }          /* bcf_float_missing */
	v->n_info = v->n_allele = v->n_fmt = v->n_sample = 0;
	v->shared.l = v->indiv.l = 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * BCF header: the lines in order + the three dictionaries
 * ------------------------------------------------------------------------------------------------------------------ */
KHASH_MAP_INIT_STR(vdict, bcf_idinfo_t)
typedef khash_t(vdict) vdict_t;

typedef struct { char **line; int n, m; } hdr_lines;

static void dict_add(bcf_hdr_t *h, int which, const char *key) {
	vdict_t *d = h->dict[which];
	if (kh_get(vdict, d, key) != kh_end(d)) return;
	char *k = strdup(key);
	int ret;
	const khint_t it = kh_put(vdict, d, k, &ret);
	bcf_idinfo_t *v = &kh_val(d, it);
	memset(v, 0, sizeof *v);
	v->id = h->n[which];
	h->id[which] = realloc(h->id[which], sizeof(bcf_idpair_t) * (size_t)(h->n[which] + 1));
	h->id[which][h->n[which]].key = k;
	h->id[which][h->n[which]].val = NULL;          /* the map may be rehashed: look values up through the dictionary */
	h->n[which]++;
}

static void lines_push(hdr_lines *L, const char *line, size_t len) {
	if (L->n == L->m) { L->m = L->m ? L->m * 2 : 32; L->line = realloc(L->line, sizeof(char *) * (size_t)L->m); }
	L->line[L->n] = malloc(len + 1);
	memcpy(L->line[L->n], line, len);
	L->line[L->n][len] = 0;
	L->n++;
}

/* value of ID= inside a structured line "##KEY=<ID=xxx,...": copied into out */
static int line_id(const char *line, char *out, size_t cap) {
	const char *p = strstr(line, "<ID=");
	if (!p) return -1;
	p += 4;
	size_t k = 0;
	while (p[k] && p[k] != ',' && p[k] != '>' && k + 1 < cap) { out[k] = p[k]; k++; }
	out[k] = 0;
	return k ? 0 : -1;
}

bcf_hdr_t *bcf_hdr_init(const char *mode) {
	(void)mode;
	bcf_hdr_t *h = calloc(1, sizeof(bcf_hdr_t));
	if (!h) return NULL;
	for (int i = 0; i < 3; i++) h->dict[i] = kh_init(vdict);
	h->priv = calloc(1, sizeof(hdr_lines));
	bcf_hdr_append(h, "##fileformat=VCFv4.2");
	bcf_hdr_append(h, "##FILTER=<ID=PASS,Description=\"All filters passed\">");
	return h;
}

const char *bcf_hdr_get_version(const bcf_hdr_t *hdr) {
	const hdr_lines *L = hdr->priv;
	for (int i = 0; i < L->n; i++) if (!strncmp(L->line[i], "##fileformat=", 13)) return L->line[i] + 13;
	return "VCFv4.2";
}

int bcf_hdr_append(bcf_hdr_t *h, const char *line) {
	hdr_lines *L = h->priv;
	size_t len = strlen(line);
	while (len && (line[len - 1] == '\n' || line[len - 1] == '\r')) len--;
	if (len < 2 || line[0] != '#' || line[1] != '#') return -1;
	if (!strncmp(line, "##fileformat=", 13)) {          /* one version line: a second one replaces the first */
		for (int i = 0; i < L->n; i++) if (!strncmp(L->line[i], "##fileformat=", 13)) {
			char *old = L->line[i];              /* `line` may point into the old one (bcf_hdr_get_version) */
			char *nl = malloc(len + 1);
			memcpy(nl, line, len); nl[len] = 0;
			L->line[i] = nl;
			free(old);
			return 0;
		}
	}
	lines_push(L, line, len);
	char id[256];
	const char *ln = L->line[L->n - 1];
	if (!strncmp(ln, "##contig=<", 10)) { if (!line_id(ln, id, sizeof id)) dict_add(h, BCF_DT_CTG, id); }
	else if (!strncmp(ln, "##FILTER=<", 10) || !strncmp(ln, "##INFO=<", 8) || !strncmp(ln, "##FORMAT=<", 10)) { if (!line_id(ln, id, sizeof id)) dict_add(h, BCF_DT_ID, id); }
	return 0;
}

int bcf_hdr_printf(bcf_hdr_t *h, const char *format, ...) {
	va_list ap;
	va_start(ap, format);
	char *s = NULL;
	const int n = vasprintf(&s, format, ap);
	va_end(ap);
	if (n < 0) return -1;
	const int r = bcf_hdr_append(h, s);
	free(s);
	return r;
}

int bcf_hdr_add_sample(bcf_hdr_t *h, const char *sample) {
	if (!sample) return 0;
	dict_add(h, BCF_DT_SAMPLE, sample);
	return 0;
}

void bcf_hdr_destroy(bcf_hdr_t *h) {
	if (!h) return;
	hdr_lines *L = h->priv;
	for (int i = 0; i < L->n; i++) free(L->line[i]);
	free(L->line); free(L);
	for (int w = 0; w < 3; w++) {
		for (int i = 0; i < h->n[w]; i++) free((void *)h->id[w][i].key);
		free(h->id[w]);
		kh_destroy(vdict, (vdict_t *)h->dict[w]);
	}
	free(h);
}

static void hdr_text(const bcf_hdr_t *h, kstring_t *out) {
	const hdr_lines *L = h->priv;
	for (int i = 0; i < L->n; i++) { kputs(L->line[i], out); kputc('\n', out); }
	kputs("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO", out);
	if (h->n[BCF_DT_SAMPLE]) {
		kputs("\tFORMAT", out);
		for (int i = 0; i < h->n[BCF_DT_SAMPLE]; i++) { kputc('\t', out); kputs(h->id[BCF_DT_SAMPLE][i].key, out); }
	}
	kputc('\n', out);
}

int bcf_hdr_write(htsFile *fp, bcf_hdr_t *h) {
	kstring_t t = {0, 0, NULL};
	hdr_text(h, &t);
	int r;
	if (fp->format.format == bcf) {
		const uint32_t l_text = (uint32_t)t.l + 1;
		r = bgzf_write_bytes(fp->fp.bgzf, "BCF\2\2", 5) | bgzf_write_bytes(fp->fp.bgzf, &l_text, 4) | bgzf_write_bytes(fp->fp.bgzf, t.s, t.l + 1);
	} else r = bgzf_write_bytes(fp->fp.bgzf, t.s, t.l);
	free(t.s);
	return r;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Record output: BCF bytes as they are, or the same record as a VCF text line
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct { const uint8_t *p, *end; } cur_t;

static int dec_size(cur_t *c, int *type) {
	if (c->p >= c->end) { *type = 0; return 0; }
	const uint8_t b = *c->p++;
	*type = b & 15;
	int n = b >> 4;
	if (n == 15) {
		const uint8_t tb = *c->p++;
		const int tt = tb & 15;
		if (tt == BCF_BT_INT8) { n = *(const int8_t *)c->p; c->p += 1; }
		else if (tt == BCF_BT_INT16) { int16_t v; memcpy(&v, c->p, 2); n = v; c->p += 2; }
		else { int32_t v; memcpy(&v, c->p, 4); n = v; c->p += 4; }
	}
	return n;
}
static int elt_size(int type) { return type == BCF_BT_INT8 || type == BCF_BT_CHAR ? 1 : type == BCF_BT_INT16 ? 2 : type == BCF_BT_NULL ? 0 : 4; }

/* integer element i of a typed vector; *state: 0 value, 1 missing, 2 end of vector */
static int32_t dec_int(const uint8_t *p, int type, int i, int *state) {
	int32_t v;
	if (type == BCF_BT_INT8) { v = ((const int8_t *)p)[i]; *state = v == bcf_int8_missing ? 1 : v == bcf_int8_vector_end ? 2 : 0; }
	else if (type == BCF_BT_INT16) { int16_t z; memcpy(&z, p + 2 * i, 2); v = z; *state = z == bcf_int16_missing ? 1 : z == bcf_int16_vector_end ? 2 : 0; }
	else { memcpy(&v, p + 4 * i, 4); *state = v == bcf_int32_missing ? 1 : v == bcf_int32_vector_end ? 2 : 0; }
	return v;
}

static void put_float(kstring_t *s, const uint8_t *p) {
	uint32_t u;
	float f;
	memcpy(&u, p, 4); memcpy(&f, p, 4);
	if (u == 0x7F800001u) { kputc('.', s); return; }
	char buf[48];
	snprintf(buf, sizeof buf, "%g", (double)f);
	kputs(buf, s);
}

/* one typed vector of n elements as comma-separated text */
static void put_vector(kstring_t *s, const uint8_t *p, int type, int n) {
	if (type == BCF_BT_CHAR) {
		int k = 0;
		while (k < n && p[k]) k++;
		if (k == 0) kputc('.', s); else kputsn((const char *)p, (size_t)k, s);
		return;
	}
	if (n == 0 || type == BCF_BT_NULL) { kputc('.', s); return; }
	int out = 0;
	for (int i = 0; i < n; i++) {
		if (type == BCF_BT_FLOAT) {
			uint32_t u;
			memcpy(&u, p + 4 * i, 4);
			if (u == 0x7F800002u) break;                 /* end of vector */
			if (out++) kputc(',', s);
			put_float(s, p + 4 * i);
		} else {
			int st;
			const int32_t v = dec_int(p, type, i, &st);
			if (st == 2) break;
			if (out++) kputc(',', s);
			if (st == 1) kputc('.', s);
			else { char buf[16]; snprintf(buf, sizeof buf, "%d", v); kputs(buf, s); }
		}
	}
	if (!out) kputc('.', s);
}

static const char *dict_key(const bcf_hdr_t *h, int which, int id) { return id >= 0 && id < h->n[which] ? h->id[which][id].key : "?"; }

static void vcf_line(const bcf_hdr_t *h, const bcf1_t *v, kstring_t *s) {
	char buf[64];
	kputs(dict_key(h, BCF_DT_CTG, v->rid), s);
	snprintf(buf, sizeof buf, "\t%lld\t", (long long)v->pos + 1);
	kputs(buf, s);
	cur_t c = {(const uint8_t *)v->shared.s, (const uint8_t *)v->shared.s + v->shared.l};
	int type, n;
	n = dec_size(&c, &type);                                   /* ID */
	put_vector(s, c.p, BCF_BT_CHAR, n); c.p += n;
	for (int a = 0; a < (int)v->n_allele; a++) {               /* REF, ALT */
		n = dec_size(&c, &type);
		if (a < 2) kputc('\t', s); else kputc(',', s);
		put_vector(s, c.p, BCF_BT_CHAR, n); c.p += n;
	}
	if (v->n_allele < 2) kputs("\t.", s);
	kputc('\t', s);
	put_float(s, (const uint8_t *)&v->qual);                   /* QUAL */
	kputc('\t', s);
	n = dec_size(&c, &type);                                   /* FILTER */
	if (n == 0) kputc('.', s);
	for (int i = 0; i < n; i++) { int st; const int32_t id = dec_int(c.p, type, i, &st); if (i) kputc(';', s); kputs(dict_key(h, BCF_DT_ID, id), s); }
	c.p += n * elt_size(type);
	kputc('\t', s);
	if (!v->n_info) kputc('.', s);                             /* INFO */
	for (int i = 0; i < (int)v->n_info; i++) {
		int st, kt;
		(void)dec_size(&c, &kt);
		const int32_t key = dec_int(c.p, kt, 0, &st);
		c.p += elt_size(kt);
		n = dec_size(&c, &type);
		if (i) kputc(';', s);
		kputs(dict_key(h, BCF_DT_ID, key), s);
		if (n > 0) { kputc('=', s); put_vector(s, c.p, type, n); }
		c.p += n * elt_size(type);
	}
	if (v->n_sample && v->n_fmt) {                             /* FORMAT + samples */
		const uint8_t *data[256];
		const char *keys[256];
		int types[256], sizes[256];
		cur_t d = {(const uint8_t *)v->indiv.s, (const uint8_t *)v->indiv.s + v->indiv.l};
		for (int i = 0; i < (int)v->n_fmt; i++) {
			int st, kt;
			(void)dec_size(&d, &kt);
			keys[i] = dict_key(h, BCF_DT_ID, dec_int(d.p, kt, 0, &st));
			d.p += elt_size(kt);
			sizes[i] = dec_size(&d, &types[i]);
			data[i] = d.p;
			d.p += (size_t)sizes[i] * (size_t)elt_size(types[i]) * v->n_sample;
			kputc(i ? ':' : '\t', s);
			kputs(keys[i], s);
		}
		for (uint32_t smp = 0; smp < v->n_sample; smp++) {
			for (int i = 0; i < (int)v->n_fmt; i++) {
				kputc(i ? ':' : '\t', s);
				const uint8_t *p = data[i] + (size_t)smp * (size_t)sizes[i] * (size_t)elt_size(types[i]);
				if (!strcmp(keys[i], "GT") && types[i] != BCF_BT_CHAR && types[i] != BCF_BT_FLOAT) {
					/* GT: integers (allele + 1) << 1 | phased */
					for (int a = 0; a < sizes[i]; a++) {
						int st;
						const int32_t g = dec_int(p, types[i], a, &st);
						if (st == 2) break;
						if (a) kputc((g & 1) ? '|' : '/', s);
						if (st == 1 || (g >> 1) == 0) kputc('.', s);
						else { char b2[16]; snprintf(b2, sizeof b2, "%d", (g >> 1) - 1); kputs(b2, s); }
					}
				} else put_vector(s, p, types[i], sizes[i]);
			}
		}
	}
	kputc('\n', s);
}

int bcf_write(htsFile *fp, bcf_hdr_t *h, bcf1_t *v) {
	if (fp->format.format == bcf) {
		uint32_t x[8];
		x[0] = (uint32_t)v->shared.l + 24;
		x[1] = (uint32_t)v->indiv.l;
		x[2] = (uint32_t)v->rid;
		x[3] = (uint32_t)v->pos;
		x[4] = (uint32_t)v->rlen;
		memcpy(x + 5, &v->qual, 4);
		x[6] = (uint32_t)v->n_allele << 16 | v->n_info;
		x[7] = (uint32_t)v->n_fmt << 24 | v->n_sample;
		return bgzf_write_bytes(fp->fp.bgzf, x, 32) | bgzf_write_bytes(fp->fp.bgzf, v->shared.s, v->shared.l) | bgzf_write_bytes(fp->fp.bgzf, v->indiv.s, v->indiv.l);
	}
	fp->line.l = 0;
	vcf_line(h, v, &fp->line);
	return bgzf_write_bytes(fp->fp.bgzf, fp->line.s, fp->line.l);
}
