// bsgpu_writer.cu -- the per-site derivations of the VCF/BCF writer on the device.
//
// What the reference's print thread does with a block of gt_vcf records (src/process.c:89-104 driving
// print_vcf_entry / flush_vcf_entries / _print_vcf_entry, src/print_vcf.c:32-381, 535-594): pick the call of every
// site, look at the calls and the reference codes two sites either side, derive QUAL / GQ, QD, FS, the filters, the
// genotype likelihood subset, the CpG status, and serialise one BCF record per site that is not skipped.  There the
// block is walked with a five-site sliding window on one thread; here every site is an independent function of the
// block (the formulation of oracle/bs_oracle_writer.c, which is pinned against the reference's compiled print_vcf.c),
// so the records are built by one thread per site, sized first, placed by a prefix sum, and leave the device as the
// byte stream bcf_write() would have produced -- about a quarter of the bytes of the gt_vcf records they replace.
//
//   k_bcf_measure   call of every site (u8) + length of its record (u16) + per-CTA byte / record totals
//   (k_bcf_calls    the calls alone: for the two sites a chunked caller needs from the chunk after the one it emits)
//   k_bcf_offsets   exclusive scan of the per-CTA totals (one CTA)
//   k_bcf_emit      records built in shared memory at their offsets inside the CTA, copied out as aligned words
//
// The byte encoding is BCF2 (VCF/BCF specification v4.3 section 6.3) as htslib's bcf_enc_* helpers produce it.
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "bsgpu.h"
#include "bsgpu_device.cuh"
#include "bsgpu_launch.h"

namespace bsgpu {

namespace {

constexpr int kWrThreads = 128;                       // sites per CTA
constexpr int kMaxRec = BSGPU_BCF_MAX_RECORD;         // upper bound of one record (checked per site)
constexpr int kStageBytes = 20 * 1024;                // shared-memory staging of a CTA's records: 160 B per site, a few times the
                                                      // typical load (45 B per site at 30x); a CTA with more writes straight to HBM
constexpr double kLn10 = 2.30258509299404568402;     // LOG10, include/bs_call.h:36

struct GtVcf {                 // bsgpu_gt_vcf as the kernels read it
	unsigned long long counts[8];
	int qual[8];
	double gt_prob[10];
	double fisher_strand;
	int mq, aq;
	uint8_t max_gt, pad0[7];
	uint8_t ready, skip, pad1[6];
};
static_assert(sizeof(GtVcf) == 208, "gt_vcf layout");

enum { T_INT8 = 1, T_INT16 = 2, T_INT32 = 3, T_FLOAT = 5, T_CHAR = 7 };

struct Count { static constexpr bool kCounting = true; uint32_t n = 0; __device__ __forceinline__ void put(uint32_t) { n++; } };
struct Store { static constexpr bool kCounting = false; uint8_t *p; uint32_t n = 0; __device__ __forceinline__ void put(uint32_t c) { p[n++] = (uint8_t)c; } };

template <class W> __device__ __forceinline__ void put_le(W &w, uint32_t v, int bytes) { for (int b = 0; b < bytes; b++) w.put(v >> (8 * b)); }
template <class W> __device__ void enc_size(W &w, int size, int type) {
	if (size < 15) { w.put(size << 4 | type); return; }
	w.put(15 << 4 | type);
	if (size < 128) { w.put(1 << 4 | T_INT8); w.put(size); }
	else if (size < 32768) { w.put(1 << 4 | T_INT16); put_le(w, (uint32_t)size, 2); }
	else { w.put(1 << 4 | T_INT32); put_le(w, (uint32_t)size, 4); }
}
__device__ __forceinline__ int int_type(int32_t mn, int32_t mx) {          // narrowest type; the lowest eight values are reserved
	if (mx <= 127 && mn >= -120) return T_INT8;
	if (mx <= 32767 && mn >= -32760) return T_INT16;
	return T_INT32;
}
template <class W> __device__ void enc_int1(W &w, int32_t x) {
	const int t = int_type(x, x);
	w.put(1 << 4 | t);
	put_le(w, (uint32_t)x, t == T_INT8 ? 1 : t == T_INT16 ? 2 : 4);
}
template <class W> __device__ void enc_vint(W &w, int n, const int32_t *a) {
	if (n == 1) { enc_int1(w, a[0]); return; }
	int32_t mx = INT32_MIN + 1, mn = INT32_MAX;
	for (int i = 0; i < n; i++) { mx = max(mx, a[i]); mn = min(mn, a[i]); }
	const int t = int_type(mn, mx);
	enc_size(w, n, t);
	for (int i = 0; i < n; i++) put_le(w, (uint32_t)a[i], t == T_INT8 ? 1 : t == T_INT16 ? 2 : 4);
}

// genotype index 0..9 = AA AC AG AT CC CG CT GG GT TT -> its two alleles as reference codes 1..4
__device__ __forceinline__ void alleles_of(int gt, int &a0, int &a1) {
	a0 = gt < 4 ? 1 : gt < 7 ? 2 : gt < 9 ? 3 : 4;
	a1 = gt < 4 ? 1 + gt : gt < 7 ? gt - 2 : gt < 9 ? gt - 4 : 4;
}
__device__ __forceinline__ int gl_index(int a, int b) { return a < b ? a * (9 - a) / 2 + b - 5 : b * (9 - b) / 2 + a - 5; }
__device__ __forceinline__ bool has_c(int gt) { return (0x072u >> gt) & 1u; }          // AC CC CG CT
__device__ __forceinline__ bool has_g(int gt) { return (0x1a4u >> gt) & 1u; }          // AG CG GG GT
__device__ __forceinline__ bool is_het(int gt) { return (0x16eu >> gt) & 1u; }         // AC AG AT CG CT GT

// the call the writer makes (src/print_vcf.c:579-588): first maximum of gt_prob, 0 for a skipped site
__device__ __forceinline__ int site_call(const GtVcf *v) {
	if (v->skip) return 0;
	int gt = 0;
	double z = v->gt_prob[0];
#pragma unroll
	for (int i = 1; i < 10; i++) { const double p = v->gt_prob[i]; if (p > z) { z = p; gt = i; } }
	return gt + 1;
}

struct WrArgs {
	const GtVcf *vcf;
	const uint8_t *ref;              // codes of positions x .. x + sz + 1
	uint32_t x, sz;                  // window: position of site 0, number of sites
	uint32_t i0, i1;                 // sites [i0, i1) are handled by this launch
	const uint2 *blocks;             // (first, last) site index of every block in the window, ascending; NULL = the window is one block
	uint32_t nblocks;
	int32_t ids[16];
	int32_t rid;
	uint32_t ctg_end;
	int all_positions;
	const DevConst *dc;
	uint8_t *calls;                  // per site: 0, or 1 + genotype
	uint16_t *len;                   // per site: bytes of its record
	unsigned long long *cta_bytes;   // per CTA: bytes, then (after k_bcf_offsets) offset of the CTA's first byte
	uint32_t *cta_recs;
	unsigned long long *totals;      // [0] bytes, [1] records, [2] sites whose record exceeded kMaxRec
	uint8_t *out;
	unsigned long long out_cap;
	unsigned long long *guard;       // the context's counters (guard bands), or NULL
	DbView db;                       // dbSNP entries of the contig (words == 0: none)
	uint32_t reg_start, reg_stop;    // ctg->curr_reg (0, 0: none)
	// site statistics (k_bcf_stats)
	bsgpu_site_stats *stats;
	bsgpu_ctg_site_stats *ctg_stats; // this contig's entry, or NULL
	uint32_t *stats_carry;           // two words, written in turn (carry_flip) by successive launches over one window: 0 / 1 / 2 =
	uint32_t carry_flip;             // the last site of the launch was not counted / counted unfiltered / counted with a filter set
	const uint8_t *gc;
	uint32_t gc_bins, gc_start;
};

__device__ __forceinline__ void guard_note(unsigned long long *counters, int kind, unsigned long long id) {
	atomicAdd(counters + 4 + kind, 1ull);
	const unsigned long long k = atomicAdd(counters + 8, 1ull);
	if (k < (unsigned long long)kGuardCap) counters[kGuardList + k] = (unsigned long long)kind << 56 | (id & 0x00ffffffffffffffull);
}

// block of site i: first / last site index; false when the site lies between blocks.  A site shared by two touching
// blocks is the earlier block's (the later visit is dropped by the x <= old_x test, src/print_vcf.c:127).
__device__ bool block_of(const WrArgs &a, uint32_t i, uint32_t &first, uint32_t &last) {
	if (!a.blocks) { first = 0; last = a.sz - 1; return true; }
	uint32_t lo = 0, hi = a.nblocks;
	while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a.blocks[mid].y < i) lo = mid + 1; else hi = mid; }
	if (lo == a.nblocks || a.blocks[lo].x > i) return false;
	first = a.blocks[lo].x; last = a.blocks[lo].y;
	return true;
}

// Reference context of site i.  The reference fills its window with strncpy() from a string in which N is the terminator
// (src/print_vcf.c:572-578), so once an N has been copied everything after it reads as N; the window starts at
// site - 2, except for the last two sites of a block, which reuse the window of the block's last site
// (flush_vcf_entries only shifts it, :539-543): an N up to two codes further left wipes them too.  Codes before the
// block are N.
__device__ __forceinline__ void ref_context(const WrArgs &a, uint32_t i, uint32_t first, uint32_t last, uint32_t rc[5]) {
	const int64_t ii = i, f = first, l = last;
	const int64_t wstart = ii + 2 <= l ? ii - 2 : l - 4;
	bool wiped = false;
	for (int64_t j = wstart < f ? f : wstart; j < ii - 2; j++) wiped |= a.ref[j] == 0;
#pragma unroll
	for (int k = 0; k < 5; k++) {
		const int64_t j = ii + k - 2;
		const uint32_t c = j >= f && !wiped ? a.ref[j] : 0u;
		if (c == 0 && j >= f) wiped = true;
		rc[k] = c;
	}
}

// dbSNP (src/print_vcf.c:133): what dbSNP_lookup_name() answers for a position -- 0, 1 (known) or 3 (known, always written) and the ID
__device__ __forceinline__ uint32_t dbsnp_lookup(const WrArgs &a, uint32_t pos, const uint8_t *&rs, uint32_t &rs_len) {
	rs = nullptr;
	rs_len = 0;
	if ((pos >> 6) >= a.db.words) return 0;
	const unsigned long long m = a.db.mask[pos >> 6], bit = 1ull << (pos & 63);
	if (!(m & bit)) return 0;
	const uint32_t k = a.db.cum[pos >> 6] + (uint32_t)__popcll(m & (bit - 1));
	rs = a.db.names + a.db.off[k];
	rs_len = a.db.off[k + 1] - a.db.off[k];
	return (a.db.fq[pos >> 6] & bit) ? 3u : 1u;
}

// Does site i get a record?  `own` = its call (0: skipped).  The tests build_record() makes before it starts to write
// (src/print_vcf.c:139, 154-157), on their own so that a CTA can gather its writing sites before it builds anything.
__device__ __forceinline__ bool site_writes(const WrArgs &a, const GtVcf *v, uint32_t i, uint32_t first, uint32_t last, int own) {
	if (!own) return false;
	uint32_t dp = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) dp += (uint32_t)v->counts[k];
	if (!dp) return false;
	const uint32_t pos = a.x + i;
	if (a.reg_start | a.reg_stop) { if (pos < a.reg_start || pos > a.reg_stop) return false; }
	else if (pos > a.ctg_end) return false;
	if (a.all_positions) return true;
	const int gt = own - 1;
	if (gt != 0 && gt != 9) return true;
	uint32_t rc[5];
	ref_context(a, i, first, last, rc);
	if (!((gt == 0 && rc[2] == 1) || (gt == 9 && rc[2] == 4))) return true;
	const uint8_t *rs;
	uint32_t rs_len;
	return (dbsnp_lookup(a, pos, rs, rs_len) & 2u) != 0;
}

// One record.  `g` = calls of sites i-2 .. i+2 as the writer's window holds them.  Returns bytes written (0: no record).
template <class W>
__device__ __forceinline__ uint32_t build_record(const WrArgs &a, const GtVcf *v, uint32_t i, uint32_t first, uint32_t last, const int g[5], W &w) {
	if (!g[2]) return 0;
	uint32_t dp1 = 0, dinf = 0;
#pragma unroll
	for (int k = 0; k < 4; k++) { dp1 += (uint32_t)v->counts[k]; dinf += (uint32_t)v->counts[4 + k]; }
	if (!(dp1 + dinf)) return 0;
	uint32_t rc[5];
	ref_context(a, i, first, last, rc);
	const int rfix = (int)rc[2], gt = g[2] - 1;
	const uint32_t pos = a.x + i;
	const uint8_t *rs;
	uint32_t rs_len;
	const uint32_t rs_found = dbsnp_lookup(a, pos, rs, rs_len);
	if (!a.all_positions && !(rs_found & 2u) && ((gt == 0 && rfix == 1) || (gt == 9 && rfix == 4))) return 0;      // hom-ref A / T (gt_flag, :91-102, 139)
	if (a.reg_start | a.reg_stop) { if (pos < a.reg_start || pos > a.reg_stop) return 0; }      // ctg->curr_reg clips, else the contig end (:154-157)
	else if (pos > a.ctg_end) return 0;
	// QUAL / GQ: phred of the probability that the call is wrong (:140-148)
	const MathTables *mt = &a.dc->tab.math;
	const double lp = v->gt_prob[gt] * kLn10;
	const double z1 = lp == 0.0 ? 1.0 : fast_exp(lp < -700.0 ? -700.0 : lp, mt);
	int phred;
	if (z1 >= 1.0) phred = 255;
	else {
		const double ph = -10.0 * fast_log(1.0 - z1, mt) / kLn10;
		phred = (int)ph; if (phred > 255) phred = 255;
		// Guard band (counted once, in the measuring pass): the posterior is within 1e-9 relative of the reference's, which
		// moves 1 - z1 by |lp| 1e-9 + an ulp of 1 and ph by 10 / ln 10 times that relative change; a ph that close to an
		// integer (below the 255 cap) may truncate differently.
		if (W::kCounting && a.guard && ph < 255.5) {
			const double band = 4.342944819 * (fabs(lp) * 1.0e-9 + 2.3e-16) / (1.0 - z1) + 1.0e-9 * ph + 1.0e-12;
			const double fr = ph - floor(ph);
			if (fr < band || 1.0 - fr < band) guard_note(a.guard, 2, (unsigned long long)a.x + i);
		}
	}
	const double fsv = -v->fisher_strand * 10.0 + 0.5;
	const int fs = (int)fsv;
	if (W::kCounting && a.guard && is_het(gt)) {        // FS is only written for a het call; fisher_strand is within 1e-8 relative
		const double band = fabs(fsv) * 1.0e-8 + 1.0e-12, fr = fsv - floor(fsv);
		if (fr < band || 1.0 - fr < band) guard_note(a.guard, 3, (unsigned long long)a.x + i);
	}
	const uint32_t qd = dp1 > 0 ? (uint32_t)phred / dp1 : (uint32_t)phred;
	uint32_t flt = 0;
	if (phred < 20) flt |= 1;
	if (qd < 2) flt |= 2;
	if (fs > 60) flt |= 4;
	if (v->mq < 40) flt |= 8;
	int32_t fid = a.ids[0];
	if (!flt) {
		const unsigned long long *c = v->counts;
		bool mac1 = false;
		switch (gt) {                                                                  // :190-210
		case 1: mac1 = c[1] + c[5] + c[7] <= 1 || c[0] + c[4] <= 1; break;
		case 2: mac1 = c[2] + c[6] <= 1 || c[0] <= 1; break;
		case 3: mac1 = c[3] + c[7] <= 1 || c[0] + c[4] <= 1; break;
		case 5: mac1 = c[2] + c[6] + c[4] <= 1 || c[1] + c[5] + c[7] <= 1; break;
		case 6: mac1 = c[3] <= 1 || c[1] + c[5] <= 1; break;
		case 8: mac1 = c[3] + c[7] <= 1 || c[2] + c[6] + c[4] <= 1; break;
		}
		if (mac1) fid = a.ids[2];
	} else fid = a.ids[1];
	// ALT: the alleles of the call that are not the reference base, in base order (ref_alt / all_idx, :34-45, 62-73)
	int a0, a1;
	alleles_of(gt, a0, a1);
	int alts[2] = { 0, 0 }, n_alt = 0;
	if (a0 != rfix) alts[n_alt++] = a0;
	if (a1 != rfix && a1 != a0) alts[n_alt++] = a1;
	const uint32_t base_char = 0x54474341u;            // "ACGT"
	auto bchar = [&](uint32_t c) -> uint32_t { return c ? (base_char >> (8 * (c - 1))) & 0xffu : (uint32_t)'N'; };

	// ---- the fixed part is written last (it holds the lengths); shared then indiv follow it
	const uint32_t start = w.n;
	for (int k = 0; k < 32; k++) w.put(0);
	const uint32_t sh0 = w.n;
	enc_size(w, (int)rs_len, T_CHAR);                  // ID (:163-167): the dbSNP name, or none
	for (uint32_t k = 0; k < rs_len; k++) w.put(rs[k]);
	w.put(1 << 4 | T_CHAR); w.put(bchar(rc[2]));       // REF
	for (int k = 0; k < n_alt; k++) { w.put(1 << 4 | T_CHAR); w.put(bchar((uint32_t)alts[k])); }
	enc_int1(w, fid);                                  // FILTER
	enc_int1(w, a.ids[3]);                             // INFO CX: reference context
	w.put(5 << 4 | T_CHAR);
#pragma unroll
	for (int k = 0; k < 5; k++) w.put(bchar(rc[k]));
	const uint32_t l_shared = w.n - sh0;

	const uint32_t in0 = w.n;
	uint32_t n_fmt = 11;
	// GT as the reference's gt_int table has it (:75-86): 2,2 hom-ref; 4,4 hom-alt; 2,4 het with the reference allele;
	// 4,8 -- not 4,6 -- for a het of two ALT alleles
	enc_int1(w, a.ids[4]);
	w.put(2 << 4 | T_INT8);
	if (a0 == a1) { const uint32_t c = a0 == rfix ? 2 : 4; w.put(c); w.put(c); }
	else if (a0 == rfix || a1 == rfix) { w.put(2); w.put(4); }
	else { w.put(4); w.put(8); }
	// FT (:277-301): the names of the failed filters, each WITH its terminating NUL (the copy loop steps over it), ';' between
	enc_int1(w, a.ids[5]);
	if (flt & 15u) {
		const uint32_t nl = ((flt & 1u) ? 4 : 0) + ((flt & 2u) ? 4 : 0) + ((flt & 4u) ? 5 : 0) + ((flt & 8u) ? 5 : 0) + __popc(flt & 15u) - 1;
		enc_size(w, (int)nl, T_CHAR);
		bool some = false;
		if (flt & 1u) { w.put('q'); w.put('2'); w.put('0'); w.put(0); some = true; }
		if (flt & 2u) { if (some) w.put(';'); w.put('q'); w.put('d'); w.put('2'); w.put(0); some = true; }
		if (flt & 4u) { if (some) w.put(';'); w.put('f'); w.put('s'); w.put('6'); w.put('0'); w.put(0); some = true; }
		if (flt & 8u) { if (some) w.put(';'); w.put('m'); w.put('q'); w.put('4'); w.put('0'); w.put(0); }
	} else { w.put(4 << 4 | T_CHAR); w.put('P'); w.put('A'); w.put('S'); w.put('S'); }
	enc_int1(w, a.ids[8]); enc_int1(w, (int32_t)dp1);  // DP
	enc_int1(w, a.ids[9]); enc_int1(w, v->mq);         // MQ
	enc_int1(w, a.ids[7]); enc_int1(w, phred);         // GQ
	enc_int1(w, a.ids[10]); enc_int1(w, (int32_t)qd);  // QD
	{                                                  // GL (:317-345): RR, then per ALT allele R/A (when the reference base is known) and A/A
		enc_int1(w, a.ids[6]);
		const int n = 1 + n_alt * (rfix ? 2 : 1);
		w.put(n << 4 | T_FLOAT);
		auto putf = [&](double z) { if (z < -99.999) z = -99.999; put_le(w, __float_as_uint((float)z), 4); };
		putf(rfix ? v->gt_prob[gl_index(rfix, rfix)] : -99.999);
		for (int k = 0; k < n_alt; k++) {
			if (rfix) putf(v->gt_prob[gl_index(rfix, alts[k])]);
			putf(v->gt_prob[gl_index(alts[k], alts[k])]);
		}
	}
	{                                                  // MC8, AMQ (:347-358)
		int32_t c8[8];
#pragma unroll
		for (int k = 0; k < 8; k++) c8[k] = (int32_t)v->counts[k];
		enc_int1(w, a.ids[11]);
		enc_vint(w, 8, c8);
		int32_t q8[8];
		int n = 0;
#pragma unroll
		for (int k = 0; k < 8; k++) if (v->counts[k] > 0) q8[n++] = v->qual[k];
		if (n) { enc_int1(w, a.ids[12]); enc_vint(w, n, q8); n_fmt++; }
	}
	enc_int1(w, a.ids[13]);                            // CS: strand(s) on which the call has a cytosine (:60-61)
	if (has_c(gt)) { if (has_g(gt)) { w.put(2 << 4 | T_CHAR); w.put('+'); w.put('-'); } else { w.put(1 << 4 | T_CHAR); w.put('+'); } }
	else if (has_g(gt)) { w.put(1 << 4 | T_CHAR); w.put('-'); }
	else { w.put(2 << 4 | T_CHAR); w.put('N'); w.put('A'); }
	{                                                  // CG: CpG status from the calls either side (:229-270); "CG" goes out as its first character
		const int c0 = g[2], nx = g[3], pv = g[1];
		uint32_t cg;
		if ((c0 == 5 && nx == 8) || (c0 == 8 && pv == 5)) cg = 'C';
		else if (c0 == 5) cg = nx ? (has_g(nx - 1) ? 'H' : 'N') : '?';
		else if (c0 == 8) cg = pv ? (has_c(pv - 1) ? 'H' : 'N') : '?';
		else if (has_c(c0 - 1)) cg = nx ? (has_g(nx - 1) ? 'H' : 'N') : '?';
		else if (has_g(c0 - 1)) cg = pv ? (has_c(pv - 1) ? 'H' : 'N') : '.';
		else cg = '.';
		enc_int1(w, a.ids[14]);
		w.put(1 << 4 | T_CHAR); w.put(cg);
	}
	enc_int1(w, a.ids[3]);                             // CX: context from the calls, IUPAC
	w.put(5 << 4 | T_CHAR);
#pragma unroll
	for (int k = 0; k < 5; k++) w.put((uint32_t)"NAMRWCSYGKT"[g[k]]);
	if (is_het(gt)) { enc_int1(w, a.ids[15]); enc_int1(w, fs); n_fmt++; }      // FS
	const uint32_t l_indiv = w.n - in0;

	// the eight words in front, as bcf_write lays a record out
	const uint32_t end = w.n;
	w.n = start;
	put_le(w, l_shared + 24, 4);
	put_le(w, l_indiv, 4);
	put_le(w, (uint32_t)a.rid, 4);
	put_le(w, a.x + i - 1, 4);
	put_le(w, 1u, 4);
	put_le(w, __float_as_uint((float)phred), 4);
	put_le(w, (uint32_t)(1 + n_alt) << 16 | 1u, 4);
	put_le(w, n_fmt << 24 | 1u, 4);
	w.n = end;
	return end - start;
}

__global__ void __launch_bounds__(256) k_bcf_calls(const WrArgs a) {
	const uint32_t i = a.i0 + blockIdx.x * 256 + threadIdx.x;
	if (i < a.i1) a.calls[i] = (uint8_t)site_call(a.vcf + i);
}

// the writer's window around site i: nothing before the block; beyond its end the last site's call is seen again
// (flush_vcf_entries shifts the window without clearing the slot it vacates, src/print_vcf.c:538)
__device__ __forceinline__ void window_calls(const WrArgs &a, uint32_t i, uint32_t first, uint32_t last, int g[5]) {
#pragma unroll
	for (int k = 0; k < 5; k++) {
		const int64_t j = (int64_t)i + k - 2;
		g[k] = j < (int64_t)first ? 0 : a.calls[j <= (int64_t)last ? j : (int64_t)last];
	}
}

// The writing sites of the CTA, gathered: `mine` = this thread's site writes.  Returns their number; list[k] = thread index of
// the k-th one, in site order (roughly 40 % of the sites of a WGBS window write a record).
__device__ __forceinline__ uint32_t gather_writers(bool mine, uint16_t *list, uint32_t *wcnt) {
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t bal = __ballot_sync(0xffffffffu, mine);
	if (lane == 0) wcnt[wid] = __popc(bal);
	__syncthreads();
	uint32_t base = 0, total = 0;
#pragma unroll
	for (int k = 0; k < kWrThreads / 32; k++) { if (k < wid) base += wcnt[k]; total += wcnt[k]; }
	if (mine) list[base + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)threadIdx.x;
	__syncthreads();
	return total;
}

__global__ void __launch_bounds__(kWrThreads) k_bcf_measure(const WrArgs a) {
	__shared__ uint32_t wsum[kWrThreads / 32][2];
	const uint32_t i = a.i0 + blockIdx.x * kWrThreads + threadIdx.x;
	uint32_t n = 0;
	if (i < a.i1) {
		// the site's own call is made here (and kept for the neighbours' records); a record's LENGTH does not depend on
		// the calls around it, so the window is left empty for the count.  (Gathering the writing sites first, as the single
		// pass below does, was measured here too: 1.19 against 1.15 ms per 8 M sites -- the lanes it frees were hiding the
		// latency of the scattered record loads.)
		const int own = site_call(a.vcf + i);
		a.calls[i] = (uint8_t)own;
		uint32_t first, last;
		if (block_of(a, i, first, last)) {
			const int g[5] = { 0, 0, own, 0, 0 };
			Count w;
			n = build_record(a, a.vcf + i, i, first, last, g, w);
			if (n > (uint32_t)kMaxRec) { atomicAdd(a.totals + 2, 1ull); n = 0; }
		}
		a.len[i] = (uint16_t)n;
	}
	const uint32_t bytes = __reduce_add_sync(0xffffffffu, n), recs = __reduce_add_sync(0xffffffffu, n ? 1u : 0u);
	if ((threadIdx.x & 31) == 0) { wsum[threadIdx.x >> 5][0] = bytes; wsum[threadIdx.x >> 5][1] = recs; }
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t b = 0, r = 0;
		for (int k = 0; k < kWrThreads / 32; k++) { b += wsum[k][0]; r += wsum[k][1]; }
		a.cta_bytes[blockIdx.x] = b;
		a.cta_recs[blockIdx.x] = r;
	}
}

// exclusive scan of the per-CTA byte totals, in place; totals[0] = bytes, totals[1] = records of the window
__global__ void __launch_bounds__(1024) k_bcf_offsets(unsigned long long *cta_bytes, const uint32_t *cta_recs, uint32_t nctas, unsigned long long *totals) {
	__shared__ unsigned long long wsum[32];
	__shared__ unsigned long long carry_s;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	if (tid == 0) carry_s = 0;
	unsigned long long recs = 0;
	__syncthreads();
	for (uint32_t base = 0; base < nctas; base += 1024) {
		const uint32_t j = base + tid;
		const unsigned long long v = j < nctas ? cta_bytes[j] : 0ull;
		if (j < nctas) recs += cta_recs[j];
		unsigned long long inc = v;
		for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
		if (lane == 31) wsum[wid] = inc;
		__syncthreads();
		unsigned long long pre = carry_s;
		for (int k = 0; k < wid; k++) pre += wsum[k];
		if (j < nctas) cta_bytes[j] = pre + inc - v;
		__syncthreads();
		if (tid == 1023) carry_s = pre + inc;
		__syncthreads();
	}
	for (int d = 16; d; d >>= 1) recs += __shfl_down_sync(0xffffffffu, recs, d);
	if (lane == 0) wsum[wid] = recs;
	__syncthreads();
	if (tid == 0) {
		unsigned long long r = 0;
		for (int k = 0; k < 32; k++) r += wsum[k];
		totals[0] = carry_s;
		totals[1] = r;
	}
}

// copy a CTA's staged records out: bytes up to the first aligned word of the destination, aligned words (each stitched
// from two staged words), bytes after the last one
__device__ __forceinline__ void copy_stage_out(uint8_t *dst, const uint8_t *stage, uint32_t total) {
	const uint32_t head = min(total, (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3));
	if (threadIdx.x < head) dst[threadIdx.x] = stage[threadIdx.x];
	const uint32_t nwords = (total - head) >> 2;
	uint32_t *dw = (uint32_t *)(dst + head);
	const uint32_t *sw = (const uint32_t *)stage;
	const uint32_t sh = (head & 3) * 8;
	for (uint32_t k = threadIdx.x; k < nwords; k += kWrThreads) {
		const uint32_t b = head + 4 * k;                     // staged byte offset of this word
		const uint32_t lo = sw[b >> 2], hi = sw[(b >> 2) + 1];
		dw[k] = sh ? __funnelshift_r(lo, hi, sh) : lo;
	}
	const uint32_t tail0 = head + 4 * nwords;
	if (threadIdx.x < total - tail0) dst[tail0 + threadIdx.x] = stage[tail0 + threadIdx.x];
}

__global__ void __launch_bounds__(kWrThreads) k_bcf_emit(const WrArgs a) {
	extern __shared__ __align__(16) uint8_t stage[];          // the CTA's records, back to back
	__shared__ uint32_t wsum[kWrThreads / 32];
	const uint32_t i = a.i0 + blockIdx.x * kWrThreads + threadIdx.x;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t n = i < a.i1 ? a.len[i] : 0u;
	uint32_t inc = n;
	for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
	if (lane == 31) wsum[wid] = inc;
	__syncthreads();
	uint32_t off = inc - n, total = 0;
	for (int k = 0; k < kWrThreads / 32; k++) { if (k < wid) off += wsum[k]; total += wsum[k]; }
	const unsigned long long dst0 = a.cta_bytes[blockIdx.x];
	if (dst0 + total > a.out_cap) return;                   // the host compares totals[0] with the capacity and reports
	uint8_t *dst = a.out + dst0;
	const bool staged = total <= (uint32_t)kStageBytes;
	if (n) {
		uint32_t first, last;
		block_of(a, i, first, last);
		int g[5];
		window_calls(a, i, first, last, g);
		// two instances on purpose: with the destination's address space known the byte stores are STS / STG, not generic
		if (staged) { Store w; w.p = stage + off; build_record(a, a.vcf + i, i, first, last, g, w); }
		else { Store w; w.p = dst + off; build_record(a, a.vcf + i, i, first, last, g, w); }
	}
	if (!staged) return;
	__syncthreads();
	copy_stage_out(dst, stage, total);
}

// ---------------------------------------------------------------------------------------------------------------
// One pass: the three kernels above read every 208-byte record twice, one thread per record (a quarter of every sector
// fetched is used, and the two passes together move almost twice the algorithmic bytes).  Here a CTA takes the next tile of
// 128 sites (ticket), pulls its records into shared memory with coalesced 16-byte loads, makes the calls (its own and the
// two either side), gathers the writing sites, sizes their records, builds them into the stage, and finds where its bytes go
// by a decoupled look-back over the tiles before it (tile_state: 2 status bits | bytes; tiles are handed out in order, so
// every tile a CTA waits for is already running).  Same bytes as the split kernels.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kFusedStage = 16 * 1024;
constexpr unsigned long long kTileMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long *p) { return *(const volatile unsigned long long *)p; }

__global__ void __launch_bounds__(kWrThreads) k_bcf_fused(const WrArgs a, unsigned long long *tile_state, unsigned int *ticket, uint32_t ntiles) {
	extern __shared__ __align__(16) uint8_t dyn[];
	GtVcf *recs = (GtVcf *)dyn;
	uint8_t *stage = dyn + kWrThreads * sizeof(GtVcf);
	__shared__ uint32_t wsum[kWrThreads / 32], wcnt[kWrThreads / 32];
	__shared__ uint16_t list[kWrThreads];
	__shared__ uint2 blk[kWrThreads];
	__shared__ uint8_t calls_s[kWrThreads + 4];
	__shared__ uint32_t s_tile;
	__shared__ unsigned long long s_excl;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	if (tid == 0) s_tile = atomicAdd(ticket, 1u);
	__syncthreads();
	const uint32_t tile = s_tile;
	if (tile >= ntiles) return;
	const uint32_t i0 = a.i0 + tile * kWrThreads, cnt = min((uint32_t)kWrThreads, a.i1 - i0);
	{
		const uint4 *src = (const uint4 *)(a.vcf + i0);
		uint4 *dst = (uint4 *)recs;
		for (uint32_t k = tid; k < cnt * 13u; k += kWrThreads) dst[k] = __ldg(src + k);
	}
	if (tid < 4) {      // the calls two sites either side of the tile: made here while their records are in reach, else from the calls array
		const int64_t j = tid < 2 ? (int64_t)i0 - 2 + tid : (int64_t)i0 + kWrThreads + (tid - 2);
		uint8_t c = 0;
		if (j >= 0 && j < (int64_t)a.sz) c = (j >= (int64_t)a.i0 && j < (int64_t)a.i1) ? (uint8_t)site_call(a.vcf + j) : a.calls[j];
		calls_s[tid < 2 ? tid : kWrThreads + tid] = c;
	}
	__syncthreads();
	bool mine = false;
	if ((uint32_t)tid < cnt) {
		const int own = site_call(recs + tid);
		calls_s[2 + tid] = (uint8_t)own;
		a.calls[i0 + tid] = (uint8_t)own;
		uint32_t first, last;
		if (block_of(a, i0 + tid, first, last)) {
			blk[tid] = make_uint2(first, last);
			mine = site_writes(a, recs + tid, i0 + tid, first, last, own);
		}
	} else calls_s[2 + tid] = 0;
	const uint32_t nw = gather_writers(mine, list, wcnt);
	// lengths of the records, dense: thread k sizes the k-th writing site
	uint32_t n = 0, t = 0;
	if ((uint32_t)tid < nw) {
		t = list[tid];
		const int g[5] = { 0, 0, (int)calls_s[2 + t], 0, 0 };
		Count w;
		n = build_record(a, recs + t, i0 + t, blk[t].x, blk[t].y, g, w);
		if (n > (uint32_t)kMaxRec) { atomicAdd(a.totals + 2, 1ull); n = 0; }
	}
	uint32_t inc = n;
	for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
	const uint32_t nrec_w = __popc(__ballot_sync(0xffffffffu, n != 0));
	if (lane == 31) wsum[wid] = inc;
	if (lane == 0) wcnt[wid] = nrec_w;       // (gather_writers is done with wcnt)
	__syncthreads();
	uint32_t off = inc - n, total = 0, nrec = 0;
#pragma unroll
	for (int k = 0; k < kWrThreads / 32; k++) { if (k < wid) off += wsum[k]; total += wsum[k]; nrec += wcnt[k]; }
	// let the tiles behind this one see its bytes as early as possible
	if (tid == 0) {
		atomicExch(tile_state + tile, (tile ? 1ull : 2ull) << 62 | (unsigned long long)total);
		if (nrec) atomicAdd(a.totals + 1, (unsigned long long)nrec);
	}
	auto look_back = [&]() {            // warp 0: bytes of all tiles before this one
		if (wid != 0) return;
		unsigned long long excl = 0;
		for (int64_t p = (int64_t)tile - 1; p >= 0; p -= 32) {
			const int64_t q = p - lane;
			unsigned long long v = q >= 0 ? ld_state(tile_state + q) : 2ull << 62;
			while (__any_sync(0xffffffffu, (v >> 62) == 0)) { if ((v >> 62) == 0) v = ld_state(tile_state + q); }
			const uint32_t pref = __ballot_sync(0xffffffffu, (v >> 62) == 2);
			const int stop = pref ? __ffs(pref) - 1 : 31;
			unsigned long long c = lane <= stop ? (v & kTileMask) : 0ull;
			for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
			excl += c;
			if (pref) break;
		}
		if (lane == 0) {
			s_excl = excl;
			if (tile) atomicExch(tile_state + tile, 2ull << 62 | (excl + total));
			if (tile == ntiles - 1) a.totals[0] = excl + total;
		}
	};
	auto window = [&](uint32_t tt, int g[5]) {      // window_calls() from the tile's own calls
		const int64_t i = (int64_t)i0 + tt, first = blk[tt].x, last = blk[tt].y;
#pragma unroll
		for (int k = 0; k < 5; k++) {
			const int64_t j = i + k - 2;
			g[k] = j < first ? 0 : (int)calls_s[(j <= last ? j : last) - (int64_t)i0 + 2];
		}
	};
	if (total <= (uint32_t)kFusedStage) {
		if (n) {
			int g[5];
			window(t, g);
			Store w; w.p = stage + off;
			build_record(a, recs + t, i0 + t, blk[t].x, blk[t].y, g, w);
		}
		look_back();
		__syncthreads();
		if (s_excl + total <= a.out_cap) copy_stage_out(a.out + s_excl, stage, total);      // else: the host compares totals[0] with the capacity and reports
	} else {                            // more bytes than the stage holds: straight to where they go
		look_back();
		__syncthreads();
		if (n && s_excl + total <= a.out_cap) {
			int g[5];
			window(t, g);
			Store w; w.p = a.out + s_excl + off;
			build_record(a, recs + t, i0 + t, blk[t].x, blk[t].y, g, w);
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// --report-file statistics of the sites (src/print_vcf.c:382-526; restated in oracle/bs_oracle_stats.c, which is pinned
// against the compiled print_vcf.c).  One thread per site re-derives what the record builder derives (call window, reference
// context, dbSNP hit, skip, QUAL, QD, FS, filter bits); counters that many sites share go through per-CTA histograms in shared
// memory, the sparse ones (coverage table, FS beyond 63) straight to global memory, lanes with equal keys combined first.
// The only state the reference carries from site to site -- position and filter state of the last '+' strand CpG
// (prev_cpg_x, prev_cpg_flt, :107-108) -- is a function of the site before: a '-' strand CpG at x is called GG behind a call
// CC at x - 1, which is then a '+' strand CpG itself and set that state iff it was visited and not skipped.
// The methylation posteriors of the CpG strands (101 exponentials each) are pooled per CTA and evaluated one entry per warp.
// ---------------------------------------------------------------------------------------------------------------
struct SiteDer { uint32_t rc[5]; int gt, rfix; uint32_t rs_found; bool skip; int phred, fs; uint32_t qd, flt, dp1, dinf; };

// false: the writer does not visit the site (not called, no depth)
__device__ bool derive_site(const WrArgs &a, uint32_t i, uint32_t first, uint32_t last, const int g[5], SiteDer &d) {
	if (!g[2]) return false;
	const GtVcf *v = a.vcf + i;
	d.dp1 = d.dinf = 0;
#pragma unroll
	for (int k = 0; k < 4; k++) { d.dp1 += (uint32_t)v->counts[k]; d.dinf += (uint32_t)v->counts[4 + k]; }
	if (!(d.dp1 + d.dinf)) return false;
	{
		const int64_t ii = i, f = first, l = last;
		const int64_t wstart = ii + 2 <= l ? ii - 2 : l - 4;
		bool wiped = false;
		for (int64_t j = wstart < f ? f : wstart; j < ii - 2; j++) wiped |= a.ref[j] == 0;
#pragma unroll
		for (int k = 0; k < 5; k++) {
			const int64_t j = ii + k - 2;
			const uint32_t c = j >= f && !wiped ? a.ref[j] : 0u;
			if (c == 0 && j >= f) wiped = true;
			d.rc[k] = c;
		}
	}
	d.rfix = (int)d.rc[2]; d.gt = g[2] - 1;
	const uint32_t pos = a.x + i;
	d.rs_found = 0;
	if ((pos >> 6) < a.db.words) {
		const unsigned long long bit = 1ull << (pos & 63);
		if (a.db.mask[pos >> 6] & bit) d.rs_found = (a.db.fq[pos >> 6] & bit) ? 3u : 1u;
	}
	d.skip = !a.all_positions && !(d.rs_found & 2u) && ((d.gt == 0 && d.rfix == 1) || (d.gt == 9 && d.rfix == 4));
	const MathTables *mt = &a.dc->tab.math;
	const double lp = v->gt_prob[d.gt] * kLn10;
	const double z1 = lp == 0.0 ? 1.0 : fast_exp(lp < -700.0 ? -700.0 : lp, mt);
	if (z1 >= 1.0) d.phred = 255;
	else { d.phred = (int)(-10.0 * fast_log(1.0 - z1, mt) / kLn10); if (d.phred > 255) d.phred = 255; }
	d.fs = (int)(-v->fisher_strand * 10.0 + 0.5);
	d.qd = d.dp1 > 0 ? (uint32_t)d.phred / d.dp1 : (uint32_t)d.phred;
	if (!d.skip) {
		if (a.reg_start | a.reg_stop) d.skip = pos < a.reg_start || pos > a.reg_stop;
		else d.skip = pos > a.ctg_end;
	}
	d.flt = 0;
	if (!d.skip) {
		if (d.phred < 20) d.flt |= 1;
		if (d.qd < 2) d.flt |= 2;
		if (d.fs > 60) d.flt |= 4;
		if (v->mq < 40) d.flt |= 8;
		if (!d.flt) {
			const unsigned long long *c = v->counts;
			bool mac1 = false;
			switch (d.gt) {
			case 1: mac1 = c[1] + c[5] + c[7] <= 1 || c[0] + c[4] <= 1; break;
			case 2: mac1 = c[2] + c[6] <= 1 || c[0] <= 1; break;
			case 3: mac1 = c[3] + c[7] <= 1 || c[0] + c[4] <= 1; break;
			case 5: mac1 = c[2] + c[6] + c[4] <= 1 || c[1] + c[5] + c[7] <= 1; break;
			case 6: mac1 = c[3] <= 1 || c[1] + c[5] <= 1; break;
			case 8: mac1 = c[3] + c[7] <= 1 || c[2] + c[6] + c[4] <= 1; break;
			}
			if (mac1) d.flt |= 128;
		}
	}
	return true;
}

// layout of the per-CTA histograms (32-bit counters: a CTA has 128 sites)
enum { H_MISC = 0, H_MUT = 12, H_DBMUT = 36, H_QUAL = 60, H_FLT = H_QUAL + 1024, H_QD = H_FLT + 64, H_MQ = H_QD + 512, H_FS = H_MQ + 512, H_END = H_FS + 128 };
constexpr int kFsShared = 64;            // FS values below this are counted in shared memory (0 for every homozygous call)

__device__ __forceinline__ void add_u64(uint64_t *p, uint32_t v) { atomicAdd((unsigned long long *)p, (unsigned long long)v); }

__global__ void __launch_bounds__(kWrThreads) k_bcf_stats(const WrArgs a) {
	__shared__ uint32_t h[H_END];
	__shared__ double meth[4][101];          // [ref all, ref passed, nonref all, nonref passed]
	__shared__ uint32_t cg_a[kWrThreads], cg_b[kWrThreads], cg_w[kWrThreads], cg_n;
	__shared__ double logp[100];
	const int tid = threadIdx.x, lane = tid & 31;
	for (int k = tid; k < H_END; k += kWrThreads) h[k] = 0;
	for (int k = tid; k < 404; k += kWrThreads) (&meth[0][0])[k] = 0.0;
	if (tid == 0) cg_n = 0;
	__syncthreads();
	const uint32_t i = a.i0 + blockIdx.x * kWrThreads + tid;
	bsgpu_site_stats *st = a.stats;
	bool visited = false;
	SiteDer d;
	uint32_t first = 0, last = 0;
	int g[5] = { 0, 0, 0, 0, 0 };
	if (i < a.i1 && block_of(a, i, first, last)) {
		window_calls(a, i, first, last, g);
		visited = derive_site(a, i, first, last, g, d);
	}
	// coverage table: cov[dp].all and the GC histogram of the depth (:385-398), lanes with the same key combined
	{
		const uint32_t dp = visited ? d.dp1 + d.dinf : 0xffffffffu;
		uint32_t gcv = 0xffu;
		if (visited && a.gc) {
			const int bn = (int)((a.x + i - a.gc_start) / 100u);
			if (bn >= 0 && (uint32_t)bn < a.gc_bins) gcv = a.gc[bn];
		}
		const uint32_t m1 = __match_any_sync(0xffffffffu, dp);
		if (visited && lane == __ffs(m1) - 1) {
			if (dp < (uint32_t)BSGPU_STATS_COV_MAX) add_u64(&st->cov[dp].all, __popc(m1)); else add_u64(&st->cov_overflow, __popc(m1));
		}
		const uint32_t key = visited && gcv <= 100u && dp < (uint32_t)BSGPU_STATS_COV_MAX ? dp << 7 | gcv : 0xffffffffu;
		const uint32_t m2 = __match_any_sync(0xffffffffu, key);
		if (key != 0xffffffffu && lane == __ffs(m2) - 1) add_u64(&st->cov[dp].gc_pcent[gcv], __popc(m2));
	}
	if (visited && !d.skip) {
		const GtVcf *v = a.vcf + i;
		const uint32_t dp = d.dp1 + d.dinf;
		const bool het = is_het(d.gt), pass = d.flt == 0;
		int a0, a1;
		alleles_of(d.gt, a0, a1);
		// "variant": every written site; as compiled the reference files a homozygous reference call under multi (bsgpu.h)
		const int var_slot = (a0 == a1 && a0 == d.rfix) ? 2 : 0;
		atomicAdd(&h[H_MISC + var_slot], 1u);
		if (pass) atomicAdd(&h[H_MISC + var_slot + 1], 1u);
		atomicAdd(&h[H_QUAL + 256 + d.phred], 1u);
		atomicAdd(&h[H_QUAL + d.phred], 1u);
		if (dp < (uint32_t)BSGPU_STATS_COV_MAX) add_u64(&st->cov[dp].var, 1u);
		atomicAdd(&h[H_QD + 2 * min(d.qd, 255u) + het], 1u);
		{
			const int mq = v->mq;
			if (mq >= 0 && mq < 256) atomicAdd(&h[H_MQ + 2 * mq + het], 1u);
		}
		if (d.fs >= 0 && d.fs < kFsShared) atomicAdd(&h[H_FS + 2 * d.fs + het], 1u);
		else if (d.fs >= 0 && d.fs < BSGPU_STATS_FS_MAX) add_u64(&st->fs_stats[d.fs][het], 1u);
		else add_u64(&st->fs_overflow, 1u);
		atomicAdd(&h[H_FLT + 32 * het + (d.flt & 31u)], 1u);
		if (d.rs_found) {
			atomicAdd(&h[H_MISC + 4], 1u); atomicAdd(&h[H_MISC + 6], 1u);
			if (pass) { atomicAdd(&h[H_MISC + 5], 1u); atomicAdd(&h[H_MISC + 7], 1u); }
		}
		if ((g[2] == 5 && g[3] == 8) || (g[2] == 8 && g[1] == 5)) {      // a called CpG (:229-232, 444)
			bool ref_cpg, ok = false;
			uint32_t ca = 0, cb = 0;
			if (d.gt == 4) {                       // '+' strand
				ref_cpg = d.rc[2] == 2 && d.rc[3] == 3;
				ca = (uint32_t)v->counts[5]; cb = (uint32_t)v->counts[7];
				ok = true;
			} else {                               // '-' strand (gt == 7): pairs with the '+' site before it when that one was counted
				ref_cpg = d.rc[1] == 2 && d.rc[2] == 3;
				// (a launch over a later chunk of the window cannot read the record of the site before its first one any more:
				// the launch before left what is needed in stats_carry)
				uint32_t prev;
				if (i > a.i0) {
					int gp[5];
					SiteDer dpv;
					window_calls(a, i - 1, first, last, gp);
					prev = derive_site(a, i - 1, first, last, gp, dpv) && !dpv.skip ? (dpv.flt ? 2u : 1u) : 0u;
				} else prev = a.i0 ? a.stats_carry[a.carry_flip ^ 1u] : 0u;
				if (prev) {
					atomicAdd(&h[H_MISC + (ref_cpg ? 8 : 10)], 1u);
					if (!(prev == 2u || d.flt)) atomicAdd(&h[H_MISC + (ref_cpg ? 9 : 11)], 1u);
				}
				ca = (uint32_t)v->counts[6]; cb = (uint32_t)v->counts[4];
				ok = true;
			}
			if (ok) {
				atomicAdd(&h[H_QUAL + (ref_cpg ? 512 : 768) + d.phred], 1u);
				if (dp < (uint32_t)BSGPU_STATS_COV_MAX) add_u64(&st->cov[dp].CpG[ref_cpg ? 0 : 1], 1u);
				if (d.dinf < (uint32_t)BSGPU_STATS_COV_MAX) add_u64(&st->cov[d.dinf].CpG_inf[ref_cpg ? 0 : 1], 1u); else add_u64(&st->cov_overflow, 1u);
				if (ca + cb) {
					const uint32_t e = atomicAdd(&cg_n, 1u);
					cg_a[e] = ca; cg_b[e] = cb; cg_w[e] = (ref_cpg ? 0u : 2u) | (pass ? 4u : 0u);
				}
			}
		}
		// mutation spectrum (mut_type, :47-58): the call's non-reference allele(s) against the reference base
		if (d.rfix) {
			int mut = -1;
			const int alt = a0 != d.rfix ? (a1 != d.rfix && a1 != a0 ? -1 : a0) : (a1 != d.rfix ? a1 : -1);
			// the table has an entry where exactly one allele differs from the reference base, or both are the same other base
			// stats_mut order (include/bs_call.h:46): reference base major, the three other bases in base order
			if (alt > 0) mut = 3 * (d.rfix - 1) + (alt - 1) - (alt > d.rfix ? 1 : 0);
			if (mut >= 0) {
				atomicAdd(&h[H_MUT + 2 * mut], 1u);
				if (pass) atomicAdd(&h[H_MUT + 2 * mut + 1], 1u);
				if (d.rs_found) { atomicAdd(&h[H_DBMUT + 2 * mut], 1u); if (pass) atomicAdd(&h[H_DBMUT + 2 * mut + 1], 1u); }
			}
		}
	}
	if (i + 1 == a.i1 && a.stats_carry) a.stats_carry[a.carry_flip] = visited && !d.skip ? (d.flt ? 2u : 1u) : 0u;
	__syncthreads();
	// methylation posterior of every pooled CpG strand (:484-514): one warp per entry, lanes over the 101 levels
	const uint32_t ncg = cg_n;
	if (ncg) {
		if (tid < 100) logp[tid] = log(0.01 * (double)(tid + 1));
		__syncthreads();
		for (uint32_t e = tid >> 5; e < ncg; e += kWrThreads / 32) {
			const uint32_t ca = cg_a[e], cb = cg_b[e], w = cg_w[e];
			const double konst = lgamma((double)(ca + cb + 1) + 1.0) - lgamma((double)ca + 1.0) - lgamma((double)cb + 1.0);
			const double da = (double)ca, db = (double)cb;
			double m[4], sum = 0.0;
#pragma unroll
			for (int q = 0; q < 4; q++) {
				const int k = lane + 32 * q;
				double z = 0.0;
				if (k == 0) z = ca ? 0.0 : exp(konst);
				else if (k == 100) z = cb ? 0.0 : exp(konst);
				else if (k < 100) z = exp(konst + logp[k - 1] * da + logp[99 - k] * db);
				m[q] = z; sum += z;
			}
#pragma unroll
			for (int s = 16; s; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
#pragma unroll
			for (int q = 0; q < 4; q++) {
				const int k = lane + 32 * q;
				if (k <= 100 && m[q] != 0.0) {
					const double z = m[q] / sum;
					atomicAdd(&meth[w & 2u][k], z);
					if (w & 4u) atomicAdd(&meth[(w & 2u) + 1][k], z);
				}
			}
		}
		__syncthreads();
		for (int k = tid; k < 404; k += kWrThreads) {
			const double z = (&meth[0][0])[k];
			if (z != 0.0) {
				const int row = k / 101, col = k - row * 101;
				atomicAdd(row < 2 ? &st->CpG_ref_meth[row][col] : &st->CpG_nonref_meth[row - 2][col], z);
			}
		}
	}
	// flush the histograms
	for (int k = tid; k < H_END; k += kWrThreads) {
		const uint32_t c = h[k];
		if (!c) continue;
		uint64_t *dst;
		if (k < H_MUT) {
			static_assert(offsetof(bsgpu_site_stats, multi) == 16 && offsetof(bsgpu_site_stats, CpG_nonref) == 80, "bsgpu_site_stats layout");
			// misc slots: snps 0-1, multi 2-3, dbSNP_sites 4-5, dbSNP_var 6-7, CpG_ref 8-9, CpG_nonref 10-11 = the first twelve words
			dst = (uint64_t *)st + k;
			if (a.ctg_stats) add_u64((uint64_t *)a.ctg_stats + k, c);
		} else if (k < H_DBMUT) dst = &st->mut_counts[0][0] + (k - H_MUT);
		else if (k < H_QUAL) dst = &st->dbSNP_mut_counts[0][0] + (k - H_DBMUT);
		else if (k < H_FLT) dst = &st->qual[0][0] + (k - H_QUAL);
		else if (k < H_QD) dst = &st->filter_counts[0][0] + (k - H_FLT);
		else if (k < H_MQ) dst = &st->qd_stats[0][0] + (k - H_QD);
		else if (k < H_FS) dst = &st->mq_stats[0][0] + (k - H_MQ);
		else dst = &st->fs_stats[0][0] + (k - H_FS);
		add_u64(dst, c);
	}
}

}  // namespace

// per-site scratch of a window (calls, lengths) and per-CTA scratch of one launch over `cnt` sites
size_t bcf_site_scratch_bytes(uint32_t sz) { return (((size_t)sz + 15) & ~(size_t)15) + (((size_t)sz * 2 + 15) & ~(size_t)15); }
size_t bcf_cta_scratch_bytes(uint32_t cnt) {
	const size_t nctas = ((size_t)cnt + kWrThreads - 1) / kWrThreads;
	return nctas * 8 + ((nctas * 4 + 15) & ~(size_t)15) + 16;
}

constexpr size_t kFusedSmem = kWrThreads * sizeof(GtVcf) + kFusedStage + 16;
cudaError_t configure_writer() {
	cudaError_t e = cudaFuncSetAttribute(k_bcf_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem);
	if (e != cudaSuccess) return e;
	return cudaFuncSetAttribute(k_bcf_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes + 16);
}

static WrArgs writer_args(const BcfJob &j, uint32_t i0, uint32_t cnt) {
	WrArgs a;
	a.vcf = (const GtVcf *)j.d_vcf; a.ref = (const uint8_t *)j.d_ref; a.x = j.x; a.sz = j.sz; a.i0 = i0; a.i1 = i0 + cnt;
	a.blocks = (const uint2 *)j.d_blocks; a.nblocks = j.nblocks;
	for (int k = 0; k < 16; k++) a.ids[k] = j.p.ids[k];
	a.rid = j.p.rid; a.ctg_end = j.p.ctg_end; a.all_positions = j.p.all_positions; a.dc = j.dc;
	a.calls = (uint8_t *)j.site_scratch;
	a.len = (uint16_t *)((uint8_t *)j.site_scratch + (((size_t)j.sz + 15) & ~(size_t)15));
	a.cta_bytes = nullptr; a.cta_recs = nullptr; a.totals = nullptr; a.out = nullptr; a.out_cap = 0;
	a.guard = j.guard;
	a.db = j.db; a.reg_start = j.reg_start; a.reg_stop = j.reg_stop;
	a.stats = (bsgpu_site_stats *)j.stats;
	a.ctg_stats = j.ctg_stats && j.p.rid >= 0 && (uint32_t)j.p.rid < j.n_ctg ? (bsgpu_ctg_site_stats *)j.ctg_stats + j.p.rid : nullptr;
	a.gc = j.gc; a.gc_bins = j.gc_bins; a.gc_start = j.gc_start; a.stats_carry = j.stats_carry; a.carry_flip = j.stats_carry_flip & 1u;
	return a;
}

// the writer's call of sites [i0, i0 + cnt): must have run for a site and its two neighbours either side before the
// site's record is built
cudaError_t launch_bcf_calls(const BcfJob &j, uint32_t i0, uint32_t cnt, cudaStream_t stream, int *launches) {
	if (!cnt) return cudaSuccess;
	k_bcf_calls<<<(cnt + 255) / 256, 256, 0, stream>>>(writer_args(j, i0, cnt));
	__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
	return cudaGetLastError();
}

// records of sites [i0, i0 + cnt) -> d_out (back to back, in site order); d_totals: bytes, records, oversized records
cudaError_t launch_bcf_records(const BcfJob &j, uint32_t i0, uint32_t cnt, void *cta_scratch, void *d_out, size_t out_cap,
		unsigned long long *d_totals, cudaStream_t stream, int *launches) {
	cudaError_t e = cudaMemsetAsync(d_totals, 0, 3 * sizeof(unsigned long long), stream);
	if (e != cudaSuccess || !cnt) return e;
	const uint32_t nctas = (cnt + kWrThreads - 1) / kWrThreads;
	WrArgs a = writer_args(j, i0, cnt);
	a.cta_bytes = (unsigned long long *)cta_scratch;
	a.cta_recs = (uint32_t *)((uint8_t *)cta_scratch + (size_t)nctas * 8);
	a.totals = d_totals; a.out = (uint8_t *)d_out; a.out_cap = out_cap;
	// Default: three kernels (sizes, offsets, records).  BSGPU_WRITER=fused: the single pass (k_bcf_fused) -- it moves half the
	// bytes (every record read once, coalesced) and is still slower, 1.32 against 1.19 ms per 8 M sites: neither is bound by
	// memory traffic but by the latency of the ~1500 dependent instructions of a record's serialisation with few warps at work,
	// and staging the records in shared memory leaves the fused kernel fewer of them (profiles/r02d_writer_ab.txt).
	const char *wenv = getenv("BSGPU_WRITER");
	const bool split = !(wenv && !strcmp(wenv, "fused"));
	if (split || (((uintptr_t)(a.vcf + i0)) & 15u)) {
		k_bcf_measure<<<nctas, kWrThreads, 0, stream>>>(a);
		k_bcf_offsets<<<1, 1024, 0, stream>>>(a.cta_bytes, a.cta_recs, nctas, d_totals);
		k_bcf_emit<<<nctas, kWrThreads, kStageBytes + 16, stream>>>(a);
		__atomic_fetch_add(launches, 3, __ATOMIC_RELAXED);
	} else {
		// tile states + the ticket counter live where the split kernels keep their per-CTA totals
		e = cudaMemsetAsync(cta_scratch, 0, (size_t)nctas * 8 + 16, stream);
		if (e != cudaSuccess) return e;
		k_bcf_fused<<<nctas, kWrThreads, kFusedSmem, stream>>>(a, a.cta_bytes, (unsigned int *)a.cta_recs, nctas);
		__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
	}
	if (a.stats) {                  // the calls of the sites and of their neighbours exist now
		k_bcf_stats<<<nctas, kWrThreads, 0, stream>>>(a);
		__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
	}
	return cudaGetLastError();
}

}  // namespace bsgpu
