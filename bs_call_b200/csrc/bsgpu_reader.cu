// bsgpu_reader.cu -- the reader side of the path: BAM alignment records -> templates grouped into blocks.
//
//   k_decode_records      one lane per record for the scalar part (flag / MAPQ / insert-size / orientation filters,
//                         positions, CIGAR -> event list, bisulfite strand tag), one warp per record for the bytes
//                         (4-bit sequence + qualities -> packed bytes)
//                         (what get_next_align_details does per record, src/input_sam.c:222-312)
//   frame_records         host: walks the block_size chain of the record stream (what sam_read1 does per call)
//   BlockBuilder          host: mate pairing by read name, positional duplicate removal, block cutting
//                         (read_input, src/get_template_vector.c:49-389).  Order dependent over a coordinate-sorted
//                         stream, so it runs on the host over the 52-byte descriptors the kernel produced; the
//                         decoded reads themselves stay in HBM and templates refer to them by offset.
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <functional>
#include <deque>
#include <condition_variable>
#include <thread>
#include <vector>
#include <unordered_map>
#include <string>
#include <cuda_runtime.h>
#include "bsgpu.h"
#include "bsgpu_launch.h"

namespace bsgpu {

namespace {

enum : uint32_t { F_PAIRED = 1, F_PROPER = 2, F_UNMAP = 4, F_MUNMAP = 8, F_REVERSE = 16, F_READ2 = 128, F_SECONDARY = 256,
	F_QCFAIL = 512, F_DUP = 1024, F_SUPP = 2048 };
// gt_filter_reason (include/bs_call.h:50)
enum : uint32_t { FLT_NONE = 0, FLT_UNMAPPED, FLT_QC, FLT_SECONDARY, FLT_MATE_UNMAPPED, FLT_DUPLICATE, FLT_NOPOS, FLT_NOMATEPOS,
	FLT_MISMATCH_CHR, FLT_ORIENTATION, FLT_INSERT_SIZE, FLT_NOSEQ, FLT_MAPQ, FLT_NOT_CORRECTLY_ALIGNED };

// records are not aligned in the stream: fields are assembled from bytes
__host__ __device__ __forceinline__ uint32_t ld_u16(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }
__host__ __device__ __forceinline__ uint32_t ld_u32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

// bisulfite strand from the aligner's tag (src/input_sam.c:144-220): GEM XB:A, Bowtie/Bismark XG:Z, Novoalign ZB:Z,
// BSMAP ZS:Z, bwa-meth YD:Z.  The walk keeps the reference's treatment of malformed tags (unknown type letters
// consume nothing, array sizes multiply in 32 bits).
__device__ uint32_t strand_from_tags(const uint8_t *s, const uint8_t *end) {
	enum { UNK, GEM, BOWTIE, NOVO, BSMAP, BWAMETH };
	uint32_t strand = 0;
	bool ok = true;
	while (ok && s + 4 <= end) {
		int al = UNK;
		const uint8_t t0 = s[0], t1 = s[1];
		if (t0 == 'Z') al = t1 == 'B' ? NOVO : (t1 == 'S' ? BSMAP : UNK);
		else if (t0 == 'X') al = t1 == 'G' ? BOWTIE : (t1 == 'B' ? GEM : UNK);
		else if (t0 == 'Y' && t1 == 'D') al = BWAMETH;
		const uint8_t type = s[2];
		s += 3;
		switch (type) {
		case 'A':
			if (al == GEM) { if (*s == 'C') strand = 1; else if (*s == 'G') strand = 2; }
			s++;
			break;
		case 'C': case 'c': s++; break;
		case 'S': case 's': if (s + 2 <= end) s += 2; else ok = false; break;
		case 'I': case 'i': case 'f': if (s + 4 <= end) s += 4; else ok = false; break;
		case 'd': if (s + 8 <= end) s += 8; else ok = false; break;
		case 'Z': {
			const uint8_t c = *s;
			if (al == BOWTIE || al == NOVO) { if (c == 'C') strand = 1; else if (c == 'G') strand = 2; }
			else if (al == BSMAP) { if (c == '+') strand = 1; else if (c == '-') strand = 2; }
			else if (al == BWAMETH) { if (c == 'f') strand = 1; else if (c == 'r') strand = 2; }
		}   // a string either way: skip to its terminator
		case 'H':
			while (s < end && *s) s++;
			if (s < end) s++; else ok = false;
			break;
		case 'B': {
			const uint8_t sub = *s++;
			uint32_t sz;
			switch (sub) {
			case 'A': case 'C': case 'c': sz = 1; break;
			case 's': case 'S': sz = 2; break;
			case 'i': case 'I': case 'f': sz = 4; break;
			case 'd': sz = 8; break;
			case 'Z': case 'H': case 'B': sz = sub; break;       // the reference's table holds the letter itself
			default: sz = 0;
			}
			if (s + 4 <= end && sz != 0) {
				const uint32_t bytes = ld_u32(s) * sz;
				s += 4;
				if (s + bytes <= end) s += bytes; else ok = false;
			} else ok = false;
			break; }
		default: break;
		}
	}
	return strand;
}

constexpr int kDecodeWarps = 8;

// ---- QNAME join on the device (the hash find / insert of read_input, src/get_template_vector.c:131-385, turned into a perfect
// hash): every kept, paired record is entered into an open-addressing table under the 64-bit hash of its read name
// (k_decode_records), then k_name_ids gives each such record the index of the FIRST record of the batch that carries exactly
// the same name (names compared byte by byte).  The host's block builder indexes its waiting-mate entries by that id and
// never hashes or even reads a name.  A slot keeps up to five record indices; more records under one hash (a name used by
// more than five kept records, or colliding hashes) raise the overflow flag and the host computes the ids itself.
struct NameSlot { unsigned long long key; uint32_t cnt; uint32_t idx[5]; };
static_assert(sizeof(NameSlot) == 32, "name table slot");

__host__ __device__ __forceinline__ unsigned long long name_hash(const uint8_t *s, uint32_t n) {
	unsigned long long h = 0x9e3779b97f4a7c15ull ^ n;
	for (uint32_t i = 0; i < n; i++) { h = (h ^ s[i]) * 0x100000001b3ull; }
	h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
	return h | 1ull;                      // 0 marks an empty slot
}

__device__ __forceinline__ void name_insert(NameSlot *tab, uint32_t mask, unsigned long long key, uint32_t rec, uint32_t *overflow) {
	for (uint32_t s = (uint32_t)(key >> 17) & mask;; s = (s + 1) & mask) {
		const unsigned long long old = atomicCAS(&tab[s].key, 0ull, key);
		if (old == 0ull || old == key) {
			const uint32_t k = atomicAdd(&tab[s].cnt, 1u);
			if (k < 5) tab[s].idx[k] = rec; else atomicExch(overflow, 1u);
			return;
		}
	}
}

// A warp takes 32 consecutive records.  Phase 1, one record per LANE: the fixed fields, the filter cascade, the CIGAR
// walk, the tag walk and the descriptor -- short scalar work whose cost is shared by 32 records.  Phase 2, one record
// per TRIP with all lanes: the 4-bit sequence and the qualities become packed bytes, lanes striding over the bases
// (coalesced byte stores).
__global__ void __launch_bounds__(kDecodeWarps * 32)
k_decode_records(const uint8_t *__restrict__ bam, const uint64_t *__restrict__ rec_off, const uint32_t *__restrict__ read_off,
		const uint32_t *__restrict__ mm_off, size_t nrec, uint32_t mapq_thresh, uint32_t max_tlen, int keep_unmatched, int ignore_dup,
		bsgpu_record *__restrict__ out, uint8_t *__restrict__ bases, bsgpu_misms *__restrict__ misms, uint4 *__restrict__ keys,
		NameSlot *__restrict__ names, uint32_t name_mask, uint32_t rec_base, uint32_t *__restrict__ name_overflow) {
	const size_t w0 = ((size_t)blockIdx.x * kDecodeWarps + (threadIdx.x >> 5)) * 32;
	const int lane = threadIdx.x & 31;
	if (w0 >= nrec) return;
	const size_t w = w0 + lane;
	// ---- phase 1: my record
	const uint8_t *seq = nullptr;
	int32_t l_qseq = 0;
	uint32_t boff = 0;
	bool decode = false;
	if (w < nrec) {
		const uint8_t *rec = bam + rec_off[w];
		const uint32_t block_size = ld_u32(rec);
		const uint8_t *p = rec + 4;
		const int32_t tid = (int32_t)ld_u32(p), pos = (int32_t)ld_u32(p + 4), mtid = (int32_t)ld_u32(p + 20), mpos = (int32_t)ld_u32(p + 24),
			isize = (int32_t)ld_u32(p + 28);
		l_qseq = (int32_t)ld_u32(p + 16);
		const uint32_t l_qname = p[8], mapq = p[9], n_cigar = ld_u16(p + 12), flag = ld_u16(p + 14);

		// filters and positions (src/input_sam.c:231-300)
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
		bool mis_matched = (flag & (F_MUNMAP | F_PROPER)) != F_PROPER;
		const bool reverse = flag & F_REVERSE, second = flag & F_READ2;
		const bool mult_seg = (flag & (F_PAIRED | F_MUNMAP)) == F_PAIRED;
		uint32_t fwd = reverse ? (uint32_t)(mpos + 1) : (uint32_t)(pos + 1);
		uint32_t rev = reverse ? (uint32_t)(pos + 1) : (uint32_t)(mpos + 1);
		if (mapq < mapq_thresh && !flt) flt = FLT_MAPQ;
		if (mult_seg) {
			if (tid != mtid) { if (!flt) flt = FLT_MISMATCH_CHR; if (keep_unmatched) mis_matched = true; }
			if (!flt) {
				const uint64_t is = (uint64_t)(isize < 0 ? -(int64_t)isize : (int64_t)isize);
				if (is > (uint64_t)max_tlen) { flt = FLT_INSERT_SIZE; if (keep_unmatched) mis_matched = true; }
			}
			if (reverse ? pos < mpos : pos > mpos) { if (!flt) flt = FLT_ORIENTATION; if (keep_unmatched) mis_matched = true; }
			if (mis_matched) { if (reverse) fwd = 0; else rev = 0; }
		}
		const bool dropped = flt && !(keep_unmatched && (flt == FLT_INSERT_SIZE || flt == FLT_MISMATCH_CHR || flt == FLT_ORIENTATION));

		const uint8_t *cigar = p + 32 + l_qname;
		seq = cigar + 4 * (size_t)n_cigar;
		const uint8_t *qual = seq + (((size_t)l_qseq + 1) >> 1);
		const uint8_t *aux = qual + l_qseq, *end = p + block_size;
		boff = read_off[w];
		const uint32_t moff = mm_off[w];
		decode = !dropped;

		bsgpu_record r;
		memset(&r, 0, sizeof(r));
		r.ret = dropped ? 1 : 0;
		r.filtered = flt;
		r.forward_position = fwd;
		r.reverse_position = rev;
		r.alignment_flag = (!mult_seg || mis_matched) ? flag & ~F_PAIRED : flag;
		r.reverse = reverse;
		r.orientation = ((second && reverse) || !(second || reverse)) ? 0 : 1;
		r.mapq = (uint8_t)mapq;
		r.tid = tid;
		if (!dropped) {
			// CIGAR -> events (src/input_sam.c:90-136); note CIGAR I -> DEL, D -> INS, P treated like S
			uint32_t position = 0, span = 0, n = 0;
			for (uint32_t i = 0; i < n_cigar; i++) {
				const uint32_t c = ld_u32(cigar + 4 * (size_t)i), len = c >> 4, op = c & 15u;
				uint32_t type = 0;
				switch (op) {
				case 0: case 7: case 8: position += len; span += len; break;
				case 4: case 6: type = 3; break;
				case 1: type = 2; break;
				case 2: type = 1; break;
				default: break;
				}
				if (type) {
					bsgpu_misms m;
					m.type = type; m.position = position; m.size = len;
					misms[moff + n++] = m;
					if (type == 1) span += len; else position += len;
				}
			}
			r.align_length = position;
			r.reference_span = span;
			r.mm_off = moff;
			r.mm_n = n;
			r.read_off = boff;
			r.read_len = (uint32_t)l_qseq;
			r.bs_strand = (uint8_t)strand_from_tags(aux, end);
			// the qualities get_al_qual looks at: byte k of mate k (src/al_utils.c:26)
			for (int k = 0; k < 2 && k < l_qseq; k++) {
				const uint32_t byte = seq[k >> 1], nib = k ? (byte & 15u) : (byte >> 4);
				const bool known = nib == 1 || nib == 2 || nib == 4 || nib == 8;
				r.q01[k] = known ? (uint8_t)min((uint32_t)qual[k], (uint32_t)BSGPU_MAX_QUAL) : (uint8_t)0;
			}
		}
		out[w] = r;
		if (names && !dropped && (r.alignment_flag & F_PAIRED)) name_insert(names, name_mask, name_hash(p + 32, l_qname), rec_base + (uint32_t)w, name_overflow);
		if (keys) {
			// what the scan for certain block starts needs of a record (certain_block_starts): its contig (-1: dropped), the
			// smaller of its positions if it is inserted by its flags alone (else 0), the furthest end it can give its block
			uint32_t minnz = 0;
			if (!dropped) {
				bool insert = true;
				if ((r.alignment_flag & F_PAIRED) && fwd > 0 && rev > 0) insert = fwd == rev ? false : (reverse ? fwd > rev : fwd < rev);
				if (insert) minnz = fwd == 0 ? rev : rev == 0 ? fwd : min(fwd, rev);
			}
			const unsigned long long e1 = (unsigned long long)(reverse ? rev : fwd) + r.reference_span, e2 = (unsigned long long)(fwd > 0 ? fwd : rev) + r.align_length;
			const unsigned long long e = max(e1, e2);
			keys[w] = make_uint4(dropped ? 0xffffffffu : (uint32_t)tid, minnz, (uint32_t)min(e, 0xffffffffull), 0u);
		}
	}
	// ---- phase 2: sequence and qualities (src/input_sam.c:61-88), record by record with the whole warp
	const uint32_t todo = __ballot_sync(0xffffffffu, decode);
	const unsigned long long seq_bits = (unsigned long long)(uintptr_t)seq;
	for (uint32_t m = todo; m; m &= m - 1) {
		const int j = __ffs(m) - 1;
		const uint8_t *sq = (const uint8_t *)(uintptr_t)__shfl_sync(0xffffffffu, seq_bits, j);
		const int32_t ls = __shfl_sync(0xffffffffu, l_qseq, j);
		const uint8_t *ql = sq + (((size_t)ls + 1) >> 1);
		uint8_t *dst = bases + __shfl_sync(0xffffffffu, boff, j);
		for (int32_t k = lane; k < ls; k += 32) {
			const uint32_t byte = sq[k >> 1];
			const uint32_t nib = (k & 1) ? (byte & 15u) : (byte >> 4);
			// 1 2 4 8 -> A C G T; everything else is N and becomes the zero byte
			const uint32_t base = nib == 1 ? 0u : (nib == 2 ? 1u : (nib == 4 ? 2u : 3u));
			const bool known = nib == 1 || nib == 2 || nib == 4 || nib == 8;
			const uint32_t q = min((uint32_t)ql[k], (uint32_t)BSGPU_MAX_QUAL);
			dst[k] = known ? (uint8_t)(base | q << 2) : (uint8_t)0;
		}
	}
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// device: certain block starts (see certain_block_starts below for what they are) as a bit per record.
// The host formulation is a running maximum that is reset at every contig change: a segmented max-scan over tiles of 4096
// keys (four consecutive records per thread), one CTA per tile: a take-right scan gives every thread the contig
// of the last kept record before its own, which decides where segments begin; a segmented max-scan of the records' ends
// gives it the running end before its records; with those it replays its four records exactly as the host loop would.
// The state at the end of a chunk is carried to the next launch through `carry` ({last contig, running end}).
// ---------------------------------------------------------------------------------------------------------------
namespace {
constexpr uint32_t kNoTid = 0xffffffffu;
struct SegMax { uint32_t flag; uint32_t m; };      // flag: a segment begins inside; m: maximum since the last begin
__device__ __forceinline__ SegMax seg_combine(SegMax a, SegMax b) { return SegMax{a.flag | b.flag, b.flag ? b.m : max(a.m, b.m)}; }

// Three phases so that the scan fills the device instead of one SM (the first version was ONE CTA walking the chunk: 56 ms per
// 10 M records in the reader stage against 13 ms for four host threads):
//   k_certain_tile<false>  every 4096-key tile summarises itself as if the stream simply continued into it: first / last kept
//                          contig, "a contig change inside", running end since the last change (or since the tile began);
//   k_certain_combine      one CTA chains the summaries from the carried state: the state every tile starts from;
//   k_certain_tile<true>   every tile replays its records from that state and writes its 128 mask words.
// tests/test_cpu_reader.py emulates the three phases against the sequential scan; BSGPU_CHECK_SCAN=1 compares on the device.
struct TileAgg { uint32_t first, last, flag, m; };

template <bool FINAL>
__global__ void __launch_bounds__(1024) k_certain_tile(const uint4 *__restrict__ keys, uint32_t n, const uint2 *__restrict__ tile_in,
		TileAgg *__restrict__ tile_agg, uint32_t *__restrict__ mask) {
	__shared__ uint32_t w_tid[32], w_first[32];
	__shared__ SegMax w_seg[32];
	const int tid_x = threadIdx.x, lane = tid_x & 31, wid = tid_x >> 5;
	const uint32_t i0 = blockIdx.x * 4096u + 4u * (uint32_t)tid_x;
	uint4 k[4];
#pragma unroll
	for (int e = 0; e < 4; e++) k[e] = i0 + e < n ? keys[i0 + e] : make_uint4(kNoTid, 0, 0, 0);
	// ---- contig of the last kept record before my four: take-right scan; first kept contig of the tile: take-left
	uint32_t mine = kNoTid, mine_first = kNoTid;
#pragma unroll
	for (int e = 0; e < 4; e++) if (k[e].x != kNoTid) { if (mine_first == kNoTid) mine_first = k[e].x; mine = k[e].x; }
	uint32_t inc = mine;
	for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d && inc == kNoTid) inc = o; }
	if (lane == 31) w_tid[wid] = inc;
	{
		const uint32_t have = __ballot_sync(0xffffffffu, mine_first != kNoTid);
		const uint32_t f = __shfl_sync(0xffffffffu, mine_first, have ? __ffs(have) - 1 : 0);
		if (lane == 0) w_first[wid] = have ? f : kNoTid;
	}
	__syncthreads();
	uint32_t tile_first = kNoTid, tile_last = kNoTid;
	for (int w = 0; w < 32 && tile_first == kNoTid; w++) tile_first = w_first[w];
	for (int w = 31; w >= 0 && tile_last == kNoTid; w--) tile_last = w_tid[w];
	// the state the tile starts from: phase 3 knows it; phase 1 pretends the stream continues (its own first contig, no end yet)
	const uint32_t c_tid = FINAL ? tile_in[blockIdx.x].x : tile_first, c_m = FINAL ? tile_in[blockIdx.x].y : 0u;
	uint32_t prev = __shfl_up_sync(0xffffffffu, inc, 1);
	if (lane == 0) prev = kNoTid;
	if (prev == kNoTid) { for (int w = wid - 1; w >= 0 && prev == kNoTid; w--) prev = w_tid[w]; }
	if (prev == kNoTid) prev = c_tid;
	// ---- my four records: where segments begin, and the running end since the last begin
	SegMax agg{0u, 0u};
	uint32_t pt = prev;
#pragma unroll
	for (int e = 0; e < 4; e++) {
		if (k[e].x == kNoTid) continue;
		const uint32_t f = k[e].x != pt;
		agg = seg_combine(agg, SegMax{f, k[e].z});
		pt = k[e].x;
	}
	SegMax sinc = agg;
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t of = __shfl_up_sync(0xffffffffu, sinc.flag, d), om = __shfl_up_sync(0xffffffffu, sinc.m, d);
		if (lane >= d) sinc = seg_combine(SegMax{of, om}, sinc);
	}
	if (lane == 31) w_seg[wid] = sinc;
	__syncthreads();
	SegMax before{0u, c_m};                                   // everything before my records, the incoming state included
	for (int w = 0; w < wid; w++) before = seg_combine(before, w_seg[w]);
	{
		const uint32_t of = __shfl_up_sync(0xffffffffu, sinc.flag, 1), om = __shfl_up_sync(0xffffffffu, sinc.m, 1);
		if (lane) before = seg_combine(before, SegMax{of, om});
	}
	if (!FINAL) {
		if (tid_x == 1023) {
			const SegMax total = seg_combine(before, agg);
			tile_agg[blockIdx.x] = TileAgg{tile_first, tile_last, total.flag, total.m};
		}
		return;
	}
	// ---- replay (src/get_template_vector.c:111-149 as certain_scan_seq has it)
	uint32_t m = before.m, bits = 0;
	pt = prev;
#pragma unroll
	for (int e = 0; e < 4; e++) {
		if (k[e].x == kNoTid) continue;
		if (k[e].x != pt) { bits |= 1u << e; m = 0; pt = k[e].x; }
		else if (k[e].y && (unsigned long long)k[e].y > (unsigned long long)m + 1) bits |= 1u << e;
		m = max(m, k[e].z);
	}
	// eight threads make one word of the mask
	uint32_t word = bits << (4 * (lane & 7));
	word |= __shfl_xor_sync(0xffffffffu, word, 1);
	word |= __shfl_xor_sync(0xffffffffu, word, 2);
	word |= __shfl_xor_sync(0xffffffffu, word, 4);
	if ((lane & 7) == 0 && i0 < n) mask[i0 >> 5] = word;
}

__global__ void __launch_bounds__(1024) k_certain_combine(const TileAgg *__restrict__ tile_agg, uint32_t ntiles, uint32_t *__restrict__ carry,
		uint2 *__restrict__ tile_in) {
	__shared__ TileAgg s[1024];
	__shared__ uint2 o[1024];
	__shared__ uint32_t c_tid, c_m;
	const uint32_t t = threadIdx.x;
	if (t == 0) { c_tid = carry[0]; c_m = carry[1]; }
	__syncthreads();
	for (uint32_t base = 0; base < ntiles; base += 1024) {
		if (base + t < ntiles) s[t] = tile_agg[base + t];
		__syncthreads();
		if (t == 0) {
			uint32_t tid = c_tid, m = c_m;
			const uint32_t cnt = min(1024u, ntiles - base);
			for (uint32_t i = 0; i < cnt; i++) {
				o[i] = make_uint2(tid, m);
				const TileAgg a = s[i];
				if (a.first == kNoTid) continue;             // no kept record in the tile: the state passes through
				m = (a.flag || a.first != tid) ? a.m : max(m, a.m);
				tid = a.last;
			}
			c_tid = tid; c_m = m;
		}
		__syncthreads();
		if (base + t < ntiles) tile_in[base + t] = o[t];
		__syncthreads();
	}
	if (t == 0) { carry[0] = c_tid; carry[1] = c_m; }
}
// name_id[i] = the smallest record index of the batch whose read name is that of record i (i itself when it is the first), for
// kept paired records; 0xffffffff for the others.  rec_off / rec / name_id are indexed by the batch-wide record number.
__global__ void __launch_bounds__(256)
k_name_ids(const uint8_t *__restrict__ bam, const uint64_t *__restrict__ rec_off, const bsgpu_record *__restrict__ rec, uint32_t r0, uint32_t r1,
		const NameSlot *__restrict__ tab, uint32_t mask, uint32_t *__restrict__ name_id) {
	const uint32_t i = r0 + blockIdx.x * 256 + threadIdx.x;
	if (i >= r1) return;
	uint32_t id = 0xffffffffu;
	const bsgpu_record r = rec[i];
	if (r.ret == 0 && (r.alignment_flag & F_PAIRED)) {
		const uint8_t *p = bam + rec_off[i] + 4;
		const uint32_t l = p[8];
		const unsigned long long key = name_hash(p + 32, l);
		id = i;
		for (uint32_t s = (uint32_t)(key >> 17) & mask;; s = (s + 1) & mask) {
			const unsigned long long k = tab[s].key;
			if (k == key) {
				const uint32_t cnt = min(tab[s].cnt, 5u);
				for (uint32_t c = 0; c < cnt; c++) {
					const uint32_t j = tab[s].idx[c];
					if (j >= id) continue;
					const uint8_t *q = bam + rec_off[j] + 4;
					bool same = q[8] == l;
					for (uint32_t b = 0; same && b < l; b++) same = q[32 + b] == p[32 + b];
					if (same) id = j;
				}
				break;
			}
			if (k == 0ull) break;          // cannot happen for an inserted record
		}
	}
	name_id[i] = id;
}

}  // namespace

size_t name_table_slots(size_t nrec) { size_t s = 1024; while (s < 2 * nrec) s <<= 1; return s; }
size_t name_table_bytes(size_t nrec) { return name_table_slots(nrec) * sizeof(NameSlot); }

cudaError_t launch_name_ids(const void *bam, const void *rec_off, const void *rec, uint32_t r0, uint32_t r1, const void *table, size_t slots,
		void *name_id, cudaStream_t stream, int *launches) {
	if (r1 <= r0) return cudaSuccess;
	k_name_ids<<<(r1 - r0 + 255) / 256, 256, 0, stream>>>((const uint8_t *)bam, (const uint64_t *)rec_off, (const bsgpu_record *)rec, r0, r1,
		(const NameSlot *)table, (uint32_t)(slots - 1), (uint32_t *)name_id);
	__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
	return cudaGetLastError();
}

// mask: one bit per record of the chunk (bit i of word i / 32), (n + 31) / 32 words, 4096-record tiles start on word boundaries
cudaError_t launch_certain_starts(const void *keys, uint32_t n, void *carry, void *mask, void *scratch, cudaStream_t stream, int *launches) {
	if (!n) return cudaSuccess;
	// scratch: certain_scratch_bytes(n) = per tile a 16-byte summary and the 8-byte state it starts from
	const uint32_t ntiles = (n + 4095u) / 4096u;
	TileAgg *agg = (TileAgg *)scratch;
	uint2 *tin = (uint2 *)((uint8_t *)scratch + (size_t)ntiles * sizeof(TileAgg));
	k_certain_tile<false><<<ntiles, 1024, 0, stream>>>((const uint4 *)keys, n, nullptr, agg, nullptr);
	k_certain_combine<<<1, 1024, 0, stream>>>(agg, ntiles, (uint32_t *)carry, tin);
	k_certain_tile<true><<<ntiles, 1024, 0, stream>>>((const uint4 *)keys, n, tin, nullptr, (uint32_t *)mask);
	__atomic_fetch_add(launches, 3, __ATOMIC_RELAXED);
	return cudaGetLastError();
}
size_t certain_scratch_bytes(size_t n) { return ((n + 4095) / 4096 + 1) * 24 + 16; }

cudaError_t launch_decode_records(const void *bam, const void *rec_off, const void *read_off, const void *mm_off, size_t nrec,
		uint32_t mapq_thresh, uint32_t max_tlen, int keep_unmatched, int ignore_dup, void *out, void *bases, void *misms,
		cudaStream_t stream, int *launches, void *keys, void *name_table, size_t name_slots, uint32_t rec_base, void *name_overflow) {
	if (!nrec) return cudaSuccess;
	const unsigned grid = (unsigned)((nrec + kDecodeWarps * 32 - 1) / (kDecodeWarps * 32));
	k_decode_records<<<grid, kDecodeWarps * 32, 0, stream>>>((const uint8_t *)bam, (const uint64_t *)rec_off, (const uint32_t *)read_off,
		(const uint32_t *)mm_off, nrec, mapq_thresh, max_tlen, keep_unmatched, ignore_dup, (bsgpu_record *)out, (uint8_t *)bases, (bsgpu_misms *)misms, (uint4 *)keys,
		(NameSlot *)name_table, name_table ? (uint32_t)(name_slots - 1) : 0u, rec_base, (uint32_t *)name_overflow);
	__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
	return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// host: framing
// ---------------------------------------------------------------------------------------------------------------
namespace {

// one record header at `at`: 0 ok (sizes in *bs, *lseq, *ncig), -1 malformed / truncated
inline int frame_one(const uint8_t *bam, size_t nbytes, size_t at, uint32_t *bs, uint32_t *lseq, uint32_t *ncig) {
	if (at + 36 > nbytes) return -1;
	*bs = ld_u32(bam + at);
	if (*bs < 32 || at + 4 + (size_t)*bs > nbytes) return -1;
	const uint8_t *p = bam + at + 4;
	const uint32_t l_qname = p[8];
	*ncig = ld_u16(p + 12);
	*lseq = ld_u32(p + 16);
	if ((int32_t)*lseq < 0 || 32 + (uint64_t)l_qname + 4ull * *ncig + ((uint64_t)*lseq + 1) / 2 + *lseq > *bs) return -1;
	return 0;
}

// does a record plausibly start at `at`?  Only used to GUESS where a worker starts; the stitch below never trusts it.
bool plausible_record(const uint8_t *bam, size_t nbytes, size_t at, int depth) {
	for (int d = 0; d < depth; d++) {
		if (at == nbytes) return true;
		uint32_t bs, lseq, ncig;
		if (frame_one(bam, nbytes, at, &bs, &lseq, &ncig)) return false;
		const uint8_t *p = bam + at + 4;
		const int32_t tid = (int32_t)ld_u32(p), pos = (int32_t)ld_u32(p + 4), mtid = (int32_t)ld_u32(p + 20);
		const uint32_t l_qname = p[8];
		if (tid < -1 || tid > (1 << 24) || mtid < -1 || mtid > (1 << 24) || pos < -1 || l_qname < 2 || bs > (1u << 24)) return false;
		if (p[32 + l_qname - 1] != 0) return false;
		for (uint32_t i = 0; i + 1 < l_qname && i < 12; i++) if (p[32 + i] < 33 || p[32 + i] > 126) return false;
		at += 4 + (size_t)bs;
	}
	return true;
}

}  // namespace

struct FramePiece {
	size_t start = 0, end = 0;          // first offset walked, offset after the last good record
	bool bad = false;                   // the walk stopped at a malformed record at `end`
	std::vector<uint64_t> off;
	std::vector<uint32_t> rbase, rmm;   // bases / CIGAR ops before each record, counted from the start of the piece
	uint64_t nb = 0, nm = 0;
};
struct FrameScratch { std::vector<FramePiece> piece; };      // kept by the caller between streams: no fresh pages per call

FrameScratch *frame_scratch_new() { return new FrameScratch(); }
void frame_scratch_free(FrameScratch *s) { delete s; }

namespace {

void walk_piece(const uint8_t *bam, size_t nbytes, size_t from, size_t until, FramePiece &fp) {
	size_t at = from;
	fp.start = from; fp.bad = false; fp.nb = fp.nm = 0;
	fp.off.clear(); fp.rbase.clear(); fp.rmm.clear();
	const size_t guess = (until > from ? until - from : 0) / 200 + 16;      // typical short-read records are 250-400 bytes
	fp.off.reserve(guess); fp.rbase.reserve(guess); fp.rmm.reserve(guess);
	uint64_t nb = 0, nm = 0;
	while (at < until) {
		uint32_t bs, lseq, ncig;
		if (frame_one(bam, nbytes, at, &bs, &lseq, &ncig)) { fp.bad = true; break; }
		// the chain is a dependent load per record: ask for the header a few records ahead, where it will be if sizes stay similar
		__builtin_prefetch(bam + at + 12 * (size_t)(bs + 4));
		__builtin_prefetch(bam + at + 12 * (size_t)(bs + 4) + 64);
		fp.off.push_back(at); fp.rbase.push_back((uint32_t)nb); fp.rmm.push_back((uint32_t)nm);
		nb += lseq; nm += ncig;
		at += 4 + (size_t)bs;
	}
	fp.end = at; fp.nb = nb; fp.nm = nm;
}

}  // namespace

// Framing = walking the block_size chain, a dependent load per record.  Large streams are walked by several host
// threads: worker k guesses a record boundary near k/K of the stream (a plausible header followed by plausible headers)
// and walks its piece; the pieces are then stitched from the front -- a piece is taken over only from an offset the
// chain walked so far actually lands on, otherwise the stitcher keeps walking itself until it does.  The result is the
// sequential chain whatever the guesses were.
// `framed` (may be NULL): a stream that ends inside a record is accepted and *framed receives the offset of that
// incomplete record (nbytes when the stream ends on a record boundary) -- what a reader that feeds the stream in arbitrary
// slices needs; without it a truncated stream is an error.
int frame_records(const uint8_t *bam, size_t nbytes, std::vector<uint64_t> &rec_off, std::vector<uint32_t> &read_off,
		std::vector<uint32_t> &mm_off, uint64_t *nbases, uint64_t *nmisms, FrameScratch *scratch, size_t *framed) {
	// is the record at `at` merely cut off by the end of the buffer (as opposed to malformed)?
	auto cut_off = [&](size_t at) {
		if (!framed) return false;
		if (at + 4 > nbytes) return true;
		const uint32_t bs = ld_u32(bam + at);
		return bs >= 32 && at + 4 + (size_t)bs > nbytes;
	};
	bool stopped = false;
	unsigned want = std::thread::hardware_concurrency();
	if (const char *e = getenv("BSGPU_FRAMER_THREADS")) want = (unsigned)atoi(e);
	want = std::max(1u, std::min(want, 32u));
	size_t min_bytes = 16u << 20;
	if (const char *e = getenv("BSGPU_FRAMER_MIN_BYTES")) min_bytes = (size_t)atoll(e);
	const unsigned K = nbytes >= min_bytes ? want : 1;
	FrameScratch local;
	std::vector<FramePiece> &piece = scratch ? scratch->piece : local.piece;
	if (piece.size() < K) piece.resize(K);
	if (K == 1) walk_piece(bam, nbytes, 0, nbytes, piece[0]);
	else {
		std::vector<std::thread> thr;
		for (unsigned k = 0; k < K; k++) thr.emplace_back([&, k] {
			const size_t lo = nbytes * k / K, hi = nbytes * (k + 1) / K;
			size_t from = lo;
			// BSGPU_FRAMER_BLIND (tests): start at the raw byte offset, i.e. with a wrong guess almost every time
			if (k && !getenv("BSGPU_FRAMER_BLIND")) { while (from < hi && !plausible_record(bam, nbytes, from, 4)) from++; }
			walk_piece(bam, nbytes, from, hi, piece[k]);
		});
		for (auto &t : thr) t.join();
	}
	// stitch, in two steps: decide sequentially where every piece joins the chain (cheap: a binary search per piece, a few
	// records walked by hand when a guess was wrong), then let the threads copy their pieces into place
	struct Join { size_t idx, m, dest; uint64_t db, dm; };
	struct Own { size_t dest; uint64_t off; uint32_t rb, rm; };
	std::vector<Join> join(K, Join{0, 0, 0, 0, 0});
	std::vector<Own> own;
	uint64_t nb = 0, nm = 0;
	size_t at = 0, count = 0;
	for (unsigned k = 0; k < K && !stopped; k++) {
		const FramePiece &p = piece[k];
		const size_t hi = nbytes * (k + 1) / K;
		// walk on our own until we stand on an offset the piece has (at once when the guess was right), or pass the piece
		size_t idx = 0;
		bool joined = false;
		while (at < hi) {
			const auto it = std::lower_bound(p.off.begin() + idx, p.off.end(), (uint64_t)at);
			idx = (size_t)(it - p.off.begin());
			if (it != p.off.end() && *it == at) { joined = true; break; }
			uint32_t bs, lseq, ncig;
			if (frame_one(bam, nbytes, at, &bs, &lseq, &ncig)) {
				if (cut_off(at)) { stopped = true; break; }
				return -1;
			}
			own.push_back(Own{count++, (uint64_t)at, (uint32_t)nb, (uint32_t)nm});
			nb += lseq; nm += ncig;
			at += 4 + (size_t)bs;
		}
		if (joined) {
			Join &j = join[k];
			j.idx = idx; j.m = p.off.size() - idx; j.dest = count;
			j.db = nb - p.rbase[idx]; j.dm = nm - p.rmm[idx];
			count += j.m;
			nb = p.nb + j.db; nm = p.nm + j.dm;
			at = p.end;
			if (p.bad) {
				if (cut_off(at)) stopped = true;
				else return -1;
			}
		}
	}
	// (growing a vector value-initialises the new tail; a context's vectors keep their size from stream to stream)
	rec_off.resize(count); read_off.resize(count); mm_off.resize(count);
	for (const Own &o : own) { rec_off[o.dest] = o.off; read_off[o.dest] = o.rb; mm_off[o.dest] = o.rm; }
	auto place = [&](unsigned k) {
		const FramePiece &p = piece[k];
		const Join &j = join[k];
		if (!j.m) return;
		memcpy(rec_off.data() + j.dest, p.off.data() + j.idx, j.m * sizeof(uint64_t));
		uint32_t *ro = read_off.data() + j.dest, *mo = mm_off.data() + j.dest;
		for (size_t i = 0; i < j.m; i++) { ro[i] = (uint32_t)(p.rbase[j.idx + i] + j.db); mo[i] = (uint32_t)(p.rmm[j.idx + i] + j.dm); }
	};
	if (K == 1) place(0);
	else {
		std::vector<std::thread> thr;
		for (unsigned k = 0; k < K; k++) thr.emplace_back(place, k);
		for (auto &t : thr) t.join();
	}
	if (at != nbytes && !stopped) return -1;
	if (framed) *framed = at;
	if (nb > 0xffffffffull || nm > 0xffffffffull) return -2;
	*nbases = nb;
	*nmisms = nm;
	return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// host: block builder.  Follows read_input (src/get_template_vector.c:49-389) decision by decision; a template is a
// slot in `list` holding positions and, per mate, the index of the decoded record.
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Tmpl {
	uint32_t fwd, rev, span[2];
	int64_t rec[2];
	uint8_t mapq[2], orientation, bs_strand;
};

// read_input's hash of waiting mates (uthash keyed by QNAME, src/get_template_vector.c:131-385) as an index: name_id[rec] is
// the first record of the stream that carries rec's read name (computed on the device, k_name_ids, or by host_name_ids), so
// the entry of a name is a slot of an array -- no hashing, no name bytes touched.  Entries are cleared in O(1) at block ends
// by bumping `cur`.  A builder piece starts at a certain block start; the rare name whose first record lies before the
// piece (a read name used twice on a contig) goes through a small map.
struct NameIndex {
	struct Ent { uint32_t gen = 0, ix = 0; };      // live iff gen == cur
	const uint32_t *name_id = nullptr;
	size_t base = 0;
	std::vector<Ent> near;
	std::unordered_map<uint32_t, Ent> far;
	uint32_t cur = 1;
	void init(const uint32_t *ids, size_t rbeg, size_t rend) { name_id = ids; base = rbeg; near.assign(rend - rbeg, Ent()); far.clear(); cur = 1; }
	void clear() { cur++; }
	Ent *slot(size_t rec) { const uint32_t r = name_id[rec]; return r >= base ? &near[r - base] : &far[r]; }
	Ent *find(size_t rec) { Ent *e = slot(rec); return e->gen == cur ? e : nullptr; }
	void add(size_t rec, uint32_t ix) { Ent *e = slot(rec); e->gen = cur; e->ix = ix; }
	static void kill(Ent *e) { e->gen = 0; }
};

}  // namespace

struct BlockBuilder {
	const uint8_t *bam;
	const uint64_t *rec_off;
	const bsgpu_record *rec;
	std::vector<Tmpl> list;
	std::vector<int64_t> list_name;      // per slot: record whose name keys the slot's live table entry, -1 = none ("alh_p[ix]")
	std::vector<uint32_t> list_flag;
	size_t used = 0;
	NameIndex names;
	const uint32_t *name_id = nullptr;
	std::vector<bsgpu_block> *blocks;
	bsgpu_template *out = nullptr;       // templates of this builder, in publication order (room for one per record)
	size_t nout = 0;
	uint32_t maxcap = 1;                 // largest read_len + reference span of a published mate: bounds a mate in reference coordinates
	bool before_window = false;          // a published template's absent mate lies before its block's window: the reference aborts (publish)
	uint64_t *tally = nullptr;           // --report-file: filter_cts[15] then filter_bases[15] of read_input, or none
	void count(int reason_cts, uint64_t cts, int reason_bases, uint64_t bases) { tally[reason_cts] += cts; tally[15 + reason_bases] += bases; }
	uint32_t read_len_of(const Tmpl &t, int k) const { return t.rec[k] >= 0 ? rec[t.rec[k]].read_len : 0; }

	uint32_t al_qual(const Tmpl &t) const {          // get_al_qual with its sq[k] indexing (src/al_utils.c:19-35)
		uint32_t qual = 0, n = 0;
		for (int k = 0; k < 2; k++) {
			if (t.rec[k] < 0) continue;
			const bsgpu_record &r = rec[t.rec[k]];
			const uint32_t q = r.q01[k];
			if (q != BSGPU_FLT_QUAL) { qual += q * r.read_len; n += r.read_len; }
		}
		return n ? qual / n : 0;
	}
	void put(size_t ix, const Tmpl &t, int64_t name_rec, uint32_t flag) {
		if (ix >= list.size()) { list.resize(ix + 1); list_name.resize(ix + 1); list_flag.resize(ix + 1); }
		list[ix] = t; list_name[ix] = name_rec; list_flag[ix] = flag;
		if (used <= ix) used = ix + 1;
	}
	void publish(uint32_t tid, uint32_t y) {
		if (!used) return;
		bsgpu_block b;
		memset(&b, 0, sizeof(b));
		const uint32_t first = list[0].fwd ? list[0].fwd : list[0].rev;
		b.tid = tid; b.y = y; b.x = first > 2 ? first - 2 : 1;            // src/process_template.c:24-28
		b.first_template = (uint32_t)nout; b.n_templates = (uint32_t)used;
		blocks->push_back(b);
		for (size_t i = 0; i < used; i++) {
			const Tmpl &t = list[i];
			// call_genotypes_ML asserts that the smaller position of every template lies inside the window, read or no read there
			// (src/call_genotypes.c:182-186).  With -k and -d together a lone mate is kept with the position its absent partner
			// claimed (src/get_template_vector.c:247-268), which may lie before the block: the reference aborts on such a stream.
			// (A PRESENT mate cannot start before the window of a coordinate-sorted stream.)
			for (int k = 0; k < 2; k++) { const uint32_t pk = k ? t.rev : t.fwd; if (pk && pk < b.x && t.rec[k] < 0) before_window = true; }
			bsgpu_template d;
			memset(&d, 0, sizeof(d));
			d.forward_position = t.fwd; d.reverse_position = t.rev; d.orientation = t.orientation; d.bs_strand = t.bs_strand;
			for (int k = 0; k < 2; k++) {
				d.mapq[k] = t.mapq[k];
				if (t.rec[k] < 0) continue;
				const bsgpu_record &r = rec[t.rec[k]];
				d.present[k] = 1; d.reference_span[k] = t.span[k];
				d.read_off[k] = r.read_off; d.read_len[k] = r.read_len; d.mm_off[k] = r.mm_off; d.mm_n[k] = r.mm_n;
				const uint64_t cap = (uint64_t)r.read_len + t.span[k];
				if (cap > maxcap) maxcap = (uint32_t)(cap > 0xffffffu ? 0xffffffu : cap);
			}
			out[nout++] = d;
		}
		used = 0;
	}

	// records [rbeg, nrec): rbeg must be the start of the stream or a point where read_input is certain to start a new block
	int run(size_t rbeg, size_t nrec, bool keep_unmatched, bool keep_duplicates) {
		names.init(name_id, rbeg, nrec);
		int curr_tid = -1, old_tid = -1;
		uint32_t max_pos = 0, start_pos = 0, read_idx = 0, curr_pos = 0, start_idx = 0;
		// The loop is a chain of dependent loads (record -> its name id -> the entry of that name): the entry of a record a few
		// places ahead is requested early.
		constexpr size_t kAhead = 12;
		auto look_ahead = [&](size_t j) {
			if (j >= nrec) return;
			const bsgpu_record &q = rec[j];
			if (q.ret > 0 || !(q.alignment_flag & F_PAIRED)) return;
			const uint32_t id = name_id[j];
			if (id >= rbeg) __builtin_prefetch(&names.near[id - rbeg]);
		};
		for (size_t j = rbeg; j < rbeg + kAhead; j++) look_ahead(j);
		for (size_t ri = rbeg; ri < nrec; ri++) {
			look_ahead(ri + kAhead);
			const bsgpu_record &r = rec[ri];
			if (r.ret > 0) {
				if (tally) {                 // src/get_template_vector.c:104-107: l_seq of the record that was dropped
					int32_t l_seq;
					memcpy(&l_seq, bam + rec_off[ri] + 20, 4);
					count((int)r.filtered, 1, (int)r.filtered, (uint64_t)(uint32_t)l_seq);
				}
				continue;
			}
			const int ix = r.reverse ? 1 : 0;
			Tmpl al;
			memset(&al, 0, sizeof(al));
			al.fwd = r.forward_position; al.rev = r.reverse_position; al.orientation = r.orientation; al.bs_strand = r.bs_strand;
			al.rec[0] = al.rec[1] = -1; al.rec[ix] = (int64_t)ri; al.mapq[ix] = r.mapq; al.span[ix] = r.reference_span;
			const bool paired = r.alignment_flag & F_PAIRED;
			NameIndex::Ent *waiting = nullptr;
			bool new_block = false, new_contig = false;
			if (curr_tid < 0 || curr_tid != r.tid) { new_contig = new_block = true; old_tid = curr_tid; curr_tid = r.tid; }
			bool insert = true;
			if (!new_contig) {
				if (paired && al.fwd > 0 && al.rev > 0) {
					if (al.fwd == al.rev) insert = names.find(ri) == nullptr;
					else insert = r.reverse ? al.fwd > al.rev : al.fwd < al.rev;
				}
				if (insert && start_pos > 0) {
					if (al.fwd > 0) {
						if (al.fwd > max_pos && (al.rev > max_pos || al.rev == 0) && al.fwd - max_pos > 1) new_block = true;
					} else if (al.rev > max_pos && al.rev - max_pos > 1) new_block = true;
				}
			}
			if (new_block) {
				names.clear();
				read_idx = start_idx = curr_pos = 0;
				publish((uint32_t)(new_contig ? old_tid : curr_tid), max_pos);
				max_pos = start_pos = 0;
			}
			{
				const uint32_t s0 = r.reverse ? al.rev : al.fwd, ml = s0 + r.reference_span;
				if (ml > max_pos) max_pos = ml;
				if (start_pos == 0 || start_pos > s0) start_pos = s0;
			}
			if (paired) {
				if (!insert) {
					waiting = names.find(ri);
					if (waiting) {
						Tmpl &t = list[waiting->ix];
						if (al.fwd != t.fwd || al.rev != t.rev) return -5;      // the reference asserts here (src/get_template_vector.c:239)
						t.rec[ix] = (int64_t)ri; t.mapq[ix] = r.mapq; t.span[ix] = r.reference_span;
						list_name[waiting->ix] = -1;
						NameIndex::kill(waiting);
					} else {
						if (tally) count(14, 1, 14, r.read_len);      // the partner never came (:243-246)
						bool skip = false;
						if (!keep_duplicates) { const uint32_t xx = r.reverse ? al.rev : al.fwd; if (xx >= start_pos) skip = true; }
						if (!skip && keep_unmatched) {
							const uint32_t xx = (al.fwd > 0 ? al.fwd : al.rev) + r.align_length;
							if (xx > max_pos) max_pos = xx;
							put(read_idx++, al, -1, 0);
						}
					}
				} else {
					bool skip = false;
					if (!keep_duplicates) {
						const uint32_t pos = al.fwd > 0 ? al.fwd : al.rev;
						if (pos == curr_pos) {
							for (uint32_t i = start_idx; i < read_idx; i++) {
								Tmpl &a1 = list[i];
								if (al.fwd != a1.fwd || al.rev != a1.rev || al.bs_strand != a1.bs_strand) continue;
								int maxq = 0, maxq1 = 0, kn = 0, kn1 = 0;
								for (int k = 0; k < 2; k++) {
									if (al.rec[k] >= 0 && rec[al.rec[k]].read_len > 0) { maxq += al.mapq[k]; kn++; }
									if (a1.rec[k] >= 0 && rec[a1.rec[k]].read_len > 0) { maxq1 += a1.mapq[k]; kn1++; }
								}
								maxq /= kn; maxq1 /= kn1;
								if (maxq1 < maxq || (maxq == maxq1 && al_qual(a1) < al_qual(al))) {
									// the newcomer takes the slot; the slot's table entry is re-keyed to the newcomer's name
									NameIndex::Ent *h = names.find(ri);
									if (h && list_name[i] >= 0) return -4;            // duplicate read name (fatal in the reference)
									const bool from_slot = !h && list_name[i] >= 0;
									if (from_slot) h = names.find((size_t)list_name[i]);
									const Tmpl old = a1;
									a1 = al;
									if (h) NameIndex::kill(h);
									names.add(ri, i);
									if (from_slot) { list_name[i] = (int64_t)ri; list_flag[i] = r.alignment_flag; }
									al = old;
								}
								if (tally) {                 // the template that lost (:314-319)
									const uint32_t len1 = read_len_of(al, 0), len2 = read_len_of(al, 1);
									count(5, len1 && len2 ? 2 : 1, 5, (uint64_t)len1 + len2);
								}
								skip = true;
							}
						} else { curr_pos = pos; start_idx = read_idx; }
					}
					if (!skip) {
						if (names.find(ri)) return -4;
						names.add(ri, read_idx);
						put(read_idx, al, (int64_t)ri, r.alignment_flag);
						read_idx++;
					}
				}
			} else {
				bool skip = false;
				if (!keep_duplicates) {
					const uint32_t pos = al.fwd > 0 ? al.fwd : al.rev;
					if (pos == curr_pos) {
						for (uint32_t i = start_idx; i < read_idx; i++) {
							Tmpl &a1 = list[i];
							const bool lone = list_name[i] < 0 || (list_flag[i] & 9u) == 9u || (list_flag[i] & 9u) == 0u;
							if (al.fwd == a1.fwd && al.rev == a1.rev && al.bs_strand == a1.bs_strand && lone) {
								// mapq[0] on both sides whichever strand the reads are on (reference behaviour)
								if (a1.mapq[0] < al.mapq[0] || (a1.mapq[0] == al.mapq[0] && al_qual(a1) < al_qual(al))) { const Tmpl old = a1; a1 = al; al = old; }
								if (tally) count(5, 1, 0, read_len_of(al, ix));      // a duplicate whose bases go under gt_flt_none (:361-364)
								skip = true;
							}
						}
					} else { curr_pos = pos; start_idx = read_idx; }
				}
				if (!skip) put(read_idx++, al, -1, 0);
			}
		}
		if (curr_tid >= 0) publish((uint32_t)curr_tid, max_pos);
		return before_window ? -6 : 0;       // -6: the blocks are complete (read_input itself succeeds), the reference aborts when it calls them
	}
};

// name_id[] on the host (what k_name_ids computes on the device): used by the host-only entry point bsgpu_build_blocks and
// when the device table overflowed
void host_name_ids(const uint8_t *bam, const uint64_t *rec_off, const bsgpu_record *rec, size_t nrec, uint32_t *name_id) {
	std::unordered_map<std::string, uint32_t> first;
	first.reserve(nrec);
	for (size_t i = 0; i < nrec; i++) {
		name_id[i] = 0xffffffffu;
		if (rec[i].ret > 0 || !(rec[i].alignment_flag & F_PAIRED)) continue;
		const uint8_t *p = bam + rec_off[i] + 4;
		const auto it = first.emplace(std::string((const char *)p + 32, p[8]), (uint32_t)i);
		name_id[i] = it.first->second;
	}
}

struct CertainState { int tid = -1; uint64_t maxend = 0; };

// Records at which read_input is CERTAIN to start a new block whatever its state (src/get_template_vector.c:111-149):
// the first kept record of a contig, or a record that is inserted by its flags alone and whose positions all lie more
// than one base beyond the end of every kept record before it on the contig.  The builder's state is reset there
// (:151-207), so the stream can be cut at such records and the pieces built independently.  The scan keeps its state
// in `st`, so a stream whose descriptors arrive in chunks can be scanned chunk by chunk.  (A conservative subset of
// the block starts: the running end is never reset, and mates at equal positions -- whose insertion depends on the
// name table -- are not used.)
// What the scan needs of record i, from the 56-byte descriptor or from the 16-byte key k_decode_records writes next to it
// ({contig or -1 for a dropped record, smaller position of a record inserted by its flags alone or 0, furthest end, 0})
struct RecView {
	const bsgpu_record *rec;
	bool kept(size_t i) const { return rec[i].ret <= 0; }
	int tid(size_t i) const { return rec[i].tid; }
	uint64_t minnz(size_t i) const {
		const bsgpu_record &r = rec[i];
		const uint32_t fwd = r.forward_position, rev = r.reverse_position;
		bool insert = true;
		if ((r.alignment_flag & F_PAIRED) && fwd > 0 && rev > 0) insert = fwd == rev ? false : (r.reverse ? fwd > rev : fwd < rev);
		return !insert ? 0 : fwd == 0 ? rev : rev == 0 ? fwd : std::min(fwd, rev);
	}
	uint64_t end(size_t i) const {
		const bsgpu_record &r = rec[i];
		const uint32_t fwd = r.forward_position, rev = r.reverse_position;
		return std::max((uint64_t)(r.reverse ? rev : fwd) + r.reference_span, (uint64_t)(fwd > 0 ? fwd : rev) + r.align_length);
	}
};
struct KeyView {
	const uint32_t *key;
	bool kept(size_t i) const { return key[4 * i] != 0xffffffffu; }
	int tid(size_t i) const { return (int)key[4 * i]; }
	uint64_t minnz(size_t i) const { return key[4 * i + 1]; }
	uint64_t end(size_t i) const { return key[4 * i + 2]; }
};

template <class V>
static void certain_scan_seq(const V &v, size_t rbeg, size_t rend, CertainState *st, std::vector<size_t> &starts) {
	int tid = st->tid;
	uint64_t maxend = st->maxend;
	for (size_t i = rbeg; i < rend; i++) {
		if (!v.kept(i)) continue;
		if (v.tid(i) != tid) { tid = v.tid(i); maxend = 0; starts.push_back(i); }
		else { const uint64_t m = v.minnz(i); if (m && m > maxend + 1) starts.push_back(i); }
		const uint64_t e = v.end(i);
		if (e > maxend) maxend = e;
	}
	st->tid = tid; st->maxend = maxend;
}

// The scan is a running maximum, so it splits: every host thread scans a segment from a blank state and keeps the
// records that pass against its LOCAL running end, with the smaller of their positions; whether those before the
// segment's first contig change really pass depends on the end carried in from the segments before, which a short
// sequential pass supplies.  (Descriptors and keys lie in pinned memory the device has just written: one thread streams
// them at 6-8 GB/s.)
template <class V>
static void certain_scan(const V &v, size_t rbeg, size_t rend, CertainState *st, std::vector<size_t> &starts, unsigned max_threads) {
	unsigned want = std::thread::hardware_concurrency();
	if (const char *e = getenv("BSGPU_BUILDER_THREADS")) want = (unsigned)atoi(e);
	want = std::max(1u, std::min(want, max_threads));
	const size_t n = rend - rbeg;
	size_t min_rec = 50000;
	if (const char *e = getenv("BSGPU_BUILDER_MIN_RECORDS")) min_rec = (size_t)atoll(e);
	if (want == 1 || n < min_rec || n < want) { certain_scan_seq(v, rbeg, rend, st, starts); return; }
	struct Cand { size_t i; uint64_t minpos; };
	struct SegOut {
		bool any = false, changed = false;
		size_t head = 0;                      // first kept record: decided by the carry
		int head_tid = -1, tail_tid = -1;
		uint64_t head_end = 0;                // running end of the head's contig up to the first contig change (or the segment end)
		uint64_t tail_end = 0;                // running end of the last contig at the segment end
		std::vector<Cand> pending;            // pass locally, before the first contig change: need the carried end
		std::vector<size_t> fin;              // decided: contig changes and whatever passes after the first one
	};
	std::vector<SegOut> seg(want);
	std::vector<std::thread> thr;
	for (unsigned t = 0; t < want; t++) thr.emplace_back([&, t] {
		SegOut &o = seg[t];
		const size_t lo = rbeg + n * t / want, hi = rbeg + n * (t + 1) / want;
		int tid = -1;
		uint64_t maxend = 0;
		for (size_t i = lo; i < hi; i++) {
			if (!v.kept(i)) continue;
			const int ti = v.tid(i);
			if (!o.any) { o.any = true; o.head = i; o.head_tid = tid = ti; maxend = 0; }
			else if (ti != tid) {
				if (!o.changed) { o.changed = true; o.head_end = maxend; }
				tid = ti; maxend = 0; o.fin.push_back(i);
			} else {
				const uint64_t m = v.minnz(i);
				if (m && m > maxend + 1) {
					if (o.changed) o.fin.push_back(i);
					else o.pending.push_back(Cand{i, m});
				}
			}
			const uint64_t e = v.end(i);
			if (e > maxend) maxend = e;
		}
		o.tail_tid = tid; o.tail_end = maxend;
		if (!o.changed) o.head_end = maxend;
	});
	for (auto &t : thr) t.join();
	int tid = st->tid;
	uint64_t maxend = st->maxend;
	for (const SegOut &o : seg) {
		if (!o.any) continue;
		if (o.head_tid != tid) { tid = o.head_tid; maxend = 0; starts.push_back(o.head); }
		else { const uint64_t m = v.minnz(o.head); if (m && m > maxend + 1) starts.push_back(o.head); }
		// a record that passed against the local end passes against the true one iff it also clears the carried end
		for (const Cand &cd : o.pending) if (cd.minpos > maxend + 1) starts.push_back(cd.i);
		starts.insert(starts.end(), o.fin.begin(), o.fin.end());
		if (o.changed) { tid = o.tail_tid; maxend = o.tail_end; }
		else if (o.head_end > maxend) maxend = o.head_end;
	}
	st->tid = tid; st->maxend = maxend;
}

void certain_block_starts(const bsgpu_record *rec, size_t rbeg, size_t rend, CertainState *st, std::vector<size_t> &starts) {
	certain_scan(RecView{rec}, rbeg, rend, st, starts, 16);
}
void certain_block_starts_keys(const uint32_t *keys, size_t rbeg, size_t rend, CertainState *st, std::vector<size_t> &starts) {
	// a chunk's keys are a few MB and the builder threads of the chunk before are busy on the same cores: a few threads do
	static const unsigned kt = [] { const char *e = getenv("BSGPU_SCAN_THREADS"); const int v = e ? atoi(e) : 0; return v > 0 ? (unsigned)v : 4u; }();
	certain_scan(KeyView{keys}, rbeg, rend, st, starts, kt);
}

// The builder's workers: a fixed pool (two cores are left to the thread that queues device work and to the driver's),
// fed with pieces in the order the jobs were started, so the pieces of an earlier job -- the ones the caller waits for
// first -- are always served first and several jobs in flight never oversubscribe the cores.
class BuildPool {
	std::mutex mu;
	std::condition_variable cv;
	std::deque<std::function<void()>> q;
	std::vector<std::thread> workers;
	BuildPool() {
		unsigned n = std::thread::hardware_concurrency();
		// BSGPU_BUILDER_THREADS is what a multi-rank launcher sets to share the box's cores between its ranks (bench.py)
		if (const char *e = getenv("BSGPU_POOL_THREADS")) n = (unsigned)atoi(e);
		else if (const char *e2 = getenv("BSGPU_BUILDER_THREADS")) n = (unsigned)atoi(e2);
		else n = n > 4 ? n - 2 : n;
		n = std::max(1u, std::min(n, 64u));
		for (unsigned i = 0; i < n; i++) workers.emplace_back([this] {
			for (;;) {
				std::function<void()> f;
				{
					std::unique_lock<std::mutex> lk(mu);
					cv.wait(lk, [this] { return !q.empty(); });
					f = std::move(q.front());
					q.pop_front();
				}
				f();
			}
		});
		for (auto &t : workers) t.detach();              // the pool lives as long as the process
	}
public:
	static BuildPool &get() { static BuildPool *p = new BuildPool(); return *p; }
	void submit(std::function<void()> f) {
		{ std::lock_guard<std::mutex> lk(mu); q.push_back(std::move(f)); }
		cv.notify_one();
	}
};

// A build in flight: the stream cut into pieces at certain block starts, pieces built by a pool of host threads in
// order.  Piece p's templates sit at tmpl + cuts[p] (a piece has no more templates than records) and its blocks number
// their templates from the start of the piece, so a consumer can take pieces over one by one while later ones are
// still being built.
struct BuildJob {
	std::vector<size_t> cuts;
	std::vector<std::vector<bsgpu_block>> pb;
	std::vector<size_t> pn;
	std::vector<uint32_t> pmax;          // per piece: BlockBuilder::maxcap
	std::vector<uint64_t> tally;         // 30 per piece when tallies were asked for
	std::vector<int> rc;
	std::vector<std::atomic<int>> done;
	std::vector<std::thread> thr;
	std::atomic<size_t> next{0};
	explicit BuildJob(size_t np) : pb(np), pn(np, 0), pmax(np, 1), rc(np, 0), done(np) { for (auto &d : done) d.store(0); }
};

// records [rbeg, rend); rbeg is the start of the stream or a certain block start; `starts` = the certain block starts
// inside the range (ascending), from which the cuts between pieces are chosen
BuildJob *build_blocks_start_range(const uint8_t *bam, const uint64_t *rec_off, const bsgpu_record *rec, const uint32_t *name_id, size_t rbeg, size_t rend,
		const std::vector<size_t> &starts, bool keep_unmatched, bool keep_duplicates, bsgpu_template *tmpl, unsigned pieces_per_thread,
		bool with_tally) {
	unsigned want = std::thread::hardware_concurrency();
	if (const char *e = getenv("BSGPU_BUILDER_THREADS")) want = (unsigned)atoi(e);
	want = std::max(1u, std::min(want, 32u));
	size_t min_rec = 50000;                              // below this one thread is faster than starting several
	if (const char *e = getenv("BSGPU_BUILDER_MIN_RECORDS")) min_rec = (size_t)atoll(e);
	const size_t nrec = rend - rbeg;
	std::vector<size_t> cuts{rbeg};
	if (want > 1 && nrec >= min_rec) {
		const size_t npw = (size_t)want * std::max(1u, pieces_per_thread);
		size_t si = 0;
		for (size_t k = 1; k < npw; k++) {
			const size_t target = rbeg + nrec * k / npw;
			while (si < starts.size() && starts[si] < target) si++;
			if (si < starts.size() && starts[si] > cuts.back() && starts[si] < rend) cuts.push_back(starts[si]);
		}
	}
	cuts.push_back(rend);
	const size_t np = cuts.size() - 1;
	BuildJob *job = new BuildJob(np);
	job->cuts = cuts;
	if (with_tally) job->tally.assign(np * 30, 0);
	auto build_piece = [=](size_t p) {
		BlockBuilder b;
		b.bam = bam; b.rec_off = rec_off; b.rec = rec; b.name_id = name_id; b.blocks = &job->pb[p];
		b.out = tmpl + job->cuts[p];
		if (with_tally) b.tally = job->tally.data() + p * 30;
		job->rc[p] = b.run(job->cuts[p], job->cuts[p + 1], keep_unmatched, keep_duplicates);
		job->pn[p] = b.nout;
		job->pmax[p] = b.maxcap;
		job->done[p].store(1, std::memory_order_release);
	};
	if (np == 1) build_piece(0);                          // small stream or one thread asked for: built here, now
	else for (size_t p = 0; p < np; p++) BuildPool::get().submit([=] { build_piece(p); });
	return job;
}

BuildJob *build_blocks_start(const uint8_t *bam, const uint64_t *rec_off, const bsgpu_record *rec, const uint32_t *name_id, size_t nrec, bool keep_unmatched,
		bool keep_duplicates, bsgpu_template *tmpl, unsigned pieces_per_thread, bool with_tally) {
	std::vector<size_t> starts;
	CertainState st;
	certain_block_starts(rec, 0, nrec, &st, starts);
	return build_blocks_start_range(bam, rec_off, rec, name_id, 0, nrec, starts, keep_unmatched, keep_duplicates, tmpl, pieces_per_thread, with_tally);
}

size_t build_blocks_pieces(const BuildJob *job) { return job->pb.size(); }

// waits for piece p; returns its status, its blocks (templates numbered from the start of the piece), the index of its
// first template in the template array and its number of templates
int build_blocks_piece(BuildJob *job, size_t p, const std::vector<bsgpu_block> **blocks, size_t *tmpl_base, size_t *ntmpl) {
	while (!job->done[p].load(std::memory_order_acquire)) std::this_thread::yield();
	*blocks = &job->pb[p];
	*tmpl_base = job->cuts[p];
	*ntmpl = job->pn[p];
	return job->rc[p];
}

uint32_t build_blocks_piece_maxcap(const BuildJob *job, size_t p) { return job->pmax[p]; }
bool build_blocks_piece_ready(const BuildJob *job, size_t p) { return job->done[p].load(std::memory_order_acquire) != 0; }

// read_input's tallies of piece p (valid once build_blocks_piece has returned it): 15 counts then 15 base sums, or NULL
const uint64_t *build_blocks_piece_tally(const BuildJob *job, size_t p) { return job->tally.empty() ? nullptr : job->tally.data() + p * 30; }

void build_blocks_finish(BuildJob *job) {
	// every piece has been handed to the pool (or built): wait until the last one has reported before the job goes
	for (auto &d : job->done) while (!d.load(std::memory_order_acquire)) std::this_thread::yield();
	delete job;
}

// the whole build at once: `tmpl` must have room for nrec templates; *ntmpl receives the count, templates compacted
int build_blocks_host(const uint8_t *bam, const uint64_t *rec_off, const bsgpu_record *rec, size_t nrec, bool keep_unmatched,
		bool keep_duplicates, std::vector<bsgpu_block> &blocks, bsgpu_template *tmpl, size_t *ntmpl, uint64_t *tally) {
	std::vector<uint32_t> name_id(nrec + 1);
	host_name_ids(bam, rec_off, rec, nrec, name_id.data());
	BuildJob *job = build_blocks_start(bam, rec_off, rec, name_id.data(), nrec, keep_unmatched, keep_duplicates, tmpl, 1, tally != nullptr);
	const size_t np = build_blocks_pieces(job);
	size_t at = 0;
	int rc = 0, soft = 0;                 // soft (-6): the blocks are complete, but the reference would abort when it calls one of them
	for (size_t p = 0; p < np; p++) {
		const std::vector<bsgpu_block> *pb;
		size_t base, n;
		const int r = build_blocks_piece(job, p, &pb, &base, &n);
		if (r == -6) soft = r;
		else if (r && !rc) rc = r;
		if (rc) continue;
		if (tally) { const uint64_t *pt = build_blocks_piece_tally(job, p); for (int k = 0; k < 30; k++) tally[k] += pt[k]; }
		for (bsgpu_block b : *pb) { b.first_template += (uint32_t)at; blocks.push_back(b); }
		if (n && at != base) memmove(tmpl + at, tmpl + base, n * sizeof(bsgpu_template));      // close the gap (earlier pieces are done)
		at += n;
	}
	build_blocks_finish(job);
	*ntmpl = at;
	return rc ? rc : soft;
}

}  // namespace bsgpu
