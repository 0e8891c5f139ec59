// bsgpu_normalise.cu -- template normalisation on the device (one warp per template).
//
// Does per template what the reference's process_template_vector does before it calls call_genotypes_ML
// (src/process_template.c:36-110):
//     trim_read          -L/-R marking with quality 63                  src/read_utils.c:13-26
//     trim_soft_clips    drop soft-clipped ends, rebase the event list   src/al_utils.c:122-162
//     handle_overlap     give the shared part of overlapping mates to one of them, honouring indels  src/al_utils.c:164-318
//     indel normalise    zero-fill reference bases the read lacks (CIGAR D), drop inserted bases (CIGAR I)  :66-110
// and then the mate walk of the pileup loop (src/call_genotypes.c:181-212, 224) that decides which strand index each
// mate is counted under.  Output: reads in reference coordinates in `obases` plus one segment record per mate chunk.
//
// A mate is kept as a window (start, length) into the raw byte array, so the memmove()s of the reference become
// index arithmetic; the only copy is the final write in reference coordinates.  Integer widths and wrap-around
// follow the reference (uint32 lengths, int32 overlap, int64 running adjustment), because out-of-range CIGARs take
// the same path there.
#include <algorithm>
#include <cstdint>
#include <cuda_runtime.h>
#include "bsgpu.h"
#include "bsgpu_launch.h"

namespace bsgpu {

namespace {

constexpr uint32_t kFlt = BSGPU_FLT_QUAL;
enum { EV_INS = 1, EV_DEL = 2, EV_SOFT = 3 };

struct Mate {
	const uint8_t *raw;      // original read bytes
	uint32_t rl0;            // original length (trim marks refer to it)
	uint32_t lt, rt;         // -L / -R for this mate
	uint32_t s, len;         // current window into raw
	bsgpu_misms *ev;         // private, mutable copy of the event list
	uint32_t nev;
	uint32_t tl, tr;         // bases taken off the left / right end so far (trim_left / trim_right of src/process_template.c:42-45)
	bool present;
};

// byte j of the original read after trim_read's marking (the right trim takes its base bits from the mirrored left
// index -- reference quirk, src/read_utils.c:22)
__device__ __forceinline__ uint8_t marked(const Mate &m, uint32_t j) {
	const uint32_t k1 = m.rl0 - 1 - j;
	if (k1 < m.rt) return (uint8_t)((m.raw[k1] & 3) | (kFlt << 2));
	if (j < m.lt) return (uint8_t)((m.raw[j] & 3) | (kFlt << 2));
	return m.raw[j];
}

__device__ __forceinline__ void cut_left(Mate &m, uint32_t l) {
	if (!l) return;
	if (l >= m.len) { m.len = 0; return; }
	m.s += l;
	m.len -= l;
}
__device__ __forceinline__ void cut_right(Mate &m, uint32_t l) {
	if (!l) return;
	if (l >= m.len) m.len = 0;
	else m.len -= l;
}

__device__ bool strip_soft_clips(Mate &m) {
	if (!m.present || !m.len) return true;
	const uint32_t rl = m.len, n0 = m.nev;
	uint32_t kept = 0, shift = 0, nclip = 0;
	for (uint32_t z = 0; z < n0; z++) {
		bsgpu_misms e = m.ev[z];
		if (e.type == EV_SOFT) {
			if (z && z != n0 - 1) return false;
			nclip++;
			if (!e.position) {
				if (e.size >= rl) return false;
				shift = e.size;
				cut_left(m, shift);
				m.tl = shift;
			} else {
				if (e.position + e.size != rl) return false;
				cut_right(m, e.size);
				m.tr = e.size;
			}
		} else {
			if (nclip) e.position -= shift;
			m.ev[kept++] = e;
		}
	}
	m.nev = kept;
	return true;
}

// mean quality of the bytes that are not trim-marked (src/al_utils.c:185-202), the warp summing strided bytes
__device__ uint32_t mean_untrimmed_qual(const Mate &m, int lane) {
	uint32_t tot = 0, n = 0;
	for (uint32_t i = lane; i < m.len; i += 32) {
		const uint32_t q = marked(m, m.s + i) >> 2;
		if (q != kFlt) { tot += q; n++; }
	}
	tot = __reduce_add_sync(0xffffffffu, tot);
	n = __reduce_add_sync(0xffffffffu, n);
	return n ? tot / n : 0;
}

__device__ __forceinline__ void drop_events(Mate &m, uint32_t z) {
	if (z) for (uint32_t i = z; i < m.nev; i++) m.ev[i - z] = m.ev[i];
	m.nev -= z;
}

// src/al_utils.c:164-318
__device__ void resolve_overlap(Mate mt[2], uint32_t pos[2], const uint32_t span[2], uint32_t meanq0, uint32_t meanq1) {
	if (!(mt[0].present && mt[0].len && mt[1].present && mt[1].len)) return;
	const bool rev = !(pos[0] <= pos[1]);
	const int32_t overlap = rev ? (int32_t)(span[1] + pos[1] - pos[0]) : (int32_t)(span[0] - pos[1] + pos[0]);
	if (!(pos[0] + span[0] >= pos[1])) return;
	int tr;
	if (span[0] > span[1]) tr = 1;
	else if (span[0] < span[1]) tr = 0;
	else tr = meanq0 <= meanq1 ? 0 : 1;
	const bool at_right = (rev == (tr != 0));
	if (!at_right) pos[tr] += (uint32_t)overlap;
	Mate &m = mt[tr];
	const uint32_t rl = m.len;
	const uint32_t nev = m.nev;
	if (!nev) {
		if (at_right) cut_right(m, (uint32_t)overlap);
		else cut_left(m, (uint32_t)overlap);
		return;
	}
	bool done = false;
	int64_t adj = 0;
	if (at_right) {
		const uint32_t keep = span[tr] - (uint32_t)overlap;
		for (uint32_t z = 0; z < nev; z++) {
			bsgpu_misms &e = m.ev[z];
			if ((int64_t)e.position + adj >= (int64_t)keep) {
				const int64_t trim = (int64_t)(uint32_t)(rl - keep) + adj;
				cut_right(m, (uint32_t)trim);
				m.nev = z;
				done = true;
				break;
			}
			if (e.type == EV_INS) {
				if ((int64_t)e.position + adj + (int64_t)e.size >= (int64_t)keep) {
					const uint32_t trim = rl - e.position;
					e.size = (uint32_t)((int64_t)keep - ((int64_t)e.position + adj));
					cut_right(m, trim);
					m.nev = z + 1;
					done = true;
					break;
				}
				adj += e.size;
			} else if (e.type == EV_DEL) adj -= e.size;
		}
		if (!done) cut_right(m, (uint32_t)overlap);
	} else {
		const uint32_t cut = (uint32_t)overlap;
		for (uint32_t z = 0; z < nev; z++) {
			bsgpu_misms &e = m.ev[z];
			if ((int64_t)e.position + adj >= (int64_t)cut) {
				const uint32_t trim = (uint32_t)((int64_t)overlap - adj);
				cut_left(m, trim);
				for (uint32_t z1 = z; z1 < nev; z1++) m.ev[z1].position -= trim;
				drop_events(m, z);
				done = true;
				break;
			}
			if (e.type == EV_INS) {
				if ((int64_t)e.position + adj + (int64_t)e.size >= (int64_t)cut) {
					e.size = (uint32_t)((int64_t)e.position + (int64_t)e.size + adj - (int64_t)cut);
					const uint32_t trim = e.position;
					cut_left(m, trim);
					const uint32_t z2 = e.size ? z : z + 1;
					for (uint32_t z1 = z2; z1 < nev; z1++) m.ev[z1].position -= trim;
					drop_events(m, z2);
					done = true;
					break;
				}
				adj += e.size;
			} else if (e.type == EV_DEL) adj -= e.size;
		}
		if (!done) {
			cut_left(m, (uint32_t)((int64_t)overlap - adj));
			m.nev = 0;
		}
	}
}

// ---- non-CpG conversion profile (src/meth_profile.c:48-76), a by-product of the rewrite below ----
// What one mate needs to classify its bytes: where byte 0 sits in the reference window, the strand's row of
// flt_tab (src/init_param.c:57-69) packed as four nibbles, and the map from read index to original read position.
struct ProfMate {
	int32_t r0;              // index in ref[] of the code the FSM pairs with output byte 0
	uint32_t tab;            // nibble b = flt_tab[strand][q << 2 | b] for MIN_QUAL <= q < FLT_QUAL
	uint32_t k, tl;
	int32_t posx, maxpos;
};
// per-thread tallies of stats->base_filter / filter_cts / filter_bases (src/process_template.c:52-63, src/al_utils.c:141,150,308)
struct ProfAcc { uint32_t none, trim, lowqual, clip, overlap, reads, read_bases, too_long; };

constexpr uint32_t kProfMinQual = 20;    // MIN_QUAL (include/bs_call.h:28): flt_tab is zero below it whatever -Q says

__device__ __forceinline__ uint32_t ref_code(const ProfArgs &pa, int64_t r) { return r < 0 || r >= (int64_t)pa.refn ? 0u : pa.ref[r]; }

// One byte of a normalised read.  meth_profile's FSM pairs read byte j with the reference codes (prev, cur, next) =
// R[r0+j-1 .. r0+j+1] and counts it iff cur is a C followed by A/C/T and the strand row has bit 2 for the read base, or
// cur is a G preceded by A/G/T and the row has bit 3 (rtab / btab, src/meth_profile.c:14-23,63-73).  The count goes to
// entry orig+1 of the profile; the one byte whose entry equals the template's own `used` is set aside (see
// k_profile_resolve).
__device__ __forceinline__ void profile_byte(const ProfArgs &pa, const ProfMate &pm, uint32_t *hist, size_t ti, uint32_t oj, uint32_t c, uint8_t b) {
	const uint32_t q = b >> 2;
	if (q < kProfMinQual || q == kFlt) return;
	const uint32_t xx = (pm.tab >> (4 * (b & 3u))) & 15u;
	const int64_t r = (int64_t)pm.r0 + oj;
	const uint32_t cur = ref_code(pa, r);
	bool hit = false;
	if (cur == 2 && (xx & 4u)) { const uint32_t nx = ref_code(pa, r + 1); hit = nx == 1 || nx == 2 || nx == 4; }
	else if (cur == 3 && (xx & 8u)) { const uint32_t pv = ref_code(pa, r - 1); hit = pv == 1 || pv == 3 || pv == 4; }
	if (!hit) return;
	const int32_t orig = pm.k ? pm.posx - (int32_t)c : (int32_t)(pm.tl + c);
	if (orig == pm.maxpos) { pa.cand[ti] = (uint8_t)(1u + (xx & 3u)); return; }
	const uint32_t idx = (uint32_t)(orig + 1);
	if (idx < BSGPU_PROFILE_MAX) atomicAdd(hist + idx * 4 + (xx & 3u), 1u);
}

// write the mate in reference coordinates, the lanes of the warp copying strided bytes; returns the number of bytes
// written (at most `cap`).  The event list is read after the lane that edited it has synchronised with the warp.
template <bool PROF>
__device__ uint32_t to_ref_coords(const Mate &m, uint8_t *out, uint32_t cap, int lane, const ProfArgs &pa, const ProfMate &pm, uint32_t *hist, size_t ti) {
	uint32_t o = 0, cur = 0;          // output cursor, cursor in the (windowed) read: warp-uniform
	auto copy = [&](uint32_t upto) {
		if (upto <= cur) return;
		const uint32_t n = min(upto - cur, cap - o);
		for (uint32_t j = lane; j < n; j += 32) {
			const uint8_t b = marked(m, m.s + cur + j);
			out[o + j] = b;
			if (PROF) profile_byte(pa, pm, hist, ti, o + j, cur + j, b);
		}
		o += n; cur += n;
	};
	for (uint32_t z = 0; z < m.nev; z++) {
		const bsgpu_misms e = m.ev[z];
		copy(e.position < m.len ? e.position : m.len);
		if (e.type == EV_INS) {
			const uint32_t n = min(e.size, cap - o);
			for (uint32_t j = lane; j < n; j += 32) out[o + j] = 0;
			o += n;
		} else if (e.type == EV_DEL) {
			const uint32_t nx = e.position + e.size;
			if (nx > cur) cur = nx < m.len ? nx : m.len;
		}
	}
	copy(m.len);
	return o;
}

constexpr int kNormWarps = 8;

struct NormArgs {
	const bsgpu_template *tmpl; size_t n; const uint8_t *bases; bsgpu_misms *ev_work; const uint32_t *out_off; uint8_t *obases;
	Seg *segs; uint32_t segs_per_mate, x, y, lt0, rt0, lt1, rt1; unsigned long long *counters;
	uint32_t slot;               // out_off == NULL: mate j of the launch owns bytes [j * slot, (j + 1) * slot) of obases
};
__device__ __forceinline__ uint32_t out_begin(const NormArgs &a, size_t j) { return a.out_off ? a.out_off[j] : (uint32_t)(j * a.slot); }
__device__ __forceinline__ uint32_t out_room(const NormArgs &a, size_t j) { return a.out_off ? a.out_off[j + 1] - a.out_off[j] : a.slot; }

// One WARP per template.  The event-list surgery (soft clips, overlap) is a short sequential walk done by lane 0; what
// touches every byte of the reads -- the quality means that break span ties, the rewrite into reference coordinates,
// the search for the first and last counted byte -- is done by the 32 lanes together with coalesced accesses.
template <bool PROF>
__device__ void normalise_one(const NormArgs &a, const size_t i, const int lane, const ProfArgs &pa, uint32_t *hist, ProfAcc &acc) {
	const bsgpu_template t = a.tmpl[i];          // same address in every lane: one broadcast load
	Mate mt[2];
	uint32_t pos[2] = { t.forward_position, t.reverse_position };
	const uint32_t span[2] = { t.reference_span[0], t.reference_span[1] };
	// -L/-R refer to read 1 / read 2; slot [0] holds read 1 iff the template is FORWARD (src/process_template.c:36-41)
	const int msk = t.orientation == 0 ? 0 : 1;
	for (int k = 0; k < 2; k++) {
		Mate &m = mt[k];
		m.present = t.present[k] != 0;
		m.raw = a.bases + t.read_off[k];
		m.rl0 = m.present ? t.read_len[k] : 0;
		m.s = 0;
		m.len = m.rl0;
		m.ev = a.ev_work + t.mm_off[k];
		m.nev = t.mm_n[k];
		m.tl = m.tr = 0;
		const int r = k ^ msk;          // which read (0 = R1, 1 = R2) sits in slot k
		m.lt = r ? a.lt1 : a.lt0;
		m.rt = r ? a.rt1 : a.rt0;
	}
	Seg *sg = a.segs + i * 2 * (size_t)a.segs_per_mate;
	for (uint32_t j = lane; j < 2 * a.segs_per_mate; j += 32) { Seg e; e.pos = 0; e.off = 0; e.len = 0; e.mapq = 0; e.flags = 0; e.pad = 0; sg[j] = e; }
	if (PROF && lane == 0) { pa.cand[i] = 0; pa.used16[i] = 0; }
	// ---- lane 0: soft clips; the windows it arrives at are handed to the other lanes
	uint32_t ok = 1;
	if (lane == 0) ok = strip_soft_clips(mt[0]) && strip_soft_clips(mt[1]);
	ok = __shfl_sync(0xffffffffu, ok, 0);
	if (!ok) {
		if (lane == 0) atomicAdd(a.counters + 2, 1ull);      // "Error in CIGAR" (src/al_utils.c:134-147): reported by the host
		return;
	}
	for (int k = 0; k < 2; k++) {
		mt[k].s = __shfl_sync(0xffffffffu, mt[k].s, 0);
		mt[k].len = __shfl_sync(0xffffffffu, mt[k].len, 0);
		mt[k].nev = __shfl_sync(0xffffffffu, mt[k].nev, 0);
	}
	// ---- all lanes: the two quality means (only consulted when the spans tie)
	uint32_t mq0 = 0, mq1 = 0;
	if (mt[0].present && mt[0].len && mt[1].present && mt[1].len && span[0] == span[1] && pos[0] + span[0] >= pos[1]) {
		mq0 = mean_untrimmed_qual(mt[0], lane);
		mq1 = mean_untrimmed_qual(mt[1], lane);
	}
	// ---- lane 0: overlap
	if (lane == 0) {
		const uint32_t len0 = mt[0].len, len1 = mt[1].len;
		const bool rev = !(pos[0] <= pos[1]);
		resolve_overlap(mt, pos, span, mq0, mq1);
		if (PROF) {
			// what the overlap took goes to trim_right of the left-hand mate or trim_left of the right-hand one (:309-313)
			acc.clip += mt[0].tl + mt[0].tr + mt[1].tl + mt[1].tr;
			const uint32_t d[2] = { len0 - mt[0].len, len1 - mt[1].len };
			for (int k = 0; k < 2; k++) if (d[k]) { if (rev == (k != 0)) mt[k].tr += d[k]; else mt[k].tl += d[k]; }
			acc.overlap += d[0] + d[1];
		}
	}
	for (int k = 0; k < 2; k++) {
		mt[k].s = __shfl_sync(0xffffffffu, mt[k].s, 0);
		mt[k].len = __shfl_sync(0xffffffffu, mt[k].len, 0);
		mt[k].nev = __shfl_sync(0xffffffffu, mt[k].nev, 0);
		pos[k] = __shfl_sync(0xffffffffu, pos[k], 0);
		if (PROF) {
			mt[k].tl = __shfl_sync(0xffffffffu, mt[k].tl, 0);
			mt[k].tr = __shfl_sync(0xffffffffu, mt[k].tr, 0);
		}
	}
	__syncwarp();                          // lane 0's edits of the event lists are visible to the warp
	ProfMate pm[2];
	if (PROF) {
		// original read positions (src/process_template.c:80-91): slot 0 counts up from trim_left, slot 1 down from
		// rdl + trim_right - 1; the template claims entries up to max_pos of the profile
		int32_t maxpos = 0;
		for (int k = 0; k < 2; k++) {
			if (!mt[k].present) continue;
			const int32_t mpos = k ? (int32_t)(mt[k].len + mt[k].tr) - 1 : (int32_t)(mt[k].tl + mt[k].len);
			if (mpos > maxpos) maxpos = mpos;
			// base_filter tallies over the bytes the overlap left (src/process_template.c:52-62)
			for (uint32_t j = lane; j < mt[k].len; j += 32) {
				const uint32_t q = marked(mt[k], mt[k].s + j) >> 2;
				if (q == kFlt) acc.trim++;
				else if (q < pa.min_qual) acc.lowqual++;
				else acc.none++;
			}
			if (lane == 0) { acc.reads++; acc.read_bases += mt[k].len; }
		}
		const uint32_t used = (uint32_t)maxpos + 1;
		if (lane == 0) {
			if (used > BSGPU_PROFILE_MAX) acc.too_long++;
			pa.used16[i] = (uint16_t)(used > 0xffffu ? 0xffffu : used);
			uint32_t *cm = pa.chunkmax + (i / kProfChunk);
			if (used > *(volatile uint32_t *)cm) atomicMax(cm, used);
		}
		const uint32_t row = t.bs_strand == 0 ? 0x7A6Bu : t.bs_strand == 1 ? 0x5A4Bu : 0x7869u;      // src/init_param.c:60-67
		for (int k = 0; k < 2; k++) {
			// the FSM starts one code late for a mate that begins at the block's first position (src/meth_profile.c:65),
			// which only position 1 can be: blocks start two bases before their first template (src/process_template.c:27)
			pm[k].r0 = (int32_t)(pos[k] - a.x) - (pos[k] == 1 ? 1 : 0);
			pm[k].tab = row;
			pm[k].k = (uint32_t)k;
			pm[k].tl = mt[k].tl;
			pm[k].posx = (int32_t)(mt[k].len + mt[k].tr) - 1;
			pm[k].maxpos = maxpos;
		}
		__syncwarp();                        // cand[i] = 0 above is ordered before a lane sets it
	}
	uint32_t ori = t.orientation & 1u;
	for (int k = 0; k < 2; k++) {
		if (!mt[k].present) continue;
		uint8_t *out = a.obases + out_begin(a, 2 * i + k);
		const uint32_t cap = out_room(a, 2 * i + k);
		const uint32_t rl = to_ref_coords<PROF>(mt[k], out, cap, lane, pa, pm[k], hist, i);
		if (!rl) continue;
		__syncwarp();
		// mate walk of the pileup loop: first / last byte with 0 < q != 63; a mate without one does not flip `ori`
		uint32_t first = rl, last = 0;
		for (uint32_t b0 = 0; b0 < rl; b0 += 32) {
			const uint32_t idx = b0 + lane;
			const uint32_t q = idx < rl ? out[idx] >> 2 : 0;
			const uint32_t hit = __ballot_sync(0xffffffffu, q > 0 && q != kFlt);
			if (hit) { first = b0 + (uint32_t)__ffs(hit) - 1; break; }
		}
		if (first == rl) continue;
		for (uint32_t b0 = (rl - 1) & ~31u;; b0 -= 32) {
			const uint32_t idx = b0 + lane;
			const uint32_t q = idx < rl ? out[idx] >> 2 : 0;
			const uint32_t hit = __ballot_sync(0xffffffffu, q > 0 && q != kFlt);
			if (hit) { last = b0 + 32 - (uint32_t)__clz(hit); break; }
			if (!b0) break;
		}
		if (lane == 0) {
			uint32_t p = pos[k] + first, off = out_begin(a, 2 * i + k) + first, len = last - first;
			if (p < a.x) { atomicAdd(a.counters + 3, 1ull); len = 0; }       // cannot happen for a well-formed block (assert at :186)
			if (len && p <= a.y) {
				if ((uint64_t)p + len > (uint64_t)a.y + 1) len = a.y + 1 - p;
				Seg *d = sg + (size_t)k * a.segs_per_mate;
				for (uint32_t c = 0; c < a.segs_per_mate && len; c++) {
					const uint32_t l = len > BSGPU_MAX_SEG_LEN ? BSGPU_MAX_SEG_LEN : len;
					Seg e;
					e.pos = p; e.off = off; e.len = (uint16_t)l; e.mapq = t.mapq[k]; e.flags = (uint8_t)(ori | ((uint32_t)t.bs_strand << 1)); e.pad = 0;
					d[c] = e;
					p += l; off += l; len -= l;
				}
			}
		}
		ori ^= 1u;
	}
}

__global__ void __launch_bounds__(kNormWarps * 32) k_normalise(const NormArgs a) {
	const size_t i = (size_t)blockIdx.x * kNormWarps + (threadIdx.x >> 5);
	if (i >= a.n) return;
	ProfArgs pa{};
	ProfAcc acc{};
	normalise_one<false>(a, i, threadIdx.x & 31, pa, nullptr, acc);
}

// The same with the --report-file side channels switched on: every CTA keeps the profile it gathers in shared memory
// over many templates (the grid is a few CTAs per SM, warps stride over the templates) and adds it to the context's
// profile once at the end.
__global__ void __launch_bounds__(kNormWarps * 32) k_normalise_profile(const NormArgs a, const ProfArgs pa) {
	__shared__ uint32_t hist[BSGPU_PROFILE_MAX * 4];
	for (uint32_t j = threadIdx.x; j < BSGPU_PROFILE_MAX * 4; j += blockDim.x) hist[j] = 0;
	__syncthreads();
	ProfAcc acc{};
	const int lane = threadIdx.x & 31;
	for (size_t i = (size_t)blockIdx.x * kNormWarps + (threadIdx.x >> 5); i < a.n; i += (size_t)gridDim.x * kNormWarps)
		normalise_one<true>(a, i, lane, pa, hist, acc);
	__syncthreads();
	ProfDev *pd = pa.prof;
	for (uint32_t j = threadIdx.x; j < BSGPU_PROFILE_MAX * 4; j += blockDim.x)
		if (hist[j]) atomicAdd(&pd->conv[0][0] + j, (unsigned long long)hist[j]);
	const uint32_t none = __reduce_add_sync(0xffffffffu, acc.none), trim = __reduce_add_sync(0xffffffffu, acc.trim),
			lowq = __reduce_add_sync(0xffffffffu, acc.lowqual);
	if (lane == 0) {
		if (none) atomicAdd(&pd->base_filter[0], (unsigned long long)none);
		if (trim) atomicAdd(&pd->base_filter[1], (unsigned long long)trim);
		if (acc.clip) atomicAdd(&pd->base_filter[2], (unsigned long long)acc.clip);
		if (acc.overlap) atomicAdd(&pd->base_filter[3], (unsigned long long)acc.overlap);
		if (lowq) atomicAdd(&pd->base_filter[4], (unsigned long long)lowq);
		if (acc.reads) atomicAdd(&pd->reads, (unsigned long long)acc.reads);
		if (acc.read_bases) atomicAdd(&pd->read_bases, (unsigned long long)acc.read_bases);
		if (acc.too_long) atomicAdd(&pd->too_long, (unsigned long long)acc.too_long);
	}
}

// The reference's profile vector grows as templates arrive: meth_profile() raises `used` to the template's max_pos + 1
// and gt_vector_reserve(.., true) then clears everything from the old `used` upwards (src/meth_profile.c:52-55,
// gt/src/gt_vector.c:34-37).  A count that lands on entry `used` itself -- only the first byte of a slot-1 mate whose
// original position is the template's max_pos can -- is therefore never reported unless an EARLIER template had
// already raised `used` beyond it.  k_normalise_profile sets those bytes aside (cand[i] = 1 + counter); here the running
// maximum of `used` before every template is rebuilt (carry from earlier launches, maxima of the chunks before this
// one, a scan inside the chunk) and the byte is added iff the reference would have kept it.
__global__ void __launch_bounds__(kProfChunk / 4) k_profile_resolve(const uint16_t *__restrict__ used16, const uint8_t *__restrict__ cand,
		const uint32_t *__restrict__ chunkmax, size_t n, ProfDev *__restrict__ pd, int parity) {
	constexpr int kT = kProfChunk / 4;
	__shared__ uint32_t wmax[kT / 32];
	__shared__ uint32_t base_s;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const uint32_t c = blockIdx.x;
	uint32_t m = tid == 0 ? pd->used[parity] : 0u;
	for (uint32_t j = tid; j < c; j += kT) m = max(m, chunkmax[j]);
	m = __reduce_max_sync(0xffffffffu, m);
	if (lane == 0) wmax[w] = m;
	__syncthreads();
	if (tid == 0) {
		uint32_t b = 0;
		for (int j = 0; j < kT / 32; j++) b = max(b, wmax[j]);
		base_s = b;
		if (c == gridDim.x - 1) pd->used[parity ^ 1] = max(b, chunkmax[c]);
	}
	__syncthreads();
	const uint32_t base = base_s;
	__syncthreads();
	const size_t i0 = (size_t)c * kProfChunk + (size_t)tid * 4;
	uint32_t u[4], cd[4], own = 0;
	for (int e = 0; e < 4; e++) {
		const size_t i = i0 + e;
		u[e] = i < n ? used16[i] : 0u;
		cd[e] = i < n ? cand[i] : 0u;
		own = max(own, u[e]);
	}
	// exclusive running maximum over the threads of the chunk
	uint32_t inc = own;
	for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc = max(inc, o); }
	if (lane == 31) wmax[w] = inc;
	__syncthreads();
	uint32_t run = base;
	for (int j = 0; j < w; j++) run = max(run, wmax[j]);
	const uint32_t prev = __shfl_up_sync(0xffffffffu, inc, 1);
	if (lane) run = max(run, prev);
	for (int e = 0; e < 4; e++) {
		const bool keep = cd[e] && run > u[e] && u[e] < BSGPU_PROFILE_MAX;
		const uint32_t key = keep ? (u[e] << 2 | (cd[e] - 1)) : 0xffffffffu;
		const uint32_t peers = __match_any_sync(0xffffffffu, key);
		if (keep && lane == __ffs(peers) - 1) atomicAdd(&pd->conv[0][0] + key, (unsigned long long)__popc(peers));
		run = max(run, u[e]);
	}
}

}  // namespace

cudaError_t launch_normalise(const void *tmpl, size_t n, const void *bases, void *ev_work, const void *out_off, void *obases,
		void *segs, uint32_t segs_per_mate, uint32_t x, uint32_t y, const uint32_t left_trim[2], const uint32_t right_trim[2],
		unsigned long long *counters, const ProfArgs *prof, int parity, cudaStream_t stream, int *launches, uint32_t slot) {
	if (!n) return cudaSuccess;
	NormArgs a;
	a.tmpl = (const bsgpu_template *)tmpl; a.n = n; a.bases = (const uint8_t *)bases; a.ev_work = (bsgpu_misms *)ev_work;
	a.out_off = (const uint32_t *)out_off; a.obases = (uint8_t *)obases; a.segs = (Seg *)segs; a.segs_per_mate = segs_per_mate;
	a.x = x; a.y = y; a.lt0 = left_trim[0]; a.rt0 = right_trim[0]; a.lt1 = left_trim[1]; a.rt1 = right_trim[1]; a.counters = counters;
	a.slot = slot;
	const size_t ctas = (n + kNormWarps - 1) / kNormWarps;
	if (!prof) {
		k_normalise<<<(unsigned)ctas, kNormWarps * 32, 0, stream>>>(a);
		__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
		return cudaGetLastError();
	}
	const size_t nchunks = (n + kProfChunk - 1) / kProfChunk;
	cudaError_t e = cudaMemsetAsync(prof->chunkmax, 0, nchunks * sizeof(uint32_t), stream);
	if (e != cudaSuccess) return e;
	static int sms = 0;
	if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
	k_normalise_profile<<<(unsigned)std::min(ctas, (size_t)sms * 6), kNormWarps * 32, 0, stream>>>(a, *prof);
	k_profile_resolve<<<(unsigned)nchunks, kProfChunk / 4, 0, stream>>>(prof->used16, prof->cand, prof->chunkmax, n, prof->prof, parity);
	__atomic_fetch_add(launches, 2, __ATOMIC_RELAXED);
	return cudaGetLastError();
}

}  // namespace bsgpu
