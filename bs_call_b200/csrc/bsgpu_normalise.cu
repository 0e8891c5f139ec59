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
			} else {
				if (e.position + e.size != rl) return false;
				cut_right(m, e.size);
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

// write the mate in reference coordinates, the lanes of the warp copying strided bytes; returns the number of bytes
// written (at most `cap`).  The event list is read after the lane that edited it has synchronised with the warp.
__device__ uint32_t to_ref_coords(const Mate &m, uint8_t *out, uint32_t cap, int lane) {
	uint32_t o = 0, cur = 0;          // output cursor, cursor in the (windowed) read: warp-uniform
	auto copy = [&](uint32_t upto) {
		if (upto <= cur) return;
		const uint32_t n = min(upto - cur, cap - o);
		for (uint32_t j = lane; j < n; j += 32) out[o + j] = marked(m, m.s + cur + j);
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

// One WARP per template.  The event-list surgery (soft clips, overlap) is a short sequential walk done by lane 0; what
// touches every byte of the reads -- the quality means that break span ties, the rewrite into reference coordinates,
// the search for the first and last counted byte -- is done by the 32 lanes together with coalesced accesses.
__global__ void __launch_bounds__(kNormWarps * 32)
k_normalise(const bsgpu_template *__restrict__ tmpl, size_t n, const uint8_t *__restrict__ bases,
		bsgpu_misms *__restrict__ ev_work, const uint32_t *__restrict__ out_off, uint8_t *__restrict__ obases,
		Seg *__restrict__ segs, uint32_t segs_per_mate, uint32_t x, uint32_t y,
		uint32_t lt0, uint32_t rt0, uint32_t lt1, uint32_t rt1, unsigned long long *__restrict__ counters) {
	const size_t i = (size_t)blockIdx.x * kNormWarps + (threadIdx.x >> 5);
	const int lane = threadIdx.x & 31;
	if (i >= n) return;
	const bsgpu_template t = tmpl[i];          // same address in every lane: one broadcast load
	Mate mt[2];
	uint32_t pos[2] = { t.forward_position, t.reverse_position };
	const uint32_t span[2] = { t.reference_span[0], t.reference_span[1] };
	// -L/-R refer to read 1 / read 2; slot [0] holds read 1 iff the template is FORWARD (src/process_template.c:36-41)
	const int msk = t.orientation == 0 ? 0 : 1;
	for (int k = 0; k < 2; k++) {
		Mate &m = mt[k];
		m.present = t.present[k] != 0;
		m.raw = bases + t.read_off[k];
		m.rl0 = m.present ? t.read_len[k] : 0;
		m.s = 0;
		m.len = m.rl0;
		m.ev = ev_work + t.mm_off[k];
		m.nev = t.mm_n[k];
		const int r = k ^ msk;          // which read (0 = R1, 1 = R2) sits in slot k
		m.lt = r ? lt1 : lt0;
		m.rt = r ? rt1 : rt0;
	}
	Seg *sg = segs + i * 2 * (size_t)segs_per_mate;
	for (uint32_t j = lane; j < 2 * segs_per_mate; j += 32) { Seg e; e.pos = 0; e.off = 0; e.len = 0; e.mapq = 0; e.flags = 0; e.pad = 0; sg[j] = e; }
	// ---- lane 0: soft clips; the windows it arrives at are handed to the other lanes
	uint32_t ok = 1;
	if (lane == 0) ok = strip_soft_clips(mt[0]) && strip_soft_clips(mt[1]);
	ok = __shfl_sync(0xffffffffu, ok, 0);
	if (!ok) {
		if (lane == 0) atomicAdd(counters + 2, 1ull);      // "Error in CIGAR" (src/al_utils.c:134-147): reported by the host
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
	if (lane == 0) resolve_overlap(mt, pos, span, mq0, mq1);
	for (int k = 0; k < 2; k++) {
		mt[k].s = __shfl_sync(0xffffffffu, mt[k].s, 0);
		mt[k].len = __shfl_sync(0xffffffffu, mt[k].len, 0);
		mt[k].nev = __shfl_sync(0xffffffffu, mt[k].nev, 0);
		pos[k] = __shfl_sync(0xffffffffu, pos[k], 0);
	}
	__syncwarp();                          // lane 0's edits of the event lists are visible to the warp
	uint32_t ori = t.orientation & 1u;
	for (int k = 0; k < 2; k++) {
		if (!mt[k].present) continue;
		uint8_t *out = obases + out_off[2 * i + k];
		const uint32_t cap = out_off[2 * i + k + 1] - out_off[2 * i + k];
		const uint32_t rl = to_ref_coords(mt[k], out, cap, lane);
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
			uint32_t p = pos[k] + first, off = out_off[2 * i + k] + first, len = last - first;
			if (p < x) { atomicAdd(counters + 3, 1ull); len = 0; }       // cannot happen for a well-formed block (assert at :186)
			if (len && p <= y) {
				if ((uint64_t)p + len > (uint64_t)y + 1) len = y + 1 - p;
				Seg *d = sg + (size_t)k * segs_per_mate;
				for (uint32_t c = 0; c < segs_per_mate && len; c++) {
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

}  // namespace

cudaError_t launch_normalise(const void *tmpl, size_t n, const void *bases, void *ev_work, const void *out_off, void *obases,
		void *segs, uint32_t segs_per_mate, uint32_t x, uint32_t y, const uint32_t left_trim[2], const uint32_t right_trim[2],
		unsigned long long *counters, cudaStream_t stream, int *launches) {
	if (!n) return cudaSuccess;
	k_normalise<<<(unsigned)((n + kNormWarps - 1) / kNormWarps), kNormWarps * 32, 0, stream>>>((const bsgpu_template *)tmpl, n, (const uint8_t *)bases,
			(bsgpu_misms *)ev_work, (const uint32_t *)out_off, (uint8_t *)obases, (Seg *)segs, segs_per_mate, x, y,
			left_trim[0], right_trim[0], left_trim[1], right_trim[1], counters);
	*launches += 1;
	return cudaGetLastError();
}

}  // namespace bsgpu
