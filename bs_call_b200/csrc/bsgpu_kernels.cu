// bsgpu_kernels.cu -- sm_100a kernels of libbsgpu and their launchers.
//
//   k_call_sites      pileup[] (+ref) -> gt_meth[] / gt_vcf[]            (likelihood kernel, config 2 of BASELINE.json)
//   k_bin_*           counting sort of read segments by 256-site tile    (replaces a host sort)
//   k_pileup_tile     segments -> per-site counts in registers -> either pileup[] or, fused, gt_vcf[]
//   k_synth_*         counter-based synthetic workloads generated in HBM
//
// Layout notes (DESIGN.md has the full picture):
//   * records keep the reference's AoS layouts (104 / 200 / 208 B) at the ABI; inside a CTA a tile of records is moved
//     between HBM and shared memory as ONE contiguous block with the TMA bulk-copy engine (cp.async.bulk, SASS UBLKCP),
//     and threads touch their own record in shared memory with 8-byte accesses at an odd 8-byte stride (13 or 25 words)
//     so the accesses are bank-conflict free.
//   * the pileup is a GATHER: one thread owns one site and pulls the one byte each overlapping read contributes, so
//     there are no atomics and the integer sums are order independent (bit-exact against the reference's float sums
//     inside the 2^24 envelope).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "bsgpu_device.cuh"
#include "bsgpu_launch.h"

namespace bsgpu {

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copies
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
	asm volatile(
		"{\n"
		".reg .pred P1;\n"
		"LAB_WAIT:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
		"@P1 bra DONE;\n"
		"bra LAB_WAIT;\n"
		"DONE:\n"
		"}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void *gdst, const void *smem_src, uint32_t bytes) {
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// shared-memory copy of the lookup tables (q -> {k, ln k, ln(1/2+k), ln(1+k)}, log/exp reduction tables)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_tables(Tables *st, const DevConst *__restrict__ dc, int tid, int nthr) {
	const double *src = (const double *)&dc->tab;
	double *dst = (double *)st;
	for (int i = tid; i < (int)(sizeof(Tables) / sizeof(double)); i += nthr) dst[i] = src[i];
}

// Write the CTA's staged output tile (nrec records of REC bytes, contiguous in smem) to global memory.  With the bulk
// engine the copy is asynchronous: the issuing thread (tid 0) must call tma_store_wait() before the tile buffer is
// written again or the CTA exits.
template <int REC>
__device__ __forceinline__ void store_tile(void *gdst, const uint64_t *stage, int nrec, bool bulk_ok, int tid, int nthr) {
	const uint32_t bytes = (uint32_t)nrec * REC;
	if (bulk_ok && (bytes & 15u) == 0) {
		fence_async_smem();
		__syncthreads();
		if (tid == 0) tma_store_1d(gdst, stage, bytes);
	} else {
		__syncthreads();
		uint64_t *g = (uint64_t *)gdst;
		for (uint32_t i = tid; i < bytes / 8; i += nthr) g[i] = stage[i];
	}
}

// ------------------------------------------------------------------------------------------------
// Likelihood kernel: one site per thread, kCallTile sites per tile, persistent CTAs striding over the tiles.
//   in : pileup[n] (104 B each), ref[n] (codes 0..4)
//   out: VCF ? gt_vcf[n] (208 B, ready = 1) : gt_meth[n] (200 B) + skip[n]
// Per tile: the input records arrive by one TMA bulk copy that was issued while the previous tile was being computed;
// threads lift their record into registers, the next tile's copy is issued, the warp-cooperative model runs, records
// are staged in shared memory and leave by one TMA bulk store that drains while the next tile computes.
// ------------------------------------------------------------------------------------------------
constexpr int kCallTile = 128;

template <bool VCF, int MINB>
__global__ void __launch_bounds__(kCallTile, MINB)
k_call_sites(const uint8_t *__restrict__ pileup, const uint8_t *__restrict__ ref, size_t n,
		uint8_t *__restrict__ out, uint8_t *__restrict__ skip, const DevConst *__restrict__ dc, int bulk_ok,
		unsigned long long *__restrict__ counters) {
	constexpr int REC = VCF ? 208 : 200;
	constexpr int RW = REC / 8;
	extern __shared__ __align__(128) uint8_t smem_raw[];
	uint64_t *stage = (uint64_t *)smem_raw;                              // kCallTile output records
	uint8_t *inbuf = smem_raw + kCallTile * REC;                         // kCallTile input records
	Tables *tabs = (Tables *)(inbuf + kCallTile * 104);
	__shared__ uint64_t bar;

	const int tid = threadIdx.x;
	const size_t ntiles = (n + kCallTile - 1) / kCallTile;
	load_tables(tabs, dc, tid, kCallTile);
	if (tid == 0) mbar_init(&bar, 1);
	__syncthreads();

	auto tile_bulk = [&](size_t tile) {          // can this tile's input travel by the bulk engine?
		const size_t first = tile * kCallTile;
		const uint32_t bytes = (uint32_t)min((size_t)kCallTile, n - first) * 104u;
		return bulk_ok && (bytes & 15u) == 0;
	};
	auto issue = [&](size_t tile) {
		const size_t first = tile * kCallTile;
		const uint32_t bytes = (uint32_t)min((size_t)kCallTile, n - first) * 104u;
		mbar_expect_tx(&bar, bytes);
		tma_load_1d(inbuf, pileup + first * 104, bytes, &bar);
	};
	size_t tile = blockIdx.x;
	if (tile < ntiles && tid == 0 && tile_bulk(tile)) issue(tile);
	uint32_t phase = 0, ncalled = 0;
	for (; tile < ntiles; tile += gridDim.x) {
		const size_t first = tile * kCallTile;
		const int nrec = (int)min((size_t)kCallTile, n - first);
		const int rf = tid < nrec ? ref[first + tid] : 0;
		if (tile_bulk(tile)) {
			mbar_wait(&bar, phase);
			phase ^= 1;
		} else {
			const uint32_t *g = (const uint32_t *)(pileup + first * 104);
			uint32_t *sm = (uint32_t *)inbuf;
			for (int i = tid; i < nrec * 26; i += kCallTile) sm[i] = g[i];
			__syncthreads();
		}
		SiteCounts s;
		if (tid < nrec) {
			const uint2 *rec = (const uint2 *)(inbuf + tid * 104);
			uint32_t w[26];
#pragma unroll
			for (int i = 0; i < 13; i++) { const uint2 v = rec[i]; w[2 * i] = v.x; w[2 * i + 1] = v.y; }
#pragma unroll
			for (int j = 0; j < 8; j++) { s.cnt[0][j] = w[j]; s.cnt[1][j] = w[8 + j]; s.qsum[j] = __uint_as_float(w[17 + j]); }
			s.n = w[16];
			s.mapq2 = __uint_as_float(w[25]);
		} else {
#pragma unroll
			for (int j = 0; j < 8; j++) { s.cnt[0][j] = s.cnt[1][j] = 0; s.qsum[j] = 0.0f; }
			s.n = 0;
			s.mapq2 = 0.0f;
		}
		if (tid == 0) tma_store_wait();        // the previous tile's output has left the staging buffer
		__syncthreads();                       // every thread holds its input; staging buffer free
		const size_t next = tile + gridDim.x;
		if (next < ntiles && tid == 0 && tile_bulk(next)) issue(next);
		uint64_t *rec = stage + tid * RW;
		// pooled-argument list of this warp: the (not yet written) output rows of its own 32 sites
		double *wbuf = (double *)(stage + (tid & ~31) * RW);
		const bool called = call_site(s, rf, dc, tabs, rec, wbuf, tid & 31);
		ncalled += called;
		if (VCF) rec[25] = 1ull | ((called ? 0ull : 1ull) << 8);
		else if (tid < nrec) skip[first + tid] = called ? 0 : 1;
		store_tile<REC>(out + first * REC, stage, nrec, bulk_ok, tid, kCallTile);
	}
	if (tid == 0) tma_store_wait();
	ncalled = __reduce_add_sync(0xffffffffu, ncalled);
	if ((tid & 31) == 0 && ncalled) atomicAdd(counters, (unsigned long long)ncalled);
}

// ------------------------------------------------------------------------------------------------
// Segment binning: counting sort of segments by the 256-site tile their start falls in.
// ------------------------------------------------------------------------------------------------
constexpr int kPileTile = 256;


__global__ void k_bin_count(const Seg *__restrict__ segs, size_t nseg, uint32_t x, uint32_t ntiles, uint32_t *__restrict__ counts) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nseg) return;
	if (!segs[i].len) return;                 // empty slots (the normaliser leaves them for absent mates)
	const uint32_t pos = segs[i].pos;
	// a segment that starts before the window still contributes to tile 0 (clipped there)
	uint32_t t = pos >= x ? (pos - x) / kPileTile : 0;
	if (t < ntiles) atomicAdd(counts + t, 1u);
}

// single-CTA exclusive scan (ntiles is at most a few hundred thousand per launch; runs once per block of sites)
__global__ void __launch_bounds__(1024) k_bin_scan(const uint32_t *__restrict__ counts, uint32_t ntiles, uint32_t *__restrict__ start, uint32_t *__restrict__ cursor) {
	__shared__ uint32_t warp_tot[32];
	__shared__ uint32_t carry;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	if (tid == 0) carry = 0;
	__syncthreads();
	for (uint32_t base = 0; base < ntiles; base += 1024) {
		const uint32_t i = base + tid;
		const uint32_t v = i < ntiles ? counts[i] : 0;
		uint32_t incl = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
		if (lane == 31) warp_tot[wid] = incl;
		__syncthreads();
		if (wid == 0) {
			uint32_t w = warp_tot[lane];
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += o; }
			warp_tot[lane] = w;
		}
		__syncthreads();
		const uint32_t excl = carry + (wid ? warp_tot[wid - 1] : 0) + incl - v;
		if (i < ntiles) { start[i] = excl; cursor[i] = excl; }
		__syncthreads();
		if (tid == 1023) carry = excl + v;
		__syncthreads();
	}
	if (tid == 0) start[ntiles] = carry;
}

__global__ void k_bin_scatter(const Seg *__restrict__ segs, size_t nseg, uint32_t x, uint32_t ntiles, uint32_t *__restrict__ cursor, Seg *__restrict__ sorted) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nseg) return;
	const Seg s = segs[i];
	if (!s.len) return;
	uint32_t t = s.pos >= x ? (s.pos - x) / kPileTile : 0;
	if (t < ntiles) sorted[atomicAdd(cursor + t, 1u)] = s;
}

// ------------------------------------------------------------------------------------------------
// Pileup by gather.  CTA = 128 threads = 128 consecutive sites (half a bin).  Candidate segments are the two bins
// [t-1, t] (a segment is at most 256 long).  Each warp filters 32 candidates at a time with one ballot, then walks the
// hits; a lane adds the byte at its own site into halfword-packed 64-bit register counters keyed by
// (strand index, bisulfite strand), which are widened every 1500 hits.
//   MODE 0: write pileup[] (104 B / site)       MODE 1: run the model and write gt_vcf[] (208 B / site)
// ------------------------------------------------------------------------------------------------
struct Packed {
	// four 16-bit fields indexed by base, kept as two 32-bit halves (bases A,C | G,T) so that adds and multiply-adds
	// never need a carry across the halves
	uint32_t c[2][3][2];   // [ori][bs_strand][half] : counts
	uint32_t q[3][2];      // [bs_strand][half]      : quality sums (q <= 43: good for 1524 hits)
};
constexpr uint32_t kWidenEvery = 1500;

__device__ __forceinline__ void widen(Packed &p, uint32_t cnt[2][8], uint32_t qs[8]) {
	// class of (bs_strand, base): st0 {0,1,2,3}  st1=C2T {0,5,2,7}  st2=G2A {4,1,6,3}   (src/call_genotypes.c:17-19)
	constexpr int cls[3][4] = {{0, 1, 2, 3}, {0, 5, 2, 7}, {4, 1, 6, 3}};
#pragma unroll
	for (int st = 0; st < 3; st++) {
#pragma unroll
		for (int b = 0; b < 4; b++) {
#pragma unroll
			for (int o = 0; o < 2; o++) cnt[o][cls[st][b]] += (p.c[o][st][b >> 1] >> (16 * (b & 1))) & 0xffffu;
			qs[cls[st][b]] += (p.q[st][b >> 1] >> (16 * (b & 1))) & 0xffffu;
		}
		p.c[0][st][0] = p.c[0][st][1] = p.c[1][st][0] = p.c[1][st][1] = 0;
		p.q[st][0] = p.q[st][1] = 0;
	}
}

constexpr int kPileThreads = 128;     // one CTA = 128 consecutive sites = half a bin

template <int MODE>
__global__ void __launch_bounds__(kPileThreads)
k_pileup_tile(const Seg *__restrict__ segs, const uint32_t *__restrict__ bin_start, const uint8_t *__restrict__ bases,
		const uint8_t *__restrict__ ref, uint32_t x, uint32_t sz, uint32_t tile0, uint8_t *__restrict__ out,
		const DevConst *__restrict__ dc, unsigned long long *__restrict__ counters) {
	constexpr int REC = MODE ? 208 : 104;
	constexpr int RW = REC / 8;
	extern __shared__ __align__(128) uint8_t smem_raw[];
	uint64_t *stage = (uint64_t *)smem_raw;
	Tables *tabs = (Tables *)(smem_raw + kPileThreads * REC);
	__shared__ uint4 cand[kPileThreads];      // {pos, off, len | st << 16 | ori << 18, mapq^2}

	const int tid = threadIdx.x, lane = tid & 31;
	const uint32_t site0 = tile0 * kPileTile + blockIdx.x * kPileThreads;     // first site of this CTA
	const uint32_t bin = site0 / kPileTile;
	const int nrec = (int)min((uint32_t)kPileThreads, sz - site0);
	const uint32_t mypos = x + site0 + tid;                        // 1-based reference position of this thread's site
	const uint32_t wpos0 = x + site0 + (tid & ~31);                // first position of this warp's 32 sites
	const uint32_t min_qual = (uint32_t)dc->min_qual;
	if (MODE) load_tables(tabs, dc, tid, kPileThreads);

	uint32_t cnt[2][8], qs[8], mq2 = 0;
#pragma unroll
	for (int j = 0; j < 8; j++) cnt[0][j] = cnt[1][j] = qs[j] = 0;
	Packed pk;
#pragma unroll
	for (int st = 0; st < 3; st++) { pk.c[0][st][0] = pk.c[0][st][1] = pk.c[1][st][0] = pk.c[1][st][1] = 0; pk.q[st][0] = pk.q[st][1] = 0; }
	uint32_t since_widen = 0;

	// candidates: segments that start in the previous bin or in this one (a segment is at most one bin long)
	const uint32_t c_lo = bin_start[bin ? bin - 1 : 0], c_hi = bin_start[bin + 1];
	for (uint32_t base = c_lo; base < c_hi; base += kPileThreads) {
		const uint32_t nc = min((uint32_t)kPileThreads, c_hi - base);
		__syncthreads();
		if ((uint32_t)tid < nc) {
			const Seg sg = segs[base + tid];
			const uint32_t mq = sg.mapq;
			cand[tid] = make_uint4(sg.pos, sg.off, (uint32_t)sg.len | ((uint32_t)(sg.flags >> 1) & 3u) << 16 | ((uint32_t)sg.flags & 1u) << 18, mq * mq);
		}
		__syncthreads();
		for (uint32_t g = 0; g < nc; g += 32) {
			bool hit = false;
			if (g + lane < nc) {
				const uint4 c = cand[g + lane];
				hit = c.x < wpos0 + 32 && c.x + (c.z & 0xffffu) > wpos0;
			}
			uint32_t m = __ballot_sync(0xffffffffu, hit);
			while (m) {
				// four hits per trip: their byte loads are issued back to back before any is consumed
				uint4 raw[4];
				uint32_t byte[4];
#pragma unroll
				for (int u = 0; u < 4; u++) {
					const bool valid = m != 0;
					const int b = valid ? __ffs(m) - 1 : 0;
					m &= m - 1;                                       // 0 stays 0
					raw[u] = cand[g + b];                             // broadcast read
					if (!valid) raw[u].z = 0;                         // len 0: contributes nothing
				}
#pragma unroll
				for (int u = 0; u < 4; u++) {
					const uint32_t d = mypos - raw[u].x;
					byte[u] = d < (raw[u].z & 0xffffu) ? (uint32_t)__ldg(bases + (size_t)raw[u].y + d) : 0u;
				}
#pragma unroll
				for (int u = 0; u < 4; u++) {
					const uint32_t q = byte[u] >> 2;
					const uint32_t ok = (q - min_qual) < ((uint32_t)kFltQual - min_qual) ? 1u : 0u;   // min_qual <= q < 63 (q = 0: no byte)
					const uint32_t v = ok << ((byte[u] & 1u) << 4);                   // +1 in the 16-bit field of this base ...
					const uint32_t lo = (byte[u] & 2u) ? 0u : v, hi = v - lo;         // ... of the half that holds it
					const uint32_t meta = raw[u].z;
					// (st, ori) are warp-uniform: uniform branches pick the packed registers
					switch ((meta >> 16) & 3u) {
					case 0:
						pk.q[0][0] += lo * q; pk.q[0][1] += hi * q;
						if (meta >> 18) { pk.c[1][0][0] += lo; pk.c[1][0][1] += hi; } else { pk.c[0][0][0] += lo; pk.c[0][0][1] += hi; }
						break;
					case 1:
						pk.q[1][0] += lo * q; pk.q[1][1] += hi * q;
						if (meta >> 18) { pk.c[1][1][0] += lo; pk.c[1][1][1] += hi; } else { pk.c[0][1][0] += lo; pk.c[0][1][1] += hi; }
						break;
					default:
						pk.q[2][0] += lo * q; pk.q[2][1] += hi * q;
						if (meta >> 18) { pk.c[1][2][0] += lo; pk.c[1][2][1] += hi; } else { pk.c[0][2][0] += lo; pk.c[0][2][1] += hi; }
						break;
					}
					mq2 += ok * raw[u].w;
				}
				since_widen += 4;
				if (since_widen > kWidenEvery) { widen(pk, cnt, qs); since_widen = 0; }
			}
		}
	}
	widen(pk, cnt, qs);

	SiteCounts s;
	uint32_t n = 0, qmax = 0;
#pragma unroll
	for (int j = 0; j < 8; j++) {
		s.cnt[0][j] = cnt[0][j]; s.cnt[1][j] = cnt[1][j];
		n += cnt[0][j] + cnt[1][j];
		s.qsum[j] = (float)qs[j];
		qmax = max(qmax, qs[j]);
	}
	s.n = tid < nrec ? n : 0;
	s.mapq2 = (float)mq2;
	// integer sums equal the reference's float sums only below 2^24: count the sites that leave the envelope
	if (s.n && (qmax >= (1u << 24) || mq2 >= (1u << 24))) atomicAdd(counters + 1, 1ull);

	uint64_t *rec = stage + tid * RW;
	if (MODE == 0) {
		uint32_t *w = (uint32_t *)rec;
#pragma unroll
		for (int j = 0; j < 8; j++) { w[j] = s.cnt[0][j]; w[8 + j] = s.cnt[1][j]; w[17 + j] = __float_as_uint(s.qsum[j]); }
		w[16] = s.n;
		w[25] = __float_as_uint(s.mapq2);
	} else {
		const int rf = tid < nrec ? ref[site0 + tid] : 0;
		__syncthreads();                    // tables loaded
		double *wbuf = (double *)(stage + (tid & ~31) * RW);
		const bool called = call_site(s, rf, dc, tabs, rec, wbuf, lane);
		rec[25] = 1ull | ((called ? 0ull : 1ull) << 8);
		const uint32_t nc = __syncthreads_count(called);
		if (tid == 0 && nc) atomicAdd(counters, (unsigned long long)nc);
	}
	store_tile<REC>(out + (size_t)blockIdx.x * kPileThreads * REC, stage, nrec, true, tid, kPileThreads);
	if (tid == 0) tma_store_wait();
}

// ------------------------------------------------------------------------------------------------
// Synthetic workloads (counter-based: record i depends only on (seed, i))
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
	z += 0x9e3779b97f4a7c15ull;
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
	return z ^ (z >> 31);
}
struct Rng {     // splitmix64 stream keyed by (seed, index)
	uint64_t s;
	__device__ Rng(uint64_t seed, uint64_t idx) : s(mix64(seed ^ mix64(idx))) {}
	__device__ __forceinline__ uint64_t next() { s += 0x9e3779b97f4a7c15ull; uint64_t z = s; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
	__device__ __forceinline__ float unif() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }
};

// Per-site count vectors (config 2): depth ~ Poisson(mean), bases drawn from the site genotype seen through bisulfite
// conversion on a random strand; 3 % empty sites; q in [20,43]; MAPQ 60.
__global__ void k_synth_sites(uint64_t seed, uint64_t first, size_t n, float mean_depth, uint8_t *__restrict__ pileup, uint8_t *__restrict__ ref) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	Rng r(seed, first + i);
	// reference base A/C/G/T = .295/.205/.205/.295
	float u = r.unif();
	const int rb = u < 0.295f ? 0 : (u < 0.5f ? 1 : (u < 0.705f ? 2 : 3));
	int a0 = rb, a1 = rb;
	u = r.unif();
	if (u < 0.001f) a1 = (rb + 1 + (int)(r.next() % 3)) & 3;                       // het
	else if (u < 0.0015f) a0 = a1 = (rb + 1 + (int)(r.next() % 3)) & 3;            // hom alt
	const float meth = r.unif() < 0.02f ? 0.7f : 0.01f;
	int depth = 0;
	if (r.unif() >= 0.03f) {          // Poisson by inversion
		double p = exp(-(double)mean_depth), c = p;
		const double uu = (double)(r.next() >> 11) * (1.0 / 9007199254740992.0);
		while (uu > c && depth < 1000) { depth++; p *= (double)mean_depth / depth; c += p; }
	}
	uint32_t cnt[2][8];
	uint32_t qs[8];
#pragma unroll
	for (int j = 0; j < 8; j++) cnt[0][j] = cnt[1][j] = qs[j] = 0;
	for (int k = 0; k < depth; k++) {
		const uint64_t bits = r.next();
		int b = (bits & 1) ? a1 : a0;
		const int st = 1 + (int)((bits >> 1) & 1), ori = (int)((bits >> 2) & 1);
		const float uc = (float)((bits >> 8) & 0xffffff) * (1.0f / 16777216.0f);
		const bool converts = uc >= meth && uc < meth + (1.0f - meth) * 0.99f;
		if (st == 1 && b == 1 && converts) b = 3;
		if (st == 2 && b == 2 && converts) b = 0;
		if (((bits >> 32) & 0x3ff) < 3) b = (b + 1 + (int)((bits >> 42) % 3)) & 3;  // ~0.3 % errors
		const int q = 20 + (int)((bits >> 48) % 24);
		const int cl = st == 1 ? (b == 1 ? 5 : (b == 3 ? 7 : b)) : (b == 0 ? 4 : (b == 2 ? 6 : b));
		// dynamic index into small local arrays: fine for a generator
		cnt[ori][cl]++;
		qs[cl] += q;
	}
	uint32_t *w = (uint32_t *)(pileup + i * 104);
	for (int j = 0; j < 8; j++) { w[j] = cnt[0][j]; w[8 + j] = cnt[1][j]; w[17 + j] = __float_as_uint((float)qs[j]); }
	w[16] = (uint32_t)depth;
	w[25] = __float_as_uint(3600.0f * (float)depth);
	ref[i] = (uint8_t)(rb + 1);
}

// reference code of position pos for the synthetic genome: a pure function of (seed, pos); ~0.2 % N
__device__ __forceinline__ int synth_ref_code(uint64_t seed, uint32_t pos) {
	const uint64_t h = mix64(seed ^ (0x5851f42d4c957f2dull * (uint64_t)pos));
	const uint32_t u = (uint32_t)(h >> 40);                     // 24 bits
	if ((h & 0x1ff) == 0) return 0;
	return u < 4949279u ? 1 : (u < 8388608u ? 2 : (u < 11827937u ? 3 : 4));
}

__global__ void k_synth_ref(uint64_t seed, uint32_t x, uint32_t sz, uint8_t *__restrict__ ref) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < sz) ref[i] = (uint8_t)synth_ref_code(seed, x + (uint32_t)i);
}

// one warp per read: lanes write consecutive bytes of the read
__global__ void k_synth_reads(uint64_t seed, uint32_t x, uint32_t sz, uint32_t read_len, double step, size_t nseg,
		Seg *__restrict__ segs, uint8_t *__restrict__ bases) {
	const size_t wi = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (wi >= nseg) return;
	Rng r(seed ^ 0xabcdef12345ull, wi);
	const uint64_t h = r.next();
	uint32_t len = read_len;
	if ((h & 0xff) < 13) len = read_len - (uint32_t)((h >> 8) % (read_len / 3 + 1));          // ~5 % shorter reads
	const uint32_t pos = x + 2 + (uint32_t)((double)wi * step) + (uint32_t)((h >> 20) & 7);
	const uint32_t st = 1 + (uint32_t)((h >> 24) & 1), ori = (uint32_t)((h >> 25) & 1), hap = (uint32_t)((h >> 26) & 1);
	const uint32_t mapq = ((h >> 28) & 0xf) ? 60 : 20 + (uint32_t)((h >> 32) % 40);
	const size_t off = wi * (size_t)read_len;
	if (lane == 0) {
		Seg s;
		s.pos = pos; s.off = (uint32_t)off; s.len = (uint16_t)len; s.mapq = (uint8_t)mapq; s.flags = (uint8_t)(ori | (st << 1)); s.pad = 0;
		segs[wi] = s;
	}
	for (uint32_t j = lane; j < read_len; j += 32) {
		uint8_t byte = 0;
		if (j < len) {
			const uint32_t p = pos + j;
			const int rc = synth_ref_code(seed, p);
			const uint64_t g = mix64(seed ^ 0x77aa55ull ^ ((uint64_t)p << 20));            // site genotype
			int b = rc ? rc - 1 : (int)(g & 3);
			const uint32_t gsel = (uint32_t)(g >> 40) % 3000u;
			if (gsel < 2) { if (gsel == 0 || hap) b = (b + 1 + (int)((g >> 8) % 3)) & 3; }      // hom-alt (0) / het (1)
			const uint64_t e = mix64(h ^ ((uint64_t)j * 0x9e3779b97f4a7c15ull));
			const float uc = (float)(e >> 40) * (1.0f / 16777216.0f);
			const bool cpg = st == 1 ? (rc == 2 && synth_ref_code(seed, p + 1) == 3) : (rc == 3 && synth_ref_code(seed, p - 1) == 2);
			const float meth = cpg ? 0.7f : 0.01f;
			const bool converts = uc >= meth && uc < meth + (1.0f - meth) * 0.99f;
			if (st == 1 && b == 1 && converts) b = 3;
			if (st == 2 && b == 2 && converts) b = 0;
			uint32_t q = ((e >> 8) & 0xff) < 218 ? 37u : 10u + (uint32_t)((e >> 16) % 31);
			const float perr = __expf(-0.2302585f * (float)q);
			if ((float)((e >> 24) & 0xffff) * (1.0f / 65536.0f) < perr) b = (b + 1 + (int)((e >> 4) % 3)) & 3;
			byte = (uint8_t)(b | (q << 2));
			if (((e >> 44) & 0x3ff) == 0) byte = 0;                                     // N
		}
		bases[off + j] = byte;
	}
}

// ------------------------------------------------------------------------------------------------
// launchers (called from bsgpu_api.cpp; all asynchronous on `stream`)
// ------------------------------------------------------------------------------------------------
#define LAUNCH_CHECK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return e_; } while (0)

static size_t call_smem(bool vcf) { return (size_t)kCallTile * ((vcf ? 208 : 200) + 104) + sizeof(Tables); }
static size_t pile_smem(int mode) { return (size_t)kPileThreads * (mode ? 208 : 104) + sizeof(Tables); }

static int g_sms = 148;
static int g_call_minb = 5;          // resident CTAs per SM the likelihood kernel is compiled for (5: 96 regs; 4: 120 regs)
static int g_call_ctas[2][2];        // [minb == 5][vcf] measured occupancy

template <typename K>
static cudaError_t prep(K kernel, size_t smem, int threads, int *ctas_per_sm) {
	cudaError_t e;
	if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
	if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
	if (ctas_per_sm) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, kernel, threads, smem);
	return cudaSuccess;
}

cudaError_t configure_kernels() {
	cudaError_t e;
	int dev = 0;
	if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
	if ((e = cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
	const char *env = getenv("BSGPU_CALL_MINB");          // tuning knob: 5 (default, measured faster) or 4 resident CTAs per SM
	if (env && atoi(env) == 4) g_call_minb = 4;
	if ((e = prep(k_call_sites<false, 4>, call_smem(false), kCallTile, &g_call_ctas[0][0])) != cudaSuccess) return e;
	if ((e = prep(k_call_sites<true, 4>, call_smem(true), kCallTile, &g_call_ctas[0][1])) != cudaSuccess) return e;
	if ((e = prep(k_call_sites<false, 5>, call_smem(false), kCallTile, &g_call_ctas[1][0])) != cudaSuccess) return e;
	if ((e = prep(k_call_sites<true, 5>, call_smem(true), kCallTile, &g_call_ctas[1][1])) != cudaSuccess) return e;
	if ((e = prep(k_pileup_tile<0>, pile_smem(0), kPileThreads, nullptr)) != cudaSuccess) return e;
	if ((e = prep(k_pileup_tile<1>, pile_smem(1), kPileThreads, nullptr)) != cudaSuccess) return e;
	return cudaSuccess;
}

cudaError_t launch_call_sites(const void *pileup, const void *ref, size_t n, void *out, void *skip, bool vcf,
		const DevConst *dc, unsigned long long *counters, cudaStream_t stream, int *launches) {
	if (!n) return cudaSuccess;
	const bool bulk_ok = (((uintptr_t)pileup | (uintptr_t)out) & 15u) == 0;
	const size_t ntiles = (n + kCallTile - 1) / kCallTile;
	const int five = g_call_minb == 5;
	const size_t resident = (size_t)g_sms * (size_t)(g_call_ctas[five][vcf] > 0 ? g_call_ctas[five][vcf] : 1);
	const unsigned grid = (unsigned)(ntiles < resident ? ntiles : resident);       // persistent: one CTA per resident slot
	const uint8_t *p = (const uint8_t *)pileup, *r = (const uint8_t *)ref;
	uint8_t *o = (uint8_t *)out, *sk = (uint8_t *)skip;
	if (vcf) {
		if (five) k_call_sites<true, 5><<<grid, kCallTile, call_smem(true), stream>>>(p, r, n, o, nullptr, dc, bulk_ok, counters);
		else k_call_sites<true, 4><<<grid, kCallTile, call_smem(true), stream>>>(p, r, n, o, nullptr, dc, bulk_ok, counters);
	} else {
		if (five) k_call_sites<false, 5><<<grid, kCallTile, call_smem(false), stream>>>(p, r, n, o, sk, dc, bulk_ok, counters);
		else k_call_sites<false, 4><<<grid, kCallTile, call_smem(false), stream>>>(p, r, n, o, sk, dc, bulk_ok, counters);
	}
	*launches += 1;
	LAUNCH_CHECK();
	return cudaSuccess;
}

static size_t seg_area(size_t nseg) { return (nseg * sizeof(Seg) + 255) & ~(size_t)255; }

size_t pileup_scratch_bytes(size_t nseg, uint32_t sz) {
	const size_t ntiles = (sz + kPileTile - 1) / kPileTile;
	return seg_area(nseg) + (3 * ntiles + 8) * sizeof(uint32_t) + 256;
}

// scratch layout: sorted segs | counts[ntiles] | start[ntiles+1] | cursor[ntiles]
cudaError_t launch_bin_segments(const void *segs, size_t nseg, uint32_t x, uint32_t sz, void *scratch,
		cudaStream_t stream, int *launches) {
	if (!sz) return cudaSuccess;
	const uint32_t ntiles = (sz + kPileTile - 1) / kPileTile;
	Seg *sorted = (Seg *)scratch;
	uint32_t *counts = (uint32_t *)((uint8_t *)scratch + seg_area(nseg));
	uint32_t *start = counts + ntiles;
	uint32_t *cursor = start + ntiles + 1;
	cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(uint32_t) * ntiles, stream);
	if (e != cudaSuccess) return e;
	const unsigned g = (unsigned)((nseg + 255) / 256);
	if (nseg) {
		k_bin_count<<<g, 256, 0, stream>>>((const Seg *)segs, nseg, x, ntiles, counts);
		*launches += 1;
		LAUNCH_CHECK();
	}
	k_bin_scan<<<1, 1024, 0, stream>>>(counts, ntiles, start, cursor);
	*launches += 1;
	LAUNCH_CHECK();
	if (nseg) {
		k_bin_scatter<<<g, 256, 0, stream>>>((const Seg *)segs, nseg, x, ntiles, cursor, sorted);
		*launches += 1;
		LAUNCH_CHECK();
	}
	return cudaSuccess;
}

cudaError_t launch_pileup_tiles(const void *scratch, size_t nseg, const void *bases, const void *ref, uint32_t x,
		uint32_t sz, uint32_t tile0, uint32_t ntiles, void *out, int mode, const DevConst *dc,
		unsigned long long *counters, cudaStream_t stream, int *launches) {
	if (!ntiles) return cudaSuccess;
	const uint32_t all_tiles = (sz + kPileTile - 1) / kPileTile;
	const Seg *sorted = (const Seg *)scratch;
	const uint32_t *start = (const uint32_t *)((const uint8_t *)scratch + seg_area(nseg)) + all_tiles;
	// sites covered by this launch: bins [tile0, tile0 + ntiles) clipped to the window; one CTA per 128 sites
	const uint32_t first_site = tile0 * kPileTile;
	const uint32_t nsite = min(ntiles * (uint32_t)kPileTile, sz - first_site);
	const unsigned grid = (nsite + kPileThreads - 1) / kPileThreads;
	if (mode) k_pileup_tile<1><<<grid, kPileThreads, pile_smem(1), stream>>>(sorted, start, (const uint8_t *)bases, (const uint8_t *)ref, x, sz, tile0, (uint8_t *)out, dc, counters);
	else k_pileup_tile<0><<<grid, kPileThreads, pile_smem(0), stream>>>(sorted, start, (const uint8_t *)bases, (const uint8_t *)ref, x, sz, tile0, (uint8_t *)out, dc, counters);
	*launches += 1;
	LAUNCH_CHECK();
	return cudaSuccess;
}

cudaError_t launch_synth_sites(uint64_t seed, uint64_t first, size_t n, double mean_depth, void *pileup, void *ref,
		cudaStream_t stream, int *launches) {
	if (!n) return cudaSuccess;
	k_synth_sites<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(seed, first, n, (float)mean_depth, (uint8_t *)pileup, (uint8_t *)ref);
	*launches += 1;
	LAUNCH_CHECK();
	return cudaSuccess;
}

size_t synth_block_nseg(uint32_t sz, uint32_t read_len, double depth) {
	if (sz < read_len + 16) return 0;
	return (size_t)((double)(sz - read_len - 12) * depth / (double)read_len);
}

cudaError_t launch_synth_block(uint64_t seed, uint32_t x, uint32_t sz, uint32_t read_len, double depth,
		void *segs, void *bases, void *ref, cudaStream_t stream, int *launches) {
	const size_t nseg = synth_block_nseg(sz, read_len, depth);
	k_synth_ref<<<(sz + 255) / 256, 256, 0, stream>>>(seed, x, sz, (uint8_t *)ref);
	*launches += 1;
	LAUNCH_CHECK();
	if (nseg) {
		const double step = (double)read_len / depth;
		k_synth_reads<<<(unsigned)((nseg * 32 + 255) / 256), 256, 0, stream>>>(seed, x, sz, read_len, step, nseg, (Seg *)segs, (uint8_t *)bases);
		*launches += 1;
		LAUNCH_CHECK();
	}
	return cudaSuccess;
}

}  // namespace bsgpu
