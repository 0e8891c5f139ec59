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
__device__ __forceinline__ void tma_load_1d_s32(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void *gdst, const void *smem_src, uint32_t bytes) {
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16-byte asynchronous copies global -> shared through the load / store units (LDGSTS), grouped per thread
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void *gsrc) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

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

// a site whose result sits inside a guard band: counted, and listed while the list has room
// (tie: 1 = closer than the band, 2 = equal; both are listed as kind 1.  Rare: kept out of line)
__device__ __forceinline__ void guard_flag(unsigned long long *counters, int tie, unsigned long long id) {
	atomicAdd(counters + (tie == 1 ? 4 : 5), 1ull);
	const unsigned long long k = atomicAdd(counters + 8, 1ull);
	if (k < (unsigned long long)kGuardCap) counters[kGuardList + k] = 1ull << 56 | (id & 0x00ffffffffffffffull);
}

// MODE fixes how tiles travel at compile time (the launcher picks it from the transport bits): 1 = bits 3 (bulk load and bulk
// store), 2 = bits 10 (cp.async prefetch, bulk store), 0 = whatever `bulk_arg` says at run time.  With both prefetch paths live
// in one body the loop carried the other path's address arithmetic and predicates (13.1 -> 12.8 G sites/s when the cp.async
// variant was added); the specialised bodies only hold their own.
template <bool VCF, int MINB, int MODE>
__global__ void __launch_bounds__(kCallTile, MINB)
k_call_sites(const uint8_t *__restrict__ pileup, const uint8_t *__restrict__ ref, size_t n,
		uint8_t *__restrict__ out, uint8_t *__restrict__ skip, const DevConst *__restrict__ dc, int bulk_arg,
		unsigned long long *__restrict__ counters, unsigned long long guard_base) {
	const int bulk_ok = MODE == 1 ? 3 : MODE == 2 ? 10 : bulk_arg;
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
		return (bulk_ok & 1) && (bytes & 15u) == 0;
	};
	auto issue = [&](size_t tile) {
		const size_t first = tile * kCallTile;
		const uint32_t bytes = (uint32_t)min((size_t)kCallTile, n - first) * 104u;
		mbar_expect_tx(&bar, bytes);
		tma_load_1d(inbuf, pileup + first * 104, bytes, &bar);
	};
	// the same prefetch without the bulk-copy engine (bulk_ok & 8): every thread copies its share of the tile, 16 bytes at a time
	const bool ldgsts = (bulk_ok & 8) != 0;
	auto tile_async = [&](size_t tile) {
		const size_t first = tile * kCallTile;
		const uint32_t bytes = (uint32_t)min((size_t)kCallTile, n - first) * 104u;
		return ldgsts && (bytes & 15u) == 0;
	};
	auto issue_async = [&](size_t tile) {
		const size_t first = tile * kCallTile;
		const uint32_t chunks = (uint32_t)min((size_t)kCallTile, n - first) * 104u / 16u;
		const uint8_t *src = pileup + first * 104;
		const uint32_t dst = smem_u32(inbuf);
		for (uint32_t k = tid; k < chunks; k += kCallTile) cp_async16(dst + 16u * k, src + 16u * (size_t)k);
		cp_async_commit();
	};
	size_t tile = blockIdx.x;
	if (tile < ntiles && tid == 0 && tile_bulk(tile)) issue(tile);
	if (tile < ntiles && tile_async(tile)) issue_async(tile);
	uint32_t phase = 0, ncalled = 0;
	for (; tile < ntiles; tile += gridDim.x) {
		const size_t first = tile * kCallTile;
		const int nrec = (int)min((size_t)kCallTile, n - first);
		const int rf = tid < nrec ? ref[first + tid] : 0;
		if (tile_bulk(tile)) {
			mbar_wait(&bar, phase);
			phase ^= 1;
		} else if (tile_async(tile)) {
			cp_async_wait();
			__syncthreads();
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
		if (tid == 0) { if (bulk_ok & 4) tma_store_wait_all(); else tma_store_wait(); }        // the previous tile's output has left the staging buffer
		__syncthreads();                       // every thread holds its input; staging buffer free
		const size_t next = tile + gridDim.x;
		if (next < ntiles && tid == 0 && tile_bulk(next)) issue(next);
		if (next < ntiles && tile_async(next)) issue_async(next);
		uint64_t *rec = stage + tid * RW;
		// pooled-argument list of this warp: the (not yet written) output rows of its own 32 sites
		double *wbuf = (double *)(stage + (tid & ~31) * RW);
		const bool called = call_site(s, rf, dc, tabs, rec, wbuf, tid & 31);
		ncalled += called;
		if (called && (rec[24] >> 8)) {        // guard flag of the call (rare): report it, give the padding byte back
			guard_flag(counters, (int)(rec[24] >> 8) & 3, guard_base + first + tid);
			rec[24] &= 0xffull;
		}
		if (VCF) rec[25] = 1ull | ((called ? 0ull : 1ull) << 8);
		else if (tid < nrec) skip[first + tid] = called ? 0 : 1;
		store_tile<REC>(out + first * REC, stage, nrec, (bulk_ok & 2) != 0, tid, kCallTile);
	}
	if (tid == 0) { if (bulk_ok & 4) tma_store_wait_all(); else tma_store_wait(); }
	ncalled = __reduce_add_sync(0xffffffffu, ncalled);
	if ((tid & 31) == 0 && ncalled) atomicAdd(counters, (unsigned long long)ncalled);
}

// ------------------------------------------------------------------------------------------------
// Segment binning: counting sort of segments by the 128-site tile their start falls in (count, 3-phase exclusive scan,
// scatter).  The scatter writes the pre-digested candidate record the gather kernel consumes.
// ------------------------------------------------------------------------------------------------
constexpr int kPileTile = kPileTileSites;      // 128 sites = one CTA of the gather kernel
constexpr int kBinsBack = (kMaxSegLen + kPileTile - 2) / kPileTile;   // bins before tile T that can reach into it

// {pos, off - pos (mod 2^32), len | combo << 16, mapq^2}; combo = 2 * bisulfite strand + strand index (0..5)
struct Cand { uint32_t pos, offd, meta, mq2; };

__device__ __forceinline__ uint32_t seg_bin(uint32_t pos, uint32_t x) { return pos >= x ? (pos - x) / kPileTile : 0; }

// Input segments usually arrive in coordinate order, so the lanes of a warp mostly fall into one or two tiles: lanes
// with the same tile are grouped with match.any and one of them does the atomic for the group (a deep panel puts
// hundreds of segments into every tile).
__global__ void k_bin_count(const Seg *__restrict__ segs, size_t nseg, uint32_t x, uint32_t ntiles, uint32_t *__restrict__ counts,
		unsigned long long *__restrict__ counters) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t t = 0xffffffffu;                 // no contribution: out of range, empty slot (the normaliser leaves them for absent mates)
	if (i < nseg) {
		const Seg s = segs[i];
		if (s.len > kMaxSegLen && counters) atomicAdd(counters + 9, 1ull);      // contract of the _dev entry points (bsgpu_stats.long_segments)
		// a segment that starts before the window still contributes to the first tiles (clipped there)
		if (s.len) { const uint32_t b = seg_bin(s.pos, x); if (b < ntiles) t = b; }
	}
	const uint32_t peers = __match_any_sync(0xffffffffu, t);
	if (t != 0xffffffffu && (threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(counts + t, (uint32_t)__popc(peers));
}

// exclusive scan of counts[0..n) in three phases: per-CTA scan of 1024 items, single-CTA scan of the CTA totals, add.
__device__ __forceinline__ uint32_t cta_scan_1024(uint32_t v, uint32_t *warp_tot, int tid, uint32_t *total) {
	const int lane = tid & 31, wid = tid >> 5;
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
	const uint32_t excl = (wid ? warp_tot[wid - 1] : 0) + incl - v;
	*total = warp_tot[31];
	__syncthreads();
	return excl;
}

__global__ void __launch_bounds__(1024) k_scan_local(const uint32_t *__restrict__ counts, uint32_t n, uint32_t *__restrict__ start, uint32_t *__restrict__ partial) {
	__shared__ uint32_t warp_tot[32];
	const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
	uint32_t total;
	const uint32_t excl = cta_scan_1024(i < n ? counts[i] : 0, warp_tot, threadIdx.x, &total);
	if (i < n) start[i] = excl;
	if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_partials(uint32_t *__restrict__ partial, uint32_t nb) {
	__shared__ uint32_t warp_tot[32];
	uint32_t carry = 0;
	for (uint32_t base = 0; base < nb; base += 1024) {
		const uint32_t i = base + threadIdx.x;
		uint32_t total;
		const uint32_t excl = cta_scan_1024(i < nb ? partial[i] : 0, warp_tot, threadIdx.x, &total);
		if (i < nb) partial[i] = carry + excl;
		carry += total;
	}
	if (threadIdx.x == 0) partial[nb] = carry;
}

__global__ void __launch_bounds__(1024) k_scan_add(uint32_t *__restrict__ start, uint32_t *__restrict__ cursor, uint32_t n, const uint32_t *__restrict__ partial, uint32_t nb) {
	const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
	if (i < n) { const uint32_t v = start[i] + partial[blockIdx.x]; start[i] = v; cursor[i] = v; }
	if (i == 0) start[n] = partial[nb];
}

__global__ void k_bin_scatter(const Seg *__restrict__ segs, size_t nseg, uint32_t x, uint32_t ntiles, uint32_t *__restrict__ cursor, Cand *__restrict__ sorted) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t t = 0xffffffffu;
	Seg s;
	s.pos = 0; s.off = 0; s.len = 0; s.mapq = 0; s.flags = 0; s.pad = 0;
	if (i < nseg) {
		s = segs[i];
		if (s.len) { const uint32_t b = seg_bin(s.pos, x); if (b < ntiles) t = b; }
	}
	// one atomic per group of lanes that share a tile; the leader's old value is the group's first slot
	const uint32_t peers = __match_any_sync(0xffffffffu, t);
	const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
	uint32_t first = 0;
	if (t != 0xffffffffu && lane == leader) first = atomicAdd(cursor + t, (uint32_t)__popc(peers));
	first = __shfl_sync(0xffffffffu, first, leader);
	if (t == 0xffffffffu) return;
	const uint32_t mq = s.mapq, combo = ((uint32_t)(s.flags >> 1) & 3u) * 2u + ((uint32_t)s.flags & 1u);
	Cand c;
	c.pos = s.pos; c.offd = s.off - s.pos; c.meta = (uint32_t)s.len | combo << 16; c.mq2 = mq * mq;
	sorted[first + __popc(peers & ((1u << lane) - 1u))] = c;
}

// ------------------------------------------------------------------------------------------------
// Pileup by gather.  CTA = 128 threads = 128 consecutive sites (one bin); one thread owns one site and pulls the one
// byte each overlapping segment contributes: no atomics on the counts, integer sums, order independent.
//
// Candidates are the bins [T - 2, T] (a segment is at most 256 long).  They are taken 128 at a time, one per thread,
// and dealt to the warps whose 32-site window they overlap, sorted by `combo` = (bisulfite strand, strand index): a
// counting sort in shared memory (24 counters).  A warp then walks its own hit list combo by combo, so inside the
// hit loop the class mapping is fixed and the per-hit work is: broadcast LDS of the candidate, one byte load, one LDS
// from a 256-entry table byte -> packed increment, two adds and a predicated add for mapq^2.
//
// Narrow accumulators: four 16-bit fields (one per base; 5-bit count | 11-bit quality sum) in two registers,
// flushed into the wide per-class registers at the end of a combo run or after 28 hits.
//   MODE 0: write pileup[] (104 B / site)       MODE 1: run the model and write gt_vcf[] (208 B / site)
// ------------------------------------------------------------------------------------------------
constexpr int kPileThreads = kPileTile;
constexpr int kHitTrip = 28;                   // hits between flushes: 5-bit counts hold 31, multiple of the unroll

struct WideAcc {
	uint32_t cnt[8];           // per class: strand index 0 in the low half-word, 1 in the high one
	uint32_t qs[8];            // per class quality sums
	uint32_t mq2;
};

template <int K>
__device__ __forceinline__ void flush_combo(uint32_t &lo, uint32_t &hi, WideAcc &w) {
	// class of (bs_strand, base): st0 {0,1,2,3}  st1=C2T {0,5,2,7}  st2=G2A {4,1,6,3}   (src/call_genotypes.c:17-19)
	constexpr int cls[3][4] = {{0, 1, 2, 3}, {0, 5, 2, 7}, {4, 1, 6, 3}};
	constexpr int st = K >> 1, ori = K & 1;
#pragma unroll
	for (int b = 0; b < 4; b++) {
		const uint32_t half = b < 2 ? lo : hi;
		const uint32_t f = (b & 1) ? half >> 16 : half & 0xffffu;
		w.cnt[cls[st][b]] += (f & 31u) << (16 * ori);
		w.qs[cls[st][b]] += f >> 5;
	}
	lo = hi = 0;
}

// hit record in shared memory: {pos, len, t, mapq^2}; the byte of reference position p was staged at shared address
// t + p (32-bit shared-window arithmetic), so the per-hit address is one add
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ void red_shared_add(uint32_t addr, uint32_t v) {
	asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
	return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
	uint2 v;
	asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
	return v;
}

// One run = the hits of one combo of this warp, padded to a multiple of four with empty records (len 0) so the loop
// has no remainder path.  Records carry (mapq^2 - the tile's reference value): a run in which that is zero throughout
// (COMMON) needs no per-hit mapq work -- reference x (bases counted) is added once per site at the end.
template <int K, bool COMMON>
__device__ __forceinline__ void run_hits(uint32_t hits, uint32_t hend, uint32_t mypos, uint32_t lut, WideAcc &w) {
	uint32_t lo = 0, hi = 0;
	while (hits < hend) {
		const uint32_t trip_end = min(hits + (uint32_t)kHitTrip * 16u, hend);
#pragma unroll 1
		do {
			uint4 c[4];
			uint32_t byte[4];
#pragma unroll
			for (int u = 0; u < 4; u++) c[u] = lds_v4(hits + 16 * u);
#pragma unroll
			for (int u = 0; u < 4; u++) byte[u] = (mypos - c[u].x) < c[u].y ? lds_u8(c[u].z + mypos) : 0u;
#pragma unroll
			for (int u = 0; u < 4; u++) {
				const uint2 e = lds_v2(lut + 8 * byte[u]);
				lo += e.x; hi += e.y;
				if (!COMMON) { if (e.x | e.y) w.mq2 += c[u].w; }
			}
			hits += 64;
		} while (hits < trip_end);
		flush_combo<K>(lo, hi, w);
	}
}

template <int K>
__device__ __forceinline__ void run_combo(const uint32_t *__restrict__ rs, const uint32_t *__restrict__ re, int wid,
		uint32_t whits_s, uint32_t mypos, uint32_t lut_s, WideAcc &w) {
	const uint32_t hs = rs[wid * 8 + K], he = re[wid * 8 + K];       // bit 31 of `he`: some record of the run has odd mapq
	if (hs == (he & 0x7fffffffu)) return;
	if (he >> 31) run_hits<K, false>(whits_s + 16 * hs, whits_s + 16 * (he & 0x7fffffffu), mypos, lut_s, w);
	else run_hits<K, true>(whits_s + 16 * hs, whits_s + 16 * he, mypos, lut_s, w);
}

constexpr int kDeal = 96;                      // candidates dealt per round (three warps' worth)
constexpr int kHitCap = kDeal + 24;            // a warp's hit list: kDeal hits + padding of its six runs
constexpr int kSlotBytes = 144;
__host__ __device__ constexpr size_t pile_work_bytes(int rec) {
	const size_t a = (size_t)kPileTile * rec, b = (size_t)4 * kHitCap * 16 + (size_t)kDeal * kSlotBytes;
	return ((a > b ? a : b) + 127) & ~(size_t)127;
}                // 128 bytes of a segment inside the tile + 16-byte alignment slack on both sides

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kPileThreads, MODE ? 4 : 9)
k_pileup_tile(const Cand *__restrict__ cands, const uint32_t *__restrict__ bin_start, const uint8_t *__restrict__ bases,
		const uint8_t *__restrict__ ref, uint32_t x, uint32_t sz, uint32_t tile0, uint8_t *__restrict__ out,
		const DevConst *__restrict__ dc, unsigned long long *__restrict__ counters) {
	constexpr int REC = MODE ? 208 : 104;
	constexpr int RW = REC / 8;
	extern __shared__ __align__(128) uint8_t smem_raw[];
	// the output tile is staged where the hit lists and the staged read bytes lived (they are dead by then)
	uint64_t *stage = (uint64_t *)smem_raw;
	uint4 *whits = (uint4 *)smem_raw;                            // [4][kHitCap]
	uint8_t *slots = smem_raw + 4 * kHitCap * sizeof(uint4);     // [kDeal][kSlotBytes] staged read bytes
	uint2 *lut = (uint2 *)(smem_raw + pile_work_bytes(REC));     // byte -> {low, high} packed increment
	Tables *tabs = (Tables *)(smem_raw + pile_work_bytes(REC) + 256 * sizeof(uint2));
	__shared__ uint32_t cnt[32], cur[32], rs[32], re[32];        // [warp][combo (8 slots)]; cnt bit 31: odd mapq seen
	__shared__ uint64_t bar;

	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const uint32_t tile = tile0 + blockIdx.x;
	const uint32_t site0 = tile * kPileTile;                     // first site of this CTA
	const int nrec = (int)min((uint32_t)kPileThreads, sz - site0);
	const uint32_t tpos0 = x + site0;                            // 1-based reference position of the CTA's first site
	const uint32_t mypos = tpos0 + tid;
	const uint32_t myslot = smem_u32(slots + tid * kSlotBytes);
	{
		const uint2 *src = (const uint2 *)dc->pile_lut;
		lut[tid] = src[tid];
		lut[tid + kPileThreads] = src[tid + kPileThreads];
	}
	if (MODE) load_tables(tabs, dc, tid, kPileThreads);
	if (tid == 0) mbar_init(&bar, kPileThreads);

	WideAcc w;
#pragma unroll
	for (int j = 0; j < 8; j++) w.cnt[j] = w.qs[j] = 0;
	w.mq2 = 0;
	uint32_t warp_hits = 0, phase = 0;

	const uint32_t c_lo = bin_start[tile > (uint32_t)kBinsBack ? tile - kBinsBack : 0], c_hi = bin_start[tile + 1];
	const uint32_t mq_ref = c_lo < c_hi ? cands[c_lo].mq2 : 0u;       // the tile's reference mapq^2 (any value is correct)
	const uint32_t whits_s = smem_u32(whits), lut_s = smem_u32(lut);
	for (uint32_t base = c_lo; base < c_hi; base += kDeal) {
		if (tid < 32) cnt[tid] = 0;
		__syncthreads();
		// deal: which warps does my candidate overlap?  Its bytes inside the tile start their way to shared memory now.
		uint4 c = make_uint4(0, 0, 0, 0);
		int w_lo = 1, w_hi = 0;
		uint32_t k = 0, t = 0;
		if (tid < kDeal && base + tid < c_hi) {
			c = *(const uint4 *)(cands + base + tid);
			const int rel = (int)(c.x - tpos0);                 // start relative to the tile (negative: starts before it)
			const int s_lo = max(rel, 0), s_hi = min(rel + (int)(c.z & 0xffffu) - 1, kPileThreads - 1);
			k = c.z >> 16;
			if (s_hi >= s_lo && k < 6) {
				c.w -= mq_ref;
				w_lo = s_lo >> 5; w_hi = s_hi >> 5;
				// byte of position p is bases[c.y + p]; copy the 16-byte aligned cover of positions [s_lo, s_hi]
				const uint8_t *g0 = bases + (uint32_t)(c.y + tpos0 + (uint32_t)s_lo);
				const uint8_t *a0 = (const uint8_t *)((uintptr_t)g0 & ~(uintptr_t)15);
				const uint32_t bytes = (uint32_t)(((uintptr_t)g0 + (uint32_t)(s_hi - s_lo) + 16) & ~(uintptr_t)15) - (uint32_t)(uintptr_t)a0;
				mbar_expect_tx(&bar, bytes);
				tma_load_1d_s32(myslot, a0, bytes, &bar);
				t = myslot + (uint32_t)(g0 - a0) - (tpos0 + (uint32_t)s_lo);
			}
		}
		if (w_hi < w_lo) mbar_arrive(&bar);
#pragma unroll
#pragma unroll
		for (int ww = 0; ww < 4; ww++) if (ww >= w_lo && ww <= w_hi) {
			atomicAdd(&cnt[ww * 8 + k], 1u);
			if (c.w) atomicOr(&cnt[ww * 8 + k], 0x80000000u);      // "odd mapq" flag of the run (rare)
		}
		__syncthreads();
		if (tid < 32) {
			// runs start on multiples of four and are padded with empty records
			const uint32_t vv = cnt[tid], v = vv & 0xffffu, v4 = (v + 3u) & ~3u;
			uint32_t incl = v4;
#pragma unroll
			for (int d = 1; d < 8; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d, 8); if ((tid & 7) >= d) incl += o; }
			const uint32_t first = (uint32_t)(tid >> 3) * kHitCap + incl - v4;
			rs[tid] = first; cur[tid] = first; re[tid] = (first + v4) | (vv & 0x80000000u);
			for (uint32_t i = first + v; i < first + v4; i++) whits[i] = make_uint4(0, 0, 0, 0);
		}
		__syncthreads();
		{
			const uint4 hrec = make_uint4(c.x, c.z & 0xffffu, t, c.w);
#pragma unroll
			for (int ww = 0; ww < 4; ww++) if (ww >= w_lo && ww <= w_hi) whits[atomicAdd(&cur[ww * 8 + k], 1u)] = hrec;
		}
		__syncthreads();
		mbar_wait(&bar, phase);            // the staged bytes have landed
		phase ^= 1;
		// walk this warp's hits, combo by combo
		warp_hits += (re[wid * 8 + 5] & 0x7fffffffu) - rs[wid * 8];      // includes padding: an upper bound is all the envelope check needs
		run_combo<0>(rs, re, wid, whits_s, mypos, lut_s, w);
		run_combo<1>(rs, re, wid, whits_s, mypos, lut_s, w);
		run_combo<2>(rs, re, wid, whits_s, mypos, lut_s, w);
		run_combo<3>(rs, re, wid, whits_s, mypos, lut_s, w);
		run_combo<4>(rs, re, wid, whits_s, mypos, lut_s, w);
		run_combo<5>(rs, re, wid, whits_s, mypos, lut_s, w);
		__syncthreads();
	}

	SiteCounts s;
	uint32_t nsum = 0, qmax = 0;
#pragma unroll
	for (int j = 0; j < 8; j++) {
		s.cnt[0][j] = w.cnt[j] & 0xffffu; s.cnt[1][j] = w.cnt[j] >> 16;
		nsum += w.cnt[j];
		s.qsum[j] = (float)w.qs[j];
		qmax = max(qmax, w.qs[j]);
	}
	const uint32_t n = (nsum & 0xffffu) + (nsum >> 16);
	s.n = tid < nrec ? n : 0;
	w.mq2 += mq_ref * n;               // the common runs' share (exact modulo 2^32, like the sum itself)
	s.mapq2 = (float)w.mq2;
	// integer sums equal the reference's float sums only below 2^24, and the half-word counts hold 65535:
	// count the sites that leave the envelope
	if (s.n && (qmax >= (1u << 24) || w.mq2 >= (1u << 24) || warp_hits >= 65535u)) atomicAdd(counters + 1, 1ull);

	uint64_t *rec = stage + tid * RW;
	if (MODE == 0) {
		uint32_t o[26];
#pragma unroll
		for (int j = 0; j < 8; j++) { o[j] = s.cnt[0][j]; o[8 + j] = s.cnt[1][j]; o[17 + j] = __float_as_uint(s.qsum[j]); }
		o[16] = s.n;
		o[25] = __float_as_uint(s.mapq2);
#pragma unroll
		for (int j = 0; j < 13; j++) rec[j] = (uint64_t)o[2 * j] | (uint64_t)o[2 * j + 1] << 32;
	} else {
		const int rf = tid < nrec ? ref[site0 + tid] : 0;
		__syncthreads();                    // tables loaded (the candidate loop may not have run)
		double *wbuf = (double *)(stage + (tid & ~31) * RW);
		const bool called = call_site(s, rf, dc, tabs, rec, wbuf, lane);
		if (called && (rec[24] >> 8)) {
			guard_flag(counters, (int)(rec[24] >> 8) & 3, (unsigned long long)x + site0 + tid);
			rec[24] &= 0xffull;
		}
		rec[25] = 1ull | ((called ? 0ull : 1ull) << 8);
		const uint32_t nc = __syncthreads_count(called);
		if (tid == 0 && nc) atomicAdd(counters, (unsigned long long)nc);
	}
	store_tile<REC>(out + (size_t)blockIdx.x * kPileThreads * REC, stage, nrec, true, tid, kPileThreads);
	if (tid == 0) tma_store_wait();
}

// ------------------------------------------------------------------------------------------------
// Pileup by scatter into shared memory (the formulation the reference itself uses, src/call_genotypes.c:213-222, made
// parallel): CTA = one 128-site tile as above, but the lanes of a warp run ALONG a read.  A warp takes one candidate at a
// time; every lane loads four consecutive bytes of the read's part inside the tile and adds each counted byte to its site's
// cell with one shared-memory atomic: cell (strand index, class) of a site is one 32-bit word, count in the high and quality
// sum in the low half-word, so "counts[ori][c]++, quality[c] += q" is a single add.  A 3 x 256-entry table in shared memory
// maps (bisulfite strand, byte) to {increment, byte offset of the class plane}, zero for a byte that does not count.
// Integer adds commute, so the result does not depend on the order the atomics land in: bit-identical to the gather.
//   * columns are swizzled, col(site) = (site & 3) * 32 + (site >> 2): the four bytes of a lane go to four different
//     32-column groups and, for each of them, the 32 lanes hit 32 consecutive words -- no bank conflicts;
//   * MAPQ^2: the tile's first candidate gives a reference value that is added n times at the end; only candidates with
//     another MAPQ pay a second atomic per counted base (the difference, modulo 2^32 like the sum itself);
//   * a tile with more than kScNarrowMax candidates (a half-word cell holds 65535: 1400 bases of quality 43 are safe)
//     keeps counts and quality sums in separate words (two atomics per base, 32-bit range).
// Per counted base: ~8 instructions, against ~10 per (site, hit) pair plus the dealing of the gather.
// ------------------------------------------------------------------------------------------------
constexpr int kScNarrowMax = 1400;
constexpr int kScWarps = 4;
__host__ __device__ constexpr size_t scatter_smem_bytes() {
	// cells [2 planes][16][128] (narrow uses the first plane) | mapq plane [128] | table [3][256]; the output tile (128 x 104 B)
	// is staged over the cells once they have been read into registers
	return (size_t)2 * 16 * 128 * 4 + 128 * 4 + 3 * 256 * 4;
}

template <bool WIDE>
__device__ __forceinline__ void scatter_candidates(const Cand *__restrict__ cands, uint32_t c_lo, uint32_t c_hi, const uint8_t *__restrict__ bases,
		uint32_t tpos0, uint32_t mq_ref, uint32_t cells_s, uint32_t mqd_s, uint32_t lut_s, int lane, int wid) {
	for (uint32_t ci = c_lo + (uint32_t)wid; ci < c_hi; ci += kScWarps) {
		const uint4 c = *(const uint4 *)(cands + ci);
		const int rel = (int)(c.x - tpos0);
		const int s_lo = max(rel, 0), s_hi = min(rel + (int)(c.z & 0xffffu) - 1, kPileTile - 1);
		const uint32_t k = c.z >> 16;
		if (s_hi < s_lo || k >= 6) continue;
		const uint32_t lut = lut_s + (k >> 1) * 1024u;                 // table of this bisulfite strand
		const uint32_t ori_off = (k & 1u) * 8u * 512u;                 // plane group of this strand index
		const uint32_t dmq = c.w - mq_ref;
		const uint8_t *g0 = bases + (uint32_t)(c.y + tpos0 + (uint32_t)s_lo);      // byte of site s_lo
		const uint32_t d = (uint32_t)((uintptr_t)g0 & 3u);
		const uint32_t *a0 = (const uint32_t *)(g0 - d);
		const int n = s_hi - s_lo + 1;
		// lane i, byte j of chunk q holds site s_lo + 128 q + 4 i + j - d
		for (int q0 = 0; q0 < n + (int)d; q0 += 128) {
			const int first = q0 + 4 * lane - (int)d;                   // site offset (from s_lo) of this lane's byte 0
			if (first + 3 < 0 || first >= n) continue;
			const uint32_t word = a0[(q0 >> 2) + lane];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const int so = first + j;
				const uint32_t e = lds_u32(lut + ((word >> (8 * j)) & 0xffu) * 4u);
				if ((unsigned)so < (unsigned)n && e) {
					const uint32_t site = (uint32_t)(s_lo + so);
					const uint32_t col = ((site & 3u) << 5) + (site >> 2);
					const uint32_t addr = cells_s + ori_off + (e >> 20) + col * 4u;
					if (WIDE) {
						red_shared_add(addr, 1u);
						red_shared_add(addr + 16u * 512u, e & 0xffffu);
					} else red_shared_add(addr, e & 0xfffffu);
					if (dmq) red_shared_add(mqd_s + col * 4u, dmq);
				}
			}
		}
	}
}

__global__ void __launch_bounds__(kPileThreads, 8)
k_pileup_scatter(const Cand *__restrict__ cands, const uint32_t *__restrict__ bin_start, const uint8_t *__restrict__ bases,
		uint32_t x, uint32_t sz, uint32_t tile0, uint8_t *__restrict__ out, const DevConst *__restrict__ dc,
		unsigned long long *__restrict__ counters) {
	extern __shared__ __align__(128) uint8_t smem_raw[];
	uint32_t *cells = (uint32_t *)smem_raw;                                  // [2][16][128]
	uint32_t *mqd = cells + 2 * 16 * 128;                                    // [128]
	uint32_t *lut = mqd + 128;                                               // [3][256]
	uint64_t *stage = (uint64_t *)smem_raw;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const uint32_t tile = tile0 + blockIdx.x;
	const uint32_t site0 = tile * kPileTile;
	const int nrec = (int)min((uint32_t)kPileThreads, sz - site0);
	const uint32_t tpos0 = x + site0;
	const uint32_t c_lo = bin_start[tile > (uint32_t)kBinsBack ? tile - kBinsBack : 0], c_hi = bin_start[tile + 1];
	const bool wide = c_hi - c_lo > (uint32_t)kScNarrowMax;
	// table: (strand, byte) -> class plane offset << 20 | increment; narrow: 1 << 16 | q, wide: q (the count is a separate add)
	{
		const uint32_t minq = (uint32_t)dc->min_qual;
		for (int i = tid; i < 3 * 256; i += kPileThreads) {
			const uint32_t st = (uint32_t)i >> 8, b = (uint32_t)i & 255u, q = b >> 2, base = b & 3u;
			// class of (bs_strand, base): st0 {0,1,2,3}  st1=C2T {0,5,2,7}  st2=G2A {4,1,6,3}   (src/call_genotypes.c:17-19)
			const uint32_t cls = ((st == 0 ? 0x3210u : st == 1 ? 0x7250u : 0x3614u) >> (4 * base)) & 15u;
			const bool counted = q >= minq && q != (uint32_t)BSGPU_FLT_QUAL;
			lut[i] = counted ? (cls * 512u) << 20 | (wide ? q : (1u << 16 | q)) : 0u;
		}
		for (int i = tid; i < (wide ? 2 : 1) * 16 * 128; i += kPileThreads) cells[i] = 0;
		mqd[tid] = 0;
	}
	__syncthreads();
	const uint32_t mq_ref = c_lo < c_hi ? cands[c_lo].mq2 : 0u;
	if (wide) scatter_candidates<true>(cands, c_lo, c_hi, bases, tpos0, mq_ref, smem_u32(cells), smem_u32(mqd), smem_u32(lut), lane, wid);
	else scatter_candidates<false>(cands, c_lo, c_hi, bases, tpos0, mq_ref, smem_u32(cells), smem_u32(mqd), smem_u32(lut), lane, wid);
	__syncthreads();
	// my site's cells -> the reference's pileup record
	uint32_t o[26];
	{
		const uint32_t col = (((uint32_t)tid & 3u) << 5) + ((uint32_t)tid >> 2);
		uint32_t n = 0, qmax = 0;
#pragma unroll
		for (int j = 0; j < 8; j++) {
			uint32_t c0, c1, qs;
			if (wide) {
				c0 = cells[j * 128 + col]; c1 = cells[(8 + j) * 128 + col];
				qs = cells[(16 + j) * 128 + col] + cells[(24 + j) * 128 + col];
			} else {
				const uint32_t a = cells[j * 128 + col], b = cells[(8 + j) * 128 + col];
				c0 = a >> 16; c1 = b >> 16; qs = (a & 0xffffu) + (b & 0xffffu);
			}
			o[j] = c0; o[8 + j] = c1; o[17 + j] = __float_as_uint((float)qs);
			n += c0 + c1;
			qmax = max(qmax, qs);
		}
		const uint32_t mq2 = mq_ref * n + mqd[col];          // exact modulo 2^32, like the sum itself
		if (tid >= nrec) {
#pragma unroll
			for (int j = 0; j < 26; j++) o[j] = 0;
			n = 0;
		}
		o[16] = n;
		o[25] = __float_as_uint((float)mq2);
		// integer sums equal the reference's float sums only below 2^24: count the sites that leave the envelope
		if (n && (qmax >= (1u << 24) || mq2 >= (1u << 24))) atomicAdd(counters + 1, 1ull);
	}
	__syncthreads();                       // everybody has read its cells: the staging tile may overwrite them
	uint64_t *rec = stage + tid * 13;
#pragma unroll
	for (int j = 0; j < 13; j++) rec[j] = (uint64_t)o[2 * j] | (uint64_t)o[2 * j + 1] << 32;
	store_tile<104>(out + (size_t)blockIdx.x * kPileThreads * 104, stage, nrec, true, tid, kPileThreads);
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

// Synthetic coordinate-sorted BAM record stream (config 3 shape) written straight into HBM: one warp per record.
// Records are fixed size (qname of 15 characters + NUL, one CIGAR op, GEM strand tag XB:A), so record i of the sorted
// stream sits at i * record_bytes; `rank` (computed by the caller with a sort over the start positions) says where the
// record of (template, mate) goes.  Template t copies the reads of template src[t] (a positional duplicate when
// src[t] != t) under its own name.
__host__ __device__ constexpr uint32_t synth_bam_record_bytes(uint32_t read_len) { return 4 + 32 + 16 + 4 + (read_len + 1) / 2 + read_len + 4; }

__global__ void k_synth_bam(uint64_t seed, size_t ntemplates, uint32_t read_len, const uint32_t *__restrict__ pos_f,
		const uint32_t *__restrict__ pos_r, const uint32_t *__restrict__ src, const uint32_t *__restrict__ rank, uint8_t *__restrict__ out) {
	const size_t wi = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (wi >= 2 * ntemplates) return;
	const uint32_t k = wi >= ntemplates ? 1u : 0u;                 // 0: forward-strand mate, 1: reverse-strand mate
	const size_t t = wi - (k ? ntemplates : 0), ts = src[t];
	const uint32_t rb = synth_bam_record_bytes(read_len);
	uint8_t *rec = out + (size_t)rank[wi] * rb;
	Rng r(seed ^ 0x51ed27fULL, ts);
	const uint64_t h = r.next();
	const uint32_t st = 1 + (uint32_t)((h >> 24) & 1);             // C2T: top-strand fragment, G2A: bottom strand
	const uint32_t hap = (uint32_t)((h >> 26) & 1);
	const uint32_t fpos = pos_f[t], rpos = pos_r[t], pos = k ? rpos : fpos, mpos = k ? fpos : rpos;
	const uint32_t mapq = ((h >> 28) & 0x1f) ? 60 : 5 + (uint32_t)((h >> 34) % 50);
	const int32_t tlen = (int32_t)(rpos + read_len - fpos);
	// directional library: read 1 is the forward mate of a top-strand fragment, the reverse mate of a bottom-strand one
	const bool read1 = (k == 0) == (st == 1);
	const uint32_t flag = 1u | 2u | (k ? 16u : 32u) | (read1 ? 64u : 128u);
	if (lane == 0) {
		auto put32 = [&](uint32_t o, uint32_t v) { rec[o] = (uint8_t)v; rec[o + 1] = (uint8_t)(v >> 8); rec[o + 2] = (uint8_t)(v >> 16); rec[o + 3] = (uint8_t)(v >> 24); };
		put32(0, rb - 4);
		put32(4, 0);                                  // refID
		put32(8, pos - 1);
		rec[12] = 16; rec[13] = (uint8_t)mapq; rec[14] = 0x48; rec[15] = 0x12;      // l_read_name, mapq, bin (unused)
		rec[16] = 1; rec[17] = 0; rec[18] = (uint8_t)flag; rec[19] = (uint8_t)(flag >> 8);
		put32(20, read_len);
		put32(24, 0);                                 // next refID
		put32(28, mpos - 1);
		put32(32, (uint32_t)(k ? -tlen : tlen));
		uint64_t id = t;
		rec[36] = 't';
		for (int i = 14; i >= 1; i--) { rec[36 + i] = (uint8_t)('0' + id % 10); id /= 10; }
		rec[51] = 0;
		put32(52, read_len << 4);                     // <read_len>M
		uint8_t *aux = rec + 56 + (read_len + 1) / 2 + read_len;
		aux[0] = 'X'; aux[1] = 'B'; aux[2] = 'A'; aux[3] = st == 1 ? 'C' : 'G';
	}
	uint8_t *seq = rec + 56, *qual = seq + (read_len + 1) / 2;
	for (uint32_t j2 = lane; j2 < (read_len + 1) / 2; j2 += 32) {
		uint32_t nibs = 0;
		for (uint32_t u = 0; u < 2; u++) {
			const uint32_t j = 2 * j2 + u;
			uint32_t nib = 0, q = 0;
			if (j < read_len) {
				const uint32_t p = pos + j;
				const int rc = synth_ref_code(seed, p);
				const uint64_t g = mix64(seed ^ 0x77aa55ull ^ ((uint64_t)p << 20));            // site genotype
				int b = rc ? rc - 1 : (int)(g & 3);
				const uint32_t gsel = (uint32_t)(g >> 40) % 3000u;
				if (gsel < 2) { if (gsel == 0 || hap) b = (b + 1 + (int)((g >> 8) % 3)) & 3; }
				const uint64_t e = mix64(h ^ ((uint64_t)(j + 1000 * k) * 0x9e3779b97f4a7c15ull));
				const float uc = (float)(e >> 40) * (1.0f / 16777216.0f);
				const bool cpg = st == 1 ? (rc == 2 && synth_ref_code(seed, p + 1) == 3) : (rc == 3 && synth_ref_code(seed, p - 1) == 2);
				const float meth = cpg ? 0.7f : 0.01f;
				const bool converts = uc >= meth && uc < meth + (1.0f - meth) * 0.99f;
				if (st == 1 && b == 1 && converts) b = 3;
				if (st == 2 && b == 2 && converts) b = 0;
				q = ((e >> 8) & 0xff) < 218 ? 37u : 10u + (uint32_t)((e >> 16) % 31);
				const float perr = __expf(-0.2302585f * (float)q);
				if ((float)((e >> 24) & 0xffff) * (1.0f / 65536.0f) < perr) b = (b + 1 + (int)((e >> 4) % 3)) & 3;
				nib = 1u << b;
				if (((e >> 44) & 0x3ff) == 0) nib = 15;                                    // N
				qual[j] = (uint8_t)q;
			}
			nibs = (nibs << 4) | nib;
		}
		seq[j2] = (uint8_t)nibs;
	}
}

// ------------------------------------------------------------------------------------------------
// launchers (called from bsgpu_api.cpp; all asynchronous on `stream`)
// ------------------------------------------------------------------------------------------------
#define LAUNCH_CHECK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return e_; } while (0)

static size_t call_smem(bool vcf) { return (size_t)kCallTile * ((vcf ? 208 : 200) + 104) + sizeof(Tables); }
static size_t pile_smem(int mode) { return pile_work_bytes(mode ? 208 : 104) + 256 * sizeof(uint2) + (mode ? sizeof(Tables) : 0); }

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
	if ((e = prep(k_call_sites<false, 4, 0>, call_smem(false), kCallTile, &g_call_ctas[0][0])) != cudaSuccess) return e;
	if ((e = prep(k_call_sites<true, 4, 0>, call_smem(true), kCallTile, &g_call_ctas[0][1])) != cudaSuccess) return e;
	if ((e = prep(k_call_sites<false, 5, 0>, call_smem(false), kCallTile, &g_call_ctas[1][0])) != cudaSuccess) return e;
	if ((e = prep(k_call_sites<true, 5, 0>, call_smem(true), kCallTile, &g_call_ctas[1][1])) != cudaSuccess) return e;
	if ((e = prep(k_call_sites<false, 5, 1>, call_smem(false), kCallTile, nullptr)) != cudaSuccess) return e;
	if ((e = prep(k_call_sites<true, 5, 1>, call_smem(true), kCallTile, nullptr)) != cudaSuccess) return e;
	if ((e = prep(k_call_sites<false, 5, 2>, call_smem(false), kCallTile, nullptr)) != cudaSuccess) return e;
	if ((e = prep(k_call_sites<true, 5, 2>, call_smem(true), kCallTile, nullptr)) != cudaSuccess) return e;
	if ((e = prep(k_pileup_tile<0>, pile_smem(0), kPileThreads, nullptr)) != cudaSuccess) return e;
	if ((e = prep(k_pileup_tile<1>, pile_smem(1), kPileThreads, nullptr)) != cudaSuccess) return e;
	if ((e = prep(k_pileup_scatter, scatter_smem_bytes(), kPileThreads, nullptr)) != cudaSuccess) return e;
	return cudaSuccess;
}

cudaError_t launch_call_sites(const void *pileup, const void *ref, size_t n, void *out, void *skip, bool vcf,
		const DevConst *dc, unsigned long long *counters, cudaStream_t stream, int *launches, unsigned long long guard_base, bool overlap_safe) {
	if (!n) return cudaSuccess;
	// How a tile comes in and goes out, as bits: 1 input by one bulk copy (cp.async.bulk + mbarrier), 2 output by one bulk copy,
	// 4 wait for the completion of the output copy instead of for its reads of shared memory, 8 (without 1) input by 16-byte
	// cp.async copies of every thread.  Default 3.  overlap_safe (the caller runs other kernels next to this one: a second
	// context on the device, decode kernels on their own stream, two chunk streams) asks for 10: on B200 the bulk LOAD of this
	// persistent kernel ends in "illegal memory access" when CTAs of other kernels share its SMs (bisected on the genome leg
	// with two sessions per GPU: bits 1, 3, 7 fault within seconds, 0, 2, 10 never did; profiles/r02c_fault_bisect.md), and the
	// cp.async prefetch costs 6 % of the kernel (12.0 against 12.8 G sites/s).  BSGPU_BULK=<bits> / BSGPU_NO_BULK=1 override both.
	static const int env_bits = [] { if (getenv("BSGPU_NO_BULK")) return 0; const char *e = getenv("BSGPU_BULK"); return e ? atoi(e) & 15 : -1; }();
	const int bulk_bits = env_bits >= 0 ? env_bits : overlap_safe ? 10 : 3;
	const int bulk_ok = (((uintptr_t)pileup | (uintptr_t)out) & 15u) == 0 ? bulk_bits : 0;
	const size_t ntiles = (n + kCallTile - 1) / kCallTile;
	const int five = g_call_minb == 5;
	const size_t resident = (size_t)g_sms * (size_t)(g_call_ctas[five][vcf] > 0 ? g_call_ctas[five][vcf] : 1);
	const unsigned grid = (unsigned)(ntiles < resident ? ntiles : resident);       // persistent: one CTA per resident slot
	const uint8_t *p = (const uint8_t *)pileup, *r = (const uint8_t *)ref;
	uint8_t *o = (uint8_t *)out, *sk = (uint8_t *)skip;
	static const bool generic = getenv("BSGPU_CALL_GENERIC") != nullptr;      // A/B switch: the run-time transport body for every launch
	const int mode = (!five || generic) ? 0 : bulk_ok == 3 ? 1 : bulk_ok == 10 ? 2 : 0;
#define BSGPU_CALL(V, M, MD) k_call_sites<V, M, MD><<<grid, kCallTile, call_smem(V), stream>>>(p, r, n, o, V ? nullptr : sk, dc, bulk_ok, counters, guard_base)
	if (vcf) {
		if (mode == 1) BSGPU_CALL(true, 5, 1); else if (mode == 2) BSGPU_CALL(true, 5, 2); else if (five) BSGPU_CALL(true, 5, 0); else BSGPU_CALL(true, 4, 0);
	} else {
		if (mode == 1) BSGPU_CALL(false, 5, 1); else if (mode == 2) BSGPU_CALL(false, 5, 2); else if (five) BSGPU_CALL(false, 5, 0); else BSGPU_CALL(false, 4, 0);
	}
#undef BSGPU_CALL
	__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
	LAUNCH_CHECK();
	return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------
// Wire records (bsgpu_wire.h): the 120 informative bytes of each gt_meth / gt_vcf record, for the trip over PCIe.
// One thread per record; the records were written a moment ago by k_call_sites and are read back from L2.  A field that
// does not fit its wire width raises the chunk's flag word (wire[n * 15]), which the launcher clears beforehand.
// ------------------------------------------------------------------------------------------------
template <int REC>
__global__ void __launch_bounds__(256) k_wire_pack(const uint64_t *__restrict__ rec, const uint8_t *__restrict__ skip, uint32_t n,
		uint64_t *__restrict__ wire) {
	const uint32_t i = blockIdx.x * 256u + threadIdx.x;
	if (i >= n) return;
	const uint64_t *r = rec + (size_t)i * (REC / 8);
	uint64_t *w = wire + (size_t)i * 15;
	uint64_t c[2] = {0, 0}, q = 0, wide = 0;
#pragma unroll
	for (int j = 0; j < 8; j++) {
		const uint64_t v = r[j];
		wide |= v >> 16;
		c[j >> 2] |= (v & 0xffff) << 16 * (j & 3);
	}
#pragma unroll
	for (int j = 0; j < 4; j++) {
		const uint64_t v = r[8 + j];
		wide |= (v & 0xffffff00ffffff00ull);
		q |= (v & 0xff) << 16 * j | (v >> 32 & 0xff) << (16 * j + 8);
	}
#pragma unroll
	for (int g = 0; g < 11; g++) w[g] = r[12 + g];
	const uint64_t m = r[23], b = r[24];
	wide |= (m & 0xffffff00ffffff00ull);
	uint64_t sk;
	if constexpr (REC == 208) sk = r[25] >> 8 & 0xff;
	else sk = skip[i];
	w[11] = c[0];
	w[12] = c[1];
	w[13] = q;
	w[14] = (m & 0xff) | (m >> 32 & 0xff) << 8 | (b & 0xff) << 16 | sk << 24;
	if (wide) atomicOr((unsigned long long *)(wire + (size_t)n * 15), 1ull);
}

// records (gt_meth + skip[], or gt_vcf when skip == NULL) -> wire[n * 15 + 1]; the last word is the chunk's flag
cudaError_t launch_wire_pack(const void *rec, const void *skip, size_t n, void *wire, cudaStream_t stream, int *launches) {
	if (!n) return cudaSuccess;
	cudaError_t e = cudaMemsetAsync((uint64_t *)wire + n * 15, 0, 8, stream);
	if (e != cudaSuccess) return e;
	const unsigned grid = (unsigned)((n + 255) / 256);
	if (skip) k_wire_pack<200><<<grid, 256, 0, stream>>>((const uint64_t *)rec, (const uint8_t *)skip, (uint32_t)n, (uint64_t *)wire);
	else k_wire_pack<208><<<grid, 256, 0, stream>>>((const uint64_t *)rec, nullptr, (uint32_t)n, (uint64_t *)wire);
	__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
	LAUNCH_CHECK();
	return cudaSuccess;
}

static size_t seg_area(size_t nseg) { return (nseg * sizeof(Cand) + 255) & ~(size_t)255; }
static uint32_t scan_ctas(uint32_t ntiles) { return (ntiles + 1023) / 1024; }

size_t pileup_scratch_bytes(size_t nseg, uint32_t sz) {
	const size_t ntiles = (sz + kPileTile - 1) / kPileTile;
	return seg_area(nseg) + (3 * ntiles + scan_ctas((uint32_t)ntiles) + 16) * sizeof(uint32_t) + 256;
}

// scratch layout: sorted candidates | counts[ntiles] | start[ntiles+1] | cursor[ntiles] | partial[ctas+1]
cudaError_t launch_bin_segments(const void *segs, size_t nseg, uint32_t x, uint32_t sz, void *scratch,
		cudaStream_t stream, int *launches, unsigned long long *counters) {
	if (!sz) return cudaSuccess;
	const uint32_t ntiles = (sz + kPileTile - 1) / kPileTile;
	Cand *sorted = (Cand *)scratch;
	uint32_t *counts = (uint32_t *)((uint8_t *)scratch + seg_area(nseg));
	uint32_t *start = counts + ntiles;
	uint32_t *cursor = start + ntiles + 1;
	uint32_t *partial = cursor + ntiles;
	cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(uint32_t) * ntiles, stream);
	if (e != cudaSuccess) return e;
	const unsigned g = (unsigned)((nseg + 255) / 256);
	if (nseg) {
		k_bin_count<<<g, 256, 0, stream>>>((const Seg *)segs, nseg, x, ntiles, counts, counters);
		__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
		LAUNCH_CHECK();
	}
	const uint32_t nb = scan_ctas(ntiles);
	k_scan_local<<<nb, 1024, 0, stream>>>(counts, ntiles, start, partial);
	k_scan_partials<<<1, 1024, 0, stream>>>(partial, nb);
	k_scan_add<<<nb, 1024, 0, stream>>>(start, cursor, ntiles, partial, nb);
	__atomic_fetch_add(launches, 3, __ATOMIC_RELAXED);
	LAUNCH_CHECK();
	if (nseg) {
		k_bin_scatter<<<g, 256, 0, stream>>>((const Seg *)segs, nseg, x, ntiles, cursor, sorted);
		__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
		LAUNCH_CHECK();
	}
	return cudaSuccess;
}

cudaError_t launch_pileup_tiles(const void *scratch, size_t nseg, const void *bases, const void *ref, uint32_t x,
		uint32_t sz, uint32_t tile0, uint32_t ntiles, void *out, int mode, const DevConst *dc,
		unsigned long long *counters, cudaStream_t stream, int *launches) {
	if (!ntiles) return cudaSuccess;
	const uint32_t all_tiles = (sz + kPileTile - 1) / kPileTile;
	const Cand *sorted = (const Cand *)scratch;
	const uint32_t *start = (const uint32_t *)((const uint8_t *)scratch + seg_area(nseg)) + all_tiles;
	// sites covered by this launch: tiles [tile0, tile0 + ntiles) clipped to the window; one CTA per tile
	const unsigned grid = min(ntiles, all_tiles - tile0);
	// BSGPU_PILEUP=scatter: build pileup[] (mode 0) with the shared-memory-atomics formulation instead of the gather.  Measured
	// on B200 (bit-identical results): 30x / 150 bp 4.62 ms per 50 M sites against 2.22 ms, 500x panel 15.0 ms per 10 M sites
	// against 6.15 ms -- a RED.shared per counted base costs more than the gather's ten instructions per (site, hit) pair.
	static const bool scatter = [] { const char *e = getenv("BSGPU_PILEUP"); return e ? e[0] == 's' : false; }();
	if (mode) k_pileup_tile<1><<<grid, kPileThreads, pile_smem(1), stream>>>(sorted, start, (const uint8_t *)bases, (const uint8_t *)ref, x, sz, tile0, (uint8_t *)out, dc, counters);
	else if (scatter) k_pileup_scatter<<<grid, kPileThreads, scatter_smem_bytes(), stream>>>(sorted, start, (const uint8_t *)bases, x, sz, tile0, (uint8_t *)out, dc, counters);
	else k_pileup_tile<0><<<grid, kPileThreads, pile_smem(0), stream>>>(sorted, start, (const uint8_t *)bases, (const uint8_t *)ref, x, sz, tile0, (uint8_t *)out, dc, counters);
	__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
	LAUNCH_CHECK();
	return cudaSuccess;
}

cudaError_t launch_synth_sites(uint64_t seed, uint64_t first, size_t n, double mean_depth, void *pileup, void *ref,
		cudaStream_t stream, int *launches) {
	if (!n) return cudaSuccess;
	k_synth_sites<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(seed, first, n, (float)mean_depth, (uint8_t *)pileup, (uint8_t *)ref);
	__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
	LAUNCH_CHECK();
	return cudaSuccess;
}

cudaError_t launch_synth_ref(uint64_t seed, uint32_t x, uint32_t sz, void *ref, cudaStream_t stream, int *launches) {
	if (!sz) return cudaSuccess;
	k_synth_ref<<<(sz + 255) / 256, 256, 0, stream>>>(seed, x, sz, (uint8_t *)ref);
	__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
	LAUNCH_CHECK();
	return cudaSuccess;
}

size_t synth_bam_bytes(size_t ntemplates, uint32_t read_len) { return 2 * ntemplates * (size_t)synth_bam_record_bytes(read_len); }

cudaError_t launch_synth_bam(uint64_t seed, size_t ntemplates, uint32_t read_len, const void *pos_f, const void *pos_r, const void *src,
		const void *rank, void *out, cudaStream_t stream, int *launches) {
	if (!ntemplates) return cudaSuccess;
	k_synth_bam<<<(unsigned)((2 * ntemplates * 32 + 255) / 256), 256, 0, stream>>>(seed, ntemplates, read_len, (const uint32_t *)pos_f,
		(const uint32_t *)pos_r, (const uint32_t *)src, (const uint32_t *)rank, (uint8_t *)out);
	__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
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
	__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
	LAUNCH_CHECK();
	if (nseg) {
		const double step = (double)read_len / depth;
		k_synth_reads<<<(unsigned)((nseg * 32 + 255) / 256), 256, 0, stream>>>(seed, x, sz, read_len, step, nseg, (Seg *)segs, (uint8_t *)bases);
		__atomic_fetch_add(launches, 1, __ATOMIC_RELAXED);
		LAUNCH_CHECK();
	}
	return cudaSuccess;
}

}  // namespace bsgpu
