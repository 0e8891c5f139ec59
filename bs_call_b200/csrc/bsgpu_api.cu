// bsgpu_api.cu -- the C ABI declared in include/bsgpu.h: context, staging, pipelines.  No CPU compute path:
// every entry point that produces results launches kernels, and bsgpu_init fails if no sm_100 device opens.
#include <algorithm>
#include <atomic>
#include <thread>
#include <functional>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <unordered_map>
#include <map>
#include <set>
#include <deque>
#include <mutex>
#include <condition_variable>
#include <string>
#include <cuda_runtime.h>

#include "bsgpu.h"
#include "bsgpu_device.cuh"
#include "bsgpu_launch.h"
#include "bsgpu_session.h"
#include "bsgpu_wire.h"

static_assert(sizeof(bsgpu_pileup) == 104, "pileup layout (include/bs_call.h:174-182)");
static_assert(sizeof(bsgpu_gt_meth) == 200, "gt_meth layout (include/bs_call.h:152-160)");
static_assert(sizeof(bsgpu_gt_vcf) == 208, "gt_vcf layout (include/bs_call.h:162-166)");
static_assert(sizeof(bsgpu_seg) == 16, "segment layout");
static_assert(BSGPU_MAX_SEG_LEN == bsgpu::kMaxSegLen, "segment length limit");
static_assert(sizeof(bsgpu_template) == 56, "template layout");
static_assert(offsetof(bsgpu_gt_meth, gt_prob) == 96 && offsetof(bsgpu_gt_meth, fisher_strand) == 176 &&
		offsetof(bsgpu_gt_meth, mq) == 184 && offsetof(bsgpu_gt_meth, max_gt) == 192, "gt_meth offsets");
static_assert(offsetof(bsgpu_pileup, n) == 64 && offsetof(bsgpu_pileup, quality) == 68 && offsetof(bsgpu_pileup, mapq2) == 100, "pileup offsets");

using namespace bsgpu;

namespace bsgpu {      // host helpers of bsgpu_reader.cu
struct FrameScratch;
FrameScratch *frame_scratch_new();
void frame_scratch_free(FrameScratch *s);
int frame_records(const uint8_t *bam, size_t nbytes, std::vector<uint64_t> &rec_off, std::vector<uint32_t> &read_off,
		std::vector<uint32_t> &mm_off, uint64_t *nbases, uint64_t *nmisms, FrameScratch *scratch, size_t *framed = nullptr);
int build_blocks_host(const uint8_t *bam, const uint64_t *rec_off, const bsgpu_record *rec, size_t nrec, bool keep_unmatched,
		bool keep_duplicates, std::vector<bsgpu_block> &blocks, bsgpu_template *tmpl, size_t *ntmpl, uint64_t *tally);
struct BuildJob;
struct CertainState { int tid = -1; uint64_t maxend = 0; };
void certain_block_starts(const bsgpu_record *rec, size_t rbeg, size_t rend, CertainState *st, std::vector<size_t> &starts);
void certain_block_starts_keys(const uint32_t *keys, size_t rbeg, size_t rend, CertainState *st, std::vector<size_t> &starts);
BuildJob *build_blocks_start_range(const uint8_t *bam, const uint64_t *rec_off, const bsgpu_record *rec, const uint32_t *name_id, size_t rbeg, size_t rend,
		const std::vector<size_t> &starts, bool keep_unmatched, bool keep_duplicates, bsgpu_template *tmpl, unsigned pieces_per_thread,
		bool with_tally);
void host_name_ids(const uint8_t *bam, const uint64_t *rec_off, const bsgpu_record *rec, size_t nrec, uint32_t *name_id);
const uint64_t *build_blocks_piece_tally(const BuildJob *job, size_t p);
uint32_t build_blocks_piece_maxcap(const BuildJob *job, size_t p);
bool build_blocks_piece_ready(const BuildJob *job, size_t p);
size_t build_blocks_pieces(const BuildJob *job);
int build_blocks_piece(BuildJob *job, size_t p, const std::vector<bsgpu_block> **blocks, size_t *tmpl_base, size_t *ntmpl);
void build_blocks_finish(BuildJob *job);
}
static_assert(sizeof(bsgpu_record) == 56, "record descriptor layout");
static_assert(sizeof(bsgpu_block) == 32, "block layout");

static thread_local char g_err[512] = "";
// contexts alive per device: with more than one, kernels of different contexts share the SMs (see launch_call_sites)
static std::atomic<int> g_ctx_on_device[64];

static int fail(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return BSGPU_FAIL;
}

// BSGPU_SYNC_CHECK=1 (debugging aid; compute-sanitizer is not available on every pool): the device is synchronised after
// every runtime call made through CU(), so that an asynchronous fault is reported at the call that caused it
static const bool g_sync_check = getenv("BSGPU_SYNC_CHECK") != nullptr;
// BSGPU_SYNC_AFTER=<text> (debugging aid): the device is synchronised BEFORE AND AFTER every launcher call whose source text
// contains <text> (e.g. launch_pileup, launch_bcf), so that those kernels never run next to any other kernel of the process
static const char *g_sync_after = getenv("BSGPU_SYNC_AFTER");
#define CU(call) do { const bool iso_ = g_sync_after && strstr(#call, g_sync_after); if (iso_) cudaDeviceSynchronize(); \
	cudaError_t e_ = (call); if (e_ == cudaSuccess && (g_sync_check || iso_)) e_ = cudaDeviceSynchronize(); \
	if (e_ != cudaSuccess) return fail("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// grow-on-demand device / pinned buffers
//
// BSGPU_REDZONE=1 (debugging aid; compute-sanitizer is not available on every box): every device buffer of the library is
// allocated at exactly the size asked for (rounded up to 16 bytes) instead of with 12.5 % head-room, followed by a 4-KiB red
// zone filled with 0xC3.  bsgpu_debug_redzones() -- and every release of a buffer -- reads the zones back and counts the ones a
// kernel or copy wrote into, so out-of-bounds WRITES of any kernel into library-owned memory are caught without a tool.
static const bool g_redzone = getenv("BSGPU_REDZONE") != nullptr;
constexpr size_t kRedZone = 4096;
struct DevBuf;
static std::mutex g_red_mu;
static std::set<DevBuf *> g_red_live;
static unsigned long long g_red_checked = 0, g_red_corrupt = 0, g_red_first_bytes = 0;
struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	bool zone_intact() const {           // red-zone mode only; caller holds g_red_mu
		static std::vector<uint8_t> h(kRedZone);
		if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(h.data(), (const uint8_t *)p + cap, kRedZone, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
		g_red_checked++;
		for (size_t i = 0; i < kRedZone; i++) if (h[i] != 0xC3) { g_red_corrupt++; if (!g_red_first_bytes) g_red_first_bytes = cap; return false; }
		return true;
	}
	cudaError_t reserve(size_t bytes) {
		if (bytes <= cap) return cudaSuccess;
		release();
		size_t want = g_redzone ? ((bytes + 15) & ~(size_t)15) : bytes + bytes / 8 + 256;
		cudaError_t e = cudaMalloc(&p, want + (g_redzone ? kRedZone : 0));
		if (e != cudaSuccess) { p = nullptr; return e; }
		cap = want;
		if (g_redzone) {
			if ((e = cudaMemset((uint8_t *)p + want, 0xC3, kRedZone)) != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) return e;
			std::lock_guard<std::mutex> lk(g_red_mu);
			g_red_live.insert(this);
		}
		return cudaSuccess;
	}
	void release() {
		if (p) {
			if (g_redzone) { std::lock_guard<std::mutex> lk(g_red_mu); zone_intact(); g_red_live.erase(this); }
			cudaFree(p);
		}
		p = nullptr; cap = 0;
	}
};

// red zones of all live device buffers of the process: *checked zones looked at so far (live ones now + every buffer released
// before), *corrupt the ones found written into.  Returns BSGPU_OK when none was, BSGPU_FAIL otherwise (or without BSGPU_REDZONE).
extern "C" int bsgpu_debug_redzones(unsigned long long *checked, unsigned long long *corrupt) {
	if (!g_redzone) { if (checked) *checked = 0; if (corrupt) *corrupt = 0; return BSGPU_FAIL; }
	std::lock_guard<std::mutex> lk(g_red_mu);
	for (DevBuf *b : g_red_live) b->zone_intact();
	if (checked) *checked = g_red_checked;
	if (corrupt) *corrupt = g_red_corrupt;
	if (g_red_corrupt) fprintf(stderr, "bsgpu: %llu red zone(s) written into (first: behind a buffer of %llu bytes)\n", g_red_corrupt, g_red_first_bytes);
	return g_red_corrupt ? BSGPU_FAIL : BSGPU_OK;
}

struct PinBuf {                // grow-on-demand page-locked host buffer
	void *p = nullptr;
	size_t cap = 0;
	cudaError_t reserve(size_t bytes) {
		if (bytes <= cap) return cudaSuccess;
		if (p) cudaFreeHost(p);
		p = nullptr; cap = 0;
		const size_t want = bytes + bytes / 8 + 256;
		cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Slot {                  // one stage of the host-buffer pipelines
	cudaStream_t stream = nullptr;
	cudaEvent_t done = nullptr;
	DevBuf in, ref, out, skip, wire;
};

struct bsgpu_ctx {
	int device = 0;
	bsgpu_params params;
	DevConst *d_const = nullptr;
	unsigned long long *d_counters = nullptr;    // [0] sites called by the fused kernel, [1] envelope overflows
	cudaStream_t stream = nullptr;               // context stream (block path, _dev default)
	cudaStream_t copy_stream = nullptr;          // D2H of finished windows
	Slot slot[2];
	DevBuf segs, bases, ref, scratch, vcf, tmpl, misms, obases, ooff, pile;
	// Reader side of one run over a record stream: the stream itself, its framing, the decoded arrays, what comes home of them.
	// Two sets, so that a streaming session can have the reader stage of batch k + 1 under way (framing, upload, decode,
	// certain-start scan, block builder) while the windows of batch k are still being queued and run.
	struct ReaderSet {
		DevBuf rd_bam, rd_recoff, rd_readoff, rd_mmoff, rd_rec, rd_bases, rd_misms, rd_key, rd_mask, rd_scan, rd_names, rd_nameid;
		PinBuf h_nameid;                         // name ids of the records coming home (QNAME join on the device)
		std::vector<size_t> mask_off;            // word offset of every chunk's certain-start mask in rd_mask / h_mask
		std::vector<uint64_t> rec_off;           // framing of the stream
		std::vector<uint32_t> read_off, mm_off;
		FrameScratch *frame_scratch = nullptr;
		PinBuf h_rec, h_off, h_tmpl, h_key, h_mask;      // pinned staging: descriptors coming back, offset tables and templates going up
		std::vector<cudaEvent_t> rd_up, rd_done;         // byte piece uploaded, chunk descriptors home
		cudaStream_t up = nullptr;               // the set's upload stream
	} rs[2];
	// window stage turnstile of pipelined runs: run `ticket` queues its windows when win_done == ticket
	std::mutex win_mu;
	std::condition_variable win_cv;
	uint64_t win_done = 0, next_ticket = 0;
	std::vector<uint32_t> off_tmp;
	std::vector<uint8_t> ref_tmp;
	bool fused = false;                          // BSGPU_FUSED=1: one fused pileup+model kernel instead of two kernels
	bool block_overlap = false;                  // BSGPU_BLOCK_OVERLAP=1: gather of part k + 1 on its own stream next to the model of part k
	bool overlap_kernels = false;                // the context runs kernels on more than one stream (decode stream of the reader stage)
	// overlapped block path (BSGPU_BLOCK_OVERLAP): the pileup of slab i + 1 on its own stream under the model of slab i
	cudaStream_t pile_stream = nullptr;
	cudaEvent_t pile_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};      // piled[2], consumed[2], binned
	DevBuf pile2;
	uint32_t pile_pos = 0;
	bool pile_used[2] = {false, false};
	std::vector<cudaEvent_t> win_events;         // output ring
	uint32_t ring_pos = 0;                       // next output ring slot of the deferred runs
	bool ring_busy[3] = {false, false, false};   // output ring slot may still be copying out (deferred block_run)
	bsgpu_stats stats;
	// results of the host-buffer entry points come home as wire records (bsgpu_wire.h): pool of rebuilding threads, pinned landing buffers
	WireExpander *expander = nullptr;
	PinBuf wire_pin[6];
	cudaEvent_t wire_landed[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	int wire_mode = 0;                           // BSGPU_WIRE: 0 full records only (default), 1 adaptive (a chunk goes out in full when no landing buffer is free), 2 wire only
	// dbSNP / region annotation: per contig (bsgpu_set_contig_annotation) and of the last single-contig call (params->dbsnp)
	struct DbDev { DevBuf mask, fq, cum, off, names; uint32_t words = 0; uint32_t reg_start = 0, reg_stop = 0; uint64_t key[4] = {0, 0, 0, 0}; };
	std::map<int, DbDev *> contig_ann;
	DbDev call_db;
	// --report-file statistics of the sites (bsgpu_site_stats_enable): device image, per-contig table, GC bins per contig (by rid)
	bool site_stats_on = false;
	DevBuf st_main, st_ctg;
	uint32_t st_nctg = 0;
	struct GcDev { DevBuf bins; uint32_t n = 0, start = 1; };
	std::map<int, GcDev *> contig_gc;
	bool zero_decoded = false;                   // bsgpu_decode_records hands the decoded arrays out: slots of dropped records are cleared
	uint64_t guard_base[4] = {0, 0, 0, 0};      // guard counters of the device before the last bsgpu_guard_read(reset)
	int launches = 0;
	// --report-file side channels (bsgpu_profile_enable)
	bool profile_on = false;
	ProfDev *d_prof = nullptr;
	// writer side
	DevBuf wr_site, wr_cta, wr_out, wr_vcf, wr_ref, wr_blocks, wr_ring[3];
	cudaEvent_t wr_built[3] = {nullptr, nullptr, nullptr}, wr_copied[3] = {nullptr, nullptr, nullptr};
	unsigned long long *d_wr_totals = nullptr;   // 8 slots of (bytes, records, oversized)
	unsigned long long *h_wr_totals = nullptr;   // pinned mirror
	double tm_prep = 0, tm_queue = 0, tm_collect = 0;      // BSGPU_TIMING: host time of bsgpu_call_bam's window loop
	uint64_t reader_tally[30] = {0};             // read_input's filter_cts[15] | filter_bases[15] (host side, bsgpu_call_bam)
	DevBuf prof_scratch;                         // used16 | cand | chunkmax of the window being normalised
	int prof_parity = 0;                         // which ProfDev::used[] holds the running value
	ProfArgs prof_args;
};

// profile arguments of a normalise launch over n templates and a window of sz sites (NULL while the profile is off)
static const ProfArgs *profile_for(bsgpu_ctx *c, size_t n, uint32_t sz, cudaError_t *err) {
	*err = cudaSuccess;
	if (!c->profile_on || !n) return nullptr;
	const size_t nchunks = (n + kProfChunk - 1) / kProfChunk;
	const size_t a = (n * 2 + 15) & ~(size_t)15, b = (n + 15) & ~(size_t)15;
	*err = c->prof_scratch.reserve(a + b + nchunks * 4);
	if (*err != cudaSuccess) return nullptr;
	ProfArgs &pa = c->prof_args;
	pa.ref = (const uint8_t *)c->ref.p;
	pa.refn = sz + 1;
	pa.min_qual = c->params.min_qual;
	pa.prof = c->d_prof;
	pa.used16 = (uint16_t *)c->prof_scratch.p;
	pa.cand = (uint8_t *)c->prof_scratch.p + a;
	pa.chunkmax = (uint32_t *)((uint8_t *)c->prof_scratch.p + a + b);
	return &pa;
}

extern "C" {

int bsgpu_version(void) { return 100; }

const char *bsgpu_last_error(void) { return g_err; }

void bsgpu_default_params(bsgpu_params *p) {
	memset(p, 0, sizeof(*p));
	p->under_conv = 0.01;      // include/bs_call.h:16-18
	p->over_conv = 0.05;
	p->ref_bias = 2.0;
	p->min_qual = 20;          // include/bs_call.h:26
	p->device = 0;
}

int bsgpu_init(const bsgpu_params *p, bsgpu_ctx **out) {
	if (!p || !out) return fail("bsgpu_init: null argument");
	if (p->min_qual < 1 || p->min_qual > BSGPU_MAX_QUAL) return fail("bsgpu_init: min_qual must be in [1,%d]", BSGPU_MAX_QUAL);
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) return fail("bsgpu_init: no CUDA device (%s); there is no CPU fallback", e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
	if (p->device < 0 || p->device >= ndev) return fail("bsgpu_init: device %d out of range (have %d)", p->device, ndev);
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, p->device));
	if (prop.major != 10) return fail("bsgpu_init: device %d is sm_%d%d; this library carries sm_100a code only", p->device, prop.major, prop.minor);
	CU(cudaSetDevice(p->device));
#define CUI(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { bsgpu_destroy(c); return fail("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
	bsgpu_ctx *c = new bsgpu_ctx();
	c->device = p->device;
	c->params = *p;
	g_ctx_on_device[c->device & 63].fetch_add(1);
	memset(&c->stats, 0, sizeof(c->stats));
	// host-side tables, computed with the C library exactly as the reference does
	std::vector<DevConst> hbuf(1);          // (48 KB: not on the stack, not shared between threads that initialise contexts)
	DevConst &h = hbuf[0];
	memset(&h, 0, sizeof(h));
	for (int q = 0; q <= kMaxQual; q++) {          // src/genotype_model.c:10-21
		double er = exp(-.1 * (double)q * kLn10);
		if (er > .5) er = .5;
		const double k = er / (3.0 - 4.0 * er);
		h.tab.qp[q][0] = k;
		h.tab.qp[q][1] = log(k);
		h.tab.qp[q][2] = log(0.5 + k);
		h.tab.qp[q][3] = log(1.0 + k);
	}
	build_math_tables(&h.tab.math);
	double acc = 0.0;                              // src/stats_utils.c:14-21
	h.lfact[0] = h.lfact[1] = 0.0;
	for (int i = 2; i < 256; i++) { acc += log((double)i); h.lfact[i] = acc; }
	h.l = 1.0 - p->under_conv;                     // src/genotype_model.c:47-48
	h.t = p->over_conv;
	h.lrb = log(p->ref_bias);                      // src/genotype_model.c:88-89
	h.lrb1 = log(0.5 * (1.0 + p->ref_bias));
	h.min_qual = p->min_qual;
	for (int b = 0; b < 256; b++) {                // a base counts iff min_qual <= q != 63 (src/call_genotypes.c:217)
		const uint32_t q = (uint32_t)b >> 2, base = (uint32_t)b & 3u;
		const uint32_t inc = (q >= p->min_qual && q != BSGPU_FLT_QUAL) ? (1u | q << 5) << (16 * (base & 1u)) : 0u;
		h.pile_lut[b][0] = base < 2 ? inc : 0u;
		h.pile_lut[b][1] = base < 2 ? 0u : inc;
	}
	CUI(cudaMalloc(&c->d_const, sizeof(DevConst)));
	CUI(cudaMemcpy(c->d_const, &h, sizeof(h), cudaMemcpyHostToDevice));
	CUI(cudaMalloc(&c->d_counters, (kGuardList + kGuardCap) * sizeof(unsigned long long)));
	CUI(cudaMemset(c->d_counters, 0, (kGuardList + kGuardCap) * sizeof(unsigned long long)));
	CUI(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	CUI(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
	for (int i = 0; i < 2; i++) {
		CUI(cudaStreamCreateWithFlags(&c->slot[i].stream, cudaStreamNonBlocking));
		CUI(cudaEventCreateWithFlags(&c->slot[i].done, cudaEventDisableTiming));
	}
	CUI(configure_kernels());
	CUI(configure_writer());
	CUI(cudaMalloc(&c->d_wr_totals, 8 * 3 * sizeof(unsigned long long)));
	CUI(cudaHostAlloc(&c->h_wr_totals, 8 * 3 * sizeof(unsigned long long), cudaHostAllocDefault));
	{ const char *e = getenv("BSGPU_FUSED"); c->fused = e && atoi(e) == 1; }
	{ const char *e = getenv("BSGPU_BLOCK_OVERLAP"); c->block_overlap = e && atoi(e) != 0; }
#undef CUI
	*out = c;
	return BSGPU_OK;
}

void bsgpu_destroy(bsgpu_ctx *c) {
	if (!c) return;
	cudaSetDevice(c->device);
	cudaDeviceSynchronize();
	g_ctx_on_device[c->device & 63].fetch_sub(1);
	c->st_main.release(); c->st_ctg.release();
	for (auto &kv : c->contig_gc) if (kv.second) { kv.second->bins.release(); delete kv.second; }
	c->contig_gc.clear();
	for (int i = 0; i < 2; i++) {
		c->slot[i].in.release(); c->slot[i].ref.release(); c->slot[i].out.release(); c->slot[i].skip.release(); c->slot[i].wire.release();
		if (c->slot[i].stream) cudaStreamDestroy(c->slot[i].stream);
		if (c->slot[i].done) cudaEventDestroy(c->slot[i].done);
	}
	delete c->expander;
	for (PinBuf &b : c->wire_pin) b.release();
	for (cudaEvent_t ev : c->wire_landed) if (ev) cudaEventDestroy(ev);
	for (cudaEvent_t ev : c->win_events) cudaEventDestroy(ev);
	for (cudaEvent_t ev : c->pile_ev) if (ev) cudaEventDestroy(ev);
	if (c->pile_stream) cudaStreamDestroy(c->pile_stream);
	c->pile2.release();
	for (bsgpu_ctx::ReaderSet &R : c->rs) {
		for (cudaEvent_t ev : R.rd_up) cudaEventDestroy(ev);
		for (cudaEvent_t ev : R.rd_done) cudaEventDestroy(ev);
		if (R.frame_scratch) frame_scratch_free(R.frame_scratch);
		R.h_rec.release(); R.h_off.release(); R.h_tmpl.release(); R.h_key.release(); R.rd_key.release(); R.h_mask.release(); R.rd_mask.release(); R.rd_scan.release();
		R.rd_names.release(); R.rd_nameid.release(); R.h_nameid.release();
		R.rd_bam.release(); R.rd_recoff.release(); R.rd_readoff.release(); R.rd_mmoff.release(); R.rd_rec.release(); R.rd_bases.release(); R.rd_misms.release();
		if (R.up && R.up != c->slot[0].stream) cudaStreamDestroy(R.up);
	}
	c->segs.release(); c->bases.release(); c->ref.release(); c->scratch.release(); c->vcf.release(); c->tmpl.release(); c->misms.release(); c->obases.release(); c->ooff.release(); c->pile.release();
	if (c->stream) cudaStreamDestroy(c->stream);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	if (c->d_const) cudaFree(c->d_const);
	if (c->d_counters) cudaFree(c->d_counters);
	if (c->d_prof) cudaFree(c->d_prof);
	c->prof_scratch.release();
	c->wr_site.release(); c->wr_cta.release(); c->wr_out.release(); c->wr_vcf.release(); c->wr_ref.release(); c->wr_blocks.release();
	for (int i = 0; i < 3; i++) { c->wr_ring[i].release(); if (c->wr_built[i]) cudaEventDestroy(c->wr_built[i]); if (c->wr_copied[i]) cudaEventDestroy(c->wr_copied[i]); }
	if (c->d_wr_totals) cudaFree(c->d_wr_totals);
	if (c->h_wr_totals) cudaFreeHost(c->h_wr_totals);
	for (auto &kv : c->contig_ann) if (kv.second) { for (DevBuf *b : {&kv.second->mask, &kv.second->fq, &kv.second->cum, &kv.second->off, &kv.second->names}) b->release(); delete kv.second; }
	for (DevBuf *b : {&c->call_db.mask, &c->call_db.fq, &c->call_db.cum, &c->call_db.off, &c->call_db.names}) b->release();
	delete c;
}

int bsgpu_get_stats(bsgpu_ctx *c, bsgpu_stats *out) {
	if (!c || !out) return fail("bsgpu_get_stats: null argument");
	CU(cudaSetDevice(c->device));
	unsigned long long h[16];
	CU(cudaMemcpy(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
	c->stats.long_segments = h[9];
	c->stats.kernel_launches = (uint64_t)c->launches;
	c->stats.sites_called = h[0];
	c->stats.qsum_overflow = h[1];
	c->stats.near_tie_sites = h[4] + c->guard_base[0];
	c->stats.exact_tie_sites = h[5] + c->guard_base[1];
	c->stats.near_qual_sites = h[6] + c->guard_base[2];
	c->stats.near_fs_sites = h[7] + c->guard_base[3];
	*out = c->stats;
	return BSGPU_OK;
}

void *bsgpu_host_alloc(size_t bytes) {
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { fail("bsgpu_host_alloc(%zu) failed", bytes); return nullptr; }
	return p;
}

void bsgpu_host_free(void *p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------------------------------------
// device-pointer entry points
// ------------------------------------------------------------------------------------------------
// does a likelihood-kernel launch of this context have to expect other kernels on its SMs?  `own`: the caller itself runs
// kernels on more than one stream
static bool overlap_expected(const bsgpu_ctx *c, bool own) { return own || c->overlap_kernels || g_ctx_on_device[c->device & 63].load() > 1; }

int bsgpu_call_sites_dev(bsgpu_ctx *c, const void *d_pileup, const void *d_ref, size_t n, void *d_out, void *d_skip, void *stream) {
	if (!c) return fail("bsgpu_call_sites_dev: null context");
	if (n && (!d_pileup || !d_ref || !d_out || !d_skip)) return fail("bsgpu_call_sites_dev: null buffer");
	if (((uintptr_t)d_pileup | (uintptr_t)d_out) & 7u) return fail("bsgpu_call_sites_dev: record arrays must be 8-byte aligned");
	CU(cudaSetDevice(c->device));
	CU(launch_call_sites(d_pileup, d_ref, n, d_out, d_skip, false, c->d_const, c->d_counters, stream ? (cudaStream_t)stream : c->stream, &c->launches, 0, overlap_expected(c, false)));
	__atomic_fetch_add(&c->stats.sites, (uint64_t)(n), __ATOMIC_RELAXED);
	return BSGPU_OK;
}

int bsgpu_call_sites_vcf_dev(bsgpu_ctx *c, const void *d_pileup, const void *d_ref, size_t n, void *d_vcf, void *stream) {
	if (!c) return fail("bsgpu_call_sites_vcf_dev: null context");
	if (n && (!d_pileup || !d_ref || !d_vcf)) return fail("bsgpu_call_sites_vcf_dev: null buffer");
	if (((uintptr_t)d_pileup | (uintptr_t)d_vcf) & 7u) return fail("bsgpu_call_sites_vcf_dev: record arrays must be 8-byte aligned");
	CU(cudaSetDevice(c->device));
	CU(launch_call_sites(d_pileup, d_ref, n, d_vcf, nullptr, true, c->d_const, c->d_counters, stream ? (cudaStream_t)stream : c->stream, &c->launches, 0, overlap_expected(c, false)));
	__atomic_fetch_add(&c->stats.sites, (uint64_t)(n), __ATOMIC_RELAXED);
	return BSGPU_OK;
}

size_t bsgpu_block_scratch_bytes(size_t nseg, uint32_t sz) { return pileup_scratch_bytes(nseg, sz); }

static int call_bins(bsgpu_ctx *c, size_t nseg, const void *d_bases, const void *d_ref, uint32_t x, uint32_t sz,
		uint32_t t0, uint32_t nt, void *dout, cudaStream_t st);

static int block_dev(bsgpu_ctx *c, const void *d_segs, size_t nseg, const void *d_bases, const void *d_ref, uint32_t x, uint32_t sz,
		void *d_out, int mode, void *d_scratch, void *stream) {
	if (!c) return fail("bsgpu block: null context");
	if (!sz) return BSGPU_OK;
	if (!d_out || (nseg && (!d_segs || !d_bases)) || (mode && !d_ref)) return fail("bsgpu block: null buffer");
	if ((uintptr_t)d_out & 15u) return fail("bsgpu block: output array must be 16-byte aligned");
	if (nseg && ((uintptr_t)d_bases & 15u)) return fail("bsgpu block: bases[] must be 16-byte aligned (and end with 16 readable bytes of slack)");
	CU(cudaSetDevice(c->device));
	cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
	void *scratch = d_scratch;
	if (!scratch) {
		// growing the context's scratch frees the old allocation: make sure nothing queued earlier still uses it
		if (pileup_scratch_bytes(nseg, sz) > c->scratch.cap) CU(cudaDeviceSynchronize());
		CU(c->scratch.reserve(pileup_scratch_bytes(nseg, sz)));
		scratch = c->scratch.p;
	}
	CU(launch_bin_segments(d_segs, nseg, x, sz, scratch, st, &c->launches, c->d_counters));
	const uint32_t ntiles = (sz + kPileTileSites - 1) / kPileTileSites;
	if (mode && !c->fused) {
		const size_t need = (size_t)sz * sizeof(bsgpu_pileup) + 16;
		if (need > c->pile.cap) CU(cudaDeviceSynchronize());
		CU(c->pile.reserve(need));
		if (call_bins(c, nseg, d_bases, d_ref, x, sz, 0, ntiles, d_out, st) != BSGPU_OK) return BSGPU_FAIL;
	} else {
		CU(launch_pileup_tiles(scratch, nseg, d_bases, d_ref, x, sz, 0, ntiles, d_out, mode, c->d_const, c->d_counters, st, &c->launches));
	}
	__atomic_fetch_add(&c->stats.sites, (uint64_t)(sz), __ATOMIC_RELAXED);
	return BSGPU_OK;
}

int bsgpu_pileup_block_dev(bsgpu_ctx *c, const void *d_segs, size_t nseg, const void *d_bases, uint32_t x, uint32_t sz, void *d_pileup_out, void *stream) {
	return block_dev(c, d_segs, nseg, d_bases, nullptr, x, sz, d_pileup_out, 0, nullptr, stream);
}

int bsgpu_call_block_dev(bsgpu_ctx *c, const void *d_segs, size_t nseg, const void *d_bases, const void *d_ref, uint32_t x, uint32_t sz, void *d_vcf_out, void *stream) {
	return block_dev(c, d_segs, nseg, d_bases, d_ref, x, sz, d_vcf_out, 1, nullptr, stream);
}

int bsgpu_synth_sites_dev(bsgpu_ctx *c, uint64_t seed, uint64_t first_site, size_t n, double mean_depth, void *d_pileup, void *d_ref, void *stream) {
	if (!c) return fail("bsgpu_synth_sites_dev: null context");
	CU(cudaSetDevice(c->device));
	CU(launch_synth_sites(seed, first_site, n, mean_depth, d_pileup, d_ref, stream ? (cudaStream_t)stream : c->stream, &c->launches));
	return BSGPU_OK;
}

size_t bsgpu_synth_block_nseg(uint32_t sz, uint32_t read_len, double depth) { return synth_block_nseg(sz, read_len, depth); }

size_t bsgpu_synth_bam_bytes(size_t ntemplates, uint32_t read_len) { return synth_bam_bytes(ntemplates, read_len); }

int bsgpu_synth_bam_dev(bsgpu_ctx *c, uint64_t seed, size_t ntemplates, uint32_t read_len, const void *d_pos_f, const void *d_pos_r,
		const void *d_src, const void *d_rank, void *d_out, void *stream) {
	if (!c) return fail("bsgpu_synth_bam_dev: null context");
	if (read_len < 16 || read_len > BSGPU_MAX_SEG_LEN) return fail("bsgpu_synth_bam_dev: read_len must be in [16,%d]", BSGPU_MAX_SEG_LEN);
	if (ntemplates && (!d_pos_f || !d_pos_r || !d_src || !d_rank || !d_out)) return fail("bsgpu_synth_bam_dev: null buffer");
	CU(cudaSetDevice(c->device));
	CU(launch_synth_bam(seed, ntemplates, read_len, d_pos_f, d_pos_r, d_src, d_rank, d_out, stream ? (cudaStream_t)stream : c->stream, &c->launches));
	return BSGPU_OK;
}

int bsgpu_synth_ref_dev(bsgpu_ctx *c, uint64_t seed, uint32_t x, uint32_t sz, void *d_ref, void *stream) {
	if (!c) return fail("bsgpu_synth_ref_dev: null context");
	CU(cudaSetDevice(c->device));
	CU(launch_synth_ref(seed, x, sz, d_ref, stream ? (cudaStream_t)stream : c->stream, &c->launches));
	return BSGPU_OK;
}

int bsgpu_synth_block_dev(bsgpu_ctx *c, uint64_t seed, uint32_t x, uint32_t sz, uint32_t read_len, double depth,
		void *d_segs, size_t seg_cap, void *d_bases, size_t base_cap, void *d_ref, size_t *nseg, size_t *nbases, void *stream) {
	if (!c) return fail("bsgpu_synth_block_dev: null context");
	if (read_len < 16 || read_len > BSGPU_MAX_SEG_LEN) return fail("bsgpu_synth_block_dev: read_len must be in [16,%d]", BSGPU_MAX_SEG_LEN);
	const size_t ns = synth_block_nseg(sz, read_len, depth);
	if (ns > seg_cap || ns * read_len > base_cap) return fail("bsgpu_synth_block_dev: need %zu segments / %zu bases", ns, ns * read_len);
	if (ns * (size_t)read_len > 0xffffffffull) return fail("bsgpu_synth_block_dev: more than 4 GiB of bases in one block");
	CU(cudaSetDevice(c->device));
	CU(launch_synth_block(seed, x, sz, read_len, depth, d_segs, d_bases, d_ref, stream ? (cudaStream_t)stream : c->stream, &c->launches));
	if (nseg) *nseg = ns;
	if (nbases) *nbases = ns * read_len;
	return BSGPU_OK;
}

// host evaluation of the device's table-driven log / exp (same source, both FMA-exact): lets CPU tests measure them
int bsgpu_math_probe(const double *x, size_t n, double *out_log, double *out_exp) {
	static MathTables mt;
	static bool built = false;
	if (!built) { build_math_tables(&mt); built = true; }
	for (size_t i = 0; i < n; i++) {
		if (out_log) out_log[i] = fast_log(x[i], &mt);
		if (out_exp) out_exp[i] = fast_exp(x[i], &mt);
	}
	return BSGPU_OK;
}

// host side of the wire records (bsgpu_wire.h), exported for tests and for hosts that want the rebuilding pass on its own
int bsgpu_wire_pack(const void *records, size_t n, size_t rec_bytes, const uint8_t *skip, void *wire) {
	if (rec_bytes != sizeof(bsgpu_gt_meth) && rec_bytes != sizeof(bsgpu_gt_vcf)) return fail("bsgpu_wire_pack: rec_bytes must be 200 or 208");
	if (n && (!records || !wire || (rec_bytes == sizeof(bsgpu_gt_meth) && !skip))) return fail("bsgpu_wire_pack: null buffer");
	return wire_pack_host((const uint8_t *)records, n, rec_bytes, skip, (uint64_t *)wire) ? BSGPU_OK : fail("bsgpu_wire_pack: a field does not fit its wire width");
}

int bsgpu_wire_expand(const void *wire, size_t n, size_t rec_bytes, void *records, uint8_t *skip, int threads) {
	if (rec_bytes != sizeof(bsgpu_gt_meth) && rec_bytes != sizeof(bsgpu_gt_vcf)) return fail("bsgpu_wire_expand: rec_bytes must be 200 or 208");
	if (n && (!records || !wire || (rec_bytes == sizeof(bsgpu_gt_meth) && !skip))) return fail("bsgpu_wire_expand: null buffer");
	if (((uintptr_t)wire | (uintptr_t)records) & 7) return fail("bsgpu_wire_expand: buffers must be 8-byte aligned");
	if (!n) return BSGPU_OK;
	if (threads <= 1) { wire_expand((const uint64_t *)wire, n, (uint8_t *)records, rec_bytes, skip); return BSGPU_OK; }
	// through the pool, as the entry points do: one job per 256 Ki sites
	WireExpander pool((unsigned)threads);
	const size_t chunk = (size_t)1 << 18;
	pool.set_buffers((int)((n + chunk - 1) / chunk));
	size_t id = 0;
	for (size_t a = 0; a < n; a += chunk, id++) {
		WireExpander::Job j;
		j.wire = (const uint64_t *)wire + a * kWireWords; j.n = n - a < chunk ? n - a : chunk; j.out = (uint8_t *)records + a * rec_bytes;
		j.skip = skip ? skip + a : nullptr; j.rec_bytes = rec_bytes; j.buf = pool.acquire(); j.chunk_id = id;
		pool.submit(j);
	}
	pool.wait_idle();
	return BSGPU_OK;
}

// Sites whose result sits inside a guard band since the last reset: ids[k] = kind << 56 | site id, where the id is the
// index of the site in the bsgpu_call_sites / _bcf call, or its position for the block and reader entry points.
int bsgpu_guard_read(bsgpu_ctx *c, uint64_t *ids, size_t cap, size_t *n, int reset) {
	if (!c || !n || (cap && !ids)) return fail("bsgpu_guard_read: null argument");
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());
	unsigned long long h[kGuardList];
	CU(cudaMemcpy(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
	const size_t have = (size_t)std::min<unsigned long long>(h[8], (unsigned long long)kGuardCap);
	*n = std::min(have, cap);
	if (*n) CU(cudaMemcpy(ids, c->d_counters + kGuardList, *n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
	if (reset) {
		for (int k = 0; k < 4; k++) c->guard_base[k] += h[4 + k];
		CU(cudaMemset(c->d_counters + 4, 0, 5 * sizeof(unsigned long long)));      // the four counters and the list length ([9]: long_segments stays)
	}
	return BSGPU_OK;
}

int bsgpu_sync(bsgpu_ctx *c) {
	if (!c) return fail("bsgpu_sync: null context");
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());
	return BSGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// host-buffer entry points
// ------------------------------------------------------------------------------------------------
// The pool that rebuilds records from wire chunks, made on first use: BSGPU_EXPAND_THREADS, else half of the cores (2..12)
static WireExpander *wire_pool(bsgpu_ctx *c) {
	if (!c->expander) {
		const char *e = getenv("BSGPU_EXPAND_THREADS");
		const int hw = (int)std::thread::hardware_concurrency();
		int t = e ? atoi(e) : 0;
		if (t <= 0) t = std::max(2, std::min(hw / 2, 12));
		c->expander = new WireExpander((unsigned)t);
		const char *m = getenv("BSGPU_WIRE");
		c->wire_mode = m ? atoi(m) : 0;
	}
	return c->expander;
}

static void wire_wait_event(void *ev) { cudaEventSynchronize((cudaEvent_t)ev); }      // the pool's watcher thread: a chunk has landed

// pileup[] -> gt_meth[] + skip[]: chunks ping-pong between two slots, each with its own stream, so the H2D of chunk
// i+1 and the D2H of chunk i-1 overlap the kernel of chunk i (PCIe is full duplex).  The way home is the narrow one
// (201 against 105 bytes per site).  BSGPU_WIRE=1|2 sends a chunk's records as 120-byte wire records instead (k_wire_pack):
// they land in one of a few pinned buffers and are rebuilt in `out` / `skip` by the expander pool while later chunks are
// under way; with 1, a chunk goes home as full records when every landing buffer is still being rebuilt.
// Measured (profiles/r02d_wire_ab.txt, 16 M sites, 16-core host of the B200 box): full records 238 - 248 M sites/s
// (PCIe-bound, 50 GB/s home); wire 220 - 276 M whatever the pool size -- the rebuilding pass alone does 357 M sites/s on 8
// threads and 464 M on 16 (149 GB/s of host memory traffic), but next to the two DMA streams the host's memory system is
// the limit: wire records cost it 545 bytes per site (105 read by the H2D DMA, 120 written by the D2H DMA, 120 read and
// 200 written by the pool) against 306 for full records.  Bit-identical either way; off by default.
int bsgpu_call_sites(bsgpu_ctx *c, const bsgpu_pileup *pileup, const uint8_t *ref, size_t n, bsgpu_gt_meth *out, uint8_t *skip) {
	if (!c) return fail("bsgpu_call_sites: null context");
	if (!n) return BSGPU_OK;
	if (!pileup || !ref || !out || !skip) return fail("bsgpu_call_sites: null buffer");
	CU(cudaSetDevice(c->device));
	static const size_t chunk = [] { const char *e = getenv("BSGPU_SITES_CHUNK"); const long long v = e ? atoll(e) : 0; return v > 0 ? (size_t)v : (size_t)1 << 18; }();      // 256 Ki sites: 252 M sites/s against 243 M with 1 Mi (the pipeline fills sooner)
	WireExpander *pool = wire_pool(c);
	const int nland = (int)(sizeof(c->wire_pin) / sizeof(c->wire_pin[0]));
	const size_t wire_bytes = chunk * kWireBytes + 8;
	if (c->wire_mode) {
		pool->wait_idle();
		for (PinBuf &b : c->wire_pin) CU(b.reserve(wire_bytes));
		for (cudaEvent_t &ev : c->wire_landed) if (!ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
		pool->set_buffers(nland);
		pool->failed();
	}
	uint64_t wire_sites = 0, d2h = 0;
	auto run_chunk = [&](size_t ci, size_t first, bool allow_wire) -> int {
		Slot &s = c->slot[ci & 1];
		const size_t m = n - first < chunk ? n - first : chunk;
		CU(cudaStreamSynchronize(s.stream));      // the slot's previous chunk has fully left the device
		CU(s.in.reserve(m * sizeof(bsgpu_pileup)));
		CU(s.ref.reserve(m));
		CU(s.out.reserve(m * sizeof(bsgpu_gt_meth)));
		CU(s.skip.reserve(m));
		int land = -1;
		if (allow_wire && c->wire_mode) {
			land = pool->acquire();
			while (land < 0 && c->wire_mode == 2) { std::this_thread::yield(); land = pool->acquire(); }
		}
		CU(cudaMemcpyAsync(s.in.p, pileup + first, m * sizeof(bsgpu_pileup), cudaMemcpyHostToDevice, s.stream));
		CU(cudaMemcpyAsync(s.ref.p, ref + first, m, cudaMemcpyHostToDevice, s.stream));
		CU(launch_call_sites(s.in.p, s.ref.p, m, s.out.p, s.skip.p, false, c->d_const, c->d_counters, s.stream, &c->launches, first, true));      // two chunk streams
		if (land >= 0) {
			cudaError_t e = s.wire.reserve(wire_bytes);
			if (e == cudaSuccess) e = launch_wire_pack(s.out.p, s.skip.p, m, s.wire.p, s.stream, &c->launches);
			if (e == cudaSuccess) e = cudaMemcpyAsync(c->wire_pin[land].p, s.wire.p, m * kWireBytes + 8, cudaMemcpyDeviceToHost, s.stream);
			if (e != cudaSuccess) { pool->give_back(land); return fail("bsgpu_call_sites: wire stage: %s", cudaGetErrorString(e)); }
			WireExpander::Job job;
			job.wire = (const uint64_t *)c->wire_pin[land].p;
			job.flag = job.wire + m * kWireWords;
			job.n = m;
			job.out = (uint8_t *)(out + first);
			job.skip = skip + first;
			job.rec_bytes = sizeof(bsgpu_gt_meth);
			job.buf = land;
			job.chunk_id = ci;
			job.wait_landed = wire_wait_event;
			job.wait_arg = c->wire_landed[land];
			e = cudaEventRecord(c->wire_landed[land], s.stream);
			if (e != cudaSuccess) { pool->give_back(land); return fail("bsgpu_call_sites: cudaEventRecord: %s", cudaGetErrorString(e)); }
			pool->submit(job);
			wire_sites += m;
			d2h += m * kWireBytes + 8;
		} else {
			CU(cudaMemcpyAsync(out + first, s.out.p, m * sizeof(bsgpu_gt_meth), cudaMemcpyDeviceToHost, s.stream));
			CU(cudaMemcpyAsync(skip + first, s.skip.p, m, cudaMemcpyDeviceToHost, s.stream));
			d2h += m * (sizeof(bsgpu_gt_meth) + 1);
		}
		__atomic_fetch_add(&c->stats.h2d_bytes, (uint64_t)(m * (sizeof(bsgpu_pileup) + 1)), __ATOMIC_RELAXED);
		return BSGPU_OK;
	};
	size_t ci = 0;
	int rc = BSGPU_OK;
	for (size_t first = 0; first < n && rc == BSGPU_OK; first += chunk, ci++) rc = run_chunk(ci, first, true);
	cudaError_t e0 = cudaStreamSynchronize(c->slot[0].stream), e1 = cudaStreamSynchronize(c->slot[1].stream);
	if (c->wire_mode) pool->wait_idle();          // also on failure: no thread may still be writing into the caller's arrays
	if (rc != BSGPU_OK) return rc;
	CU(e0);
	CU(e1);
	// chunks with a field too wide for the wire: once more, as full records
	if (c->wire_mode) for (size_t bad : pool->failed()) {
		const size_t m = n - bad * chunk < chunk ? n - bad * chunk : chunk;
		if (run_chunk(bad, bad * chunk, false) != BSGPU_OK) return BSGPU_FAIL;
		CU(cudaStreamSynchronize(c->slot[bad & 1].stream));
		wire_sites -= m;
		__atomic_fetch_add(&c->stats.wire_refetched_chunks, (uint64_t)1, __ATOMIC_RELAXED);
	}
	__atomic_fetch_add(&c->stats.d2h_bytes, d2h, __ATOMIC_RELAXED);
	__atomic_fetch_add(&c->stats.wire_sites, wire_sites, __ATOMIC_RELAXED);
	__atomic_fetch_add(&c->stats.sites, (uint64_t)(n), __ATOMIC_RELAXED);
	return BSGPU_OK;
}

// pileup + model for bins [t0, t0 + nt) of a binned block, into `dout` (gt_vcf records).  Default: the gather kernel
// writes pileup[] to an HBM scratch and the persistent likelihood kernel turns it into gt_vcf[] -- measured 1.3x faster
// than the single fused kernel (profiles/), whose register footprint (the model's) starves the gather of resident warps.
static int call_bins(bsgpu_ctx *c, size_t nseg, const void *d_bases, const void *d_ref, uint32_t x, uint32_t sz,
		uint32_t t0, uint32_t nt, void *dout, cudaStream_t st) {
	if (c->fused) {
		CU(launch_pileup_tiles(c->scratch.p, nseg, d_bases, d_ref, x, sz, t0, nt, dout, 1, c->d_const, c->d_counters, st, &c->launches));
		return BSGPU_OK;
	}
	// The count vectors go from the pileup kernel to the model through an HBM scratch.  BSGPU_SUBSLAB_TILES=n hands them over in
	// sub-slabs of n tiles that fit the L2 (126 MB), alternating between two halves of the scratch.  Measured (50 M sites, 30x):
	// one launch each 5.54 ms, n = 8192: 6.59, 4096: 7.31, 2048: 8.95 -- the tails of the smaller launches cost more than the
	// cached read-back saves (neither kernel is HBM-bound), so the default is 0: one launch each.
	static const uint32_t sub = [] { const char *e = getenv("BSGPU_SUBSLAB_TILES"); const long v = e ? atol(e) : 0; return (uint32_t)(v < 0 ? 0 : v); }();
	// BSGPU_BLOCK_OVERLAP=1 (see block_run): a window that comes in one piece (the _dev entry points; c->pile then holds the count
	// vectors of the whole window) is cut into parts of 16 Ki tiles; the gather of part k + 1 runs on the context's pile stream
	// next to the model of part k.
	const uint32_t part = (2u << 20) / kPileTileSites;
	if (c->block_overlap && !sub && nt > part && c->pile.cap >= ((size_t)nt * kPileTileSites) * sizeof(bsgpu_pileup)) {
		if (!c->pile_stream) {
			CU(cudaStreamCreateWithFlags(&c->pile_stream, cudaStreamNonBlocking));
			for (cudaEvent_t &ev : c->pile_ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
		}
		CU(cudaEventRecord(c->pile_ev[4], st));
		CU(cudaStreamWaitEvent(c->pile_stream, c->pile_ev[4], 0));
		for (uint32_t k = 0, a = 0; a < nt; a += part, k++) {
			const uint32_t m = nt - a < part ? nt - a : part, ta = t0 + a;
			const size_t site0 = (size_t)ta * kPileTileSites, rel0 = (size_t)a * kPileTileSites;
			const size_t nsite = (size_t)sz - site0 < (size_t)m * kPileTileSites ? (size_t)sz - site0 : (size_t)m * kPileTileSites;
			uint8_t *pile = (uint8_t *)c->pile.p + rel0 * sizeof(bsgpu_pileup);
			CU(launch_pileup_tiles(c->scratch.p, nseg, d_bases, d_ref, x, sz, ta, m, pile, 0, c->d_const, c->d_counters, c->pile_stream, &c->launches));
			CU(cudaEventRecord(c->pile_ev[k & 1], c->pile_stream));
			CU(cudaStreamWaitEvent(st, c->pile_ev[k & 1], 0));
			CU(launch_call_sites(pile, (const uint8_t *)d_ref + site0, nsite, (uint8_t *)dout + rel0 * sizeof(bsgpu_gt_vcf), nullptr, true,
					c->d_const, c->d_counters, st, &c->launches, (unsigned long long)x + site0, true));
		}
		// whatever is queued on `st` next (the next window's binning rewrites the scratch the gather reads) comes after the last model
		// launch, which waited for the last gather
		return BSGPU_OK;
	}
	const uint32_t step = sub ? sub : nt;
	for (uint32_t k = 0, a = 0; a < nt; a += step, k++) {
		const uint32_t m = nt - a < step ? nt - a : step, ta = t0 + a;
		const size_t site0 = (size_t)ta * kPileTileSites;
		const size_t nsite = (size_t)sz - site0 < (size_t)m * kPileTileSites ? (size_t)sz - site0 : (size_t)m * kPileTileSites;
		// the scratch holds a whole slab (c->pile is sized by the callers): sub-slab k uses half (k & 1) of its first 2 * step tiles
		uint8_t *pile = (uint8_t *)c->pile.p + (sub ? (size_t)(k & 1) * step * kPileTileSites * sizeof(bsgpu_pileup) : 0);
		CU(launch_pileup_tiles(c->scratch.p, nseg, d_bases, d_ref, x, sz, ta, m, pile, 0, c->d_const, c->d_counters, st, &c->launches));
		CU(launch_call_sites(pile, (const uint8_t *)d_ref + site0, nsite, (uint8_t *)dout + (site0 - (size_t)t0 * kPileTileSites) * sizeof(bsgpu_gt_vcf), nullptr, true,
				c->d_const, c->d_counters, st, &c->launches, (unsigned long long)x + site0, overlap_expected(c, false)));
	}
	return BSGPU_OK;
}

// Device-resident segments / bases / ref -> host pileup[] or gt_vcf[].  The window is processed in slabs of tiles so
// that the D2H of slab i (copy stream) overlaps the kernel of slab i+1 (context stream); outputs are ~6x the inputs.
// `defer`: return without waiting for the device (the caller synchronises both streams before it touches `out` or lets
// go of the inputs); successive deferred runs overlap the D2H of one window with the kernels of the next
static int block_run(bsgpu_ctx *c, const void *d_segs, size_t nseg, const void *d_bases, const void *d_ref,
		uint32_t x, uint32_t sz, void *out, int mode, bool defer) {
	const size_t rec = mode ? sizeof(bsgpu_gt_vcf) : sizeof(bsgpu_pileup);
	CU(c->scratch.reserve(pileup_scratch_bytes(nseg, sz)));
	CU(launch_bin_segments(d_segs, nseg, x, sz, c->scratch.p, c->stream, &c->launches, c->d_counters));
	const uint32_t ntiles = (sz + kPileTileSites - 1) / kPileTileSites;
	const uint32_t slab = (2u << 20) / kPileTileSites;      // tiles per slab = 2 Mi sites = 436 MB of gt_vcf
	const uint32_t nslab = (ntiles + slab - 1) / slab;
	// ring of output slabs on the device.  Deferred runs rotate through all three slots with a fixed slot size, so the
	// kernels of one window never wait for the copy-out of the window before it.
	const uint32_t resident = defer ? 3 : (nslab < 3 ? nslab : 3);
	const uint32_t slab_tiles = defer ? slab : (ntiles < slab ? ntiles : slab);
	CU(c->vcf.reserve((size_t)resident * slab_tiles * kPileTileSites * rec));
	if (mode && !c->fused) CU(c->pile.reserve((size_t)(defer || ntiles >= slab ? slab : ntiles) * kPileTileSites * sizeof(bsgpu_pileup) + 16));
	// BSGPU_BLOCK_OVERLAP=1: the gather of slab i + 1 runs on a stream of its own next to the model of slab i (two count-vector
	// scratches, ping-pong).  The two kernels stall on different things -- the gather on instruction issue of integer / shared-memory
	// work, the model on the latency of dependent FP64 chains -- so CTAs of both on one SM fill each other's gaps.  The model
	// then takes its tiles by cp.async (launch_call_sites, overlap_safe).
	static const bool no_sub = getenv("BSGPU_SUBSLAB_TILES") == nullptr;
	const bool overlap = c->block_overlap && no_sub && mode && !c->fused && nslab > 1;
	if (overlap) {
		if (!c->pile_stream) {
			CU(cudaStreamCreateWithFlags(&c->pile_stream, cudaStreamNonBlocking));
			for (cudaEvent_t &ev : c->pile_ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
		}
		CU(c->pile2.reserve((size_t)slab * kPileTileSites * sizeof(bsgpu_pileup) + 16));
		CU(cudaEventRecord(c->pile_ev[4], c->stream));              // inputs uploaded, segments binned
		CU(cudaStreamWaitEvent(c->pile_stream, c->pile_ev[4], 0));
	}
	while (c->win_events.size() < 6) {
		cudaEvent_t ev;
		CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
		c->win_events.push_back(ev);
	}
	for (uint32_t si = 0; si < nslab; si++) {
		const uint32_t r = defer ? (c->ring_pos + si) % 3 : si % resident;
		cudaEvent_t computed = c->win_events[2 * r], copied = c->win_events[2 * r + 1];
		const uint32_t t0 = si * slab, nt = ntiles - t0 < slab ? ntiles - t0 : slab;
		const size_t site0 = (size_t)t0 * kPileTileSites;
		const size_t nsite = (size_t)sz - site0 < (size_t)nt * kPileTileSites ? (size_t)sz - site0 : (size_t)nt * kPileTileSites;
		uint8_t *dslab = (uint8_t *)c->vcf.p + (size_t)r * slab_tiles * kPileTileSites * rec;
		// ring slot drained: by this run, or by an earlier deferred run whose copy may still be in flight
		if (si >= resident || c->ring_busy[r]) CU(cudaStreamWaitEvent(c->stream, copied, 0));
		if (overlap) {
			const uint32_t b = c->pile_pos++ & 1;
			uint8_t *pile = (uint8_t *)(b ? c->pile2.p : c->pile.p);
			if (c->pile_used[b]) CU(cudaStreamWaitEvent(c->pile_stream, c->pile_ev[2 + b], 0));      // the model two slabs back has read this scratch
			CU(launch_pileup_tiles(c->scratch.p, nseg, d_bases, d_ref, x, sz, t0, nt, pile, 0, c->d_const, c->d_counters, c->pile_stream, &c->launches));
			CU(cudaEventRecord(c->pile_ev[b], c->pile_stream));
			CU(cudaStreamWaitEvent(c->stream, c->pile_ev[b], 0));
			CU(launch_call_sites(pile, (const uint8_t *)d_ref + site0, nsite, dslab, nullptr, true, c->d_const, c->d_counters, c->stream, &c->launches,
					(unsigned long long)x + site0, true));
			CU(cudaEventRecord(c->pile_ev[2 + b], c->stream));
			c->pile_used[b] = true;
		} else if (mode) { if (call_bins(c, nseg, d_bases, d_ref, x, sz, t0, nt, dslab, c->stream) != BSGPU_OK) return BSGPU_FAIL; }
		else CU(launch_pileup_tiles(c->scratch.p, nseg, d_bases, d_ref, x, sz, t0, nt, dslab, 0, c->d_const, c->d_counters, c->stream, &c->launches));
		CU(cudaEventRecord(computed, c->stream));
		CU(cudaStreamWaitEvent(c->copy_stream, computed, 0));
		CU(cudaMemcpyAsync((uint8_t *)out + site0 * rec, dslab, nsite * rec, cudaMemcpyDeviceToHost, c->copy_stream));
		CU(cudaEventRecord(copied, c->copy_stream));
		c->ring_busy[r] = defer;
		__atomic_fetch_add(&c->stats.d2h_bytes, (uint64_t)(nsite * rec), __ATOMIC_RELAXED);
	}
	if (defer) c->ring_pos = (c->ring_pos + nslab) % 3;
	__atomic_fetch_add(&c->stats.sites, (uint64_t)(sz), __ATOMIC_RELAXED);
	if (defer) return BSGPU_OK;
	CU(cudaStreamSynchronize(c->copy_stream));
	CU(cudaStreamSynchronize(c->stream));
	return BSGPU_OK;
}

// host segments + bases (+ ref) -> pileup[] or gt_vcf[]: inputs go up in one piece, then block_run
static int block_host(bsgpu_ctx *c, const bsgpu_seg *segs, size_t nseg, const uint8_t *bases, size_t nbases, const uint8_t *ref,
		uint32_t x, uint32_t sz, void *out, int mode) {
	if (!c) return fail("bsgpu block: null context");
	if (!sz) return BSGPU_OK;
	if (!out || (nseg && (!segs || !bases)) || (mode && !ref)) return fail("bsgpu block: null buffer");
	if (nbases > 0xffffffffull) return fail("bsgpu block: more than 4 GiB of bases in one block; split the window");
	for (size_t i = 0; i < nseg; i++) {
		if (segs[i].len > BSGPU_MAX_SEG_LEN) return fail("bsgpu block: segment %zu longer than %d (use bsgpu_stage_templates)", i, BSGPU_MAX_SEG_LEN);
		if ((size_t)segs[i].off + segs[i].len > nbases) return fail("bsgpu block: segment %zu points outside bases[]", i);
	}
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaStreamSynchronize(c->copy_stream));
	CU(c->segs.reserve(nseg * sizeof(bsgpu_seg) + 16));
	CU(c->bases.reserve(nbases + 16));
	CU(c->ref.reserve((size_t)sz + 16));
	if (nseg) {
		CU(cudaMemcpyAsync(c->segs.p, segs, nseg * sizeof(bsgpu_seg), cudaMemcpyHostToDevice, c->stream));
		CU(cudaMemcpyAsync(c->bases.p, bases, nbases, cudaMemcpyHostToDevice, c->stream));
	}
	if (mode) CU(cudaMemcpyAsync(c->ref.p, ref, sz, cudaMemcpyHostToDevice, c->stream));
	__atomic_fetch_add(&c->stats.h2d_bytes, (uint64_t)(nseg * sizeof(bsgpu_seg) + nbases + (mode ? sz : 0)), __ATOMIC_RELAXED);
	return block_run(c, c->segs.p, nseg, c->bases.p, c->ref.p, x, sz, out, mode, false);
}

// raw templates -> gt_vcf[]: normalisation, mate walk, pileup and model on the device (src/process_template.c:18-126)
int bsgpu_process_block(bsgpu_ctx *c, const bsgpu_template *t, size_t n, const uint8_t *bases, size_t nbases,
		const bsgpu_misms *misms, size_t nmisms, const uint8_t *ref, uint32_t y, uint32_t *x_out, bsgpu_gt_vcf *out) {
	if (!c) return fail("bsgpu_process_block: null context");
	if (!n) return fail("bsgpu_process_block: empty template list");       // the reference asserts ix > 0 (:22)
	if (!t || !bases || !ref || !out || (nmisms && !misms)) return fail("bsgpu_process_block: null buffer");
	// window: two positions before the first template's start (:24-28)
	uint32_t x = t[0].forward_position ? t[0].forward_position : t[0].reverse_position;
	if (!x || x > y) return fail("bsgpu_process_block: first template starts at %u, block ends at %u", x, y);
	x = x > 2 ? x - 2 : 1;
	const uint32_t sz = y - x + 1;
	// per-mate output slots: read length + reference bases the read lacks (CIGAR D -> INS events)
	std::vector<uint32_t> off(2 * n + 1);
	uint64_t tot = 0;
	uint32_t maxcap = 1;
	for (size_t i = 0; i < n; i++) for (int k = 0; k < 2; k++) {
		off[2 * i + k] = (uint32_t)tot;
		{   // the smaller position of a template must lie inside the window, read or no read there (src/call_genotypes.c:182-186)
			const uint32_t pk = k ? t[i].reverse_position : t[i].forward_position;
			if (pk && pk < x && !t[i].present[k]) return fail("bsgpu_process_block: template %zu: its absent mate lies before the window (%u < %u); the reference asserts there", i, pk, x);
		}
		if (!t[i].present[k]) continue;
		if ((size_t)t[i].read_off[k] + t[i].read_len[k] > nbases) return fail("bsgpu_process_block: template %zu read outside bases[]", i);
		if ((size_t)t[i].mm_off[k] + t[i].mm_n[k] > nmisms) return fail("bsgpu_process_block: template %zu events outside misms[]", i);
		if (t[i].bs_strand > 2) return fail("bsgpu_process_block: template %zu has bs_strand %u", i, t[i].bs_strand);
		uint64_t cap = t[i].read_len[k];
		for (uint32_t z = 0; z < t[i].mm_n[k]; z++) if (misms[t[i].mm_off[k] + z].type == 1) cap += misms[t[i].mm_off[k] + z].size;
		tot += cap;
		if (cap > maxcap) maxcap = (uint32_t)(cap > 0xffffffu ? 0xffffffu : cap);
	}
	off[2 * n] = (uint32_t)tot;
	if (tot > 0xffffffffull) return fail("bsgpu_process_block: more than 4 GiB of bases in one block; split the window");
	const uint32_t spm = (maxcap + BSGPU_MAX_SEG_LEN - 1) / BSGPU_MAX_SEG_LEN;
	const size_t nseg = n * 2 * (size_t)spm;
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaStreamSynchronize(c->copy_stream));
	CU(c->tmpl.reserve(n * sizeof(bsgpu_template)));
	CU(c->misms.reserve(nmisms * sizeof(bsgpu_misms) + 16));
	CU(c->bases.reserve(nbases + 16));
	CU(c->obases.reserve(tot + 16));
	CU(c->ooff.reserve((2 * n + 1) * sizeof(uint32_t)));
	CU(c->segs.reserve(nseg * sizeof(bsgpu_seg) + 16));
	CU(c->ref.reserve((size_t)sz + 16));
	CU(cudaMemcpyAsync(c->tmpl.p, t, n * sizeof(bsgpu_template), cudaMemcpyHostToDevice, c->stream));
	if (nmisms) CU(cudaMemcpyAsync(c->misms.p, misms, nmisms * sizeof(bsgpu_misms), cudaMemcpyHostToDevice, c->stream));
	if (nbases) CU(cudaMemcpyAsync(c->bases.p, bases, nbases, cudaMemcpyHostToDevice, c->stream));
	CU(cudaMemcpyAsync(c->ooff.p, off.data(), (2 * n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
	const size_t nref = (size_t)sz + (c->profile_on ? 1 : 0);      // the profile looks one code past the block (src/meth_profile.c:70)
	CU(cudaMemcpyAsync(c->ref.p, ref, nref, cudaMemcpyHostToDevice, c->stream));
	CU(cudaStreamSynchronize(c->stream));          // `off` is a pageable temporary: the copy must finish before it dies
	__atomic_fetch_add(&c->stats.h2d_bytes, (uint64_t)(n * sizeof(bsgpu_template) + nmisms * sizeof(bsgpu_misms) + nbases + (2 * n + 1) * 4 + nref), __ATOMIC_RELAXED);
	unsigned long long before[4], after[4];
	CU(cudaMemcpy(before, c->d_counters, sizeof(before), cudaMemcpyDeviceToHost));
	cudaError_t perr;
	const ProfArgs *pa = profile_for(c, n, sz, &perr);
	CU(perr);
	CU(launch_normalise(c->tmpl.p, n, c->bases.p, c->misms.p, c->ooff.p, c->obases.p, c->segs.p, spm, x, y,
			c->params.left_trim, c->params.right_trim, c->d_counters, pa, c->prof_parity, c->stream, &c->launches));
	if (pa) c->prof_parity ^= 1;
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaMemcpy(after, c->d_counters, sizeof(after), cudaMemcpyDeviceToHost));
	if (after[2] != before[2]) return fail("bsgpu_process_block: Error in CIGAR - illegal soft clip in %llu template(s)", after[2] - before[2]);
	if (after[3] != before[3]) return fail("bsgpu_process_block: %llu mate(s) start before the block window", after[3] - before[3]);
	if (x_out) *x_out = x;
	return block_run(c, c->segs.p, nseg, c->obases.p, c->ref.p, x, sz, out, 1, false);
}

// ------------------------------------------------------------------------------------------------
// writer side: gt_vcf[] -> BCF records
// ------------------------------------------------------------------------------------------------
void bsgpu_default_bcf_params(bsgpu_bcf_params *p) {
	memset(p, 0, sizeof(*p));
	for (int k = 0; k < 16; k++) p->ids[k] = k;      // the order print_vcf_header() adds them to an empty header
	p->ctg_end = 0xffffffffu;
}

static int check_totals(const unsigned long long *t, size_t out_cap, const char *who) {
	if (t[2]) return fail("%s: %llu record(s) longer than %d bytes", who, t[2], BSGPU_BCF_MAX_RECORD);
	if (t[0] > out_cap) return fail("%s: output buffer too small (%llu bytes needed, %zu given)", who, t[0], out_cap);
	return BSGPU_OK;
}

// dbSNP entries of a contig -> the device tables of DbView (synchronous; done once per contig / per distinct table)
static int db_upload(bsgpu_ctx *c, const bsgpu_dbsnp *db, bsgpu_ctx::DbDev &d, const char *who) {
	d.words = 0;
	if (!db || !db->n) return BSGPU_OK;
	if (!db->pos || !db->flags || !db->name_off || !db->names) return fail("%s: dbSNP table with a null array", who);
	const uint32_t n = db->n;
	for (uint32_t k = 0; k < n; k++) {
		if (!db->pos[k] || (k && db->pos[k] <= db->pos[k - 1])) return fail("%s: dbSNP positions must be 1-based, ascending and unique (entry %u)", who, k);
		if (db->name_off[k + 1] < db->name_off[k] || db->name_off[k + 1] - db->name_off[k] > BSGPU_DBSNP_MAX_ID) return fail("%s: dbSNP entry %u has an ID longer than %d bytes", who, k, BSGPU_DBSNP_MAX_ID);
	}
	const uint32_t words = (db->pos[n - 1] >> 6) + 1;
	std::vector<unsigned long long> mask(words, 0), fq(words, 0);
	std::vector<uint32_t> cum(words + 1, 0);
	for (uint32_t k = 0; k < n; k++) {
		const uint32_t p = db->pos[k];
		mask[p >> 6] |= 1ull << (p & 63);
		if (db->flags[k] & 2) fq[p >> 6] |= 1ull << (p & 63);
		cum[(p >> 6) + 1]++;
	}
	for (uint32_t w = 0; w < words; w++) cum[w + 1] += cum[w];
	const size_t nbytes = db->name_off[n];
	CU(cudaDeviceSynchronize());              // the old tables may still be in use
	CU(d.mask.reserve(words * 8)); CU(d.fq.reserve(words * 8)); CU(d.cum.reserve((words + 1) * 4));
	CU(d.off.reserve(((size_t)n + 1) * 4)); CU(d.names.reserve(nbytes + 16));
	CU(cudaMemcpy(d.mask.p, mask.data(), words * 8, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d.fq.p, fq.data(), words * 8, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d.cum.p, cum.data(), (words + 1) * 4, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d.off.p, db->name_off, ((size_t)n + 1) * 4, cudaMemcpyHostToDevice));
	if (nbytes) CU(cudaMemcpy(d.names.p, db->names, nbytes, cudaMemcpyHostToDevice));
	__atomic_fetch_add(&c->stats.h2d_bytes, (uint64_t)(words * 20 + ((size_t)n + 1) * 4 + nbytes), __ATOMIC_RELAXED);
	d.words = words;
	return BSGPU_OK;
}

static DbView db_view(const bsgpu_ctx::DbDev &d) {
	DbView v;
	if (d.words) { v.mask = (const unsigned long long *)d.mask.p; v.fq = (const unsigned long long *)d.fq.p; v.cum = (const uint32_t *)d.cum.p; v.off = (const uint32_t *)d.off.p; v.names = (const uint8_t *)d.names.p; v.words = d.words; }
	return v;
}

// region and dbSNP table of a single-contig call (params->reg_*, params->dbsnp) into the job; the table is uploaded again only
// when it is not the one of the call before
static int job_annotate(bsgpu_ctx *c, const bsgpu_bcf_params *p, BcfJob &j, const char *who) {
	j.reg_start = p->reg_start; j.reg_stop = p->reg_stop;
	j.db = DbView();
	const bsgpu_dbsnp *db = p->dbsnp;
	if (!db || !db->n) return BSGPU_OK;
	const uint64_t key[4] = {(uint64_t)(uintptr_t)db->pos, db->n, (uint64_t)(uintptr_t)db->names, ((uint64_t)db->pos[0] << 32) | db->pos[db->n - 1]};
	if (!c->call_db.words || memcmp(key, c->call_db.key, sizeof(key))) {
		if (db_upload(c, db, c->call_db, who) != BSGPU_OK) return BSGPU_FAIL;
		memcpy(c->call_db.key, key, sizeof(key));
	}
	j.db = db_view(c->call_db);
	return BSGPU_OK;
}

int bsgpu_set_contig_annotation(bsgpu_ctx *c, int tid, uint32_t reg_start, uint32_t reg_stop, const bsgpu_dbsnp *db) {
	if (!c || tid < 0) return fail("bsgpu_set_contig_annotation: bad argument");
	CU(cudaSetDevice(c->device));
	bsgpu_ctx::DbDev *&d = c->contig_ann[tid];
	if (!d) d = new bsgpu_ctx::DbDev();
	d->reg_start = reg_start; d->reg_stop = reg_stop;
	return db_upload(c, db, *d, "bsgpu_set_contig_annotation");
}

// the statistics side channel of a writer job: where the counters live and the GC bins of the job's contig (j.p.rid is set)
static void job_stats(bsgpu_ctx *c, BcfJob &j) {
	if (!c->site_stats_on) return;
	j.stats = c->st_main.p; j.ctg_stats = c->st_nctg ? c->st_ctg.p : nullptr; j.n_ctg = c->st_nctg;
	j.stats_carry = (uint32_t *)((uint8_t *)c->st_main.p + sizeof(bsgpu_site_stats));
	const auto it = c->contig_gc.find((int)j.p.rid);
	if (it != c->contig_gc.end() && it->second->n) { j.gc = (const uint8_t *)it->second->bins.p; j.gc_bins = it->second->n; j.gc_start = it->second->start; }
}

int bsgpu_site_stats_enable(bsgpu_ctx *c, int on, int n_contigs) {
	if (!c || n_contigs < 0) return fail("bsgpu_site_stats_enable: bad argument");
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());
	if (on) {
		CU(c->st_main.reserve(sizeof(bsgpu_site_stats) + 16));          // + the carry word of chunked launches
		CU(cudaMemset(c->st_main.p, 0, sizeof(bsgpu_site_stats) + 16));
		if (n_contigs) {
			CU(c->st_ctg.reserve((size_t)n_contigs * sizeof(bsgpu_ctg_site_stats)));
			CU(cudaMemset(c->st_ctg.p, 0, (size_t)n_contigs * sizeof(bsgpu_ctg_site_stats)));
		}
		c->st_nctg = (uint32_t)n_contigs;
	}
	c->site_stats_on = on != 0;
	return BSGPU_OK;
}

int bsgpu_set_contig_gc(bsgpu_ctx *c, int rid, uint32_t start_pos, const uint8_t *gc, uint32_t nbins) {
	if (!c || rid < 0 || (nbins && !gc)) return fail("bsgpu_set_contig_gc: bad argument");
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());              // the old bins may still be in use
	bsgpu_ctx::GcDev *&d = c->contig_gc[rid];
	if (!d) d = new bsgpu_ctx::GcDev();
	d->n = 0; d->start = start_pos ? start_pos : 1;
	if (nbins) {
		CU(d->bins.reserve(nbins));
		CU(cudaMemcpy(d->bins.p, gc, nbins, cudaMemcpyHostToDevice));
		__atomic_fetch_add(&c->stats.h2d_bytes, (uint64_t)nbins, __ATOMIC_RELAXED);
		d->n = nbins;
	}
	return BSGPU_OK;
}

int bsgpu_site_stats_read(bsgpu_ctx *c, bsgpu_site_stats *out, bsgpu_ctg_site_stats *ctg, int n_contigs, int reset) {
	if (!c || !out || n_contigs < 0 || (n_contigs && !ctg)) return fail("bsgpu_site_stats_read: bad argument");
	if (!c->st_main.p) return fail("bsgpu_site_stats_read: bsgpu_site_stats_enable was never called");
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());
	CU(cudaMemcpy(out, c->st_main.p, sizeof(bsgpu_site_stats), cudaMemcpyDeviceToHost));
	const size_t nc = std::min<size_t>((size_t)n_contigs, c->st_nctg);
	if (n_contigs) memset(ctg, 0, (size_t)n_contigs * sizeof(bsgpu_ctg_site_stats));
	if (nc) CU(cudaMemcpy(ctg, c->st_ctg.p, nc * sizeof(bsgpu_ctg_site_stats), cudaMemcpyDeviceToHost));
	__atomic_fetch_add(&c->stats.d2h_bytes, (uint64_t)(sizeof(bsgpu_site_stats) + nc * sizeof(bsgpu_ctg_site_stats)), __ATOMIC_RELAXED);
	if (reset) {
		CU(cudaMemset(c->st_main.p, 0, sizeof(bsgpu_site_stats)));
		if (c->st_nctg) CU(cudaMemset(c->st_ctg.p, 0, (size_t)c->st_nctg * sizeof(bsgpu_ctg_site_stats)));
	}
	return BSGPU_OK;
}

// one block, everything resident: records into d_out, sizes back through the pinned totals (slot 0); waits for `st`
static int bcf_run(bsgpu_ctx *c, const void *d_vcf, const void *d_ref, uint32_t x, uint32_t sz, const bsgpu_bcf_params *p,
		void *d_out, size_t out_cap, size_t *nbytes, size_t *nrec, cudaStream_t st, const char *who) {
	if (bcf_site_scratch_bytes(sz) > c->wr_site.cap || bcf_cta_scratch_bytes(sz) > c->wr_cta.cap) CU(cudaDeviceSynchronize());
	CU(c->wr_site.reserve(bcf_site_scratch_bytes(sz)));
	CU(c->wr_cta.reserve(bcf_cta_scratch_bytes(sz)));
	BcfJob j;
	j.d_vcf = d_vcf; j.d_ref = d_ref; j.x = x; j.sz = sz; j.d_blocks = nullptr; j.nblocks = 0; j.p = *p; j.dc = c->d_const; j.guard = c->d_counters;
	j.site_scratch = c->wr_site.p;
	if (job_annotate(c, p, j, who) != BSGPU_OK) return BSGPU_FAIL;
	job_stats(c, j);
	CU(launch_bcf_records(j, 0, sz, c->wr_cta.p, d_out, out_cap, c->d_wr_totals, st, &c->launches));
	CU(cudaMemcpyAsync(c->h_wr_totals, c->d_wr_totals, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	if (check_totals(c->h_wr_totals, out_cap, who) != BSGPU_OK) return BSGPU_FAIL;
	*nbytes = (size_t)c->h_wr_totals[0];
	*nrec = (size_t)c->h_wr_totals[1];
	return BSGPU_OK;
}

int bsgpu_bcf_block_dev(bsgpu_ctx *c, const void *d_vcf, const void *d_ref, uint32_t x, uint32_t sz, const bsgpu_bcf_params *p,
		void *d_out, size_t out_cap, size_t *nbytes, size_t *nrec, void *stream) {
	if (!c || !p || !nbytes || !nrec) return fail("bsgpu_bcf_block_dev: null argument");
	*nbytes = *nrec = 0;
	if (!sz) return BSGPU_OK;
	if (!d_vcf || !d_ref || !d_out) return fail("bsgpu_bcf_block_dev: null buffer");
	if ((uintptr_t)d_vcf & 7u) return fail("bsgpu_bcf_block_dev: gt_vcf[] must be 8-byte aligned");
	CU(cudaSetDevice(c->device));
	return bcf_run(c, d_vcf, d_ref, x, sz, p, d_out, out_cap, nbytes, nrec, stream ? (cudaStream_t)stream : c->stream, "bsgpu_bcf_block_dev");
}

int bsgpu_bcf_block(bsgpu_ctx *c, const bsgpu_gt_vcf *vcf, const uint8_t *ref, uint32_t x, uint32_t sz, const bsgpu_bcf_params *p,
		uint8_t *out, size_t out_cap, size_t *nbytes, size_t *nrec) {
	if (!c || !p || !nbytes || !nrec) return fail("bsgpu_bcf_block: null argument");
	*nbytes = *nrec = 0;
	if (!sz) return BSGPU_OK;
	if (!vcf || !ref || !out) return fail("bsgpu_bcf_block: null buffer");
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	const size_t dcap = (size_t)sz * BSGPU_BCF_MAX_RECORD < out_cap ? (size_t)sz * BSGPU_BCF_MAX_RECORD : out_cap;
	CU(c->wr_vcf.reserve((size_t)sz * sizeof(bsgpu_gt_vcf)));
	CU(c->wr_ref.reserve((size_t)sz + 16));
	CU(c->wr_out.reserve(dcap + 16));
	CU(cudaMemcpyAsync(c->wr_vcf.p, vcf, (size_t)sz * sizeof(bsgpu_gt_vcf), cudaMemcpyHostToDevice, c->stream));
	CU(cudaMemcpyAsync(c->wr_ref.p, ref, (size_t)sz + 2, cudaMemcpyHostToDevice, c->stream));
	__atomic_fetch_add(&c->stats.h2d_bytes, (uint64_t)((size_t)sz * (sizeof(bsgpu_gt_vcf) + 1) + 2), __ATOMIC_RELAXED);
	if (bcf_run(c, c->wr_vcf.p, c->wr_ref.p, x, sz, p, c->wr_out.p, dcap, nbytes, nrec, c->stream, "bsgpu_bcf_block") != BSGPU_OK) return BSGPU_FAIL;
	if (*nbytes) CU(cudaMemcpy(out, c->wr_out.p, *nbytes, cudaMemcpyDeviceToHost));
	__atomic_fetch_add(&c->stats.d2h_bytes, (uint64_t)(*nbytes + 24), __ATOMIC_RELAXED);
	return BSGPU_OK;
}

static int block_dev(bsgpu_ctx *c, const void *d_segs, size_t nseg, const void *d_bases, const void *d_ref, uint32_t x, uint32_t sz,
		void *d_out, int mode, void *d_scratch, void *stream);

int bsgpu_call_block_bcf(bsgpu_ctx *c, const bsgpu_seg *segs, size_t nseg, const uint8_t *bases, size_t nbases, const uint8_t *ref,
		uint32_t x, uint32_t sz, const bsgpu_bcf_params *p, uint8_t *out, size_t out_cap, size_t *nbytes, size_t *nrec) {
	if (!c || !p || !nbytes || !nrec) return fail("bsgpu_call_block_bcf: null argument");
	*nbytes = *nrec = 0;
	if (!sz) return BSGPU_OK;
	if (!ref || !out || (nseg && (!segs || !bases))) return fail("bsgpu_call_block_bcf: null buffer");
	if (nbases > 0xffffffffull) return fail("bsgpu_call_block_bcf: more than 4 GiB of bases in one block; split the window");
	for (size_t i = 0; i < nseg; i++) {
		if (segs[i].len > BSGPU_MAX_SEG_LEN) return fail("bsgpu_call_block_bcf: segment %zu longer than %d (use bsgpu_stage_templates)", i, BSGPU_MAX_SEG_LEN);
		if ((size_t)segs[i].off + segs[i].len > nbases) return fail("bsgpu_call_block_bcf: segment %zu points outside bases[]", i);
	}
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaStreamSynchronize(c->copy_stream));
	const size_t dcap = (size_t)sz * BSGPU_BCF_MAX_RECORD < out_cap ? (size_t)sz * BSGPU_BCF_MAX_RECORD : out_cap;
	CU(c->segs.reserve(nseg * sizeof(bsgpu_seg) + 16));
	CU(c->bases.reserve(nbases + 16));
	CU(c->wr_ref.reserve((size_t)sz + 16));
	CU(c->wr_vcf.reserve((size_t)sz * sizeof(bsgpu_gt_vcf)));
	CU(c->wr_out.reserve(dcap + 16));
	if (nseg) {
		CU(cudaMemcpyAsync(c->segs.p, segs, nseg * sizeof(bsgpu_seg), cudaMemcpyHostToDevice, c->stream));
		CU(cudaMemcpyAsync(c->bases.p, bases, nbases, cudaMemcpyHostToDevice, c->stream));
	}
	CU(cudaMemcpyAsync(c->wr_ref.p, ref, (size_t)sz + 2, cudaMemcpyHostToDevice, c->stream));
	__atomic_fetch_add(&c->stats.h2d_bytes, (uint64_t)(nseg * sizeof(bsgpu_seg) + nbases + sz + 2), __ATOMIC_RELAXED);
	if (block_dev(c, c->segs.p, nseg, c->bases.p, c->wr_ref.p, x, sz, c->wr_vcf.p, 1, nullptr, c->stream) != BSGPU_OK) return BSGPU_FAIL;
	if (bcf_run(c, c->wr_vcf.p, c->wr_ref.p, x, sz, p, c->wr_out.p, dcap, nbytes, nrec, c->stream, "bsgpu_call_block_bcf") != BSGPU_OK) return BSGPU_FAIL;
	if (*nbytes) CU(cudaMemcpy(out, c->wr_out.p, *nbytes, cudaMemcpyDeviceToHost));
	__atomic_fetch_add(&c->stats.d2h_bytes, (uint64_t)(*nbytes + 24), __ATOMIC_RELAXED);
	return BSGPU_OK;
}

// pileup[] of one block of n sites -> BCF records.  Chunks of sites go up on one stream, through the model and the
// writer on the context stream and down on the copy stream; the writer runs one chunk behind the model because a
// site's record looks at the calls of the two sites after it.
int bsgpu_call_sites_bcf(bsgpu_ctx *c, const bsgpu_pileup *pileup, const uint8_t *ref, size_t n, uint32_t x, const bsgpu_bcf_params *p,
		uint8_t *out, size_t out_cap, size_t *nbytes, size_t *nrec) {
	if (!c || !p || !nbytes || !nrec) return fail("bsgpu_call_sites_bcf: null argument");
	*nbytes = *nrec = 0;
	if (!n) return BSGPU_OK;
	if (!pileup || !ref || !out) return fail("bsgpu_call_sites_bcf: null buffer");
	if (n > 0xfffffff0ull) return fail("bsgpu_call_sites_bcf: more than 2^32 sites in one block");
	CU(cudaSetDevice(c->device));
	cudaStream_t up = c->slot[0].stream, st = c->stream, down = c->copy_stream;
	CU(cudaStreamSynchronize(up)); CU(cudaStreamSynchronize(st)); CU(cudaStreamSynchronize(down));
	static const size_t chunk = [] { const char *e = getenv("BSGPU_SITES_CHUNK"); const long long v = e ? atoll(e) : 0; return v > 0 ? (size_t)v : (size_t)1 << 18; }();      // 478 M sites/s against 444 M with 1 Mi
	const size_t K = (n + chunk - 1) / chunk;
	const size_t ocap = chunk * BSGPU_BCF_MAX_RECORD;          // one chunk's records at most
	CU(c->slot[0].in.reserve(chunk * sizeof(bsgpu_pileup)));
	CU(c->slot[1].in.reserve(chunk * sizeof(bsgpu_pileup)));
	CU(c->wr_vcf.reserve(2 * chunk * sizeof(bsgpu_gt_vcf)));
	CU(c->wr_out.reserve(2 * ocap + 16));
	CU(c->wr_ref.reserve(n + 16));
	CU(c->wr_site.reserve(bcf_site_scratch_bytes((uint32_t)n)));
	CU(c->wr_cta.reserve(2 * bcf_cta_scratch_bytes((uint32_t)chunk)));
	// per chunk: uploaded, modelled, records built, records copied out.  Whatever way the function is left, the three
	// streams are drained before the events go
	struct Events {
		std::vector<cudaEvent_t> v; cudaStream_t s[3];
		~Events() { for (cudaStream_t q : s) cudaStreamSynchronize(q); for (cudaEvent_t e : v) if (e) cudaEventDestroy(e); }
	} evs{std::vector<cudaEvent_t>(4 * K, nullptr), {up, st, down}};
	std::vector<cudaEvent_t> &ev = evs.v;
	for (auto &e : ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	auto E = [&](size_t k, int what) { return ev[4 * k + what]; };
	BcfJob j;
	j.d_ref = c->wr_ref.p; j.x = x; j.sz = (uint32_t)n; j.d_blocks = nullptr; j.nblocks = 0; j.p = *p; j.dc = c->d_const; j.guard = c->d_counters;
	j.site_scratch = c->wr_site.p;
	if (job_annotate(c, p, j, "bsgpu_call_sites_bcf") != BSGPU_OK) return BSGPU_FAIL;
	job_stats(c, j);
	CU(cudaMemcpyAsync(c->wr_ref.p, ref, n + 2, cudaMemcpyHostToDevice, up));
	size_t at = 0, recs = 0;
	int ret = BSGPU_OK;
	// build the records of chunk k (its model and that of chunk k + 1 are queued), start their way home when sized
	auto queue_records = [&](size_t k) -> int {
		const size_t lo = k * chunk, m = n - lo < chunk ? n - lo : chunk;
		if (k >= 2) CU(cudaStreamWaitEvent(st, E(k - 2, 3), 0));      // the output buffer of chunk k - 2 has left
		// a.vcf + i must address the ring slot of chunk k for the sites of chunk k (no other site's record is read)
		j.d_vcf = (const uint8_t *)c->wr_vcf.p + (k & 1) * chunk * sizeof(bsgpu_gt_vcf) - lo * sizeof(bsgpu_gt_vcf);
		j.stats_carry_flip = (uint32_t)(k & 1);
		CU(launch_bcf_records(j, (uint32_t)lo, (uint32_t)m, (uint8_t *)c->wr_cta.p + (k & 1) * bcf_cta_scratch_bytes((uint32_t)chunk),
				(uint8_t *)c->wr_out.p + (k & 1) * ocap, ocap, c->d_wr_totals + 3 * (k & 7), st, &c->launches));
		CU(cudaMemcpyAsync(c->h_wr_totals + 3 * (k & 7), c->d_wr_totals + 3 * (k & 7), 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
		CU(cudaEventRecord(E(k, 2), st));
		return BSGPU_OK;
	};
	auto collect = [&](size_t k) -> int {
		CU(cudaEventSynchronize(E(k, 2)));
		const unsigned long long *t = c->h_wr_totals + 3 * (k & 7);
		if (t[2]) return fail("bsgpu_call_sites_bcf: %llu record(s) longer than %d bytes", t[2], BSGPU_BCF_MAX_RECORD);
		if (at + t[0] > out_cap) return fail("bsgpu_call_sites_bcf: output buffer too small (%zu bytes given)", out_cap);
		if (t[0]) CU(cudaMemcpyAsync(out + at, (uint8_t *)c->wr_out.p + (k & 1) * ocap, t[0], cudaMemcpyDeviceToHost, down));
		CU(cudaEventRecord(E(k, 3), down));
		at += t[0]; recs += t[1];
		__atomic_fetch_add(&c->stats.d2h_bytes, (uint64_t)(t[0] + 24), __ATOMIC_RELAXED);
		return BSGPU_OK;
	};
	for (size_t k = 0; k < K && ret == BSGPU_OK; k++) {
		const size_t lo = k * chunk, m = n - lo < chunk ? n - lo : chunk;
		if (k >= 2) CU(cudaStreamWaitEvent(up, E(k - 2, 1), 0));      // the model has consumed the input buffer
		CU(cudaMemcpyAsync(c->slot[k & 1].in.p, pileup + lo, m * sizeof(bsgpu_pileup), cudaMemcpyHostToDevice, up));
		CU(cudaEventRecord(E(k, 0), up));
		__atomic_fetch_add(&c->stats.h2d_bytes, (uint64_t)(m * (sizeof(bsgpu_pileup) + 1)), __ATOMIC_RELAXED);
		CU(cudaStreamWaitEvent(st, E(k, 0), 0));
		CU(launch_call_sites(c->slot[k & 1].in.p, (const uint8_t *)c->wr_ref.p + lo, m, (uint8_t *)c->wr_vcf.p + (k & 1) * chunk * sizeof(bsgpu_gt_vcf),
				nullptr, true, c->d_const, c->d_counters, st, &c->launches, 0, overlap_expected(c, false)));
		CU(cudaEventRecord(E(k, 1), st));
		j.d_vcf = (const uint8_t *)c->wr_vcf.p + (k & 1) * chunk * sizeof(bsgpu_gt_vcf) - lo * sizeof(bsgpu_gt_vcf);
		// the records of chunk k - 1 look at the calls of the first two sites of this chunk (the rest of this chunk's calls are
		// written when its own records are measured)
		CU(launch_bcf_calls(j, (uint32_t)lo, (uint32_t)std::min<size_t>(m, 2), st, &c->launches));
		if (k >= 1) ret = queue_records(k - 1);
		if (ret == BSGPU_OK && k >= 2) ret = collect(k - 2);
	}
	if (ret == BSGPU_OK) ret = queue_records(K - 1);
	if (ret == BSGPU_OK && K >= 2) ret = collect(K - 2);
	if (ret == BSGPU_OK) ret = collect(K - 1);
	CU(cudaStreamSynchronize(up)); CU(cudaStreamSynchronize(st)); CU(cudaStreamSynchronize(down));
	if (ret != BSGPU_OK) return ret;
	__atomic_fetch_add(&c->stats.sites, (uint64_t)(n), __ATOMIC_RELAXED);
	*nbytes = at; *nrec = recs;
	return BSGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// --report-file side channels
// ------------------------------------------------------------------------------------------------
int bsgpu_profile_enable(bsgpu_ctx *c, int on) {
	if (!c) return fail("bsgpu_profile_enable: null context");
	CU(cudaSetDevice(c->device));
	if (on && !c->d_prof) {
		CU(cudaMalloc(&c->d_prof, sizeof(ProfDev)));
		CU(cudaMemset(c->d_prof, 0, sizeof(ProfDev)));
		c->prof_parity = 0;
	}
	c->profile_on = on != 0;
	return BSGPU_OK;
}

int bsgpu_profile_read(bsgpu_ctx *c, bsgpu_profile *out, int reset) {
	if (!c || !out) return fail("bsgpu_profile_read: null argument");
	if (!c->d_prof) return fail("bsgpu_profile_read: the profile was never enabled");
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	std::vector<ProfDev> hbuf(1);
	ProfDev &h = hbuf[0];
	CU(cudaMemcpy(&h, c->d_prof, sizeof(h), cudaMemcpyDeviceToHost));
	__atomic_fetch_add(&c->stats.d2h_bytes, (uint64_t)(sizeof(h)), __ATOMIC_RELAXED);
	if (h.too_long) return fail("bsgpu profile: %llu template(s) reach beyond original read position %d", h.too_long, BSGPU_PROFILE_MAX - 2);
	memset(out, 0, sizeof(*out));
	out->used = h.used[c->prof_parity];
	for (uint32_t i = 0; i < out->used && i < BSGPU_PROFILE_MAX; i++) for (int k = 0; k < 4; k++) out->conv_cts[i][k] = h.conv[i][k];
	for (int k = 0; k < 5; k++) out->base_filter[k] = h.base_filter[k];
	for (int k = 0; k < 15; k++) { out->filter_cts[k] = c->reader_tally[k]; out->filter_bases[k] = c->reader_tally[15 + k]; }
	out->filter_cts[0] += h.reads;
	out->filter_bases[0] += h.read_bases;
	if (reset) {
		CU(cudaMemset(c->d_prof, 0, sizeof(ProfDev)));
		c->prof_parity = 0;
		memset(c->reader_tally, 0, sizeof(c->reader_tally));
	}
	return BSGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// reader side
// ------------------------------------------------------------------------------------------------
void bsgpu_default_reader_params(bsgpu_reader_params *p) {
	memset(p, 0, sizeof(*p));
	p->max_template_len = 1000;     // include/bs_call.h:20
	p->mapq_thresh = 20;            // include/bs_call.h:14
}

// frame + upload + decode, in chunks: the stream goes up in byte pieces on the upload stream while the host frames it;
// chunk k = the records that end inside pieces 0..k, decoded on the decode stream as soon as piece k has landed, its
// descriptors copied back to the pinned array h_rec.  On return everything is queued; chunk_end[k] = one past the last
// record of chunk k and R.rd_done[k] fires when its descriptors are on the host.  Descriptors, packed reads and events
// stay resident.  *nb / *nm = sizes of the decoded arrays.
static int decode_queue(bsgpu_ctx *c, bsgpu_ctx::ReaderSet &R, const uint8_t *bam, size_t nbytes, const bsgpu_reader_params *rp, unsigned want_chunks,
		size_t *nrec, uint64_t *nb, uint64_t *nm, std::vector<size_t> &chunk_end, size_t *framed = nullptr, bool pipelined = false) {
	std::vector<uint32_t> &read_off = R.read_off, &mm_off = R.mm_off;
	if (!R.frame_scratch) R.frame_scratch = frame_scratch_new();
	CU(cudaSetDevice(c->device));
	if (!R.up) {
		if (&R == &c->rs[0]) R.up = c->slot[0].stream;
		else CU(cudaStreamCreateWithFlags(&R.up, cudaStreamNonBlocking));
	}
	// debugging switches: BSGPU_ONE_STREAM=1 puts upload, decode and the windows on one stream (no kernel of the reader stage
	// overlaps a kernel of the window stage); BSGPU_NO_NAME_JOIN=1 leaves the QNAME join to the host
	static const bool one_stream = getenv("BSGPU_ONE_STREAM") != nullptr, no_join = getenv("BSGPU_NO_NAME_JOIN") != nullptr;
	// the decode / name-join kernels run on a stream of their own, next to the window kernels of the pieces (and, in a session,
	// of the batch) before: the context then launches its likelihood kernel in the overlap-safe form (launch_call_sites).
	// BSGPU_DECODE_STREAM=0 queues them on the context stream instead.
	static const bool dec_own = [] { const char *e = getenv("BSGPU_DECODE_STREAM"); return !e || atoi(e) != 0; }();
	cudaStream_t up = one_stream ? c->stream : R.up, dec = one_stream || !dec_own ? c->stream : c->slot[1].stream;
	if (dec != c->stream) c->overlap_kernels = true;
	if (!pipelined) {
		// nothing of an earlier call is in flight any more.  (A pipelined run of a session starts while the run before it is
		// still at its window stage: its reader set was last used two runs ago, and the session has waited for that one.)
		CU(cudaStreamSynchronize(c->stream));
		CU(cudaStreamSynchronize(c->copy_stream));
		CU(cudaStreamSynchronize(dec));
	}
	CU(cudaStreamSynchronize(up));
	size_t chunk_min = 64u << 20;                       // smaller streams go up and are decoded in one piece
	if (const char *e = getenv("BSGPU_READER_CHUNK_MIN_BYTES")) chunk_min = (size_t)atoll(e);
	const unsigned K = nbytes >= chunk_min ? std::max(1u, std::min(want_chunks, 8u)) : 1;
	while (R.rd_up.size() < K) {
		cudaEvent_t e1, e2;
		CU(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
		R.rd_up.push_back(e1); R.rd_done.push_back(e2);
	}
	// The stream itself starts its way up while the host frames it -- the first half only: the copy engine serves
	// transfers in the order they were submitted, and the offset tables (known after framing) must not queue behind the
	// whole stream, or the first chunk could not be decoded before the last byte has arrived.
	CU(R.rd_bam.reserve(nbytes + 16));
	std::vector<size_t> piece_end(K);
	for (unsigned k = 0; k < K; k++) piece_end[k] = nbytes * (k + 1) / K;
	auto upload_piece = [&](unsigned k) -> int {
		const size_t lo = nbytes * k / K, hi = piece_end[k];
		if (hi > lo) CU(cudaMemcpyAsync((uint8_t *)R.rd_bam.p + lo, bam + lo, hi - lo, cudaMemcpyHostToDevice, up));
		CU(cudaEventRecord(R.rd_up[k], up));
		return BSGPU_OK;
	};
	const unsigned k_early = (K + 1) / 2;
	for (unsigned k = 0; k < k_early; k++) if (upload_piece(k) != BSGPU_OK) return BSGPU_FAIL;
	const int fr = frame_records(bam, nbytes, R.rec_off, read_off, mm_off, nb, nm, R.frame_scratch, framed);
	if (fr) cudaStreamSynchronize(up);
	if (fr == -1) return fail("bsgpu reader: truncated or malformed BAM record stream");
	if (fr == -2) return fail("bsgpu reader: more than 4 Gi bases in one stream; split the input");
	const size_t n = R.rec_off.size();
	*nrec = n;
	chunk_end.clear();
	if (!n) { CU(cudaStreamSynchronize(up)); return BSGPU_OK; }
	CU(R.rd_recoff.reserve(n * 8));
	CU(R.rd_readoff.reserve(n * 4));
	CU(R.rd_mmoff.reserve(n * 4));
	CU(R.rd_rec.reserve(n * sizeof(bsgpu_record)));
	CU(R.rd_bases.reserve(*nb + 16));
	CU(R.rd_misms.reserve((*nm + 1) * sizeof(bsgpu_misms)));
	CU(R.h_rec.reserve(n * sizeof(bsgpu_record)));
	if (c->zero_decoded) {                          // k_decode_records skips dropped records: what the caller copies out is defined everywhere
		CU(cudaMemsetAsync(R.rd_bases.p, 0, *nb + 16, dec));
		CU(cudaMemsetAsync(R.rd_misms.p, 0, (*nm + 1) * sizeof(bsgpu_misms), dec));
	}
	// certain block starts: the device scans the keys the decode kernel writes (launch_certain_starts: three phases over all
	// SMs) and one bit per record comes home; BSGPU_HOST_SCAN=1 brings the keys home instead (16 bytes per record) and host
	// threads scan them -- the default before the device scan was spread over the SMs --, BSGPU_CHECK_SCAN=1 does both and compares
	const bool dev_scan = getenv("BSGPU_HOST_SCAN") == nullptr || getenv("BSGPU_CHECK_SCAN") != nullptr;
	const bool host_scan = getenv("BSGPU_HOST_SCAN") != nullptr || getenv("BSGPU_CHECK_SCAN") != nullptr;
	CU(R.rd_key.reserve(n * 16));
	if (host_scan) CU(R.h_key.reserve(n * 16));
	// QNAME join: table of kept paired records by name hash, name ids per record (+ one word: the overflow flag)
	const size_t name_slots = name_table_slots(n);
	CU(R.rd_names.reserve(name_table_bytes(n)));
	CU(R.rd_nameid.reserve((n + 1) * 4));
	CU(R.h_nameid.reserve((n + 1 + K) * 4));      // + the overflow flag as it stood after every chunk
	memset((uint32_t *)R.h_nameid.p + n, 0, (1 + K) * 4);
	CU(cudaMemsetAsync(R.rd_names.p, 0, name_table_bytes(n), dec));
	CU(cudaMemsetAsync((uint32_t *)R.rd_nameid.p + n, 0, 4, dec));
	CU(R.rd_mask.reserve((n / 32 + 2 * K + 8) * 4 + 16));
	if (dev_scan) CU(R.rd_scan.reserve(certain_scratch_bytes(n)));
	CU(R.h_mask.reserve((n / 32 + 2 * K + 8) * 4));
	R.mask_off.assign(K + 1, 0);
	uint32_t *d_carry = (uint32_t *)R.rd_mask.p;          // first four words of the buffer: {last contig, running end}
	CU(cudaMemsetAsync(d_carry, 0xff, 4, dec));
	CU(cudaMemsetAsync(d_carry + 1, 0, 12, dec));
	size_t mask_words = 4;
	// the offset tables go up from pinned staging (pageable vectors would serialise the copies behind the big ones)
	CU(R.h_off.reserve(n * 16));
	uint8_t *ho = (uint8_t *)R.h_off.p;
	{
		// 16 bytes per record into page-locked memory: a few threads, each a slice of the three tables
		const unsigned T = n >= (1u << 18) ? std::max(1u, std::min(8u, std::thread::hardware_concurrency())) : 1;
		auto part = [&](unsigned t) {
			const size_t lo = n * t / T, hi = n * (t + 1) / T;
			memcpy(ho + lo * 8, R.rec_off.data() + lo, (hi - lo) * 8);
			memcpy(ho + n * 8 + lo * 4, read_off.data() + lo, (hi - lo) * 4);
			memcpy(ho + n * 12 + lo * 4, mm_off.data() + lo, (hi - lo) * 4);
		};
		if (T == 1) part(0);
		else {
			std::vector<std::thread> thr;
			for (unsigned t = 0; t < T; t++) thr.emplace_back(part, t);
			for (auto &t : thr) t.join();
		}
	}
	CU(cudaMemcpyAsync(R.rd_recoff.p, ho, n * 8, cudaMemcpyHostToDevice, dec));
	CU(cudaMemcpyAsync(R.rd_readoff.p, ho + n * 8, n * 4, cudaMemcpyHostToDevice, dec));
	CU(cudaMemcpyAsync(R.rd_mmoff.p, ho + n * 12, n * 4, cudaMemcpyHostToDevice, dec));
	for (unsigned k = k_early; k < K; k++) if (upload_piece(k) != BSGPU_OK) return BSGPU_FAIL;
	__atomic_fetch_add(&c->stats.h2d_bytes, (uint64_t)(nbytes + n * 16), __ATOMIC_RELAXED);
	size_t r0 = 0;
	for (unsigned k = 0; k < K; k++) {
		// records that end inside pieces 0..k (the last chunk takes the rest)
		size_t r1 = n;
		if (k + 1 < K) {
			const auto it = std::upper_bound(R.rec_off.begin() + r0, R.rec_off.end(), (uint64_t)piece_end[k]);
			r1 = (size_t)(it - R.rec_off.begin());
			if (r1 > r0) r1--;                          // the record that starts before the boundary may end after it
			while (r1 > r0 && R.rec_off[r1 - 1] + 4 + (uint64_t)(bam[R.rec_off[r1 - 1]] | bam[R.rec_off[r1 - 1] + 1] << 8 | bam[R.rec_off[r1 - 1] + 2] << 16 | (uint64_t)bam[R.rec_off[r1 - 1] + 3] << 24) > piece_end[k]) r1--;
		}
		CU(cudaStreamWaitEvent(dec, R.rd_up[k], 0));
		if (r1 > r0) {
			CU(launch_decode_records(R.rd_bam.p, (const uint64_t *)R.rd_recoff.p + r0, (const uint32_t *)R.rd_readoff.p + r0,
					(const uint32_t *)R.rd_mmoff.p + r0, r1 - r0, rp->mapq_thresh, rp->max_template_len, rp->keep_unmatched,
					rp->ignore_duplicates, (bsgpu_record *)R.rd_rec.p + r0, R.rd_bases.p, R.rd_misms.p, dec, &c->launches, (uint8_t *)R.rd_key.p + r0 * 16,
					no_join ? nullptr : R.rd_names.p, name_slots, (uint32_t)r0, (uint32_t *)R.rd_nameid.p + n));
			if (no_join) CU(cudaMemsetAsync((uint32_t *)R.rd_nameid.p + n, 0xff, 4, dec));      // "overflow": the host computes the ids
			else CU(launch_name_ids(R.rd_bam.p, R.rd_recoff.p, R.rd_rec.p, (uint32_t)r0, (uint32_t)r1, R.rd_names.p, name_slots, R.rd_nameid.p, dec, &c->launches));
			CU(cudaMemcpyAsync((uint32_t *)R.h_nameid.p + r0, (const uint32_t *)R.rd_nameid.p + r0, (r1 - r0) * 4, cudaMemcpyDeviceToHost, dec));
			CU(cudaMemcpyAsync((uint32_t *)R.h_nameid.p + n + 1 + k, (const uint32_t *)R.rd_nameid.p + n, 4, cudaMemcpyDeviceToHost, dec));
			if (host_scan) CU(cudaMemcpyAsync((uint8_t *)R.h_key.p + r0 * 16, (const uint8_t *)R.rd_key.p + r0 * 16, (r1 - r0) * 16, cudaMemcpyDeviceToHost, dec));
			if (dev_scan) {
				const size_t words = (r1 - r0 + 31) / 32;
				CU(launch_certain_starts((const uint8_t *)R.rd_key.p + r0 * 16, (uint32_t)(r1 - r0), d_carry, (uint32_t *)R.rd_mask.p + mask_words, R.rd_scan.p, dec, &c->launches));
				CU(cudaMemcpyAsync((uint32_t *)R.h_mask.p + mask_words, (const uint32_t *)R.rd_mask.p + mask_words, words * 4, cudaMemcpyDeviceToHost, dec));
				R.mask_off[k] = mask_words;
				mask_words += words;
			}
			CU(cudaMemcpyAsync((bsgpu_record *)R.h_rec.p + r0, (const bsgpu_record *)R.rd_rec.p + r0, (r1 - r0) * sizeof(bsgpu_record),
					cudaMemcpyDeviceToHost, dec));
		}
		CU(cudaEventRecord(R.rd_done[k], dec));
		chunk_end.push_back(r1);
		r0 = r1;
	}
	__atomic_fetch_add(&c->stats.d2h_bytes, (uint64_t)(n * (sizeof(bsgpu_record) + 4 + (host_scan ? 16 : 0)) + mask_words * 4), __ATOMIC_RELAXED);
	return BSGPU_OK;
}

// the whole stream decoded and its descriptors on the host (h_rec)
static int decode_resident(bsgpu_ctx *c, const uint8_t *bam, size_t nbytes, const bsgpu_reader_params *rp, size_t *nrec, uint64_t *nb, uint64_t *nm) {
	std::vector<size_t> chunk_end;
	bsgpu_ctx::ReaderSet &R = c->rs[0];
	if (decode_queue(c, R, bam, nbytes, rp, 1, nrec, nb, nm, chunk_end) != BSGPU_OK) return BSGPU_FAIL;
	CU(cudaStreamSynchronize(c->slot[0].stream));
	CU(cudaStreamSynchronize(c->slot[1].stream));
	CU(cudaStreamSynchronize(c->stream));          // the decode kernels run on the context stream
	return BSGPU_OK;
}

int bsgpu_decode_records(bsgpu_ctx *c, const uint8_t *bam, size_t nbytes, const bsgpu_reader_params *rp,
		bsgpu_record *rec_out, size_t rec_cap, size_t *nrec, uint8_t *bases_out, size_t bases_cap, size_t *nbases,
		bsgpu_misms *misms_out, size_t misms_cap, size_t *nmisms) {
	if (!c || !rp || !nrec || (nbytes && !bam)) return fail("bsgpu_decode_records: null argument");
	size_t n = 0;
	uint64_t nb = 0, nm = 0;
	c->zero_decoded = bases_out != nullptr || misms_out != nullptr;
	const int drc = decode_resident(c, bam, nbytes, rp, &n, &nb, &nm);
	c->zero_decoded = false;
	bsgpu_ctx::ReaderSet &R = c->rs[0];
	if (drc != BSGPU_OK) return BSGPU_FAIL;
	if (n > rec_cap || (bases_out && nb > bases_cap) || (misms_out && nm > misms_cap)) return fail("bsgpu_decode_records: need room for %zu records, %llu bases, %llu events", n, (unsigned long long)nb, (unsigned long long)nm);
	if (n) {
		if (rec_out) memcpy(rec_out, R.h_rec.p, n * sizeof(bsgpu_record));
		// events of dropped records are never written: clear what the caller will see
		if (bases_out && nb) CU(cudaMemcpyAsync(bases_out, R.rd_bases.p, nb, cudaMemcpyDeviceToHost, c->stream));
		if (misms_out && nm) CU(cudaMemcpyAsync(misms_out, R.rd_misms.p, nm * sizeof(bsgpu_misms), cudaMemcpyDeviceToHost, c->stream));
		CU(cudaStreamSynchronize(c->stream));
		__atomic_fetch_add(&c->stats.d2h_bytes, (uint64_t)((bases_out ? nb : 0) + (misms_out ? nm * sizeof(bsgpu_misms) : 0)), __ATOMIC_RELAXED);
	}
	*nrec = n;
	if (nbases) *nbases = nb;
	if (nmisms) *nmisms = nm;
	return BSGPU_OK;
}

int bsgpu_build_blocks(const uint8_t *bam, size_t nbytes, const bsgpu_record *rec, size_t nrec, const bsgpu_reader_params *rp,
		bsgpu_block *blocks, size_t block_cap, size_t *nblocks, bsgpu_template *tmpl, size_t tmpl_cap, size_t *ntmpl) {
	return bsgpu_build_blocks_tally(bam, nbytes, rec, nrec, rp, blocks, block_cap, nblocks, tmpl, tmpl_cap, ntmpl, nullptr, nullptr);
}

int bsgpu_build_blocks_tally(const uint8_t *bam, size_t nbytes, const bsgpu_record *rec, size_t nrec, const bsgpu_reader_params *rp,
		bsgpu_block *blocks, size_t block_cap, size_t *nblocks, bsgpu_template *tmpl, size_t tmpl_cap, size_t *ntmpl,
		uint64_t *filter_cts, uint64_t *filter_bases) {
	if ((filter_cts == nullptr) != (filter_bases == nullptr)) return fail("bsgpu_build_blocks_tally: give both tally arrays or neither");
	uint64_t tally[30] = {0};
	if (!rp || !nblocks || !ntmpl || (nrec && (!bam || !rec))) return fail("bsgpu_build_blocks: null argument");
	std::vector<uint64_t> rec_off;
	std::vector<uint32_t> ro, mo;
	uint64_t nb, nm;
	if (frame_records(bam, nbytes, rec_off, ro, mo, &nb, &nm, nullptr) || rec_off.size() != nrec) return fail("bsgpu_build_blocks: the record stream does not frame into %zu records", nrec);
	std::vector<bsgpu_block> b;
	bsgpu_template *t = tmpl_cap >= nrec ? tmpl : (bsgpu_template *)malloc((nrec + 1) * sizeof(bsgpu_template));      // build in place when there is room
	if (!t) return fail("bsgpu_build_blocks: out of memory");
	size_t nt = 0;
	int rc = build_blocks_host(bam, rec_off.data(), rec, nrec, rp->keep_unmatched, rp->keep_duplicates, b, t, &nt, filter_cts ? tally : nullptr);
	if (rc == -6) rc = 0;       // a template whose absent mate lies before its block: read_input itself succeeds (the callers of the blocks refuse)
	if (!rc && filter_cts) for (int k = 0; k < 15; k++) { filter_cts[k] += tally[k]; filter_bases[k] += tally[15 + k]; }
	int ret = BSGPU_OK;
	if (rc == -4) ret = fail("bsgpu_build_blocks: duplicate read name among waiting mates");
	else if (rc == -5) ret = fail("bsgpu_build_blocks: the two mates of a template disagree about their positions");
	else if (rc) ret = fail("bsgpu_build_blocks: failed (%d)", rc);
	else if (b.size() > block_cap || nt > tmpl_cap) ret = fail("bsgpu_build_blocks: need room for %zu blocks, %zu templates", b.size(), nt);
	else {
		if (!b.empty()) memcpy(blocks, b.data(), b.size() * sizeof(bsgpu_block));
		if (t != tmpl && nt) memcpy(tmpl, t, nt * sizeof(bsgpu_template));
		*nblocks = b.size();
		*ntmpl = nt;
	}
	if (t != tmpl) free(t);
	return ret;
}

static int block_run(bsgpu_ctx *c, const void *d_segs, size_t nseg, const void *d_bases, const void *d_ref,
		uint32_t x, uint32_t sz, void *out, int mode, bool defer);

// What a streaming session (bsgpu_bam_*) asks of a run over the bytes it has staged so far
struct BamRunOpts {
	bool partial = false;                    // stop at the last CERTAIN block start of the buffer; a record cut off by its end is fine
	// Pipelined runs of a session: run k + 1 may do its reader stage (reader set (k + 1) & 1) while run k is still queueing
	// windows; it takes its turn at the window stage when run k has finished.  `ticket` numbers the runs of the context from
	// the value of bsgpu_ctx::win_done at the first one; on_scanned is called (from the run's scanner thread) as soon as the
	// bytes the run accounts for are known -- the staged bytes are not read by the run after that.
	bool pipelined = false;
	int reader_set = 0;
	uint64_t ticket = 0;
	void (*on_scanned)(void *user, size_t consumed, size_t records) = nullptr;
	void *scan_user = nullptr;
	size_t consumed = 0;                     // out: bytes of the buffer whose records were turned into results
	size_t records = 0;                      // out: records in those bytes
	std::vector<bsgpu_block> *blocks_vec = nullptr;      // blocks go here (grown as needed) instead of into blocks[]
	// results go into a buffer the session owns; when it is too small `grow(user, keep, need)` returns a larger one that
	// holds the first `keep` bytes of the old one (no copy into the old one is in flight when it is called)
	uint8_t *(*grow)(void *user, size_t keep, size_t need, size_t *new_cap) = nullptr;
	void *user = nullptr;
	// a contig whose codes are missing when its first window is due (bsgpu_bam_on_contig): asks the host, then looks again
	int (*need_contig)(void *user, int tid) = nullptr;
	void *contig_user = nullptr;
};

// templates [tm, tm + nt) of ONE contig (their reads and events are resident from the decode) -> gt_vcf[] of window [x, y]
// where the records of bsgpu_call_bam_bcf go: windows are queued on the context stream, their sizes come home through
// pinned memory, and the copy-out of window w is issued once its size is known -- by then window w + 1 is queued
struct BcfSink {
	bsgpu_bcf_params p;
	const int32_t *vcf_rid;
	uint8_t *out;
	size_t out_cap, at = 0, recs = 0;
	uint64_t queued = 0, collected = 0;
	size_t ring_cap[3] = {0, 0, 0};
	BamRunOpts *opts = nullptr;              // session runs: the output buffer can grow
};

static int sink_collect(bsgpu_ctx *c, BcfSink *k) {
	const int slot = (int)(k->collected % 3);
	const double tc0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
	CU(cudaEventSynchronize(c->wr_built[slot]));
	c->tm_collect += std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() - tc0;
	const unsigned long long *t = c->h_wr_totals + 3 * (k->collected & 7);
	if (t[2]) return fail("bsgpu_call_bam_bcf: %llu record(s) longer than %d bytes", t[2], BSGPU_BCF_MAX_RECORD);
	if (t[0] > k->ring_cap[slot]) return fail("bsgpu_call_bam_bcf: the records of one window (%llu bytes) exceed the device staging buffer", t[0]);
	if (k->at + t[0] > k->out_cap) {
		if (!k->opts || !k->opts->grow) return fail("bsgpu_call_bam_bcf: output buffer too small (%zu bytes given)", k->out_cap);
		CU(cudaStreamSynchronize(c->copy_stream));      // earlier windows' records have landed in the old buffer
		size_t ncap = 0;
		uint8_t *nbuf = k->opts->grow(k->opts->user, k->at, k->at + t[0], &ncap);
		if (!nbuf) return fail("bsgpu_bam: cannot grow the result buffer to %llu bytes", (unsigned long long)(k->at + t[0]));
		k->out = nbuf; k->out_cap = ncap;
	}
	if (t[0]) CU(cudaMemcpyAsync(k->out + k->at, c->wr_ring[slot].p, t[0], cudaMemcpyDeviceToHost, c->copy_stream));
	CU(cudaEventRecord(c->wr_copied[slot], c->copy_stream));
	k->at += t[0]; k->recs += t[1];
	__atomic_fetch_add(&c->stats.d2h_bytes, (uint64_t)(t[0] + 24), __ATOMIC_RELAXED);
	k->collected++;
	return BSGPU_OK;
}

struct TmSpan { const bsgpu_template *p; size_t n; };      // the templates of a window: runs in pinned host memory, in order

// wb[0 .. nwb): the blocks inside the window (BCF sink only)
static int call_window(bsgpu_ctx *c, bsgpu_ctx::ReaderSet &R, const TmSpan *span, size_t nspan, size_t nt, uint32_t tid, uint32_t ctg_len, const uint8_t *codes,
		uint32_t x, uint32_t y, bsgpu_gt_vcf *out, uint32_t maxcap_hint = 0, BcfSink *sink = nullptr, const bsgpu_block *wb = nullptr, size_t nwb = 0) {
	const uint32_t sz = y - x + 1;
	auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const double tp0 = now();
	// Per-mate output slots: read length + reference span bounds a mate in reference coordinates.  The block builder
	// reports the largest such bound of its piece; when that is moderate every mate simply gets a slot of that size and no
	// offset table is built or uploaded.  Otherwise (a few very long spans) exact offsets are laid out here.
	std::vector<uint32_t> &off = c->off_tmp;
	uint64_t tot = 0;
	uint32_t maxcap = 1, slot = 0;
	static const bool exact_slots = getenv("BSGPU_EXACT_SLOTS") != nullptr;      // tests: force the offset-table path
	if (!exact_slots && maxcap_hint && maxcap_hint <= 1024 && 2ull * nt * ((maxcap_hint + 15u) & ~15u) <= 0xffffffffull) {
		slot = (maxcap_hint + 15u) & ~15u;
		maxcap = maxcap_hint;
		tot = 2ull * nt * slot;
	} else {
		off.resize(2 * nt + 1);
		size_t i = 0;
		for (size_t sp = 0; sp < nspan; sp++) for (size_t j = 0; j < span[sp].n; j++, i++) for (int k = 0; k < 2; k++) {
			off[2 * i + k] = (uint32_t)tot;
			const bsgpu_template &t = span[sp].p[j];
			if (!t.present[k]) continue;
			const uint64_t cap = (uint64_t)t.read_len[k] + t.reference_span[k];
			tot += cap;
			if (cap > maxcap) maxcap = (uint32_t)(cap > 0xffffffu ? 0xffffffu : cap);
		}
		off[2 * nt] = (uint32_t)tot;
	}
	if (tot > 0xffffffffull) return fail("bsgpu_call_bam: more than 4 GiB of bases in one window of contig %u; split the input", tid);
	const uint32_t spm = (maxcap + BSGPU_MAX_SEG_LEN - 1) / BSGPU_MAX_SEG_LEN;
	const size_t nseg = nt * 2 * (size_t)spm;
	// reference window [x, y + 2] (one code past the window for the conversion profile, src/meth_profile.c:70, two for the
	// writer): straight from the contig's codes, N from the contig's last position on (src/get_sequence.c:36-41)
	const uint64_t navail = (uint64_t)x < ctg_len ? (uint64_t)ctg_len - x : 0;          // positions x .. ctg_len - 1
	const size_t ncopy = (size_t)std::min<uint64_t>(navail, (uint64_t)sz + 2);
	const double tp1 = now();
	c->tm_prep += tp1 - tp0;
	struct Acc { double &d; double t0; std::function<double()> f; ~Acc() { d += f() - t0; } } acc_{c->tm_queue, tp1, now};
	// Everything below is queued behind the previous window on the context stream; nothing waits on the host.  (Growing a
	// device buffer frees the old one, which synchronises the device; `off` / `refw` are pageable, so their copies are
	// staged before cudaMemcpyAsync returns and the vectors can be refilled for the next window.)
	CU(c->tmpl.reserve(nt * sizeof(bsgpu_template)));
	CU(c->obases.reserve(tot + 16));
	if (!slot) CU(c->ooff.reserve((2 * nt + 1) * sizeof(uint32_t)));
	CU(c->segs.reserve(nseg * sizeof(bsgpu_seg) + 16));
	CU(c->ref.reserve((size_t)sz + 16));
	{
		size_t at = 0;
		for (size_t sp = 0; sp < nspan; sp++) {
			if (span[sp].n) CU(cudaMemcpyAsync((bsgpu_template *)c->tmpl.p + at, span[sp].p, span[sp].n * sizeof(bsgpu_template), cudaMemcpyHostToDevice, c->stream));
			at += span[sp].n;
		}
	}
	if (!slot) CU(cudaMemcpyAsync(c->ooff.p, off.data(), (2 * nt + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
	if (ncopy) CU(cudaMemcpyAsync(c->ref.p, codes + x - 1, ncopy, cudaMemcpyHostToDevice, c->stream));
	if (ncopy < (size_t)sz + 2) CU(cudaMemsetAsync((uint8_t *)c->ref.p + ncopy, 0, (size_t)sz + 2 - ncopy, c->stream));
	__atomic_fetch_add(&c->stats.h2d_bytes, (uint64_t)(nt * sizeof(bsgpu_template) + (slot ? 0 : (2 * nt + 1) * 4) + ncopy), __ATOMIC_RELAXED);
	cudaError_t perr;
	const ProfArgs *pa = profile_for(c, nt, sz, &perr);
	CU(perr);
	CU(launch_normalise(c->tmpl.p, nt, R.rd_bases.p, R.rd_misms.p, slot ? nullptr : c->ooff.p, c->obases.p, c->segs.p, spm, x, y,
			c->params.left_trim, c->params.right_trim, c->d_counters, pa, c->prof_parity, c->stream, &c->launches, slot));
	if (pa) c->prof_parity ^= 1;
	if (!sink) return block_run(c, c->segs.p, nseg, c->obases.p, c->ref.p, x, sz, out, 1, true);
	// ---- records instead of gt_vcf[]: the window's gt_vcf[] stays on the device, the writer kernels turn it into BCF
	const int rslot = (int)(sink->queued % 3);
	for (int i = 0; i < 3; i++) if (!c->wr_built[i]) {
		CU(cudaEventCreateWithFlags(&c->wr_built[i], cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&c->wr_copied[i], cudaEventDisableTiming));
	}
	const size_t rcap = (size_t)sz * 256 + 4096;
	const bool grow = (size_t)sz * sizeof(bsgpu_gt_vcf) > c->wr_vcf.cap || rcap + 16 > c->wr_ring[rslot].cap || bcf_site_scratch_bytes(sz) > c->wr_site.cap ||
			bcf_cta_scratch_bytes(sz) > c->wr_cta.cap || (nwb + 1) * 8 > c->wr_blocks.cap;
	if (grow) CU(cudaDeviceSynchronize());          // growing frees buffers that queued work may still use
	CU(c->wr_vcf.reserve((size_t)sz * sizeof(bsgpu_gt_vcf)));
	CU(c->wr_ring[rslot].reserve(rcap + 16));
	CU(c->wr_site.reserve(bcf_site_scratch_bytes(sz)));
	CU(c->wr_cta.reserve(bcf_cta_scratch_bytes(sz)));
	CU(c->wr_blocks.reserve((nwb + 1) * 8));
	sink->ring_cap[rslot] = rcap;
	std::vector<uint32_t> pairs(2 * nwb);
	for (size_t b = 0; b < nwb; b++) { pairs[2 * b] = wb[b].x - x; pairs[2 * b + 1] = wb[b].y - x; }
	CU(cudaMemcpyAsync(c->wr_blocks.p, pairs.data(), pairs.size() * 4, cudaMemcpyHostToDevice, c->stream));      // pageable: staged before the call returns
	if (block_dev(c, c->segs.p, nseg, c->obases.p, c->ref.p, x, sz, c->wr_vcf.p, 1, nullptr, c->stream) != BSGPU_OK) return BSGPU_FAIL;
	BcfJob j;
	j.d_vcf = c->wr_vcf.p; j.d_ref = c->ref.p; j.x = x; j.sz = sz; j.d_blocks = c->wr_blocks.p; j.nblocks = (uint32_t)nwb;
	j.p = sink->p; j.p.rid = sink->vcf_rid ? sink->vcf_rid[tid] : (int32_t)tid; j.p.ctg_end = ctg_len;
	j.dc = c->d_const; j.guard = c->d_counters; j.site_scratch = c->wr_site.p;
	{
		const auto it = c->contig_ann.find((int)tid);
		if (it != c->contig_ann.end() && it->second) { j.db = db_view(*it->second); j.reg_start = it->second->reg_start; j.reg_stop = it->second->reg_stop; }
	}
	job_stats(c, j);
	if (sink->queued >= 3) CU(cudaStreamWaitEvent(c->stream, c->wr_copied[rslot], 0));      // the slot's previous records have left
	CU(launch_bcf_records(j, 0, sz, c->wr_cta.p, c->wr_ring[rslot].p, rcap, c->d_wr_totals + 3 * (sink->queued & 7), c->stream, &c->launches));
	CU(cudaMemcpyAsync(c->h_wr_totals + 3 * (sink->queued & 7), c->d_wr_totals + 3 * (sink->queued & 7), 3 * sizeof(unsigned long long),
			cudaMemcpyDeviceToHost, c->stream));
	CU(cudaEventRecord(c->wr_built[rslot], c->stream));
	sink->queued++;
	while (sink->queued - sink->collected > 1) if (sink_collect(c, sink) != BSGPU_OK) return BSGPU_FAIL;
	return BSGPU_OK;
}

static int call_bam_impl(bsgpu_ctx *c, const uint8_t *bam, size_t nbytes, int n_targets, const uint32_t *target_len,
		const uint8_t *const *ctg_codes, const bsgpu_reader_params *rp, bsgpu_block *blocks, size_t block_cap, size_t *nblocks,
		bsgpu_gt_vcf *vcf, size_t vcf_cap, size_t *nvcf, BcfSink *sink, BamRunOpts *opts = nullptr) {
	if (!c || !rp || !nblocks || !nvcf || !target_len || !ctg_codes || (nbytes && !bam)) return fail("bsgpu_call_bam: null argument");
	size_t n = 0;
	uint64_t nb = 0, nm = 0;
	*nblocks = 0; *nvcf = 0;
	auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const double t0 = now();
	std::vector<size_t> chunk_end;
	const bool partial = opts && opts->partial;
	const bool pipelined = opts && opts->pipelined;
	bsgpu_ctx::ReaderSet &R = c->rs[opts ? opts->reader_set & 1 : 0];
	// a pipelined run passes the window stage on to the next run whatever way it ends -- after it has had its own turn
	struct Turn {
		bsgpu_ctx *c; uint64_t ticket; bool on;
		void wait() { if (!on) return; std::unique_lock<std::mutex> lk(c->win_mu); c->win_cv.wait(lk, [&] { return c->win_done >= ticket; }); }
		~Turn() { if (!on) return; wait(); std::unique_lock<std::mutex> lk(c->win_mu); if (c->win_done == ticket) c->win_done = ticket + 1; c->win_cv.notify_all(); }
	} turn{c, opts ? opts->ticket : 0, pipelined};
	// the session is told how far the run goes as soon as that is known (the staged bytes are free from then on); with the
	// --report-file tallies on the block builder still reads record lengths from the stream, so it is told at the very end
	bool told = false;
	auto tell = [&](size_t consumed_, size_t records_) { if (opts && opts->on_scanned && !told) { told = true; opts->on_scanned(opts->scan_user, consumed_, records_); } };
	struct Teller { std::function<void()> f; ~Teller() { f(); } } teller{[&] { if (opts) tell(opts->consumed, opts->records); }};
	size_t framed = nbytes;
	if (decode_queue(c, R, bam, nbytes, rp, 4, &n, &nb, &nm, chunk_end, opts ? &framed : nullptr, pipelined) != BSGPU_OK) return BSGPU_FAIL;
	if (opts && !partial && framed != nbytes) return fail("bsgpu_bam: the record stream ends inside a record (%zu bytes left over)", nbytes - framed);
	if (sink) sink->opts = opts;
	const double t1 = now();
	c->stats.bam_decode_s += t1 - t0;
	if (!n) return BSGPU_OK;
	const bsgpu_record *rec = (const bsgpu_record *)R.h_rec.p;
	CU(R.h_tmpl.reserve((n + 1) * sizeof(bsgpu_template)));
	bsgpu_template *tm = (bsgpu_template *)R.h_tmpl.p;
	unsigned long long before[4], after[4];       // normalisation failures are counted on the device: compared at the end
	CU(cudaMemcpy(before, c->d_counters, sizeof(before), cudaMemcpyDeviceToHost));
	// Chunk by chunk, as the descriptors come home: the part of the stream up to the last CERTAIN block start seen so far
	// is cut into pieces that a pool of host threads builds (read_input's state is reset at such records); the caller's
	// thread takes the pieces over in order and queues their device work, so the upload and decode of later chunks, the
	// builders, the kernels and the D2H of earlier windows all overlap.  A piece is processed as one window per contig it
	// touches; the windows of a contig tile the span from its first block's x to its last block's y.  Pileup counts are
	// additive per site and the model is per site, so every block's records are exactly what a per-block run gives;
	// positions between blocks come back as skip records.
	size_t ov = 0, nbk = 0, ntm = 0, built = 0, scanned = 0;
	int cur_tid = -1;
	uint32_t ctg_x0 = 0, ctg_end = 0;       // current contig: x of its first block, last position already written
	size_t ctg_ov = 0;                      // vcf index of position ctg_x0
	int ret = BSGPU_OK;
	double t_wait = 0;
	CertainState cst;
	std::vector<size_t> starts;
	// the builder threads of a job read the pinned arrays of the context: whatever way this function is left, the job is
	// joined first and the device is quiet
	struct JobGuard {
		bsgpu_ctx *c; std::vector<BuildJob *> jobs; std::thread scanner;
		~JobGuard() {
			if (scanner.joinable()) scanner.join();
			for (BuildJob *j : jobs) if (j) build_blocks_finish(j);
			cudaStreamSynchronize(c->slot[0].stream); cudaStreamSynchronize(c->slot[1].stream);
			cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->copy_stream);
			for (bool &b : c->ring_busy) b = false;
		}
	} guard{c};
	// Chunk by chunk as the keys and descriptors come home: the certain block starts of the chunk are marked (all host
	// threads over the 16-byte keys), a builder job is started for the records up to the last certain start seen so far, and
	// while its threads work the caller's thread takes over the pieces of the job BEFORE it and queues their windows -- the
	// builder of chunk k runs under the windows of chunk k - 1 and under the upload and decode of chunk k + 1.  Pieces that
	// are already built when the caller gets there share a window (consume below).  BSGPU_SINGLE_JOB=1: one job over the
	// whole stream once all chunks are home (29 ms per 1.6 M records against 27).
	double tm_rd = 0, tm_cert = 0, tm_piece = 0, tm_win = 0;
	c->tm_prep = c->tm_queue = c->tm_collect = 0;
	static const unsigned ppt = [] { const char *e = getenv("BSGPU_BUILDER_PIECES"); const int v = e ? atoi(e) : 0; return v > 0 ? (unsigned)v : 6u; }();
	// Pieces are taken over in stream order; the pieces that are ALREADY built when the caller gets there go together (up to
	// kGroup of them), so small pieces keep the builder's latency low without multiplying the windows and their launches.
	auto consume = [&](size_t ji) -> int {
		BuildJob *&job = guard.jobs[ji];
		const size_t np = build_blocks_pieces(job);
		constexpr size_t kGroup = 6;
		std::vector<bsgpu_block> gb;                       // blocks of the group, templates numbered from the start of the group
		std::vector<TmSpan> gs;                            // one run of templates per piece
		std::vector<size_t> gs_first;                      // group number of the first template of every run
		for (size_t p = 0; p < np && ret == BSGPU_OK;) {
			gb.clear(); gs.clear(); gs_first.clear();
			size_t gnt = 0;
			uint32_t gmax = 1;
			const size_t p0 = p;
			while (p < np && p - p0 < kGroup && (p == p0 || build_blocks_piece_ready(job, p))) {
				const std::vector<bsgpu_block> *pb;
				size_t base, nt_piece;
				const double w0 = now();
				const int rc = build_blocks_piece(job, p, &pb, &base, &nt_piece);
				t_wait += now() - w0;
				tm_piece += now() - w0;
				if (rc == -4) return fail("bsgpu_call_bam: duplicate read name among waiting mates");
				if (rc == -5) return fail("bsgpu_call_bam: the two mates of a template disagree about their positions");
				if (rc == -6) return fail("bsgpu_call_bam: a template's absent mate lies before its block's window (-k with -d): the reference asserts there (src/call_genotypes.c:186)");
				if (rc) return fail("bsgpu_call_bam: block builder failed (%d)", rc);
				if (const uint64_t *pt = build_blocks_piece_tally(job, p)) for (int k = 0; k < 30; k++) c->reader_tally[k] += pt[k];
				for (bsgpu_block o : *pb) { o.first_template += (uint32_t)gnt; gb.push_back(o); }
				gs.push_back(TmSpan{tm + base, nt_piece});
				gs_first.push_back(gnt);
				gnt += nt_piece;
				gmax = std::max(gmax, build_blocks_piece_maxcap(job, p));
				p++;
			}
			const double w2 = now();
			struct WinAcc { double &d; double t0; std::function<double()> f; ~WinAcc() { d += f() - t0; } } wacc_{tm_win, w2, now};
			if (opts && opts->blocks_vec) { opts->blocks_vec->resize(nbk + gb.size()); blocks = opts->blocks_vec->data(); }
			else if (nbk + gb.size() > block_cap) return fail("bsgpu_call_bam: blocks[] too small");
			for (size_t b0 = 0; b0 < gb.size() && ret == BSGPU_OK;) {
				size_t b1 = b0;
				while (b1 < gb.size() && gb[b1].tid == gb[b0].tid) b1++;
				const uint32_t tid = gb[b0].tid;
				if ((int)tid >= n_targets) return fail("bsgpu_call_bam: record on contig %u, only %d contigs given", tid, n_targets);
				if (!ctg_codes[tid] && opts && opts->need_contig && opts->need_contig(opts->contig_user, (int)tid) != BSGPU_OK)
					return fail("bsgpu_bam: the host's contig callback failed for contig %u", tid);
				if (!ctg_codes[tid]) return fail("bsgpu_call_bam: no reference codes were given for contig %u", tid);
				if ((int)tid != cur_tid) { cur_tid = (int)tid; ctg_x0 = gb[b0].x; ctg_end = ctg_x0 - 1; ctg_ov = ov; }
				// windows tile the contig; a block may begin ON the last site of the block before it (its x is two before its
				// first template, src/process_template.c:27), and the writer's context of its first sites reaches back there:
				// the record windows then overlap by that one uncovered site
				const uint32_t x = sink ? std::min(ctg_end + 1, gb[b0].x) : ctg_end + 1, y = gb[b1 - 1].y;
				const size_t t_lo = gb[b0].first_template, t_hi = (size_t)gb[b1 - 1].first_template + gb[b1 - 1].n_templates;
				if (y >= x) {
					const uint32_t sz = y - x + 1;
					if (!sink && ov + sz > vcf_cap) {
						if (!opts || !opts->grow) return fail("bsgpu_call_bam: vcf[] too small (contig %u needs %u more records)", tid, sz);
						CU(cudaStreamSynchronize(c->copy_stream));      // the windows queued so far have landed in the old buffer
						size_t ncap = 0;
						uint8_t *nbuf = opts->grow(opts->user, ov * sizeof(bsgpu_gt_vcf), (ov + sz) * sizeof(bsgpu_gt_vcf), &ncap);
						if (!nbuf) return fail("bsgpu_bam: cannot grow the result buffer to %zu records", ov + sz);
						vcf = (bsgpu_gt_vcf *)nbuf; vcf_cap = ncap / sizeof(bsgpu_gt_vcf);
					}
					// the runs of the group that hold templates [t_lo, t_hi)
					std::vector<TmSpan> ws;
					for (size_t r = 0; r < gs.size(); r++) {
						const size_t lo = std::max(t_lo, gs_first[r]), hi = std::min(t_hi, gs_first[r] + gs[r].n);
						if (hi > lo) ws.push_back(TmSpan{gs[r].p + (lo - gs_first[r]), hi - lo});
					}
					ret = call_window(c, R, ws.data(), ws.size(), t_hi - t_lo, tid, target_len[tid], ctg_codes[tid], x, y, sink ? nullptr : vcf + ov, gmax,
							sink, gb.data() + b0, b1 - b0);
					ov += sz;
					ctg_end = y;
				}
				for (size_t b = b0; b < b1; b++) {
					bsgpu_block o = gb[b];
					o.first_template = (uint32_t)(ntm + o.first_template);
					o.vcf_off = sink ? 0 : ctg_ov + (o.x - ctg_x0);
					blocks[nbk++] = o;
				}
				b0 = b1;
			}
			ntm += gnt;
		}
		build_blocks_finish(job);
		job = nullptr;
		return ret;
	};
	const bool trace = getenv("BSGPU_TIMING") != nullptr && atoi(getenv("BSGPU_TIMING")) > 1;
	const bool per_chunk = getenv("BSGPU_SINGLE_JOB") == nullptr;      // default: a builder job per chunk, started as the chunk comes home
	// The scan for certain starts and the start of the builder jobs run on a thread of their own, chunk by chunk as the
	// descriptors come home; the caller's thread does nothing but take finished pieces over and queue their windows.
	guard.jobs.reserve(chunk_end.size() + 1);            // the scanner appends, the caller reads entries below njobs: no reallocation
	std::atomic<size_t> njobs{0};
	std::atomic<int> scan_state{0};                      // 0 running, 1 done, -1 a CUDA call failed
	double sc_rd = 0, sc_cert = 0;
	cudaError_t scan_err = cudaSuccess;
	const bool check_scan = getenv("BSGPU_CHECK_SCAN") != nullptr, use_host_scan = check_scan || getenv("BSGPU_HOST_SCAN") != nullptr;
	std::atomic<int> scan_mismatch{0};
	std::unordered_map<std::string, uint32_t> host_names;
	size_t host_names_done = 0, name_fallbacks = 0;
	guard.scanner = std::thread([&] {
		cudaSetDevice(c->device);
		for (size_t ck = 0; ck < chunk_end.size(); ck++) {
			const double w0 = now();
			scan_err = cudaEventSynchronize(R.rd_done[ck]);
			if (scan_err != cudaSuccess) { scan_state.store(-1, std::memory_order_release); return; }
			const double w1 = now();
			sc_rd += w1 - w0;
			// the device's name table overflowed (a read name on more than five kept records, or hashes colliding): from this
			// chunk on the ids are computed here, over a map that first catches up with the records before
			if (((const uint32_t *)R.h_nameid.p)[n + 1 + ck] && chunk_end[ck] > scanned) {
				uint32_t *ids = (uint32_t *)R.h_nameid.p;
				for (size_t i = host_names_done; i < chunk_end[ck]; i++) {
					uint32_t id = 0xffffffffu;
					if (rec[i].ret <= 0 && (rec[i].alignment_flag & 1u)) {
						const uint8_t *p = bam + R.rec_off[i] + 4;
						id = host_names.emplace(std::string((const char *)p + 32, p[8]), (uint32_t)i).first->second;
					}
					if (i >= scanned) ids[i] = id;
				}
				host_names_done = chunk_end[ck];
				name_fallbacks++;
			}
			if (use_host_scan) certain_block_starts_keys((const uint32_t *)R.h_key.p, scanned, chunk_end[ck], &cst, starts);
			if (!use_host_scan || check_scan) {
				// the device's mask of the chunk: bit i of word i / 32 <-> record scanned + i
				std::vector<size_t> dev;
				const uint32_t *mw = (const uint32_t *)R.h_mask.p + R.mask_off[ck];
				const size_t cn = chunk_end[ck] - scanned;
				for (size_t w = 0; w < (cn + 31) / 32; w++) for (uint32_t b = mw[w]; b; b &= b - 1) dev.push_back(scanned + w * 32 + (size_t)__builtin_ctz(b));
				if (check_scan) {
					const size_t had = starts.size() - dev.size();
					if (starts.size() < dev.size() || !std::equal(dev.begin(), dev.end(), starts.begin() + had)) scan_mismatch.store(1);
				} else starts.insert(starts.end(), dev.begin(), dev.end());
			}
			sc_cert += now() - w1;
			scanned = chunk_end[ck];
			const bool last = ck + 1 == chunk_end.size();
			if (!per_chunk && !last) continue;
			// build up to the last certain start (everything when the stream is complete); starts[] holds those > built
			// (a session's run stops at the last certain start even when the buffer is complete: the records after it wait
			// for the bytes that follow)
			const bool to_end = last && !partial;
			size_t upto = to_end ? n : built;
			if (!to_end) for (size_t i = starts.size(); i-- > 0;) if (starts[i] > built) { upto = starts[i]; break; }
			if (upto > built) {
				std::vector<size_t> inside;
				for (size_t v : starts) if (v > built && v < upto) inside.push_back(v);
				guard.jobs.push_back(build_blocks_start_range(bam, R.rec_off.data(), rec, (const uint32_t *)R.h_nameid.p, built, upto, inside, rp->keep_unmatched, rp->keep_duplicates, tm, ppt, c->profile_on));
				njobs.store(guard.jobs.size(), std::memory_order_release);
				std::vector<size_t> keep;
				for (size_t v : starts) if (v >= upto) keep.push_back(v);
				starts.swap(keep);
				built = upto;
			}
			if (trace) fprintf(stderr, "  chunk %zu: descriptors home %.2f ms, job started %.2f ms (records %zu)\n", ck, (w1 - t0) * 1e3, (now() - t0) * 1e3, chunk_end[ck]);
		}
		if (!c->profile_on) tell(built == n ? framed : (size_t)R.rec_off[built], built);
		scan_state.store(1, std::memory_order_release);
	});
	turn.wait();                                          // pipelined: the run before this one has left the window stage
	for (size_t ji = 0; ret == BSGPU_OK;) {
		const double w0 = now();
		while (njobs.load(std::memory_order_acquire) <= ji && scan_state.load(std::memory_order_acquire) == 0) std::this_thread::yield();
		t_wait += now() - w0;
		tm_piece += now() - w0;
		if (njobs.load(std::memory_order_acquire) > ji) {
			ret = consume(ji++);
			if (trace) fprintf(stderr, "  job %zu consumed at %.2f ms\n", ji - 1, (now() - t0) * 1e3);
		}
		else break;                                       // the scanner has finished (or failed) and every job is consumed
	}
	guard.scanner.join();
	tm_rd += sc_rd; tm_cert += sc_cert;
	if (scan_state.load() < 0) { CU(scan_err); }
	if (scan_mismatch.load()) return fail("bsgpu_call_bam: BSGPU_CHECK_SCAN: the device's certain block starts differ from the host scan");
	while (ret == BSGPU_OK && sink && sink->collected < sink->queued) ret = sink_collect(c, sink);
	CU(cudaStreamSynchronize(c->slot[0].stream));
	CU(cudaStreamSynchronize(c->slot[1].stream));
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaStreamSynchronize(c->copy_stream));
	if (ret != BSGPU_OK) return ret;
	CU(cudaMemcpy(after, c->d_counters, sizeof(after), cudaMemcpyDeviceToHost));
	if (after[2] != before[2]) return fail("bsgpu_call_bam: Error in CIGAR - illegal soft clip in %llu template(s)", after[2] - before[2]);
	if (after[3] != before[3]) return fail("bsgpu_call_bam: %llu mate(s) start before their contig window", after[3] - before[3]);
	*nblocks = nbk;
	*nvcf = ov;
	if (opts) {
		opts->records = built;
		opts->consumed = built == n ? framed : (size_t)R.rec_off[built];
	}
	const double t2 = now();
	if (getenv("BSGPU_TIMING"))
		fprintf(stderr, "bsgpu_call_bam: decode_queue %.2f ms | wait descriptors %.2f, certain starts %.2f, wait pieces %.2f, windows %.2f (prep %.2f, queue %.2f, collect %.2f) | total %.2f ms\n",
				(t1 - t0) * 1e3, tm_rd * 1e3, tm_cert * 1e3, tm_piece * 1e3, tm_win * 1e3, c->tm_prep * 1e3, c->tm_queue * 1e3, c->tm_collect * 1e3, (t2 - t0) * 1e3);
	c->stats.bam_build_s += t_wait;
	c->stats.bam_call_s += t2 - t1 - t_wait;
	return BSGPU_OK;
}

int bsgpu_call_bam(bsgpu_ctx *c, const uint8_t *bam, size_t nbytes, int n_targets, const uint32_t *target_len,
		const uint8_t *const *ctg_codes, const bsgpu_reader_params *rp, bsgpu_block *blocks, size_t block_cap, size_t *nblocks,
		bsgpu_gt_vcf *vcf, size_t vcf_cap, size_t *nvcf) {
	return call_bam_impl(c, bam, nbytes, n_targets, target_len, ctg_codes, rp, blocks, block_cap, nblocks, vcf, vcf_cap, nvcf, nullptr);
}

int bsgpu_call_bam_bcf(bsgpu_ctx *c, const uint8_t *bam, size_t nbytes, int n_targets, const uint32_t *target_len,
		const uint8_t *const *ctg_codes, const bsgpu_reader_params *rp, const bsgpu_bcf_params *p, const int32_t *vcf_rid,
		bsgpu_block *blocks, size_t block_cap, size_t *nblocks, uint8_t *out, size_t out_cap, size_t *nbytes_out, size_t *nrec) {
	if (!p || !nbytes_out || !nrec || (out_cap && !out)) return fail("bsgpu_call_bam_bcf: null argument");
	BcfSink sink;
	sink.p = *p; sink.vcf_rid = vcf_rid; sink.out = out; sink.out_cap = out_cap;
	size_t nvcf = 0;
	*nbytes_out = *nrec = 0;
	if (call_bam_impl(c, bam, nbytes, n_targets, target_len, ctg_codes, rp, blocks, block_cap, nblocks, nullptr, 0, &nvcf, &sink) != BSGPU_OK) return BSGPU_FAIL;
	*nbytes_out = sink.at; *nrec = sink.recs;
	return BSGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// streaming session: read_input's O(block) memory behaviour for a stream of any length
// (src/get_template_vector.c:86-110 reads one record at a time and hands blocks on as they close).
// The host logic -- staging, batching, worker thread, hand-over of results -- is bsgpu_session.h; here are its hooks:
// page-locked memory and the run of one batch through call_bam_impl.
// ------------------------------------------------------------------------------------------------
struct bsgpu_bam_session {
	bsgpu_ctx *c = nullptr;
	int n_targets = 0;
	std::vector<uint32_t> target_len;
	std::vector<const uint8_t *> codes;
	bsgpu_reader_params rp;
	bool bcf = false;
	bsgpu_bcf_params bp;
	std::vector<int32_t> rid;
	int (*on_contig)(void *user, int tid) = nullptr;
	void *on_contig_user = nullptr;
	std::mutex contig_mu;
	Session core;
};

static uint8_t *sess_pin(size_t bytes) {
	void *p = nullptr;
	return cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess ? (uint8_t *)p : nullptr;
}
static void sess_unpin(uint8_t *p) { cudaFreeHost(p); }
static void sess_thread_init(void *user) { cudaSetDevice(((bsgpu_bam_session *)user)->c->device); }

static int sess_run(void *user, const uint8_t *data, size_t len, bool whole, SessResult *r, uint64_t seq,
		void (*scanned)(void *sess, uint64_t seq, size_t consumed, size_t records), void *sess,
		size_t *consumed, size_t *records, std::string *err) {
	bsgpu_bam_session *s = (bsgpu_bam_session *)user;
	Session::GrowCtx g{&s->core, r};
	struct Told { void (*f)(void *, uint64_t, size_t, size_t); void *sess; uint64_t seq; } told{scanned, sess, seq};
	BamRunOpts o;
	o.partial = !whole; o.blocks_vec = &r->blocks; o.grow = Session::grow; o.user = &g;
	if (s->on_contig) {
		o.need_contig = [](void *u, int tid) {
			bsgpu_bam_session *ss = (bsgpu_bam_session *)u;
			std::lock_guard<std::mutex> lk(ss->contig_mu);
			return ss->codes[tid] ? (int)BSGPU_OK : ss->on_contig(ss->on_contig_user, tid);
		};
		o.contig_user = s;
	}
	// runs of a session overlap: reader stage of this one under the window stage of the one before (BSGPU_SESSION_PIPELINE=0: one at a time)
	static const bool pipeline = [] { const char *e = getenv("BSGPU_SESSION_PIPELINE"); return !e || atoi(e) != 0; }();
	o.pipelined = true;
	o.reader_set = (int)(seq & 1);
	{ std::unique_lock<std::mutex> lk(s->c->win_mu); o.ticket = s->c->next_ticket++; }
	if (pipeline) {
		o.on_scanned = [](void *u, size_t consumed_, size_t records_) { Told *t = (Told *)u; t->f(t->sess, t->seq, consumed_, records_); };
		o.scan_user = &told;
	}
	size_t nblk = 0, nvcf = 0;
	int rc;
	// BSGPU_SERIALIZE_SESSIONS=1 (debugging aid): the batches of all sessions of the process run one at a time
	static std::mutex serial;
	static const bool serialize = getenv("BSGPU_SERIALIZE_SESSIONS") != nullptr;
	std::unique_lock<std::mutex> sl(serial, std::defer_lock);
	if (serialize) sl.lock();
	if (s->bcf) {
		BcfSink sink;
		sink.p = s->bp; sink.vcf_rid = s->rid.empty() ? nullptr : s->rid.data(); sink.out = r->buf; sink.out_cap = r->cap;
		rc = call_bam_impl(s->c, data, len, s->n_targets, s->target_len.data(), s->codes.data(), &s->rp, nullptr, 0, &nblk, nullptr, 0, &nvcf, &sink, &o);
		r->nbytes = sink.at; r->nrec = sink.recs;
	} else {
		rc = call_bam_impl(s->c, data, len, s->n_targets, s->target_len.data(), s->codes.data(), &s->rp, nullptr, 0, &nblk, (bsgpu_gt_vcf *)r->buf,
				r->cap / sizeof(bsgpu_gt_vcf), &nvcf, nullptr, &o);
		r->nbytes = nvcf * sizeof(bsgpu_gt_vcf); r->nrec = nvcf;
	}
	r->blocks.resize(nblk);
	*consumed = o.consumed; *records = o.records;
	if (rc != BSGPU_OK) { *err = g_err; return -1; }
	return 0;
}

int bsgpu_bam_open(bsgpu_ctx *c, int n_targets, const uint32_t *target_len, const uint8_t *const *ctg_codes, const bsgpu_reader_params *rp,
		const bsgpu_bcf_params *bcf, const int32_t *vcf_rid, size_t batch_bytes, bsgpu_bam_session **out) {
	if (!c || !rp || !out || n_targets <= 0 || !target_len || !ctg_codes) return fail("bsgpu_bam_open: null argument");
	if (cudaSetDevice(c->device) != cudaSuccess) return fail("bsgpu_bam_open: cannot select device %d", c->device);
	bsgpu_bam_session *s = new bsgpu_bam_session();
	s->c = c; s->n_targets = n_targets; s->rp = *rp;
	s->target_len.assign(target_len, target_len + n_targets);
	s->codes.assign(ctg_codes, ctg_codes + n_targets);
	s->bcf = bcf != nullptr;
	if (bcf) s->bp = *bcf;
	if (vcf_rid) s->rid.assign(vcf_rid, vcf_rid + n_targets);
	if (!batch_bytes) { const char *e = getenv("BSGPU_BATCH_BYTES"); batch_bytes = e && atoll(e) > 0 ? (size_t)atoll(e) : (size_t)384 << 20; }
	SessHooks hk;
	hk.user = s; hk.alloc = sess_pin; hk.release = sess_unpin; hk.thread_init = sess_thread_init; hk.run = sess_run;
	// first guess of a batch's results: BCF records are about as many bytes as the BAM records they come from, gt_vcf[] four times that
	if (!s->core.open(hk, batch_bytes, s->bcf ? 1.0 : 4.2)) {
		s->core.close();
		delete s;
		return fail("bsgpu_bam_open: cannot allocate page-locked staging for batches of %zu bytes", batch_bytes);
	}
	*out = s;
	return BSGPU_OK;
}

int bsgpu_bam_set_contig(bsgpu_bam_session *s, int tid, const uint8_t *codes) {
	if (!s || tid < 0 || tid >= s->n_targets) return fail("bsgpu_bam_set_contig: contig %d out of range", tid);
	s->codes[tid] = codes;               // read by the worker only for contigs whose records are in a batch
	return BSGPU_OK;
}

int bsgpu_bam_on_contig(bsgpu_bam_session *s, int (*fn)(void *user, int tid), void *user) {
	if (!s) return fail("bsgpu_bam_on_contig: null session");
	s->on_contig = fn; s->on_contig_user = user;
	return BSGPU_OK;
}

int bsgpu_bam_feed(bsgpu_bam_session *s, const uint8_t *bytes, size_t nbytes, size_t *accepted) {
	if (!s || (nbytes && !bytes)) return fail("bsgpu_bam_feed: null argument");
	std::string err;
	size_t acc = 0;
	const bool ok = s->core.feed(bytes, nbytes, accepted != nullptr, &acc, &err);
	if (accepted) *accepted = acc;
	return ok ? BSGPU_OK : fail("%s", err.c_str());
}

int bsgpu_bam_reserve(bsgpu_bam_session *s, uint8_t **ptr, size_t *avail, int wait) {
	if (!s || !ptr || !avail) return fail("bsgpu_bam_reserve: null argument");
	std::string err;
	return s->core.reserve(ptr, avail, wait != 0, &err) ? BSGPU_OK : fail("%s", err.c_str());
}

int bsgpu_bam_commit(bsgpu_bam_session *s, size_t nbytes) {
	if (!s) return fail("bsgpu_bam_commit: null argument");
	std::string err;
	return s->core.commit(nbytes, &err) ? BSGPU_OK : fail("%s", err.c_str());
}

int bsgpu_bam_finish(bsgpu_bam_session *s) {
	if (!s) return fail("bsgpu_bam_finish: null argument");
	std::string err;
	return s->core.mark(true, &err) ? BSGPU_OK : fail("%s", err.c_str());
}

int bsgpu_bam_cut(bsgpu_bam_session *s) {
	if (!s) return fail("bsgpu_bam_cut: null argument");
	std::string err;
	return s->core.mark(false, &err) ? BSGPU_OK : fail("%s", err.c_str());
}

int bsgpu_bam_rewind(bsgpu_bam_session *s) {
	if (!s) return fail("bsgpu_bam_rewind: null argument");
	std::string err;
	return s->core.rewind(&err) ? BSGPU_OK : fail("%s", err.c_str());
}

int bsgpu_bam_drain(bsgpu_bam_session *s, int wait, bsgpu_bam_result *res) {
	if (!s || !res) return fail("bsgpu_bam_drain: null argument");
	memset(res, 0, sizeof(*res));
	std::string err;
	SessResult *r = nullptr;
	bool done = false;
	if (!s->core.drain(wait != 0, &r, &done, &err)) return fail("%s", err.c_str());
	res->finished = done ? 1 : 0;
	if (!r) return BSGPU_OK;
	res->id = r->id; res->blocks = r->blocks.data(); res->nblocks = r->blocks.size();
	res->data = r->buf; res->nbytes = r->nbytes; res->nrec = r->nrec; res->bytes_in = r->bytes_in; res->records_in = r->records_in;
	return BSGPU_OK;
}

int bsgpu_bam_release(bsgpu_bam_session *s, uint64_t id) {
	if (!s) return fail("bsgpu_bam_release: null argument");
	return s->core.give_back(id) ? BSGPU_OK : fail("bsgpu_bam_release: no result %llu is lent out", (unsigned long long)id);
}

int bsgpu_bam_progress(bsgpu_bam_session *s, bsgpu_bam_progress_t *out) {
	if (!s || !out) return fail("bsgpu_bam_progress: null argument");
	s->core.progress(out);
	return BSGPU_OK;
}

int bsgpu_bam_close(bsgpu_bam_session *s) {
	if (!s) return BSGPU_OK;
	cudaSetDevice(s->c->device);
	s->core.close();
	delete s;
	return BSGPU_OK;
}

int bsgpu_pileup_block(bsgpu_ctx *c, const bsgpu_seg *segs, size_t nseg, const uint8_t *bases, size_t nbases, uint32_t x, uint32_t sz, bsgpu_pileup *out) {
	return block_host(c, segs, nseg, bases, nbases, nullptr, x, sz, out, 0);
}

int bsgpu_call_block(bsgpu_ctx *c, const bsgpu_seg *segs, size_t nseg, const uint8_t *bases, size_t nbases, const uint8_t *ref, uint32_t x, uint32_t sz, bsgpu_gt_vcf *out) {
	return block_host(c, segs, nseg, bases, nbases, ref, x, sz, out, 1);
}

// ------------------------------------------------------------------------------------------------
// host staging: normalised templates -> segments.  Mirrors the walk at the top of the reference's pileup loop
// (src/call_genotypes.c:181-212, 224): a mate is emitted when it is present, non-empty and holds at least one base
// with 0 < q != 63; only such a mate flips the strand index for the next one.
// ------------------------------------------------------------------------------------------------
size_t bsgpu_stage_bound(const bsgpu_template *t, size_t n) {
	size_t b = 0;
	for (size_t i = 0; i < n; i++) for (int k = 0; k < 2; k++)
		if (t[i].present[k]) b += (t[i].read_len[k] + BSGPU_MAX_SEG_LEN - 1) / BSGPU_MAX_SEG_LEN;
	return b;
}

int bsgpu_stage_templates(const bsgpu_template *t, size_t n, const uint8_t *bases, uint32_t x, uint32_t y, bsgpu_seg *segs, size_t *nseg) {
	if (!nseg || (n && (!t || !bases || !segs))) return fail("bsgpu_stage_templates: null argument");
	size_t ns = 0;
	for (size_t i = 0; i < n; i++) {
		uint32_t ori = t[i].orientation & 1u;
		const uint32_t st = t[i].bs_strand;
		if (st > 2) return fail("bsgpu_stage_templates: template %zu has bs_strand %u", i, st);
		for (int k = 0; k < 2; k++) {
			if (!t[i].present[k] || !t[i].read_len[k]) continue;
			const uint8_t *sp = bases + t[i].read_off[k];
			const uint32_t rl = t[i].read_len[k];
			uint32_t first = 0, last = rl;
			while (first < rl) { const uint8_t q = sp[first] >> 2; if (q > 0 && q != BSGPU_FLT_QUAL) break; first++; }
			if (first == rl) continue;
			while (true) { const uint8_t q = sp[last - 1] >> 2; if (q > 0 && q != BSGPU_FLT_QUAL) break; last--; }
			uint32_t pos = (k ? t[i].reverse_position : t[i].forward_position) + first;
			if (pos < x) return fail("bsgpu_stage_templates: template %zu starts before the window (%u < %u)", i, pos, x);
			uint32_t off = t[i].read_off[k] + first, len = last - first;
			if (pos <= y) {
				if ((uint64_t)pos + len > (uint64_t)y + 1) len = y + 1 - pos;          // pos <= y clip (:213)
				while (len) {
					const uint32_t l = len > BSGPU_MAX_SEG_LEN ? BSGPU_MAX_SEG_LEN : len;
					bsgpu_seg &s = segs[ns++];
					s.pos = pos; s.off = off; s.len = (uint16_t)l; s.mapq = t[i].mapq[k]; s.flags = (uint8_t)(ori | (st << 1)); s.pad_ = 0;
					pos += l; off += l; len -= l;
				}
			}
			ori ^= 1u;
		}
	}
	*nseg = ns;
	return BSGPU_OK;
}

}  // extern "C"
