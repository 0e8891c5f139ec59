// bsgpu_wire.h -- the compact wire record of a called site and its host side (internal; host-only code, no CUDA).
//
// A gt_meth record (include/bs_call.h:152-160) is 200 bytes of which 120 carry information: the eight counts are stored
// as uint64 but are sums of two uint32, the eight mean qualities / mq / aq are ints below 64, and seven bytes are padding.
// The host-buffer entry points are bound by PCIe on the way home, so results cross it as wire
// records and a pool of host threads rebuilds the reference's records in the caller's array ("rehydration"):
//
//     word  0- 9   gt_prob[10]          (doubles, bit for bit)
//     word 10      fisher_strand
//     word 11-12   counts[8]   as uint16
//     word 13      qual[8]     as uint8
//     word 14      mq | aq << 8 | max_gt << 16 | skip << 24      (upper half 0)
//
// A chunk in which a field does not fit its wire width (a count above 65535, a quality above 255: not reachable from BAM
// qualities, but the count vectors are the caller's) raises the chunk's flag word and is fetched again as full records.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include <emmintrin.h>

namespace bsgpu {

constexpr int kWireWords = 15;
constexpr int kWireBytes = kWireWords * 8;      // 120

// n wire records -> n records of `rec_bytes` (200: gt_meth + skip[] byte array; 208: gt_vcf with ready = 1 and skip inside).
// Streaming stores: the destination is written once, in order, and is far larger than any cache.
static inline void wire_expand(const uint64_t *w, size_t n, uint8_t *out, size_t rec_bytes, uint8_t *skip) {
	long long *o = (long long *)out;
	const bool vcf = rec_bytes == 208;
	for (size_t i = 0; i < n; i++, w += kWireWords) {
		const uint64_t c0 = w[11], c1 = w[12], q = w[13], m = w[14];
		_mm_stream_si64(o + 0, (long long)(c0 & 0xffff));
		_mm_stream_si64(o + 1, (long long)(c0 >> 16 & 0xffff));
		_mm_stream_si64(o + 2, (long long)(c0 >> 32 & 0xffff));
		_mm_stream_si64(o + 3, (long long)(c0 >> 48));
		_mm_stream_si64(o + 4, (long long)(c1 & 0xffff));
		_mm_stream_si64(o + 5, (long long)(c1 >> 16 & 0xffff));
		_mm_stream_si64(o + 6, (long long)(c1 >> 32 & 0xffff));
		_mm_stream_si64(o + 7, (long long)(c1 >> 48));
		_mm_stream_si64(o + 8, (long long)((q & 0xff) | (q >> 8 & 0xff) << 32));
		_mm_stream_si64(o + 9, (long long)((q >> 16 & 0xff) | (q >> 24 & 0xff) << 32));
		_mm_stream_si64(o + 10, (long long)((q >> 32 & 0xff) | (q >> 40 & 0xff) << 32));
		_mm_stream_si64(o + 11, (long long)((q >> 48 & 0xff) | (q >> 56) << 32));
		for (int g = 0; g < 11; g++) _mm_stream_si64(o + 12 + g, (long long)w[g]);
		_mm_stream_si64(o + 23, (long long)((m & 0xff) | (m >> 8 & 0xff) << 32));
		_mm_stream_si64(o + 24, (long long)(m >> 16 & 0xff));
		const uint64_t sk = m >> 24 & 0xff;
		if (vcf) { _mm_stream_si64(o + 25, (long long)(1ull | sk << 8)); o += 26; }
		else { skip[i] = (uint8_t)sk; o += 25; }
	}
	_mm_sfence();
}

// The inverse, for tests and for hosts that want to see what the device sends: n records -> wire; returns false if a
// field does not fit (the device raises the chunk's flag in that case).
static inline bool wire_pack_host(const uint8_t *rec, size_t n, size_t rec_bytes, const uint8_t *skip, uint64_t *w) {
	bool ok = true;
	for (size_t i = 0; i < n; i++, rec += rec_bytes, w += kWireWords) {
		uint64_t r[26];
		memcpy(r, rec, rec_bytes);
		uint64_t c[2] = {0, 0}, q = 0;
		for (int j = 0; j < 8; j++) {
			if (r[j] > 0xffff) ok = false;
			c[j >> 2] |= (r[j] & 0xffff) << 16 * (j & 3);
			const uint32_t qj = (uint32_t)(r[8 + (j >> 1)] >> 32 * (j & 1));
			if (qj > 0xff) ok = false;
			q |= (uint64_t)(qj & 0xff) << 8 * j;
		}
		const uint32_t mq = (uint32_t)r[23], aq = (uint32_t)(r[23] >> 32);
		if (mq > 0xff || aq > 0xff) ok = false;
		for (int g = 0; g < 11; g++) w[g] = r[12 + g];
		w[11] = c[0]; w[12] = c[1]; w[13] = q;
		const uint64_t sk = rec_bytes == 208 ? (r[25] >> 8 & 0xff) : skip[i];
		w[14] = (mq & 0xff) | (uint64_t)(aq & 0xff) << 8 | (r[24] & 0xff) << 16 | sk << 24;
	}
	return ok;
}

// Pool of host threads that rebuild records from wire chunks as they come home.  The submitting thread asks for a pinned
// wire buffer (`acquire`: none free -> the caller sends that chunk as full records instead, which balances the PCIe
// link against the host's memory system whatever their ratio is) and submits the chunk together with a function that
// waits for its device-to-host copy (`submit`); a watcher thread releases landed chunks to the workers.  Pieces of a chunk
// go to whichever thread is free.
class WireExpander {
public:
	struct Job {
		const uint64_t *wire = nullptr;      // n records
		const uint64_t *flag = nullptr;      // the chunk's flag word (non-zero: a field did not fit, nothing is rebuilt), or NULL
		size_t n = 0;
		uint8_t *out = nullptr, *skip = nullptr;
		size_t rec_bytes = 200;
		int buf = -1;                        // pinned buffer to give back
		size_t chunk_id = 0;                 // reported in failed() when the flag word is set
		void (*wait_landed)(void *) = nullptr;      // blocks until the chunk has landed in `wire` (NULL: it is there already)
		void *wait_arg = nullptr;
		size_t next = 0, done = 0;           // pieces handed out / finished
		bool landed = false;
	};
	static constexpr size_t kPiece = 8192;   // sites per piece: 2.6 MB of traffic

	explicit WireExpander(unsigned threads) {
		for (unsigned t = 0; t < (threads ? threads : 1); t++) workers_.emplace_back([this] { run(); });
		workers_.emplace_back([this] { watch(); });
	}
	~WireExpander() {
		{ std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
		cv_.notify_all();
		watch_cv_.notify_all();
		for (auto &t : workers_) t.join();
	}
	unsigned threads() const { return (unsigned)workers_.size() - 1; }
	void set_buffers(int n) { std::lock_guard<std::mutex> lk(mu_); free_.clear(); for (int i = 0; i < n; i++) free_.push_back(i); nbuf_ = n; }
	// a free pinned buffer, or -1
	int acquire() {
		std::lock_guard<std::mutex> lk(mu_);
		if (free_.empty()) return -1;
		const int b = free_.back();
		free_.pop_back();
		return b;
	}
	void give_back(int b) { std::lock_guard<std::mutex> lk(mu_); free_.push_back(b); }
	void submit(const Job &j) {      // j.n > 0
		{ std::lock_guard<std::mutex> lk(mu_); jobs_.push_back(j); jobs_.back().landed = !j.wait_landed; pending_++; }
		if (j.wait_landed) watch_cv_.notify_one();
		else cv_.notify_all();
	}
	// every submitted job finished
	void wait_idle() {
		std::unique_lock<std::mutex> lk(mu_);
		idle_cv_.wait(lk, [this] { return pending_ == 0; });
	}
	// chunks whose flag word was set (not expanded); cleared by the call
	std::vector<size_t> failed() { std::lock_guard<std::mutex> lk(mu_); std::vector<size_t> f; f.swap(failed_); return f; }
	uint64_t sites_expanded() const { return expanded_.load(std::memory_order_relaxed); }

private:
	// the watcher: waits for submitted chunks to land, in submission order, and releases them to the workers (a host function
	// queued on the stream would do, but the stream stalls for as long as the driver takes to get round to calling it)
	void watch() {
		std::unique_lock<std::mutex> lk(mu_);
		for (;;) {
			watch_cv_.wait(lk, [this] { if (stop_) return true; for (const Job &j : jobs_) if (!j.landed) return true; return false; });
			if (stop_) return;
			size_t k = 0;
			while (jobs_[k].landed) k++;
			const Job snap = jobs_[k];
			lk.unlock();
			snap.wait_landed(snap.wait_arg);
			lk.lock();
			for (Job &j : jobs_) if (j.buf == snap.buf && j.chunk_id == snap.chunk_id) { j.landed = true; break; }
			cv_.notify_all();
		}
	}
	void run() {
		std::unique_lock<std::mutex> lk(mu_);
		for (;;) {
			cv_.wait(lk, [this] { return stop_ || has_piece(); });
			if (stop_) return;
			// first job that still has pieces to hand out
			size_t k = 0;
			while (!jobs_[k].landed || jobs_[k].next * kPiece >= jobs_[k].n) k++;
			Job &j = jobs_[k];
			const size_t piece = j.next++;
			const size_t a = piece * kPiece, b = a + kPiece < j.n ? a + kPiece : j.n;
			const uint64_t flag = j.flag ? *j.flag : 0;
			const Job snap = j;
			lk.unlock();
			if (!flag && b > a) wire_expand(snap.wire + a * kWireWords, b - a, snap.out + a * snap.rec_bytes, snap.rec_bytes, snap.skip ? snap.skip + a : nullptr);
			lk.lock();
			// the deque may have lost finished jobs at its front meanwhile: find ours by its buffer
			for (size_t i = 0; i < jobs_.size(); i++) if (jobs_[i].buf == snap.buf && jobs_[i].chunk_id == snap.chunk_id) {
				Job &mine = jobs_[i];
				mine.done++;
				const size_t pieces = (mine.n + kPiece - 1) / kPiece;
				if (mine.done == pieces) {
					if (flag) failed_.push_back(mine.chunk_id);
					else expanded_.fetch_add(mine.n, std::memory_order_relaxed);
					free_.push_back(mine.buf);
					jobs_.erase(jobs_.begin() + (long)i);
					if (--pending_ == 0) idle_cv_.notify_all();
				}
				break;
			}
		}
	}
	bool has_piece() const {
		for (const Job &j : jobs_) if (j.landed && j.next * kPiece < j.n) return true;
		return false;
	}
	std::mutex mu_;
	std::condition_variable cv_, idle_cv_, watch_cv_;
	std::deque<Job> jobs_;
	std::vector<int> free_;
	std::vector<size_t> failed_;
	std::vector<std::thread> workers_;
	std::atomic<uint64_t> expanded_{0};
	size_t pending_ = 0;
	int nbuf_ = 0;
	bool stop_ = false;
};

}  // namespace bsgpu
