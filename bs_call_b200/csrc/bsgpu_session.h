// bsgpu_session.h -- host logic of the streaming session (bsgpu_bam_open / _feed / _drain ..., include/bsgpu.h): staging,
// batching, the worker thread, the hand-over of results.  No CUDA in here: page-locked memory and the run of one batch
// (decode -> blocks -> normalisation -> pileup -> model [-> writer], bsgpu_api.cu) come in through hooks, so that the same
// code is exercised on the CPU by tests/session_harness.cpp with a stand-in runner.
//
// What it mirrors: read_input() reads one record at a time and hands a block on as soon as it closes
// (src/get_template_vector.c:86-110, 140-189), so bs_call's memory is O(block) whatever the length of the stream.  Here
// the stream arrives in arbitrary slices.  Two staging buffers alternate: while the caller fills one, the worker runs the
// one before it; a run turns everything up to the last CERTAIN block start of the batch into results, the records from
// there on (read_input's state is blank at such a record, so nothing is lost by starting over there) are the `carry`,
// copied in front of the bytes the caller has fed meanwhile.  Results of a batch sit in a buffer of the session that is
// lent to the caller until it is released.  Memory: two batches + the results not yet released.
//
// Two runs can be in flight: a run reports how far it goes (`scanned`) as soon as its reader stage knows -- the staged bytes
// are free from then on and the carry can be placed -- and the next batch starts its own reader stage while the run before
// it is still at its window stage (the runner serialises the window stages and keeps them in order).
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "bsgpu.h"

namespace bsgpu {

struct SessResult {
	uint8_t *buf = nullptr;
	size_t cap = 0, nbytes = 0, nrec = 0, bytes_in = 0, records_in = 0;
	std::vector<bsgpu_block> blocks;
	uint64_t id = 0;
	bool last = false;
};

struct SessHooks {
	void *user = nullptr;
	uint8_t *(*alloc)(size_t bytes) = nullptr;        // page-locked memory
	void (*release)(uint8_t *p) = nullptr;
	void (*thread_init)(void *user) = nullptr;        // first thing the worker thread does
	// One batch.  `whole`: the bytes end where a block ends (end of stream / bsgpu_bam_cut) -- everything is processed and a
	// record cut off by the end of the buffer is an error; otherwise the run stops at the last certain block start.  Fills
	// r->buf (growing it with Session::grow) / nbytes / nrec / blocks, *consumed (bytes accounted for) and *records.
	// seq numbers the runs of the session; runs seq and seq + 1 may overlap as described above.  scanned(sess, seq, consumed,
	// records) must be called exactly once per successful run, at the latest before it returns; after the call the run must not
	// read data[] any more.
	int (*run)(void *user, const uint8_t *data, size_t len, bool whole, SessResult *r, uint64_t seq,
			void (*scanned)(void *sess, uint64_t seq, size_t consumed, size_t records), void *sess,
			size_t *consumed, size_t *records, std::string *err) = nullptr;
};

class Session {
public:
	struct Stage { uint8_t *p = nullptr; size_t cap = 0, head = 0, start = 0, fill = 0; };      // data = [start, fill); fed bytes from `head` on
	SessHooks hk;
	size_t batch_bytes = 0;
	double result_ratio = 1.0;               // first guess of a batch's result bytes per input byte
	Stage stage[2];
	int w = 0;                               // staging buffer the caller fills
	bool filling = false;                    // the caller is writing into stage[w] outside the lock
	int carry_waiters = 0;                   // runs waiting for the caller's reservation to close so that they can move their carry
	std::mutex mu;
	std::condition_variable cv;
	struct Run { std::thread th; int buf = 0; bool final = false, whole = false, scanned = false, returned = false; uint64_t seq = 0; size_t len = 0; const uint8_t *data = nullptr; };
	std::deque<Run *> runs;                  // in flight, oldest first
	std::vector<Run *> dead;                 // returned: joined at the next opportunity
	uint64_t next_seq = 0, next_result_seq = 0;
	bool scan_pending = false;               // the newest run has not said how far it goes: the carry of the next batch is unknown
	bool finishing = false, final_submitted = false, finished = false, closing = false, failed = false;
	bool cut_pending = false;
	size_t max_outstanding = 3;              // runs in flight + results not yet drained
	std::string errmsg;
	std::deque<SessResult *> ready;
	std::vector<SessResult *> pool, lent;
	size_t max_ready = 2;
	uint64_t next_id = 1, bytes_fed = 0, bytes_done = 0, records_done = 0, batches = 0, carry_bytes = 0, empty_batches = 0;

	// result buffer too small: a larger one holding the first `keep` bytes.  For the runner (signature of BamRunOpts::grow).
	struct GrowCtx { Session *s; SessResult *r; };
	static uint8_t *grow(void *user, size_t keep, size_t need, size_t *new_cap) {
		GrowCtx *g = (GrowCtx *)user;
		SessResult *r = g->r;
		const size_t want = std::max(need + need / 2, r->cap * 2) + 4096;
		uint8_t *nb = g->s->hk.alloc(want);
		if (!nb) return nullptr;
		if (keep) memcpy(nb, r->buf, keep);
		if (r->buf) g->s->hk.release(r->buf);
		r->buf = nb; r->cap = want;
		*new_cap = want;
		return nb;
	}

	bool open(const SessHooks &hooks, size_t batch, double ratio) {
		hk = hooks;
		batch_bytes = std::max<size_t>(batch, 4096);
		result_ratio = ratio;
		for (int b = 0; b < 2; b++) if (!stage_fit(b, std::max<size_t>(batch_bytes / 8, 4096))) return false;
		return true;
	}

	// takes bytes until they are all in, or (nowait) until it would have to wait for the worker.  false: the session has failed
	bool feed(const uint8_t *bytes, size_t nbytes, bool nowait, size_t *accepted, std::string *err) {
		std::unique_lock<std::mutex> lk(mu);
		*accepted = 0;
		reap(lk);
		if (finishing) { *err = "bsgpu_bam_feed: the stream was finished"; return false; }
		if (filling) { *err = "bsgpu_bam_feed: a reservation is open"; return false; }
		size_t off = 0;
		while (off < nbytes) {
			if (carry_waiters) cv.wait(lk, [&] { return carry_waiters == 0 || failed || closing; });
			if (!room(lk, nowait)) break;
			Stage &st = stage[w];
			const size_t m = std::min(nbytes - off, batch_bytes - (st.fill - st.head));
			uint8_t *dst = st.p + st.fill;
			filling = true;
			lk.unlock();
			copy(dst, bytes + off, m);
			lk.lock();
			filling = false;
			stage[w].fill += m;
			off += m;
			bytes_fed += m;
			try_submit();
			cv.notify_all();
		}
		*accepted = off;
		if (failed) { *err = errmsg; return false; }
		return true;
	}

	bool reserve(uint8_t **ptr, size_t *avail, bool wait, std::string *err) {
		std::unique_lock<std::mutex> lk(mu);
		*ptr = nullptr; *avail = 0;
		if (finishing) { *err = "bsgpu_bam_reserve: the stream was finished"; return false; }
		if (filling) { *err = "bsgpu_bam_reserve: the previous reservation was not committed"; return false; }
		if (carry_waiters) cv.wait(lk, [&] { return carry_waiters == 0 || failed || closing; });
		if (!room(lk, !wait)) { if (failed) { *err = errmsg; return false; } return true; }
		Stage &st = stage[w];
		*ptr = st.p + st.fill;
		*avail = batch_bytes - (st.fill - st.head);
		filling = true;
		return true;
	}

	bool commit(size_t nbytes, std::string *err) {
		std::unique_lock<std::mutex> lk(mu);
		if (!filling) { *err = "bsgpu_bam_commit: nothing reserved"; return false; }
		Stage &st = stage[w];
		if (nbytes > batch_bytes - (st.fill - st.head)) { *err = "bsgpu_bam_commit: more bytes committed than were reserved"; return false; }
		filling = false;
		st.fill += nbytes;
		bytes_fed += nbytes;
		try_submit();
		cv.notify_all();
		if (failed) { *err = errmsg; return false; }
		return true;
	}

	bool mark(bool end_of_stream, std::string *err) {          // bsgpu_bam_cut / bsgpu_bam_finish
		std::unique_lock<std::mutex> lk(mu);
		if (filling) { *err = "bsgpu_bam: a reservation is still open"; return false; }
		if (finishing) { *err = "bsgpu_bam: the stream was finished already"; return false; }
		if (end_of_stream) finishing = true; else cut_pending = true;
		try_submit();
		cv.notify_all();
		if (failed) { *err = errmsg; return false; }
		return true;
	}

	bool rewind(std::string *err) {
		std::unique_lock<std::mutex> lk(mu);
		if (failed) { *err = errmsg; return false; }
		if (!finished || !runs.empty() || !ready.empty()) { *err = "bsgpu_bam_rewind: the stream has not been finished and drained"; return false; }
		reap(lk);
		finishing = final_submitted = finished = false;
		for (Stage &st : stage) st.start = st.fill = st.head;
		return true;
	}

	// next result in stream order (NULL: none ready; *done says whether any can still come)
	bool drain(bool wait, SessResult **out, bool *done, std::string *err) {
		std::unique_lock<std::mutex> lk(mu);
		*out = nullptr;
		reap(lk);
		// something will come as long as a batch is with the worker or is about to be handed to it
		if (wait) cv.wait(lk, [&] { return !ready.empty() || failed || finished || !(!runs.empty() || finishing || cut_pending); });
		if (ready.empty()) {
			if (failed) { *err = errmsg; return false; }
			*done = finished;
			return true;
		}
		SessResult *r = ready.front();
		ready.pop_front();
		lent.push_back(r);
		*out = r;
		*done = r->last && ready.empty();
		try_submit();                         // a batch may have been waiting for room among the outstanding results
		cv.notify_all();
		return true;
	}

	bool give_back(uint64_t id) {
		std::unique_lock<std::mutex> lk(mu);
		for (size_t i = 0; i < lent.size(); i++) if (lent[i]->id == id) {
			pool.push_back(lent[i]);
			lent.erase(lent.begin() + i);
			return true;
		}
		return false;
	}

	void progress(bsgpu_bam_progress_t *out) {
		std::unique_lock<std::mutex> lk(mu);
		out->bytes_fed = bytes_fed; out->bytes_done = bytes_done; out->records_done = records_done; out->batches = batches;
		out->carry_bytes = carry_bytes; out->empty_batches = empty_batches;
		size_t pinned = stage[0].cap + stage[1].cap;
		for (SessResult *r : pool) pinned += r->cap;
		for (SessResult *r : lent) pinned += r->cap;
		for (SessResult *r : ready) pinned += r->cap;
		out->pinned_bytes = pinned;
	}

	void close() {
		{
			std::unique_lock<std::mutex> lk(mu);
			closing = true;
			cv.notify_all();
		}
		{                                         // batches in flight run to their end first
			std::unique_lock<std::mutex> lk(mu);
			cv.wait(lk, [&] { return runs.empty(); });
			reap(lk);
		}
		for (Stage &st : stage) if (st.p) { hk.release(st.p); st.p = nullptr; }
		for (auto *v : {&pool, &lent}) { for (SessResult *r : *v) { if (r->buf) hk.release(r->buf); delete r; } v->clear(); }
		for (SessResult *r : ready) { if (r->buf) hk.release(r->buf); delete r; }
		ready.clear();
	}

private:
	// a slice of the stream into the staging buffer: one thread streams ~10 GB/s, a feed of many MB is split
	static void copy(uint8_t *dst, const uint8_t *src, size_t n) {
		// (BSGPU_FEED_THREADS; default: a quarter of the cores, 2..8 -- 8 threads: 299 M sites/s on the genome leg, 4: 282 M)
		static const unsigned T = [] {
			const char *e = getenv("BSGPU_FEED_THREADS");
			const int v = e ? atoi(e) : 0, hw = (int)std::thread::hardware_concurrency();
			return (unsigned)std::max(1, std::min(v > 0 ? v : std::max(2, std::min(hw / 4, 8)), 16));
		}();
		if (n < (8u << 20) || T == 1) { memcpy(dst, src, n); return; }
		std::vector<std::thread> thr;
		for (unsigned t = 1; t < T; t++) thr.emplace_back([=] { const size_t lo = n * t / T, hi = n * (t + 1) / T; memcpy(dst + lo, src + lo, hi - lo); });
		memcpy(dst, src, n / T);
		for (auto &t : thr) t.join();
	}

	// staging buffer b must have `head` bytes of room in front of its fed bytes, and room for batch_bytes of those:
	// (re)allocates, keeping [start, fill).  Lock held, nobody writing into the buffer.
	bool stage_fit(int b, size_t head) {
		Stage &st = stage[b];
		const size_t fed = st.fill - st.head;
		if (st.p && head <= st.head && st.head + std::max(batch_bytes, fed) + 64 <= st.cap) return true;
		const size_t nhead = std::max(head + head / 4, st.head);
		const size_t ncap = nhead + std::max(batch_bytes, fed) + 64;
		uint8_t *np = hk.alloc(ncap);
		if (!np) return false;
		const size_t live = st.fill - st.start, front = st.head - st.start;
		if (live) memcpy(np + nhead - front, st.p + st.start, live);
		if (st.p) hk.release(st.p);
		st.p = np; st.cap = ncap;
		st.start = nhead - front; st.fill = st.start + live; st.head = nhead;
		return true;
	}

	// lock held: joins the threads of runs that have returned
	void reap(std::unique_lock<std::mutex> &lk) {
		std::vector<Run *> d;
		d.swap(dead);
		if (d.empty()) return;
		lk.unlock();
		for (Run *r : d) { if (r->th.joinable()) r->th.join(); delete r; }
		lk.lock();
	}

	// lock held: start a run over stage[w] when the buffer is full (or a cut / the end of the stream is marked), the run before
	// it has said how far it goes, and there is room among the outstanding results
	void try_submit() {
		if (scan_pending || final_submitted || failed || filling || closing) return;
		if (runs.size() >= 2 || runs.size() + ready.size() >= max_outstanding) return;
		Stage &st = stage[w];
		const bool full = st.fill - st.head >= batch_bytes;
		if (!full && !finishing && !cut_pending) return;
		const bool whole = finishing || cut_pending;      // the staged bytes end on a block boundary: nothing is carried over
		if (finishing) {
			if (st.fill == st.start) {               // nothing left to run: the stream is over once the runs in flight have returned
				if (!runs.empty()) return;
				final_submitted = true; finished = true; cut_pending = false;
				cv.notify_all();
				return;
			}
			final_submitted = true;
		} else if (cut_pending && st.fill == st.start) { cut_pending = false; cv.notify_all(); return; }
		cut_pending = false;
		Run *r = new Run();
		r->buf = w; r->final = finishing; r->whole = whole; r->seq = next_seq++;
		r->data = st.p + st.start; r->len = st.fill - st.start;
		runs.push_back(r);
		scan_pending = true;
		w ^= 1;
		Stage &nx = stage[w];
		nx.start = nx.fill = nx.head;
		r->th = std::thread([this, r] { run_batch(r); });
		cv.notify_all();
	}

	// lock held: stage[w] has room, or the call gives up (false) -- it never waits when `nowait`
	bool room(std::unique_lock<std::mutex> &lk, bool nowait) {
		for (;;) {
			if (failed) return false;
			if (cut_pending) try_submit();
			if (!cut_pending) {
				if (stage[w].fill - stage[w].head < batch_bytes) return true;
				try_submit();
				if (stage[w].fill - stage[w].head < batch_bytes) return true;
			}
			if (nowait) return false;
			cv.wait(lk);
		}
	}

	// the run says how far it goes: the carry moves in front of what the caller has fed into the other buffer meanwhile, and
	// the run's own staging buffer is free
	static void scanned_cb(void *sess, uint64_t seq, size_t consumed, size_t records) { ((Session *)sess)->on_scanned(seq, consumed, records); }
	void on_scanned(uint64_t seq, size_t consumed, size_t records) {
		std::unique_lock<std::mutex> lk(mu);
		Run *r = nullptr;
		for (Run *q : runs) if (q->seq == seq) r = q;
		if (!r || r->scanned) return;
		r->scanned = true;
		const size_t carry = r->len - consumed;
		Stage &nx = stage[r->buf ^ 1];
		if (carry && !failed) {
			// a caller that reserves again the moment it has committed (bulk input) would win the mutex every time: it yields
			// to waiting carries in reserve() / feed()
			carry_waiters++;
			cv.wait(lk, [&] { return !filling || closing; });
			carry_waiters--;
			if (carry > nx.start && !stage_fit(r->buf ^ 1, carry + (nx.head - nx.start))) { failed = true; errmsg = "bsgpu_bam: cannot allocate staging memory for the carried records"; }
			if (!failed) { memcpy(nx.p + nx.start - carry, r->data + consumed, carry); nx.start -= carry; }
		}
		if (r->whole && carry && !failed) { failed = true; errmsg = "bsgpu_bam: internal: a whole batch left records behind"; }
		bytes_done += consumed; records_done += records; batches++; carry_bytes += carry;
		if (!consumed) empty_batches++;
		scan_pending = false;
		try_submit();
		cv.notify_all();
	}

	void run_batch(Run *run) {
		if (hk.thread_init) hk.thread_init(hk.user);
		std::unique_lock<std::mutex> lk(mu);
		// a result buffer from the pool: the smallest that holds the guess, else the largest there is (it grows)
		const size_t guess = (size_t)((double)run->len * result_ratio) + 4096;
		SessResult *r = nullptr;
		{
			size_t pick = pool.size();
			for (size_t i = 0; i < pool.size(); i++) {
				if (pick == pool.size()) { pick = i; continue; }
				const size_t a = pool[i]->cap, c = pool[pick]->cap;
				if (c >= guess ? (a >= guess && a < c) : a > c) pick = i;
			}
			if (pick < pool.size()) { r = pool[pick]; pool.erase(pool.begin() + pick); }
		}
		lk.unlock();
		if (!r) r = new SessResult();
		std::string err;
		int rc = 0;
		if (!r->buf) {
			r->cap = guess;
			r->buf = hk.alloc(r->cap);
			if (!r->buf) { r->cap = 0; rc = -1; err = "bsgpu_bam: cannot allocate page-locked result memory"; }
		}
		size_t consumed = 0, records = 0;
		r->blocks.clear(); r->nbytes = r->nrec = 0;
		if (!rc) rc = hk.run(hk.user, run->data, run->len, run->whole, r, run->seq, scanned_cb, this, &consumed, &records, &err);
		if (!rc) on_scanned(run->seq, consumed, records);      // (no-op when the run has told already)
		lk.lock();
		// results leave in the order of the runs
		cv.wait(lk, [&] { return next_result_seq == run->seq; });
		if (rc) {
			if (!failed) { failed = true; errmsg = err; }
			pool.push_back(r);
			if (!run->scanned) { run->scanned = true; scan_pending = false; }
		} else {
			r->bytes_in = consumed; r->records_in = records; r->id = next_id++; r->last = run->final;
			if (consumed || run->whole) ready.push_back(r); else pool.push_back(r);      // no certain start in the batch: nothing to show yet
		}
		next_result_seq++;
		for (size_t i = 0; i < runs.size(); i++) if (runs[i] == run) { runs.erase(runs.begin() + i); break; }
		dead.push_back(run);
		if (run->final) finished = true;
		try_submit();
		cv.notify_all();
	}
};

}  // namespace bsgpu
