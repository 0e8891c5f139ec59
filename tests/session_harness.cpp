// session_harness.cpp -- the streaming session's host logic (bs_call_b200/csrc/bsgpu_session.h) on the CPU, with a
// stand-in for the device run of a batch.  Built and run by tests/test_session_host.py (g++ -pthread).
//
// Stream: records of [u32 length][payload]; payload[0] == 1 marks a record at which a block is certain to start (what
// read_input's state reset is for the real stream).  The stand-in runner "processes" a batch by copying the bytes it
// accounts for into the result buffer: everything up to the last marked record (or everything, for a whole batch).  So the
// results of a session, concatenated in order, must be the stream itself -- whatever the batch size, the slicing, the
// number of threads, and wherever cuts are placed.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <unistd.h>

#include "bsgpu_session.h"

using namespace bsgpu;

static std::atomic<long> g_live_allocs{0};
static uint8_t *h_alloc(size_t n) { g_live_allocs++; return (uint8_t *)malloc(n); }
static void h_free(uint8_t *p) { g_live_allocs--; free(p); }

struct Runner {
	Session *s = nullptr;
	std::mt19937 rng{7};
	int max_sleep_us = 0;
};

static std::mutex g_run_mu;            // the stand-in serialises its "window stages" like the real runner does
static int fake_run(void *user, const uint8_t *data, size_t len, bool whole, SessResult *r, uint64_t seq,
		void (*scanned)(void *sess, uint64_t seq, size_t consumed, size_t records), void *sess,
		size_t *consumed, size_t *records, std::string *err) {
	Runner *rn = (Runner *)user;
	size_t at = 0, nrec = 0, last_start = 0, rec_before_start = 0;
	while (at + 4 <= len) {
		uint32_t l;
		memcpy(&l, data + at, 4);
		if (l < 1 || l > (1u << 20)) { *err = "malformed record"; return -1; }
		if (at + 4 + l > len) break;
		if (at && data[at + 4] == 1) { last_start = at; rec_before_start = nrec; }
		nrec++;
		at += 4 + (size_t)l;
	}
	if (whole && at != len) { *err = "the record stream ends inside a record"; return -1; }
	const size_t take = whole ? len : last_start;
	const size_t nr = whole ? nrec : rec_before_start;
	if (take > r->cap) {
		Session::GrowCtx g{rn->s, r};
		size_t ncap;
		if (!Session::grow(&g, 0, take, &ncap)) { *err = "grow failed"; return -1; }
	}
	memcpy(r->buf, data, take);           // (the run reads the staged bytes BEFORE it tells how far it goes)
	r->nbytes = take; r->nrec = nr;
	*consumed = take; *records = nr;
	unsigned coin;
	{ std::lock_guard<std::mutex> lk(g_run_mu); coin = rn->rng(); }
	if (coin % 3) scanned(sess, seq, take, nr);        // two runs in three tell early; the others leave it to their return
	if (rn->max_sleep_us) usleep(coin % rn->max_sleep_us);
	return 0;
}

static std::vector<uint8_t> make_stream(std::mt19937 &rng, size_t nrec, double p_start, std::vector<size_t> *starts) {
	std::vector<uint8_t> v;
	for (size_t i = 0; i < nrec; i++) {
		const uint32_t l = 1 + rng() % 300;
		const bool st = i == 0 || (rng() % 10000) < p_start * 10000;
		if (st && starts) starts->push_back(v.size());
		const size_t at = v.size();
		v.resize(at + 4 + l);
		memcpy(&v[at], &l, 4);
		v[at + 4] = st ? 1 : 0;
		for (uint32_t k = 1; k < l; k++) v[at + 4 + k] = (uint8_t)rng();
	}
	return v;
}

#define CHECK(c, ...) do { if (!(c)) { fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); exit(1); } } while (0)

struct Collected { std::vector<uint8_t> bytes; std::vector<size_t> ends; size_t nrec = 0; bool done = false; };

static void take(Session &s, SessResult *r, Collected &c) {
	c.bytes.insert(c.bytes.end(), r->buf, r->buf + r->nbytes);
	c.ends.push_back(c.bytes.size());
	c.nrec += r->nrec;
	CHECK(r->bytes_in == r->nbytes, "bytes_in %zu != nbytes %zu", r->bytes_in, r->nbytes);
	CHECK(s.give_back(r->id), "release");
}

// one thread: non-blocking feeds, draining when a feed comes up short; cuts after the offsets in `cuts`
static void drive_single(Session &s, const std::vector<uint8_t> &v, std::mt19937 &rng, size_t max_slice, const std::vector<size_t> &cuts, Collected &c, bool use_reserve) {
	std::string err;
	size_t at = 0, ci = 0;
	while (at < v.size()) {
		size_t lim = v.size();
		if (ci < cuts.size()) lim = cuts[ci];
		if (at == lim) { CHECK(s.mark(false, &err), "cut: %s", err.c_str()); ci++; continue; }
		const size_t m = std::min(lim - at, (size_t)(1 + rng() % max_slice));
		size_t took = 0;
		if (use_reserve) {
			uint8_t *p; size_t avail;
			CHECK(s.reserve(&p, &avail, false, &err), "reserve: %s", err.c_str());
			took = std::min(avail, m);
			if (avail) { memcpy(p, v.data() + at, took); CHECK(s.commit(took, &err), "commit: %s", err.c_str()); }
		} else CHECK(s.feed(v.data() + at, m, true, &took, &err), "feed: %s", err.c_str());
		at += took;
		if (took < m) {
			SessResult *r; bool done;
			CHECK(s.drain(true, &r, &done, &err), "drain: %s", err.c_str());
			if (r) take(s, r, c);
		}
	}
	CHECK(s.mark(true, &err), "finish: %s", err.c_str());
	for (;;) {
		SessResult *r; bool done;
		CHECK(s.drain(true, &r, &done, &err), "drain: %s", err.c_str());
		if (r) take(s, r, c);
		if (done) break;
	}
	c.done = true;
}

// two threads as in the reference: this one feeds (blocking), a second one drains
static void drive_threads(Session &s, const std::vector<uint8_t> &v, std::mt19937 &rng, size_t max_slice, const std::vector<size_t> &cuts, Collected &c) {
	std::thread printer([&] {
		std::string err;
		for (;;) {
			SessResult *r; bool done;
			CHECK(s.drain(true, &r, &done, &err), "drain: %s", err.c_str());
			if (r) take(s, r, c);
			else if (!done) usleep(200);
			if (done) break;
		}
		c.done = true;
	});
	std::string err;
	size_t at = 0, ci = 0;
	while (at < v.size()) {
		size_t lim = v.size();
		if (ci < cuts.size()) lim = cuts[ci];
		if (at == lim) { CHECK(s.mark(false, &err), "cut: %s", err.c_str()); ci++; continue; }
		const size_t m = std::min(lim - at, (size_t)(1 + rng() % max_slice));
		size_t took = 0;
		CHECK(s.feed(v.data() + at, m, false, &took, &err) && took == m, "feed: %s", err.c_str());
		at += m;
	}
	CHECK(s.mark(true, &err), "finish: %s", err.c_str());
	printer.join();
}

// the bulk-input pattern of bsgpu_seam_reader.c: the feeder reserves (blocking), fills a slice in place, commits and reserves again
// at once -- without a pause in which a run waiting to move its carry could get the session's mutex (reserve() yields to it)
static void drive_threads_reserve(Session &s, const std::vector<uint8_t> &v, std::mt19937 &rng, size_t max_slice, const std::vector<size_t> &cuts, Collected &c) {
	std::thread printer([&] {
		std::string err;
		for (;;) {
			SessResult *r; bool done;
			CHECK(s.drain(true, &r, &done, &err), "drain: %s", err.c_str());
			if (r) take(s, r, c);
			else if (!done) usleep(200);
			if (done) break;
		}
		c.done = true;
	});
	std::string err;
	size_t at = 0, ci = 0;
	while (at < v.size()) {
		size_t lim = v.size();
		if (ci < cuts.size()) lim = cuts[ci];
		if (at == lim) { CHECK(s.mark(false, &err), "cut: %s", err.c_str()); ci++; continue; }
		uint8_t *p = nullptr;
		size_t avail = 0;
		CHECK(s.reserve(&p, &avail, true, &err), "reserve: %s", err.c_str());
		CHECK(p != nullptr && avail > 0, "a blocking reserve came back without room");
		const size_t m = std::min(std::min(lim - at, avail), (size_t)(1 + rng() % max_slice));
		memcpy(p, v.data() + at, m);
		CHECK(s.commit(m, &err), "commit: %s", err.c_str());
		at += m;
	}
	CHECK(s.mark(true, &err), "finish: %s", err.c_str());
	printer.join();
}

int main(int argc, char **argv) {
	const int rounds = argc > 1 ? atoi(argv[1]) : 60;
	alarm(240);                               // a deadlock ends the test instead of hanging it
	std::mt19937 rng(12345);
	for (int it = 0; it < rounds; it++) {
		std::vector<size_t> starts;
		const size_t nrec = 50 + rng() % 4000;
		const double p_start = (it % 5 == 0) ? 0.0 : (it % 5 == 1) ? 0.001 : 0.02 + (rng() % 100) / 500.0;
		std::vector<uint8_t> v = make_stream(rng, nrec, p_start, &starts);
		// cuts: a few of the certain starts (a cut must be where a block ends)
		std::vector<size_t> cuts;
		if (it % 3 == 1) for (size_t k = 1; k < starts.size(); k++) if (rng() % 7 == 0) cuts.push_back(starts[k]);
		const size_t batch = (it % 4 == 0) ? 4096 : 4096 + rng() % 60000;
		Runner rn;
		Session s;
		rn.s = &s; rn.max_sleep_us = it % 2 ? 300 : 0;
		SessHooks hk;
		hk.user = &rn; hk.alloc = h_alloc; hk.release = h_free; hk.run = fake_run;
		CHECK(s.open(hk, batch, it % 2 ? 0.01 : 1.0), "open");
		Collected c;
		const size_t max_slice = (it % 3 == 0) ? 700 : (it % 3 == 1) ? 20000 : 1000000;
		if (it % 4 == 3) drive_threads_reserve(s, v, rng, max_slice, cuts, c);
		else if (it % 2) drive_threads(s, v, rng, max_slice, cuts, c);
		else drive_single(s, v, rng, max_slice, cuts, c, it % 4 == 2);
		CHECK(c.bytes.size() == v.size(), "round %d: %zu bytes back, %zu fed", it, c.bytes.size(), v.size());
		CHECK(!memcmp(c.bytes.data(), v.data(), v.size()), "round %d: results differ from the stream", it);
		CHECK(c.nrec == nrec, "round %d: %zu records, want %zu", it, c.nrec, nrec);
		for (size_t cut : cuts) CHECK(std::find(c.ends.begin(), c.ends.end(), cut) != c.ends.end(), "round %d: a batch of results spans the cut at %zu", it, cut);
		bsgpu_bam_progress_t pr;
		s.progress(&pr);
		CHECK(pr.bytes_fed == v.size() && pr.bytes_done == v.size() && pr.records_done == nrec, "round %d: progress", it);
		if (it % 6 == 0) {                  // a second stream through the same session
			std::string err;
			CHECK(s.rewind(&err), "rewind: %s", err.c_str());
			Collected c2;
			std::vector<uint8_t> v2 = make_stream(rng, 300, 0.05, nullptr);
			drive_single(s, v2, rng, 5000, {}, c2, false);
			CHECK(c2.bytes == v2, "round %d: second stream differs", it);
		}
		s.close();
		CHECK(g_live_allocs.load() == 0, "round %d: %ld buffers leaked", it, g_live_allocs.load());
	}
	// a stream that ends inside a record is refused at the end
	{
		std::vector<uint8_t> v = make_stream(rng, 100, 0.05, nullptr);
		v.resize(v.size() - 3);
		Runner rn; Session s; rn.s = &s;
		SessHooks hk; hk.user = &rn; hk.alloc = h_alloc; hk.release = h_free; hk.run = fake_run;
		CHECK(s.open(hk, 1 << 20, 1.0), "open");
		std::string err; size_t took;
		CHECK(s.feed(v.data(), v.size(), true, &took, &err) && took == v.size(), "feed");
		CHECK(s.mark(true, &err), "finish");
		bool failed = false;
		for (int k = 0; k < 100 && !failed; k++) { SessResult *r; bool done; if (!s.drain(true, &r, &done, &err)) failed = true; else if (r) s.give_back(r->id); else if (done) break; }
		CHECK(failed && err.find("inside a record") != std::string::npos, "truncated stream was not refused (%s)", err.c_str());
		s.close();
	}
	// feeding after the end, committing without a reservation
	{
		Runner rn; Session s; rn.s = &s;
		SessHooks hk; hk.user = &rn; hk.alloc = h_alloc; hk.release = h_free; hk.run = fake_run;
		CHECK(s.open(hk, 4096, 1.0), "open");
		std::string err; size_t took; uint8_t b[8] = {4, 0, 0, 0, 1, 2, 3, 4};
		CHECK(!s.commit(1, &err), "commit without reserve");
		CHECK(s.mark(true, &err), "finish");
		CHECK(!s.feed(b, 8, true, &took, &err), "feed after finish");
		SessResult *r; bool done;
		CHECK(s.drain(true, &r, &done, &err) && !r && done, "empty stream");
		s.close();
	}
	printf("session harness ok: %d rounds\n", rounds);
	return 0;
}
