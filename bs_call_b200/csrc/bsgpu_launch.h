// bsgpu_launch.h -- launcher prototypes shared between bsgpu_kernels.cu and bsgpu_api.cu (internal).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "bsgpu.h"

namespace bsgpu {

struct DevConst;

// device-side view of bsgpu_seg (include/bsgpu.h)
struct Seg { uint32_t pos, off; uint16_t len; uint8_t mapq, flags; uint32_t pad; };
static_assert(sizeof(Seg) == 16, "segment record is 16 bytes");

cudaError_t configure_kernels();

// guard_base: id of site 0 of this launch in the guard list (bsgpu_guard_read)
cudaError_t launch_call_sites(const void *pileup, const void *ref, size_t n, void *out, void *skip, bool vcf,
		const DevConst *dc, unsigned long long *counters, cudaStream_t stream, int *launches, unsigned long long guard_base = 0,
		bool overlap_safe = false);      // overlap_safe: other kernels may run next to this one (see the launcher)

// n records of gt_meth (+ skip[]) or gt_vcf (skip == NULL) -> n wire records + the chunk's flag word (bsgpu_wire.h)
cudaError_t launch_wire_pack(const void *rec, const void *skip, size_t n, void *wire, cudaStream_t stream, int *launches);

// scratch needed by the segment binning of one block
size_t pileup_scratch_bytes(size_t nseg, uint32_t sz);

// bins the segments of a block by 128-site tile (count, scan, scatter) into `scratch`
cudaError_t launch_bin_segments(const void *segs, size_t nseg, uint32_t x, uint32_t sz, void *scratch,
		cudaStream_t stream, int *launches, unsigned long long *counters = nullptr);

// pileup (mode 0 -> pileup[]) or fused pileup + model (mode 1 -> gt_vcf[]) for tiles [tile0, tile0 + ntiles) of a
// block previously binned into `scratch`; `out` points at the record of site tile0 * kPileTileSites
cudaError_t launch_pileup_tiles(const void *scratch, size_t nseg, const void *bases, const void *ref, uint32_t x,
		uint32_t sz, uint32_t tile0, uint32_t ntiles, void *out, int mode, const DevConst *dc,
		unsigned long long *counters, cudaStream_t stream, int *launches);

cudaError_t launch_synth_sites(uint64_t seed, uint64_t first, size_t n, double mean_depth, void *pileup, void *ref,
		cudaStream_t stream, int *launches);

size_t synth_block_nseg(uint32_t sz, uint32_t read_len, double depth);
cudaError_t launch_synth_block(uint64_t seed, uint32_t x, uint32_t sz, uint32_t read_len, double depth,
		void *segs, void *bases, void *ref, cudaStream_t stream, int *launches);

size_t synth_bam_bytes(size_t ntemplates, uint32_t read_len);
cudaError_t launch_synth_bam(uint64_t seed, size_t ntemplates, uint32_t read_len, const void *pos_f, const void *pos_r, const void *src,
		const void *rank, void *out, cudaStream_t stream, int *launches);
cudaError_t launch_synth_ref(uint64_t seed, uint32_t x, uint32_t sz, void *ref, cudaStream_t stream, int *launches);

// device-resident totals of the --report-file side channels (bsgpu_profile) and what a launch needs to add to them
struct ProfDev {
	unsigned long long conv[BSGPU_PROFILE_MAX][4];
	unsigned long long base_filter[5];
	unsigned long long reads, read_bases, too_long;
	uint32_t used[2];                // running `used` of the profile vector, double buffered by launch parity
};
struct ProfArgs {
	const uint8_t *ref;              // codes of the block window, index 0 = position x
	uint32_t refn;                   // codes available (window size + 1)
	uint32_t min_qual;
	ProfDev *prof;
	uint16_t *used16;                // per template: max_pos + 1
	uint8_t *cand;                   // per template: 0, or 1 + counter of the byte that lands on entry `used`
	uint32_t *chunkmax;              // maximum of used16 over every kProfChunk templates
};
constexpr uint32_t kProfChunk = 4096;

// raw templates -> reads in reference coordinates + segment records (bsgpu_normalise.cu); prof != NULL: also the
// profile (two more launches); out_off == NULL: every mate owns `slot` bytes of obases
cudaError_t launch_normalise(const void *tmpl, size_t n, const void *bases, void *ev_work, const void *out_off, void *obases,
		void *segs, uint32_t segs_per_mate, uint32_t x, uint32_t y, const uint32_t left_trim[2], const uint32_t right_trim[2],
		unsigned long long *counters, const ProfArgs *prof, int parity, cudaStream_t stream, int *launches, uint32_t slot = 0);

// writer side (bsgpu_writer.cu): gt_vcf[] of a window -> BCF records
// device view of a contig's dbSNP entries: one bit per position (bit pos & 63 of word pos >> 6) for "known" and for "always
// written", the number of entries before each word, and the ID bytes of entry k at names[off[k] .. off[k + 1])
struct DbView {
	const unsigned long long *mask = nullptr, *fq = nullptr;
	const uint32_t *cum = nullptr, *off = nullptr;
	const uint8_t *names = nullptr;
	uint32_t words = 0;
};

struct BcfJob {
	const void *d_vcf;               // gt_vcf[sz]
	const void *d_ref;               // sz + 2 codes, index 0 = position x
	uint32_t x, sz;
	const void *d_blocks;            // (first, last) site index of every block, ascending; NULL = the window is one block
	uint32_t nblocks;
	bsgpu_bcf_params p;
	const DevConst *dc;
	void *site_scratch;              // bcf_site_scratch_bytes(sz)
	DbView db;                       // dbSNP entries of the window's contig
	uint32_t reg_start = 0, reg_stop = 0;    // ctg->curr_reg (0, 0: none -> the contig end clips)
	unsigned long long *guard = nullptr;      // the context's counters: sites whose QUAL / FS sit inside their guard band are counted and listed
	// --report-file statistics of the sites (bsgpu_site_stats_enable): device image of bsgpu_site_stats, the per-contig table
	// (entry j.p.rid is used when it is below n_ctg) and the GC bins of the contig
	void *stats = nullptr, *ctg_stats = nullptr;
	uint32_t *stats_carry = nullptr;         // chunked launches over one window: what the last site of the chunk before left (see k_bcf_stats)
	uint32_t stats_carry_flip = 0;           // ... alternating between the two carry words: chunk number & 1
	uint32_t n_ctg = 0;
	const uint8_t *gc = nullptr;
	uint32_t gc_bins = 0, gc_start = 1;
};
size_t bcf_site_scratch_bytes(uint32_t sz);
size_t bcf_cta_scratch_bytes(uint32_t cnt);
cudaError_t configure_writer();
cudaError_t launch_bcf_calls(const BcfJob &j, uint32_t i0, uint32_t cnt, cudaStream_t stream, int *launches);
cudaError_t launch_bcf_records(const BcfJob &j, uint32_t i0, uint32_t cnt, void *cta_scratch, void *d_out, size_t out_cap,
		unsigned long long *d_totals, cudaStream_t stream, int *launches);

// reader side (bsgpu_reader.cu)
cudaError_t launch_decode_records(const void *bam, const void *rec_off, const void *read_off, const void *mm_off, size_t nrec,
		uint32_t mapq_thresh, uint32_t max_tlen, int keep_unmatched, int ignore_dup, void *out, void *bases, void *misms,
		cudaStream_t stream, int *launches, void *keys = nullptr, void *name_table = nullptr, size_t name_slots = 0, uint32_t rec_base = 0,
		void *name_overflow = nullptr);
// QNAME join (bsgpu_reader.cu): the table k_decode_records fills, and the kernel that turns it into per-record name ids
size_t name_table_slots(size_t nrec);
size_t name_table_bytes(size_t nrec);
cudaError_t launch_name_ids(const void *bam, const void *rec_off, const void *rec, uint32_t r0, uint32_t r1, const void *table, size_t slots,
		void *name_id, cudaStream_t stream, int *launches);

// certain block starts of a chunk of records as a bit mask, from the keys launch_decode_records wrote (bsgpu_reader.cu)
cudaError_t launch_certain_starts(const void *keys, uint32_t n, void *carry, void *mask, void *scratch, cudaStream_t stream, int *launches);
size_t certain_scratch_bytes(size_t n);       // scratch of launch_certain_starts for n records

constexpr int kPileTileSites = 128;      // sites per tile of the gather kernel (= its CTA size)
constexpr int kMaxSegLen = 256;          // BSGPU_MAX_SEG_LEN

}  // namespace bsgpu
