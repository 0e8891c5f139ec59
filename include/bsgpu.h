/*
 * bsgpu.h -- C ABI of libbsgpu: the B200 (sm_100a) pileup + bisulfite genotype-likelihood path of bs_call.
 *
 * Plain pointers and sizes only.  Every entry point returns BSGPU_OK (1) or BSGPU_FAIL (-1), the values of the
 * reference's gt_status (gt/include/gt_error.h:23-24); bsgpu_last_error() gives the reason.  There is no CPU
 * fallback: if no sm_100 device can be opened bsgpu_init fails.
 *
 * Record layouts are the reference's own (include/bs_call.h) so that host code can hand its arrays over
 * unchanged:
 *     bsgpu_pileup   = pileup    include/bs_call.h:174-182   (104 B)
 *     bsgpu_gt_meth  = gt_meth   include/bs_call.h:152-160   (200 B)
 *     bsgpu_gt_vcf   = gt_vcf    include/bs_call.h:162-166   (208 B)
 *
 * Which reference interface each entry point replaces:
 *     bsgpu_init / bsgpu_destroy     init_calc_threads / join_calc_threads     src/call_genotypes.c:124-153
 *                                    + fill_base_prob_table src/genotype_model.c:10, lfact_store_init src/stats_utils.c:14
 *     bsgpu_call_sites[_dev]         the per-site body of call_thread           src/call_genotypes.c:43-115
 *                                    (summarise, calc_gt_prob src/genotype_model.c:44, fisher src/stats_utils.c:25)
 *     bsgpu_pileup_block[_dev]       the pileup loop of call_genotypes_ML       src/call_genotypes.c:172-226
 *     bsgpu_call_block[_dev]         call_genotypes_ML + call_thread, fused     src/call_genotypes.c:155-272, 21-122
 *     bsgpu_stage_templates          the (position, read bytes, mapq, strand) walk at the top of that loop,
 *                                    src/call_genotypes.c:181-212, turned into sorted segments
 *     bsgpu_process_block            process_template_vector                    src/process_template.c:18-126
 *                                    (trim_read, trim_soft_clips, handle_overlap, indel normalisation on device)
 *     bsgpu_decode_records           get_next_align_details                     src/input_sam.c:222-312
 *                                    (get_seq_and_qual :61, get_bam_misms :90, get_bs_strand :144, the filters :234-300)
 *     bsgpu_build_blocks             read_input                                 src/get_template_vector.c:49-389
 *     bsgpu_call_bam                 read_input -> process_thread -> call_genotypes_ML chained (src/process.c:43-72, 172)
 *     bsgpu_bcf_block[_dev], bsgpu_call_block_bcf, bsgpu_call_sites_bcf, bsgpu_call_bam_bcf
 *                                    print_vcf_entry / flush_vcf_entries / _print_vcf_entry as print_thread drives them
 *                                    (src/print_vcf.c:32-381, 535-594; src/process.c:89-104): the BCF records themselves
 *     bsgpu_bam_open / _feed / _reserve / _commit / _finish / _drain / _release / _close
 *                                    the same chain with read_input's streaming behaviour: one sam_read1() at a time in,
 *                                    blocks out as they close, O(block) memory (src/get_template_vector.c:86-110, 140-189)
 *     bsgpu_profile_enable / _read   meth_profile + mprof_thread (src/meth_profile.c:48-76, src/process.c:20-41) and the
 *                                    bs_stats tallies of process_template_vector / read_input (--report-file)
 * The link-compatible replacements for the three reference symbols themselves (call_genotypes_ML,
 * init_calc_threads, join_calc_threads; include/bs_call.h:358-360) are in bs_call_b200/csrc/bsgpu_dropin.c and
 * are built on top of this ABI; see INTEGRATION.md.
 */
#ifndef BSGPU_H
#define BSGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSGPU_OK 1
#define BSGPU_FAIL (-1)

#define BSGPU_MAX_QUAL 43      /* include/bs_call.h:25 */
#define BSGPU_FLT_QUAL 63      /* include/bs_call.h:28 */

typedef struct {
	uint32_t counts[2][8];     /* [orientation][class]; classes 0-3 non-informative ACGT, 4-7 informative ACGT */
	uint32_t n;
	float quality[8];          /* sum of base qualities per class (integer valued) */
	float mapq2;               /* sum of MAPQ^2 */
} bsgpu_pileup;

typedef struct {
	uint64_t counts[8];
	int32_t qual[8];
	double gt_prob[10];        /* log10 posteriors: AA AC AG AT CC CG CT GG GT TT */
	double fisher_strand;
	int32_t mq;
	int32_t aq;
	uint8_t max_gt;
	uint8_t pad_[7];
} bsgpu_gt_meth;

typedef struct {
	bsgpu_gt_meth gtm;
	uint8_t ready;             /* C99 bool in the reference */
	uint8_t skip;
	uint8_t pad_[6];
} bsgpu_gt_vcf;

/* One mate of one template after normalisation: `len` packed bytes (base | qual<<2, src/input_sam.c:76-86),
 * one per reference position starting at `pos` (1-based).  flags: bit0 = strand index `ori` this mate is
 * counted under (src/call_genotypes.c:185,221,224), bits1-2 = bisulfite strand (0 none, 1 C2T, 2 G2A). */
typedef struct {
	uint32_t pos;
	uint32_t off;              /* offset of the first byte in bases[] */
	uint16_t len;              /* <= BSGPU_MAX_SEG_LEN; longer mates are split by the stager */
	uint8_t mapq;
	uint8_t flags;
	uint32_t pad_;
} bsgpu_seg;                   /* 16 bytes */

#define BSGPU_MAX_SEG_LEN 256

/* Flat template record for the raw-template entry points (what the reference keeps in align_details,
 * include/bs_call.h:64-73, with the two gt_vectors flattened into offset/length pairs). */
typedef struct {
	uint32_t forward_position, reverse_position;
	uint32_t reference_span[2];
	uint32_t read_off[2];      /* into bases[] */
	uint32_t read_len[2];
	uint32_t mm_off[2];        /* into misms[] */
	uint32_t mm_n[2];
	uint8_t present[2];        /* read[k] != NULL */
	uint8_t mapq[2];
	uint8_t orientation;       /* 0 FORWARD, 1 REVERSE */
	uint8_t bs_strand;         /* 0 NON_CONVERTED, 1 STRAND_C2T, 2 STRAND_G2A */
	uint8_t pad_[2];
} bsgpu_template;              /* 56 bytes */

/* gt_misms flattened (include/bs_call.h:53-62); type codes are the reference's enum: 1 INS, 2 DEL, 3 SOFT */
typedef struct { uint32_t type, position, size; } bsgpu_misms;

typedef struct {
	double under_conv;         /* -c, default 0.01  (include/bs_call.h:16) */
	double over_conv;          /*     default 0.05 */
	double ref_bias;           /*     default 2 */
	uint32_t left_trim[2];     /* -L */
	uint32_t right_trim[2];    /* -R */
	uint8_t min_qual;          /* -Q, default 20 */
	int32_t device;            /* CUDA device ordinal */
} bsgpu_params;

/* ---- reader side ---- */
/* What get_next_align_details() yields for one BAM alignment record (src/input_sam.c:222-312): the filter verdict,
 * positions and flags always; the decoded read (read_off / read_len into the packed-base array), CIGAR events
 * (mm_off / mm_n), reference span and bisulfite strand only for records that are kept (ret == 0). */
typedef struct {
	int32_t ret;               /* 0 keep, 1 dropped by the filters */
	uint32_t filtered;         /* gt_filter_reason (include/bs_call.h:50) */
	uint32_t forward_position, reverse_position;
	uint32_t alignment_flag;   /* BAM flag, BAM_FPAIRED cleared when the mate cannot be used */
	uint32_t align_length;     /* read length the CIGAR implies */
	uint32_t reference_span;
	uint32_t read_off, read_len;
	uint32_t mm_off, mm_n;
	int32_t tid;
	uint8_t reverse, orientation, bs_strand, mapq;
	uint8_t q01[2];            /* qualities of the first two packed bytes (what get_al_qual reads, src/al_utils.c:26) */
	uint8_t pad_[2];
} bsgpu_record;                /* 56 bytes */

/* One block as read_input() hands it to process_template_vector() (src/get_template_vector.c:170-189):
 * templates [first_template, first_template + n_templates), window [x, y] on contig tid;
 * vcf_off = index of the record of position x in the gt_vcf[] array bsgpu_call_bam returns. */
typedef struct {
	uint32_t tid, x, y, first_template, n_templates, pad_;
	uint64_t vcf_off;
} bsgpu_block;                 /* 32 bytes */

typedef struct {
	uint32_t max_template_len; /* -l, default 1000 (include/bs_call.h:20) */
	uint8_t mapq_thresh;       /* -q, default 20 */
	uint8_t keep_unmatched;    /* -k */
	uint8_t ignore_duplicates; /* ignore the BAM duplicate flag */
	uint8_t keep_duplicates;   /* -d: no positional duplicate removal */
} bsgpu_reader_params;

/* The read-level side channels of --report-file that the reference gathers inside the replaced functions (bs_stats,
 * include/bs_call.h:124-146): the non-CpG conversion profile of meth_profile() (src/meth_profile.c:48-76), the base /
 * read tallies of process_template_vector() and its helpers (src/process_template.c:52-63, src/al_utils.c:141,150,308)
 * and the per-reason tallies of read_input() (src/get_template_vector.c:104-107,243-246,314-319,361-364). */
#define BSGPU_PROFILE_MAX 1024
typedef struct {
	uint64_t conv_cts[BSGPU_PROFILE_MAX][4]; /* stats->meth_profile: entry i holds meth_cts of original read position i - 1 */
	uint32_t used;                           /* gt_vector_get_used(stats->meth_profile); entries >= used are zero */
	uint32_t pad_;
	uint64_t base_filter[5];                 /* stats->base_filter[base_none, base_trim, base_clip, base_overlap, base_lowqual] */
	uint64_t filter_cts[15];                 /* stats->filter_cts / filter_bases, indexed by gt_filter_reason (include/bs_call.h:50); */
	uint64_t filter_bases[15];               /* [0] = mates / bases that reached normalisation (+ the bases of src/get_template_vector.c:363) */
} bsgpu_profile;

/* The site-level side channels of --report-file: what _print_vcf_entry() adds to bs_stats for every site it visits
 * (src/print_vcf.c:382-526; bs_stats, include/bs_call.h:124-146).  Flat image of the fields it touches: the fs / qd / mq
 * vectors (add_flt_counts, :22-27) and the coverage hash (gt_cov_stats, include/bs_call.h:87-95) are arrays indexed by value;
 * values beyond an array are counted in fs_overflow / cov_overflow.  The writer entry points (bsgpu_*_bcf, sessions that return
 * records) gather it on the device while bsgpu_site_stats_enable is on; a host that replaces the print thread (seam D) folds
 * it into stats the way bsgpu_seam_reader.c does.  All counters are exact; the two methylation posteriors are sums of
 * doubles accumulated in device order (equal to the reference's to ~1e-12 relative).  `multi` holds the written sites whose
 * call is homozygous reference, `snps` every other written site: which of the two the reference's own counters `snps` / `multi`
 * receive depends on the byte its test at :400-402 finds behind a string literal's terminator -- undefined behaviour that
 * comes out differently in different links of the same sources (hom-ref -> multi in oracle/_ref/libbsref.so, everything ->
 * snps in oracle/_ref/bs_call); the host folds accordingly (bsgpu_seam_reader.c: homref_is_multi). */
#define BSGPU_STATS_FS_MAX 4096
#define BSGPU_STATS_COV_MAX 4096
typedef struct { uint64_t var, CpG[2], CpG_inf[2], all, gc_pcent[101]; } bsgpu_cov_stats;
typedef struct {
	uint64_t snps[2], multi[2], dbSNP_sites[2], dbSNP_var[2], CpG_ref[2], CpG_nonref[2];      /* [stats_all, stats_passed] */
	uint64_t mut_counts[12][2], dbSNP_mut_counts[12][2];    /* stats_mut order: AC AG AT CA CG CT GA GC GT TA TC TG */
	uint64_t qual[4][256];                                  /* all_sites, variant_sites, CpG_ref_sites, CpG_nonref_sites by QUAL */
	uint64_t filter_counts[2][32];                          /* [het call][filter bits q20 qd2 fs60 mq40] */
	double CpG_ref_meth[2][101], CpG_nonref_meth[2][101];   /* summed posterior of the methylation level, 1 % steps */
	uint64_t qd_stats[256][2], mq_stats[256][2], fs_stats[BSGPU_STATS_FS_MAX][2];      /* [value][het call] */
	uint64_t fs_overflow, cov_overflow;
	bsgpu_cov_stats cov[BSGPU_STATS_COV_MAX];               /* by depth (CpG_inf: by informative depth) */
} bsgpu_site_stats;
/* the per-contig copy of the first six counters (gt_ctg_stats, include/bs_call.h:75-85) */
typedef struct { uint64_t snps[2], multi[2], dbSNP_sites[2], dbSNP_var[2], CpG_ref[2], CpG_nonref[2]; } bsgpu_ctg_site_stats;

/* ---- writer side: what the print thread needs besides the records to serialise a site (src/print_vcf.c) ---- */
#define BSGPU_BCF_MAX_RECORD 384
#define BSGPU_DBSNP_MAX_ID 96
/* The dbSNP entries of one contig as the writer sees them (-D; src/print_vcf.c:133 calls dbSNP_lookup_name(), src/dbSNP.c:305,
 * once per site).  The index file and its reader stay with the host; the host hands over the ANSWERS: positions (1-based,
 * ascending, unique), flags[k] = the return value (1 known, 3 known and "always written": the site gets a record even when it
 * is homozygous reference A / T, src/print_vcf.c:139), and the ID bytes exactly as lookup leaves them in rs[0 .. rs_len)
 * (at most BSGPU_DBSNP_MAX_ID each): names[name_off[k] .. name_off[k + 1]).  name_off has n + 1 entries. */
typedef struct {
	uint32_t n;
	const uint32_t *pos;
	const uint8_t *flags;
	const uint32_t *name_off;
	const uint8_t *names;
} bsgpu_dbsnp;
typedef struct {
	int32_t ids[16];           /* BCF header dictionary ids of PASS fail mac1 CX GT FT GL GQ DP MQ QD MC8 AMQ CS CG FS: work.vcf_ids
	                              as print_vcf_header() fills it (include/bs_call.h:192-207, src/print_vcf.c:751-764) */
	int32_t rid;               /* ctg->vcf_rid */
	uint32_t ctg_end;          /* ctg->end_pos: sites beyond it are not written (src/print_vcf.c:159) */
	uint8_t all_positions;     /* -A: also write homozygous-reference A / T sites */
	uint8_t pad_[3];
	uint32_t reg_start, reg_stop;  /* ctg->curr_reg (-C / region list): only sites inside are written, and the contig end no longer
	                              clips (src/print_vcf.c:154-157); 0, 0 = no region */
	const bsgpu_dbsnp *dbsnp;  /* -D: the contig's dbSNP entries, or NULL (single-contig entry points; see bsgpu_set_contig_annotation) */
} bsgpu_bcf_params;

typedef struct bsgpu_ctx bsgpu_ctx;

/* counters a context keeps; all monotonically increasing */
typedef struct {
	uint64_t kernel_launches;  /* kernels of this library launched so far */
	uint64_t sites;            /* sites processed */
	uint64_t sites_called;     /* sites with n > 0 */
	uint64_t h2d_bytes, d2h_bytes;
	uint64_t qsum_overflow;    /* sites whose integer quality / mapq^2 sums left the float-exact envelope (2^24) */
	/* wall time bsgpu_call_bam spent, accumulated: framing + H2D + record decode | descriptors D2H + block builder |
	 * normalisation + pileup + model + D2H of gt_vcf[] */
	double bam_decode_s, bam_build_s, bam_call_s;
	/* Guard bands (hard part 1).  The device's log / exp are within 1.5 ulp of libm's, so a decision
	 * that hangs on the last bits of a likelihood may fall the other way than in the reference.  Such sites are counted:
	 * near_tie_sites   the two best genotype log-likelihoods differ by <= 1e-9 relative (src/genotype_model.c:231-239): max_gt,
	 *                  hence GT, may differ; exact_tie_sites: they are EQUAL here (the reference sums the same terms in another
	 *                  order for some genotype pairs, so it need not see a tie: listed as well)
	 * near_qual_sites  QUAL / GQ before truncation lies within its error band of an integer (src/print_vcf.c:140-148)
	 * near_fs_sites    FS before truncation lies within its error band of an integer (src/print_vcf.c:151)
	 * bsgpu_guard_read lists them (up to 65536 between two resets). */
	uint64_t near_tie_sites, exact_tie_sites, near_qual_sites, near_fs_sites;
	uint64_t long_segments;    /* segments longer than BSGPU_MAX_SEG_LEN handed to a _dev entry point (contract violation) */
	/* Host-buffer entry points return results over PCIe as 120-byte wire records and rebuild the reference's 200 / 208-byte
	 * records in the caller's array on host threads (bs_call_b200/csrc/bsgpu_wire.h): sites that came home that way, and
	 * chunks fetched again in full because a field did not fit its wire width (a count above 65535, a quality above 255) */
	uint64_t wire_sites, wire_refetched_chunks;
} bsgpu_stats;

void bsgpu_default_params(bsgpu_params *p);
int bsgpu_init(const bsgpu_params *p, bsgpu_ctx **out);
void bsgpu_destroy(bsgpu_ctx *ctx);
const char *bsgpu_last_error(void);
int bsgpu_get_stats(bsgpu_ctx *ctx, bsgpu_stats *out);
int bsgpu_version(void);
int bsgpu_sync(bsgpu_ctx *ctx);              /* wait for everything queued on the context's device */
/* the flagged sites since the last reset: ids[k] = kind << 56 | id; kind 1 near tie of the call, 2 QUAL / GQ, 3 FS; id = index
 * of the site within the bsgpu_call_sites[_bcf] call, or its position for the block / reader entry points.  *n <= cap. */
int bsgpu_guard_read(bsgpu_ctx *ctx, uint64_t *ids, size_t cap, size_t *n, int reset);
/* Debugging aid (no counterpart in the reference; compute-sanitizer is not available everywhere): with BSGPU_REDZONE=1 in the
 * environment when the library is loaded, every device buffer the library owns is allocated at exactly the size asked for and
 * followed by a 4-KiB zone of a known pattern.  This reads the zones of all live buffers back (after a device synchronise) and
 * reports how many zones were looked at so far -- buffers released earlier included -- and how many were written into by a kernel
 * or a copy.  BSGPU_OK when none was; BSGPU_FAIL otherwise, and always without BSGPU_REDZONE. */
int bsgpu_debug_redzones(unsigned long long *checked, unsigned long long *corrupt);

/* page-locked host memory: arrays handed to the host-buffer entry points copy at full PCIe rate when they
 * come from here (any host pointer is accepted, pageable ones just copy slower) */
void *bsgpu_host_alloc(size_t bytes);
void bsgpu_host_free(void *p);

/* ---- host-buffer entry points: H2D, kernels, D2H all inside the call (chunked and double buffered) ---- */

/* pileup[] + ref codes -> gt_meth[] + skip[] */
int bsgpu_call_sites(bsgpu_ctx *ctx, const bsgpu_pileup *pileup, const uint8_t *ref, size_t n,
		bsgpu_gt_meth *out, uint8_t *skip);

/* sorted segments -> pileup[] for the window [x, x+sz) */
int bsgpu_pileup_block(bsgpu_ctx *ctx, const bsgpu_seg *segs, size_t nseg, const uint8_t *bases, size_t nbases,
		uint32_t x, uint32_t sz, bsgpu_pileup *out);

/* sorted segments + ref codes for [x, x+sz) -> gt_vcf[] (ready = 1 everywhere) */
int bsgpu_call_block(bsgpu_ctx *ctx, const bsgpu_seg *segs, size_t nseg, const uint8_t *bases, size_t nbases,
		const uint8_t *ref, uint32_t x, uint32_t sz, bsgpu_gt_vcf *out);

/* normalised templates -> segments sorted by pos (host side; pure function, no device work).
 * segs must have room for bsgpu_stage_bound(templates, n) entries.  *nseg receives the count. */
size_t bsgpu_stage_bound(const bsgpu_template *t, size_t n);
int bsgpu_stage_templates(const bsgpu_template *t, size_t n, const uint8_t *bases, uint32_t x, uint32_t y,
		bsgpu_seg *segs, size_t *nseg);

/* raw templates -> gt_vcf[]: normalisation, pileup and model all on the device.
 * ref holds codes for [x, y] where x = max(first template start - 2, 1) (src/process_template.c:24-28);
 * *x_out receives x; out must hold y - x + 1 records. */
int bsgpu_process_block(bsgpu_ctx *ctx, const bsgpu_template *t, size_t n, const uint8_t *bases, size_t nbases,
		const bsgpu_misms *misms, size_t nmisms, const uint8_t *ref, uint32_t y,
		uint32_t *x_out, bsgpu_gt_vcf *out);

/* Side channels of --report-file (stats != NULL in the reference).  While enabled, bsgpu_process_block and
 * bsgpu_call_bam also gather a bsgpu_profile over every template they normalise, in call order; `ref` of
 * bsgpu_process_block must then hold one more code (position y + 1; the reference's ref1 string has it,
 * src/process_template.c:29-30).  bsgpu_profile_read waits for the queued work, copies the totals out and, if `reset`,
 * starts again from zero. */
int bsgpu_profile_enable(bsgpu_ctx *ctx, int on);
int bsgpu_profile_read(bsgpu_ctx *ctx, bsgpu_profile *out, int reset);

/* Site-level side channels (bsgpu_site_stats above).  n_contigs sizes the per-contig table; a record's contig is the CHROM the
 * writer gives it (bsgpu_bcf_params.rid, vcf_rid[tid]).  bsgpu_set_contig_gc hands over ctg_stats->gc of a contig: the GC
 * percentage of every 100-base bin counted from start_pos (src/read_reference.c:120-123; values above 100 = unknown); without
 * it the gc_pcent histograms stay empty.  _read waits for the queued work; ctg may be NULL. */
int bsgpu_site_stats_enable(bsgpu_ctx *ctx, int on, int n_contigs);
int bsgpu_set_contig_gc(bsgpu_ctx *ctx, int rid, uint32_t start_pos, const uint8_t *gc, uint32_t nbins);
int bsgpu_site_stats_read(bsgpu_ctx *ctx, bsgpu_site_stats *out, bsgpu_ctg_site_stats *ctg, int n_contigs, int reset);

/* ---- writer side: a block of gt_vcf records -> the BCF records print_thread would hand to bcf_write()
 *      (print_vcf_entry / flush_vcf_entries / _print_vcf_entry, src/print_vcf.c:32-381, 535-594, driven per block by
 *      src/process.c:89-104), laid out as in a BCF file: l_shared, l_indiv, CHROM, POS, rlen, QUAL, n_allele|n_info,
 *      n_fmt|n_sample, shared, indiv.  ref holds sz + 2 codes: positions x .. x + sz + 1, the
 *      string get_sequence_string() leaves in work.ref.  *nbytes / *nrec receive the size of the output. ---- */
void bsgpu_default_bcf_params(bsgpu_bcf_params *p);      /* ids 0..15 in header order, rid 0, no contig end, -A off, no region, no dbSNP */
/* region and dbSNP entries of contig `tid` for the entry points that see many contigs (bsgpu_call_bam_bcf, sessions opened on
 * the context afterwards or before -- it is looked up per window).  db may be NULL (region only); the arrays are copied. */
int bsgpu_set_contig_annotation(bsgpu_ctx *ctx, int tid, uint32_t reg_start, uint32_t reg_stop, const bsgpu_dbsnp *db);
int bsgpu_bcf_block(bsgpu_ctx *ctx, const bsgpu_gt_vcf *vcf, const uint8_t *ref, uint32_t x, uint32_t sz, const bsgpu_bcf_params *p,
		uint8_t *out, size_t out_cap, size_t *nbytes, size_t *nrec);
/* sorted segments -> BCF records of the block: pileup, model and writer derivations on the device, only the records come back */
int bsgpu_call_block_bcf(bsgpu_ctx *ctx, const bsgpu_seg *segs, size_t nseg, const uint8_t *bases, size_t nbases, const uint8_t *ref,
		uint32_t x, uint32_t sz, const bsgpu_bcf_params *p, uint8_t *out, size_t out_cap, size_t *nbytes, size_t *nrec);
/* pileup[] of n consecutive sites starting at position x (one block) -> BCF records; chunked and pipelined like bsgpu_call_sites */
int bsgpu_call_sites_bcf(bsgpu_ctx *ctx, const bsgpu_pileup *pileup, const uint8_t *ref, size_t n, uint32_t x, const bsgpu_bcf_params *p,
		uint8_t *out, size_t out_cap, size_t *nbytes, size_t *nrec);
/* The whole path with the writer's derivations included: BAM records -> BCF records.  As bsgpu_call_bam, except that what
 * comes back is the record stream of every block in stream order (blocks[b].vcf_off is 0).  p gives the header ids and
 * -A; CHROM of a record is vcf_rid[tid] (tid itself when vcf_rid is NULL) and sites from position target_len[tid] on are not
 * written (ctg->end_pos, src/print_vcf.c:159). */
int bsgpu_call_bam_bcf(bsgpu_ctx *ctx, const uint8_t *bam, size_t nbytes, int n_targets, const uint32_t *target_len,
		const uint8_t *const *ctg_codes, const bsgpu_reader_params *rp, const bsgpu_bcf_params *p, const int32_t *vcf_rid,
		bsgpu_block *blocks, size_t block_cap, size_t *nblocks, uint8_t *out, size_t out_cap, size_t *nbytes_out, size_t *nrec);
/* device-resident variant of bsgpu_bcf_block; waits for `stream` to return the sizes */
int bsgpu_bcf_block_dev(bsgpu_ctx *ctx, const void *d_vcf, const void *d_ref, uint32_t x, uint32_t sz, const bsgpu_bcf_params *p,
		void *d_out, size_t out_cap, size_t *nbytes, size_t *nrec, void *stream);

/* ---- reader side: `bam` is the byte stream that follows the header of an uncompressed BAM file (what remains of the
 *      file after BGZF inflation: int32 block_size + record, repeated), coordinate sorted ---- */
void bsgpu_default_reader_params(bsgpu_reader_params *p);

/* get_next_align_details() for every record, on the device (src/input_sam.c:222-312).  rec_out receives one
 * descriptor per record; the decoded reads and CIGAR events stay resident in the context for bsgpu_call_bam and are
 * also copied to bases_out / misms_out when those are not NULL (tests, host-side consumers). */
int bsgpu_decode_records(bsgpu_ctx *ctx, const uint8_t *bam, size_t nbytes, const bsgpu_reader_params *rp,
		bsgpu_record *rec_out, size_t rec_cap, size_t *nrec, uint8_t *bases_out, size_t bases_cap, size_t *nbases,
		bsgpu_misms *misms_out, size_t misms_cap, size_t *nmisms);

/* read_input(): mate pairing, positional duplicate removal, block cutting (src/get_template_vector.c:49-389) over the
 * descriptors of bsgpu_decode_records.  Host side (order dependent over the sorted stream); templates refer to the
 * decoded arrays by offset.  Streams on which the reference's read_input ends the process are refused (BSGPU_FAIL, the text says
 * which): a read name that is already waiting for its mate (:327-328, fatal error) and a mate that disagrees with its waiting
 * partner about their positions (:239, assert). */
int bsgpu_build_blocks(const uint8_t *bam, size_t nbytes, const bsgpu_record *rec, size_t nrec, const bsgpu_reader_params *rp,
		bsgpu_block *blocks, size_t block_cap, size_t *nblocks, bsgpu_template *tmpl, size_t tmpl_cap, size_t *ntmpl);
/* same, also adding read_input's per-reason tallies (filtered records, mates whose partner never came, duplicates) to
 * filter_cts[15] / filter_bases[15] */
int bsgpu_build_blocks_tally(const uint8_t *bam, size_t nbytes, const bsgpu_record *rec, size_t nrec, const bsgpu_reader_params *rp,
		bsgpu_block *blocks, size_t block_cap, size_t *nblocks, bsgpu_template *tmpl, size_t tmpl_cap, size_t *ntmpl,
		uint64_t *filter_cts, uint64_t *filter_bases);

/* The whole path: BAM records -> decode -> blocks -> normalisation -> pileup -> model.  ctg_codes[tid] holds the
 * reference codes 0..4 of positions 1..target_len[tid].  vcf receives, per contig that has blocks, one gt_vcf record for
 * every position from the first block's x to the last block's y; blocks[b].vcf_off locates block b's window in it.
 * Besides the streams bsgpu_build_blocks refuses, this (and the sessions, and bsgpu_process_block on raw templates) refuses a
 * stream on which call_genotypes_ML dies on its assert that no template begins before its block's window
 * (src/call_genotypes.c:186): a lone mate kept by -k and -d together whose absent partner's position lies before the block. */
int bsgpu_call_bam(bsgpu_ctx *ctx, const uint8_t *bam, size_t nbytes, int n_targets, const uint32_t *target_len,
		const uint8_t *const *ctg_codes, const bsgpu_reader_params *rp, bsgpu_block *blocks, size_t block_cap, size_t *nblocks,
		bsgpu_gt_vcf *vcf, size_t vcf_cap, size_t *nvcf);

/* ---- streaming session: the reader side in bounded memory, for streams of any length ----
 * read_input() reads one record at a time and hands a block on as soon as it closes (src/get_template_vector.c:86-110,
 * 140-189), so bs_call's memory is O(block).  A session gives the same behaviour to the device path: the record stream is
 * fed in arbitrary slices (a slice may end anywhere, also inside a record), the session cuts it at records where read_input
 * is certain to start a new block, runs batches of about `batch_bytes` (0: 384 MiB, or BSGPU_BATCH_BYTES) through
 * decode -> blocks -> normalisation -> pileup -> model (-> writer) on a worker thread while the caller feeds the next
 * batch, and lends the results of every batch out in page-locked memory.  The three threads of the reference map onto it:
 * the reader thread feeds (src/bs_call.c main -> read_input), the session's worker is process_thread (src/process.c:43-72),
 * the print thread drains (src/process.c:74-110).  A single thread can drive it too: pass `accepted` to bsgpu_bam_feed
 * (it then never blocks and may take fewer bytes than offered) and poll bsgpu_bam_drain.
 *   bcf == NULL: results are gt_vcf[] records, block windows as in bsgpu_call_bam (blocks[b].vcf_off indexes the batch's array);
 *   bcf != NULL: results are the BCF records of the batch's blocks in stream order, as in bsgpu_call_bam_bcf.
 * ctg_codes[] must stay valid while the session is open; the context must not be used by other calls meanwhile.
 * Limits: one block (not one stream) must stay below 4 Gi decoded bases. */
typedef struct bsgpu_bam_session bsgpu_bam_session;
typedef struct {
	uint64_t id;               /* handle for bsgpu_bam_release; 0: no result was ready */
	const bsgpu_block *blocks; /* the blocks of the batch, templates numbered within the batch */
	size_t nblocks;
	const uint8_t *data;       /* BCF record stream, or gt_vcf[] */
	size_t nbytes;
	size_t nrec;               /* BCF records, or gt_vcf records */
	size_t bytes_in, records_in;   /* the part of the stream (bytes, BAM records) these results account for */
	int finished;              /* 1: the stream was finished and no further result will come */
	int pad_;
} bsgpu_bam_result;
typedef struct {
	uint64_t bytes_fed, bytes_done, records_done, batches;
	uint64_t carry_bytes;      /* bytes that were staged again because they followed a batch's last certain block start */
	uint64_t empty_batches;    /* batches that held no certain block start (the staging buffer grew instead) */
	uint64_t pinned_bytes;     /* page-locked memory the session holds right now */
} bsgpu_bam_progress_t;
int bsgpu_bam_open(bsgpu_ctx *ctx, int n_targets, const uint32_t *target_len, const uint8_t *const *ctg_codes, const bsgpu_reader_params *rp,
		const bsgpu_bcf_params *bcf, const int32_t *vcf_rid, size_t batch_bytes, bsgpu_bam_session **out);
/* the reference codes of a contig that was NULL at bsgpu_bam_open (a host that loads contigs as the stream reaches them):
 * call it before the first record of that contig is fed; the array must stay valid until the session is closed */
int bsgpu_bam_set_contig(bsgpu_bam_session *s, int tid, const uint8_t *codes);
/* For a host that feeds bytes without looking at the records (BGZF blocks inflated straight into bsgpu_bam_reserve's buffer):
 * `fn(user, tid)` is called by the session's worker when the results of contig `tid` are about to be computed and no codes
 * were set for it -- once per contig, in stream order, never concurrently with itself.  It may call bsgpu_bam_set_contig,
 * bsgpu_set_contig_annotation and bsgpu_set_contig_gc; a return value other than BSGPU_OK (or codes still missing) fails the
 * session.  Counterpart of the contig change in read_input (src/get_template_vector.c:111-124). */
int bsgpu_bam_on_contig(bsgpu_bam_session *s, int (*fn)(void *user, int tid), void *user);
/* copies the slice into the session's staging.  accepted == NULL: blocks while both staging buffers are full;
 * accepted != NULL: never blocks, *accepted <= nbytes says how much was taken (drain, then offer the rest again) */
int bsgpu_bam_feed(bsgpu_bam_session *s, const uint8_t *bytes, size_t nbytes, size_t *accepted);
/* zero-copy feeding: *ptr = where the next bytes of the stream go, *avail = how many fit (0 when wait == 0 and both
 * staging buffers are full); write there (an inflater's output buffer), then commit what was written */
int bsgpu_bam_reserve(bsgpu_bam_session *s, uint8_t **ptr, size_t *avail, int wait);
int bsgpu_bam_commit(bsgpu_bam_session *s, size_t nbytes);
/* The bytes fed so far end where read_input would close a block whatever follows (end of a contig, end of a region cut at
 * a block boundary): they run as a batch of their own, so no batch of results mixes two regions, while the session stays
 * open and the next region can be fed at once. */
int bsgpu_bam_cut(bsgpu_bam_session *s);
int bsgpu_bam_finish(bsgpu_bam_session *s);       /* no more bytes: the staged rest runs as the last batch */
int bsgpu_bam_rewind(bsgpu_bam_session *s);       /* after a finished, fully drained stream: ready for another stream (buffers kept) */
/* next batch of results in stream order; wait != 0: blocks until one is ready or none can come.  res->data / res->blocks
 * stay valid until bsgpu_bam_release(s, res->id), which returns the buffers for reuse */
int bsgpu_bam_drain(bsgpu_bam_session *s, int wait, bsgpu_bam_result *res);
int bsgpu_bam_release(bsgpu_bam_session *s, uint64_t id);
int bsgpu_bam_progress(bsgpu_bam_session *s, bsgpu_bam_progress_t *out);
int bsgpu_bam_close(bsgpu_bam_session *s);

/* ---- device-pointer entry points: everything already resident in HBM, asynchronous on `stream`
 *      (a cudaStream_t passed as void*; NULL = the context's own stream) ----
 * Contract of the caller-owned arrays (the host-buffer entry points check or arrange all of this themselves):
 *   d_bases   16-byte aligned, and readable for 16 bytes beyond its last base: the pileup kernel stages a segment with a bulk
 *             copy of the 16-byte aligned cover of its bytes, so it reads up to 15 bytes either side of the segment
 *   d_segs    every len <= BSGPU_MAX_SEG_LEN (a longer segment would be seen by the first three tiles it overlaps only);
 *             segments that break this are counted in bsgpu_stats.long_segments and the results of the call are void
 *   d_pileup / d_out / d_vcf   8-byte aligned (16 for the block entry points) */
int bsgpu_call_sites_dev(bsgpu_ctx *ctx, const void *d_pileup, const void *d_ref, size_t n,
		void *d_out, void *d_skip, void *stream);
/* same, writing gt_vcf[] (208-byte records with ready = 1 and the skip flag inside) */
int bsgpu_call_sites_vcf_dev(bsgpu_ctx *ctx, const void *d_pileup, const void *d_ref, size_t n,
		void *d_vcf, void *stream);
int bsgpu_pileup_block_dev(bsgpu_ctx *ctx, const void *d_segs, size_t nseg, const void *d_bases,
		uint32_t x, uint32_t sz, void *d_pileup_out, void *stream);
int bsgpu_call_block_dev(bsgpu_ctx *ctx, const void *d_segs, size_t nseg, const void *d_bases, const void *d_ref,
		uint32_t x, uint32_t sz, void *d_vcf_out, void *stream);

/* ---- synthetic workloads generated on the device (counter-based RNG; seed + index -> record) ---- */
/* config 2 of BASELINE.json: per-site count vectors.  Writes n pileup records and n ref codes. */
int bsgpu_synth_sites_dev(bsgpu_ctx *ctx, uint64_t seed, uint64_t first_site, size_t n, double mean_depth,
		void *d_pileup, void *d_ref, void *stream);
/* simulated WGBS reads over a window: segments + packed bases (read i at offset i*read_len) + ref codes;
 * bsgpu_synth_block_nseg gives the number of reads the generator emits for (sz, read_len, depth) */
size_t bsgpu_synth_block_nseg(uint32_t sz, uint32_t read_len, double depth);
int bsgpu_synth_block_dev(bsgpu_ctx *ctx, uint64_t seed, uint32_t x, uint32_t sz, uint32_t read_len, double depth,
		void *d_segs, size_t seg_cap, void *d_bases, size_t base_cap, void *d_ref, size_t *nseg, size_t *nbases,
		void *stream);

/* synthetic coordinate-sorted BAM record stream (paired-end, fixed-size records: bsgpu_synth_bam_bytes(n, read_len) bytes
 * for n templates).  d_pos_f / d_pos_r: 1-based start of the forward / reverse mate of every template; d_src[t]: the
 * template whose reads template t carries (t itself, or an earlier one for a positional duplicate); d_rank[i]: index in
 * the sorted stream of record i (i < n: forward mate of template i, else reverse mate of template i - n).  Reads are
 * drawn from the same synthetic genome as bsgpu_synth_block_dev; bsgpu_synth_ref_dev writes its codes for [x, x+sz). */
size_t bsgpu_synth_bam_bytes(size_t ntemplates, uint32_t read_len);
int bsgpu_synth_bam_dev(bsgpu_ctx *ctx, uint64_t seed, size_t ntemplates, uint32_t read_len, const void *d_pos_f,
		const void *d_pos_r, const void *d_src, const void *d_rank, void *d_out, void *stream);
int bsgpu_synth_ref_dev(bsgpu_ctx *ctx, uint64_t seed, uint32_t x, uint32_t sz, void *d_ref, void *stream);

/* ---- diagnostics ---- */
/* Evaluates, on the host, the table-driven log / exp the kernels use (same source; both sides are FMA-exact, so
 * these are the device's values).  out_log / out_exp may be NULL.  log: x positive normal; exp: x in [-700, 0]. */
int bsgpu_math_probe(const double *x, size_t n, double *out_log, double *out_exp);

/* ---- wire records ---- */
/* What the host-buffer entry points send over PCIe instead of the reference's records (layout in
 * bs_call_b200/csrc/bsgpu_wire.h: gt_prob[10], fisher_strand, counts as uint16, qualities as uint8, mq, aq, max_gt, skip;
 * BSGPU_WIRE_BYTES each).  Host-only functions, no device needed:
 * bsgpu_wire_pack    n records of rec_bytes (200: gt_meth, with skip[]; 208: gt_vcf, skip may be NULL) -> wire; BSGPU_FAIL if a
 *                    field does not fit its wire width (the device then sends that chunk as full records)
 * bsgpu_wire_expand  the rebuilding pass the library runs on its own threads: wire -> records (+ skip[] for rec_bytes 200),
 *                    split over `threads` threads (<= 0: one) */
#define BSGPU_WIRE_BYTES 120
int bsgpu_wire_pack(const void *records, size_t n, size_t rec_bytes, const uint8_t *skip, void *wire);
int bsgpu_wire_expand(const void *wire, size_t n, size_t rec_bytes, void *records, uint8_t *skip, int threads);

#ifdef __cplusplus
}
#endif
#endif
