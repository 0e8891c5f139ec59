/*
 * bs_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the bs_call 2.1.7 pileup + genotype-likelihood hot path, written from the
 * behaviour of the reference sources (cited per function in bs_oracle.c).  It exists so that the CUDA path
 * can be checked on machines where /root/reference is absent (the GPU boxes).  Only tests/, the smoke check
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product (libbsgpu) never does.
 *
 * PARITY PINNING: the reference ships no tests or golden vectors, so this restatement is
 * pinned against the reference's own compiled objects (oracle/_ref/libbsref.so, built by oracle/Makefile
 * from the sources where they lie under /root/reference) by tests/test_oracle_vs_reference.py, and against
 * the golden fixtures under tests/golden/ that were captured from those objects
 * (tests/golden/make_golden.py).
 */
#ifndef BS_ORACLE_H
#define BS_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same layouts as the reference's pileup / gt_meth / gt_vcf (include/bs_call.h:174-182, 152-160, 162-166) */
typedef struct {
	uint32_t counts[2][8];
	uint32_t n;
	float quality[8];
	float mapq2;
} bso_pileup;              /* 104 bytes */

typedef struct {
	uint64_t counts[8];
	int32_t qual[8];
	double gt_prob[10];
	double fisher_strand;
	int32_t mq;
	int32_t aq;
	uint8_t max_gt;
	uint8_t pad[7];
} bso_gt_meth;             /* 200 bytes */

typedef struct {
	bso_gt_meth gtm;
	uint8_t ready;
	uint8_t skip;
	uint8_t pad[6];
} bso_gt_vcf;              /* 208 bytes */

/* Flat template record (identical to bsref_template in ref_harness.c) */
typedef struct {
	uint32_t forward_position, reverse_position;
	uint32_t reference_span[2];
	uint32_t read_off[2];
	uint32_t read_len[2];
	uint32_t mm_off[2];
	uint32_t mm_n[2];
	uint8_t present[2];
	uint8_t mapq[2];
	uint8_t orientation;
	uint8_t bs_strand;
	uint8_t pad[2];
} bso_template;            /* 56 bytes */

/* misms type codes follow the reference enum gt_misms_t {MISMS, INS, DEL, SOFT} (include/bs_call.h:53);
 * note the naming inversion: CIGAR 'D' -> INS (zero-fill), CIGAR 'I' -> DEL (drop) (src/input_sam.c:117-130) */
enum { BSO_MISMS = 0, BSO_INS = 1, BSO_DEL = 2, BSO_SOFT = 3 };
typedef struct { uint32_t type, position, size; } bso_misms;

typedef struct {
	double under_conv, over_conv, ref_bias;
	uint32_t left_trim[2], right_trim[2];
	uint8_t min_qual;
} bso_params;

void bso_default_params(bso_params *p);

/* model pieces */
void bso_qprob_table(double *out /* 44 x {e,k,ln_k,ln_k_half,ln_k_one} */);
void bso_lfact_table(double *out /* 256 */);
void bso_calc_gt_prob(bso_gt_meth *gt, const bso_params *p, int rf);
double bso_fisher(const int c_in[4]);

/* per-site body of call_thread */
void bso_summarise(const bso_pileup *tp, bso_gt_meth *tg);
int  bso_strand_table(const bso_pileup *tp, int max_gt, int ftab[4]);
void bso_call_site(const bso_pileup *tp, int rf, const bso_params *p, bso_gt_meth *out, uint8_t *skip);
void bso_call_sites(const bso_pileup *tp, const uint8_t *ref, size_t n, const bso_params *p,
		bso_gt_meth *out, uint8_t *skip, int nthreads);

/* pileup loop over normalised templates */
void bso_pileup_block(const bso_template *t, size_t n, const uint8_t *bases, uint32_t x, uint32_t y,
		const bso_params *p, bso_pileup *out /* y-x+1, zeroed here */);

/* template normalisation: returns 0, or <0 on the conditions where the reference calls gt_fatal_error_msg */
int bso_normalise_block(const bso_template *t, size_t n, const uint8_t *bases, const bso_misms *mm,
		const bso_params *p, bso_template *out_t, uint8_t *out_bases, size_t out_cap, size_t *out_used);

/* whole path: raw templates -> gt_vcf[] */
int bso_process_block(const bso_template *t, size_t n, const uint8_t *bases, const bso_misms *mm,
		const uint8_t *refcodes /* codes for [x, y] */, uint32_t y, const bso_params *p,
		uint32_t *x_out, bso_pileup *pile_out, bso_gt_vcf *vcf_out);

/* --report-file side channels of this path (bs_stats, include/bs_call.h:124-146): the non-CpG conversion profile of
 * meth_profile() (src/meth_profile.c:48-76) and the tallies of process_template_vector (src/process_template.c:52-63,
 * src/al_utils.c:141,150,308) and of read_input (src/get_template_vector.c:104-107,243-246,314-319,361-364).  While enabled, bso_process_block accumulates them (process-wide, in call order) and its
 * refcodes must hold one more code (position y + 1). */
#define BSO_PROFILE_MAX 1024
typedef struct {
	uint64_t conv_cts[BSO_PROFILE_MAX][4];
	uint32_t used, pad;
	uint64_t base_filter[5];
	uint64_t filter_cts[15], filter_bases[15];   /* indexed by gt_filter_reason; [0] also takes the reads / bases that reach normalisation */
} bso_profile;
void bso_profile_enable(int on);
int bso_profile_is_on(void);
void bso_profile_tally(int reason_cts, uint64_t cts, int reason_bases, uint64_t bases);      /* used by the reader restatement */
void bso_profile_reset(void);
void bso_profile_read(bso_profile *out);

/* ---- reader side (bs_oracle_reader.c) ---- */
/* what get_next_align_details() yields for one BAM record (src/input_sam.c:222-312) */
typedef struct {
	int32_t ret;                 /* 0 keep, 1 filtered */
	uint32_t filtered;           /* gt_filter_reason, include/bs_call.h:50 */
	uint32_t forward_position, reverse_position;
	uint32_t alignment_flag, align_length;
	uint32_t reference_span;     /* of the mate this record is */
	uint32_t read_off, read_len; /* into the packed-base output (ret == 0 only) */
	uint32_t mm_off, mm_n;       /* into the event output */
	uint8_t reverse, orientation, bs_strand, mapq;
} bso_record;                  /* 48 bytes */

/* one block as read_input() hands it to process_template_vector(): templates [first, first + n), window [x, y] */
typedef struct { uint32_t tid, x, y, first_template, n_templates, pad; uint64_t vcf_off; } bso_block;

int bso_decode_records(const uint8_t *bam, size_t nbytes, int mapq_thresh, uint32_t max_template_len, int keep_unmatched,
		int ignore_dup, bso_record *out, size_t cap, size_t *nrec, uint8_t *bases_out, size_t bases_cap, size_t *nbases,
		bso_misms *misms_out, size_t misms_cap, size_t *nmisms);
void bso_set_params(const bso_params *p);      /* model parameters used when bso_read_input runs the whole chain */
int bso_read_input(const uint8_t *bam, size_t nbytes, int n_targets, const uint32_t *target_len, const uint8_t *const *ctg_codes,
		int mapq_thresh, uint32_t max_template_len, int keep_unmatched, int ignore_duplicates, int keep_duplicates, int run_chain,
		bso_block *blocks, size_t block_cap, size_t *nblocks, bso_template *tmpl, size_t tmpl_cap, size_t *ntmpl,
		uint8_t *bases, size_t bases_cap, size_t *nbases, bso_misms *misms, size_t misms_cap, size_t *nmisms,
		bso_gt_vcf *vcf, size_t vcf_cap, size_t *nvcf);

/* ---- writer side (bs_oracle_writer.c): what print_thread derives from a block of gt_vcf records and hands to bcf_write()
 *      (src/print_vcf.c:32-381, 535-594; src/process.c:89-104).  refcodes: sz + 2 codes (positions x .. x + sz + 1);
 *      vcf_ids: the 16 header dictionary ids in the order of include/bs_call.h:192-207; out receives the records in BCF
 *      layout (l_shared, l_indiv, fixed fields, shared, indiv). ---- */
#define BSO_BCF_MAX_RECORD 384
/* the dbSNP entries of one contig as the writer sees them through dbSNP_lookup_name() (src/dbSNP.c:305-346) */
typedef struct { uint32_t n; const uint32_t *pos; const uint8_t *flags; const uint32_t *name_off; const uint8_t *names; } bso_dbsnp;
size_t bso_print_site_ann(const bso_gt_vcf *vcf, uint32_t sz, const uint8_t *refcodes, uint32_t x, uint32_t i, int rid, uint32_t ctg_end,
		const int *ids, int all_positions, uint32_t reg_start, uint32_t reg_stop, const bso_dbsnp *db, uint8_t *out);
int bso_print_block_ann(const bso_gt_vcf *vcf, uint32_t sz, const uint8_t *refcodes, uint32_t x, int rid, uint32_t ctg_end,
		const int *vcf_ids, int all_positions, uint32_t reg_start, uint32_t reg_stop, const bso_dbsnp *db,
		uint8_t *out, size_t cap, size_t *nbytes, size_t *nrec);
size_t bso_print_site(const bso_gt_vcf *vcf, uint32_t sz, const uint8_t *refcodes, uint32_t x, uint32_t i, int rid, uint32_t ctg_end,
		const int *ids, int all_positions, uint8_t *out);
int bso_print_block(const bso_gt_vcf *vcf, uint32_t sz, const uint8_t *refcodes, uint32_t x, int rid, uint32_t ctg_end,
		const int *vcf_ids, int all_positions, uint8_t *out, size_t cap, size_t *nbytes, size_t *nrec);

/* ---- the writer's --report-file statistics (bs_oracle_stats.c): what _print_vcf_entry() adds to bs_stats for every site it
 *      is called for (src/print_vcf.c:382-526).  Flat image of the reference's bs_stats fields it touches: the fs / qd / mq
 *      vectors and the coverage hash become arrays indexed by value (values beyond the array are counted in *_overflow). ---- */
#define BSO_STATS_FS_MAX 4096
#define BSO_STATS_COV_MAX 4096
typedef struct { uint64_t var, CpG[2], CpG_inf[2], all, gc_pcent[101]; } bso_cov_stats;      /* gt_cov_stats, include/bs_call.h:87-95 */
typedef struct {
	uint64_t snps[2], multi[2], dbSNP_sites[2], dbSNP_var[2], CpG_ref[2], CpG_nonref[2];
	uint64_t mut_counts[12][2], dbSNP_mut_counts[12][2];
	uint64_t qual[4][256];
	uint64_t filter_counts[2][32];
	double CpG_ref_meth[2][101], CpG_nonref_meth[2][101];
	uint64_t qd_stats[256][2], mq_stats[256][2], fs_stats[BSO_STATS_FS_MAX][2];
	uint64_t fs_overflow, cov_overflow;
	bso_cov_stats cov[BSO_STATS_COV_MAX];
} bso_site_stats;
/* the reference keeps the position and filter state of the last '+' strand CpG in function statics (:107-108) */
typedef struct { uint32_t prev_cpg_x; int prev_cpg_flt; } bso_stats_state;
/* gc: GC percentage of every 100-base bin of the contig from start_pos on (ctg_stats->gc, src/read_reference.c:120-123; values
 * above 100 = not known), or NULL */
void bso_stats_block(const bso_gt_vcf *vcf, uint32_t sz, const uint8_t *refcodes, uint32_t x, uint32_t ctg_end, int all_positions,
		uint32_t reg_start, uint32_t reg_stop, const bso_dbsnp *db, const uint8_t *gc, int nbins, uint32_t start_pos,
		bso_site_stats *st, bso_stats_state *state);

/* host twin of the device generator of per-site count vectors (same draws, same records) */
void bso_synth_sites(uint64_t seed, uint64_t first, size_t n, double mean_depth, bso_pileup *out, uint8_t *ref, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
