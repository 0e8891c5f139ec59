"""numpy views of the C records declared in include/bsgpu.h (which mirror the reference's include/bs_call.h)."""
import numpy as np

PILEUP = np.dtype([("counts", "<u4", (2, 8)), ("n", "<u4"), ("quality", "<f4", (8,)), ("mapq2", "<f4")])
GT_METH = np.dtype([("counts", "<u8", (8,)), ("qual", "<i4", (8,)), ("gt_prob", "<f8", (10,)),
                    ("fisher_strand", "<f8"), ("mq", "<i4"), ("aq", "<i4"), ("max_gt", "u1"), ("pad", "u1", (7,))])
GT_VCF = np.dtype([("gtm", GT_METH), ("ready", "u1"), ("skip", "u1"), ("pad", "u1", (6,))])
SEG = np.dtype([("pos", "<u4"), ("off", "<u4"), ("len", "<u2"), ("mapq", "u1"), ("flags", "u1"), ("pad", "<u4")])
TEMPLATE = np.dtype([("forward_position", "<u4"), ("reverse_position", "<u4"), ("reference_span", "<u4", (2,)),
                     ("read_off", "<u4", (2,)), ("read_len", "<u4", (2,)), ("mm_off", "<u4", (2,)),
                     ("mm_n", "<u4", (2,)), ("present", "u1", (2,)), ("mapq", "u1", (2,)),
                     ("orientation", "u1"), ("bs_strand", "u1"), ("pad", "u1", (2,))])
MISMS = np.dtype([("type", "<u4"), ("position", "<u4"), ("size", "<u4")])

# reader side (include/bsgpu.h: bsgpu_record, bsgpu_block)
RECORD = np.dtype([("ret", "<i4"), ("filtered", "<u4"), ("forward_position", "<u4"), ("reverse_position", "<u4"),
                   ("alignment_flag", "<u4"), ("align_length", "<u4"), ("reference_span", "<u4"),
                   ("read_off", "<u4"), ("read_len", "<u4"), ("mm_off", "<u4"), ("mm_n", "<u4"), ("tid", "<i4"),
                   ("reverse", "u1"), ("orientation", "u1"), ("bs_strand", "u1"), ("mapq", "u1"),
                   ("q01", "u1", (2,)), ("pad", "u1", (2,))])
BLOCK = np.dtype([("tid", "<u4"), ("x", "<u4"), ("y", "<u4"), ("first_template", "<u4"), ("n_templates", "<u4"),
                  ("pad", "<u4"), ("vcf_off", "<u8")])
assert RECORD.itemsize == 56 and BLOCK.itemsize == 32

assert PILEUP.itemsize == 104
assert GT_METH.itemsize == 200
assert GT_VCF.itemsize == 208
assert SEG.itemsize == 16
assert TEMPLATE.itemsize == 56
assert MISMS.itemsize == 12

GENOTYPES = ("AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT")


# bsgpu_site_stats (include/bsgpu.h): the writer's --report-file statistics, flat
STATS_FS_MAX = 4096
STATS_COV_MAX = 4096
COV_STATS = np.dtype([("var", "<u8"), ("CpG", "<u8", (2,)), ("CpG_inf", "<u8", (2,)), ("all", "<u8"), ("gc_pcent", "<u8", (101,))])
SITE_STATS = np.dtype([
    ("snps", "<u8", (2,)), ("multi", "<u8", (2,)), ("dbSNP_sites", "<u8", (2,)), ("dbSNP_var", "<u8", (2,)), ("CpG_ref", "<u8", (2,)), ("CpG_nonref", "<u8", (2,)),
    ("mut_counts", "<u8", (12, 2)), ("dbSNP_mut_counts", "<u8", (12, 2)), ("qual", "<u8", (4, 256)), ("filter_counts", "<u8", (2, 32)),
    ("CpG_ref_meth", "<f8", (2, 101)), ("CpG_nonref_meth", "<f8", (2, 101)),
    ("qd_stats", "<u8", (256, 2)), ("mq_stats", "<u8", (256, 2)), ("fs_stats", "<u8", (STATS_FS_MAX, 2)),
    ("fs_overflow", "<u8"), ("cov_overflow", "<u8"), ("cov", COV_STATS, (STATS_COV_MAX,))])
CTG_SITE_STATS = np.dtype([("snps", "<u8", (2,)), ("multi", "<u8", (2,)), ("dbSNP_sites", "<u8", (2,)), ("dbSNP_var", "<u8", (2,)),
                           ("CpG_ref", "<u8", (2,)), ("CpG_nonref", "<u8", (2,))])
