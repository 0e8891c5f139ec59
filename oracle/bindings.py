"""ctypes bindings for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package (bs_call_b200) never does.

  Oracle     -> oracle/liboracle.so      the restatement in oracle/bs_oracle.c
  Reference  -> oracle/_ref/libbsref.so  the reference's own compiled sources behind oracle/ref_harness.c
                (built in the authoring container where /root/reference exists; travels prebuilt to GPU boxes)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# numpy views of the C records (layouts asserted against sizeof in the loaders)
PILEUP = np.dtype([("counts", "<u4", (2, 8)), ("n", "<u4"), ("quality", "<f4", (8,)), ("mapq2", "<f4")])
GT_METH = np.dtype([("counts", "<u8", (8,)), ("qual", "<i4", (8,)), ("gt_prob", "<f8", (10,)),
                    ("fisher_strand", "<f8"), ("mq", "<i4"), ("aq", "<i4"), ("max_gt", "u1"), ("pad", "u1", (7,))])
GT_VCF = np.dtype([("gtm", GT_METH), ("ready", "u1"), ("skip", "u1"), ("pad", "u1", (6,))])
TEMPLATE = np.dtype([("forward_position", "<u4"), ("reverse_position", "<u4"), ("reference_span", "<u4", (2,)),
                     ("read_off", "<u4", (2,)), ("read_len", "<u4", (2,)), ("mm_off", "<u4", (2,)),
                     ("mm_n", "<u4", (2,)), ("present", "u1", (2,)), ("mapq", "u1", (2,)),
                     ("orientation", "u1"), ("bs_strand", "u1"), ("pad", "u1", (2,))])
MISMS = np.dtype([("type", "<u4"), ("position", "<u4"), ("size", "<u4")])
# reader side: what get_next_align_details() produces per BAM record, and one block as read_input() hands it over
RECORD = np.dtype([("ret", "<i4"), ("filtered", "<u4"), ("forward_position", "<u4"), ("reverse_position", "<u4"),
                   ("alignment_flag", "<u4"), ("align_length", "<u4"), ("reference_span", "<u4"),
                   ("read_off", "<u4"), ("read_len", "<u4"), ("mm_off", "<u4"), ("mm_n", "<u4"),
                   ("reverse", "u1"), ("orientation", "u1"), ("bs_strand", "u1"), ("mapq", "u1")])
BLOCK = np.dtype([("tid", "<u4"), ("x", "<u4"), ("y", "<u4"), ("first_template", "<u4"), ("n_templates", "<u4"),
                  ("pad", "<u4"), ("vcf_off", "<u8")])
assert PILEUP.itemsize == 104 and GT_METH.itemsize == 200 and GT_VCF.itemsize == 208
assert TEMPLATE.itemsize == 56 and MISMS.itemsize == 12 and RECORD.itemsize == 48 and BLOCK.itemsize == 32


def _reader_caps(bam, nrec):
    """generous output capacities for the reader entry points"""
    return dict(tmpl=nrec + 8, bases=len(bam) + 64, misms=len(bam) // 4 + 64, blocks=nrec + 8)


def _decode_records(fn, bam, mapq_thresh, max_template_len, keep_unmatched, ignore_dup):
    bam = _c(bam, np.uint8)
    cap = max(len(bam) // 36 + 8, 8)
    rec = np.zeros(cap, dtype=RECORD)
    bases = np.zeros(len(bam) + 64, dtype=np.uint8)
    misms = np.zeros(len(bam) // 4 + 64, dtype=MISMS)
    n, nb, nm = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    rc = fn(_p(bam), C.c_size_t(len(bam)), C.c_int(mapq_thresh), C.c_uint32(max_template_len), C.c_int(keep_unmatched),
            C.c_int(ignore_dup), _p(rec), C.c_size_t(cap), C.byref(n), _p(bases), C.c_size_t(len(bases)), C.byref(nb),
            _p(misms), C.c_size_t(len(misms)), C.byref(nm))
    if rc:
        raise RuntimeError("decode_records failed: %d" % rc)
    return rec[:n.value].copy(), bases[:nb.value].copy(), misms[:nm.value].copy()


def _read_input(fn, bam, target_len, ctg_codes, mapq_thresh, max_template_len, keep_unmatched, ignore_duplicates,
                keep_duplicates, run_chain):
    bam = _c(bam, np.uint8)
    target_len = _c(target_len, np.uint32)
    nt = len(target_len)
    nrec = max(len(bam) // 36 + 8, 8)
    caps = _reader_caps(bam, nrec)
    blocks = np.zeros(caps["blocks"], dtype=BLOCK)
    tmpl = np.zeros(caps["tmpl"], dtype=TEMPLATE)
    bases = np.zeros(caps["bases"], dtype=np.uint8)
    misms = np.zeros(caps["misms"], dtype=MISMS)
    vcf_cap = int(target_len.sum()) + 8 if run_chain else 8
    vcf = np.zeros(vcf_cap, dtype=GT_VCF)
    codes = None
    ptrs = None
    if run_chain:
        codes = [_c(c, np.uint8) for c in ctg_codes]
        ptrs = (C.c_void_p * nt)(*[c.ctypes.data for c in codes])
    nbk, ntm, nb, nm, nv = (C.c_size_t(0) for _ in range(5))
    rc = fn(_p(bam), C.c_size_t(len(bam)), C.c_int(nt), _p(target_len), ptrs, C.c_int(mapq_thresh), C.c_uint32(max_template_len),
            C.c_int(keep_unmatched), C.c_int(ignore_duplicates), C.c_int(keep_duplicates), C.c_int(run_chain),
            _p(blocks), C.c_size_t(len(blocks)), C.byref(nbk), _p(tmpl), C.c_size_t(len(tmpl)), C.byref(ntm),
            _p(bases), C.c_size_t(len(bases)), C.byref(nb), _p(misms), C.c_size_t(len(misms)), C.byref(nm),
            _p(vcf), C.c_size_t(len(vcf)), C.byref(nv))
    if rc:
        raise RuntimeError("read_input failed: %d" % rc)
    return (blocks[:nbk.value].copy(), tmpl[:ntm.value].copy(), bases[:nb.value].copy(), misms[:nm.value].copy(),
            vcf[:nv.value].copy())


PROFILE_MAX = 1024


class BsoProfile(C.Structure):
    _fields_ = [("conv_cts", (C.c_uint64 * 4) * PROFILE_MAX), ("used", C.c_uint32), ("pad", C.c_uint32),
                ("base_filter", C.c_uint64 * 5), ("filter_cts", C.c_uint64 * 15), ("filter_bases", C.c_uint64 * 15)]


def profile_dict(used, conv, base_filter, filter_cts, filter_bases):
    """common shape of the --report-file side channels: conv[i] = meth_cts of original read position i - 1;
    filter_cts / filter_bases indexed by gt_filter_reason"""
    return dict(used=int(used), conv=np.asarray(conv, dtype=np.uint64)[:int(used)].copy(),
                base_filter=np.asarray(base_filter, dtype=np.uint64).copy(),
                filter_cts=np.asarray(filter_cts, dtype=np.uint64).copy(), filter_bases=np.asarray(filter_bases, dtype=np.uint64).copy())


VCF_IDS = tuple(range(16))      # PASS fail mac1 CX GT FT GL GQ DP MQ QD MC8 AMQ CS CG FS: ids in header order


# bso_site_stats / bsgpu_site_stats: the writer's --report-file statistics (src/print_vcf.c:382-526), flat
STATS_FS_MAX = 4096
STATS_COV_MAX = 4096
COV_STATS = np.dtype([("var", "<u8"), ("CpG", "<u8", (2,)), ("CpG_inf", "<u8", (2,)), ("all", "<u8"), ("gc_pcent", "<u8", (101,))])
SITE_STATS = np.dtype([
    ("snps", "<u8", (2,)), ("multi", "<u8", (2,)), ("dbSNP_sites", "<u8", (2,)), ("dbSNP_var", "<u8", (2,)), ("CpG_ref", "<u8", (2,)), ("CpG_nonref", "<u8", (2,)),
    ("mut_counts", "<u8", (12, 2)), ("dbSNP_mut_counts", "<u8", (12, 2)), ("qual", "<u8", (4, 256)), ("filter_counts", "<u8", (2, 32)),
    ("CpG_ref_meth", "<f8", (2, 101)), ("CpG_nonref_meth", "<f8", (2, 101)),
    ("qd_stats", "<u8", (256, 2)), ("mq_stats", "<u8", (256, 2)), ("fs_stats", "<u8", (STATS_FS_MAX, 2)),
    ("fs_overflow", "<u8"), ("cov_overflow", "<u8"), ("cov", COV_STATS, (STATS_COV_MAX,))])
assert COV_STATS.itemsize == 107 * 8


def site_stats_equal(a, b, rtol=1e-9, what="site stats"):
    """two SITE_STATS records: integer counters identical, the methylation posteriors (sums of doubles) within rtol"""
    bad = []
    for name in SITE_STATS.names:
        if name in ("CpG_ref_meth", "CpG_nonref_meth"):
            if not np.allclose(a[name], b[name], rtol=rtol, atol=1e-12):
                bad.append(name)
        elif name == "cov":
            for f in COV_STATS.names:
                if not np.array_equal(a["cov"][f], b["cov"][f]):
                    bad.append("cov." + f)
        elif not np.array_equal(a[name], b[name]):
            bad.append(name)
    assert not bad, "%s differ in %s" % (what, bad)


class BsoDbsnp(C.Structure):
    _fields_ = [("n", C.c_uint32), ("pos", C.c_void_p), ("flags", C.c_void_p), ("name_off", C.c_void_p), ("names", C.c_void_p)]


def dbsnp_arrays(entries):
    """entries: iterable of (position, flags, id bytes), any order -> (pos u32[], flags u8[], name_off u32[n + 1], names u8[])"""
    entries = sorted(entries, key=lambda e: e[0])
    pos = np.array([e[0] for e in entries], dtype=np.uint32)
    flags = np.array([e[1] for e in entries], dtype=np.uint8)
    off = np.zeros(len(entries) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(e[2]) for e in entries])
    names = np.frombuffer(b"".join(bytes(e[2]) for e in entries) + b"\0", dtype=np.uint8).copy()
    return pos, flags, off, names


def _print_block(fn, vcf, refcodes, x, rid, ctg_end, vcf_ids, all_positions, region=None, dbsnp=None, ann_fn=None, flat_db=False):
    vcf = _c(vcf, GT_VCF)
    sz = len(vcf)
    refcodes = _c(refcodes, np.uint8)
    assert len(refcodes) >= sz + 2
    ids = _c(VCF_IDS if vcf_ids is None else vcf_ids, np.int32)
    out = np.zeros(sz * 256 + 1024, dtype=np.uint8)
    nb, nr = C.c_size_t(0), C.c_size_t(0)
    if region is None and dbsnp is None:
        rc = fn(_p(vcf), C.c_uint32(sz), _p(refcodes), C.c_uint32(x), C.c_int(rid), C.c_uint32(ctg_end), _p(ids),
                C.c_int(1 if all_positions else 0), _p(out), C.c_size_t(len(out)), C.byref(nb), C.byref(nr))
    else:
        r0, r1 = region if region is not None else (0, 0)
        pos, flags, off, names = dbsnp if dbsnp is not None else dbsnp_arrays([])
        if flat_db:          # the harness takes the table as flat arguments
            rc = ann_fn(_p(vcf), C.c_uint32(sz), _p(refcodes), C.c_uint32(x), C.c_int(rid), C.c_uint32(ctg_end), _p(ids),
                        C.c_int(1 if all_positions else 0), C.c_uint32(r0), C.c_uint32(r1), C.c_uint32(len(pos)), _p(pos), _p(flags), _p(off), _p(names),
                        _p(out), C.c_size_t(len(out)), C.byref(nb), C.byref(nr))
        else:
            db = BsoDbsnp(len(pos), pos.ctypes.data, flags.ctypes.data, off.ctypes.data, names.ctypes.data)
            rc = ann_fn(_p(vcf), C.c_uint32(sz), _p(refcodes), C.c_uint32(x), C.c_int(rid), C.c_uint32(ctg_end), _p(ids),
                        C.c_int(1 if all_positions else 0), C.c_uint32(r0), C.c_uint32(r1), C.byref(db) if len(pos) else None,
                        _p(out), C.c_size_t(len(out)), C.byref(nb), C.byref(nr))
    if rc:
        raise RuntimeError("print_block failed: %d" % rc)
    return out[:nb.value].copy(), nr.value


def bcf_diff(a, b):
    """two BCF record streams -> dict(records_a, records_b, fixed_equal, identical, order_violations, first_diff)"""
    path = os.path.join(HERE, "liboracle.so")
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    a = _c(a, np.uint8)
    b = _c(b, np.uint8)
    out = (C.c_longlong * 6)()
    rc = lib.bso_bcf_diff(_p(a), C.c_size_t(len(a)), _p(b), C.c_size_t(len(b)), out)
    if rc:
        raise RuntimeError("bcf_diff: malformed record stream")
    return dict(zip(("records_a", "records_b", "fixed_equal", "identical", "order_violations", "first_diff"), [int(v) for v in out]))


class BsoParams(C.Structure):
    _fields_ = [("under_conv", C.c_double), ("over_conv", C.c_double), ("ref_bias", C.c_double),
                ("left_trim", C.c_uint32 * 2), ("right_trim", C.c_uint32 * 2), ("min_qual", C.c_uint8)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dt):
    a = np.ascontiguousarray(a, dtype=dt)
    return a


def build_oracle():
    """(Re)build liboracle.so and, when the reference tree is present, _ref/libbsref.so."""
    subprocess.run(["make", "-s", "-C", HERE], check=True, stdout=subprocess.DEVNULL)


class Oracle:
    """The C restatement (oracle/bs_oracle.c)."""

    def __init__(self, under_conv=0.01, over_conv=0.05, ref_bias=2.0, min_qual=20,
                 left_trim=(0, 0), right_trim=(0, 0)):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = C.CDLL(path)
        self.lib.bso_fisher.restype = C.c_double
        self.params = BsoParams(under_conv, over_conv, ref_bias, (C.c_uint32 * 2)(*left_trim),
                                (C.c_uint32 * 2)(*right_trim), min_qual)

    def qprob_table(self):
        t = np.zeros((44, 5), dtype=np.float64)
        self.lib.bso_qprob_table(_p(t))
        return t

    def lfact_table(self):
        t = np.zeros(256, dtype=np.float64)
        self.lib.bso_lfact_table(_p(t))
        return t

    def calc_gt_prob(self, counts, qual, rf):
        g = np.zeros(1, dtype=GT_METH)
        g["counts"][0] = counts
        g["qual"][0] = qual
        self.lib.bso_calc_gt_prob(_p(g), C.byref(self.params), C.c_int(int(rf)))
        return g[0]

    def fisher(self, tab):
        t = np.ascontiguousarray(tab, dtype=np.int32)
        return float(self.lib.bso_fisher(_p(t)))

    def call_sites(self, pileup, ref, nthreads=1):
        pileup = _c(pileup, PILEUP)
        ref = _c(ref, np.uint8)
        n = len(pileup)
        out = np.zeros(n, dtype=GT_METH)
        skip = np.zeros(n, dtype=np.uint8)
        self.lib.bso_call_sites(_p(pileup), _p(ref), C.c_size_t(n), C.byref(self.params), _p(out), _p(skip),
                                C.c_int(nthreads))
        return out, skip

    def synth_sites(self, seed, first, n, mean_depth=30.0, nthreads=1):
        p = np.zeros(n, dtype=PILEUP)
        r = np.zeros(n, dtype=np.uint8)
        self.lib.bso_synth_sites(C.c_uint64(seed), C.c_uint64(first), C.c_size_t(n), C.c_double(mean_depth), _p(p), _p(r),
                                 C.c_int(nthreads))
        return p, r

    def pileup_block(self, templates, bases, x, y):
        templates = _c(templates, TEMPLATE)
        bases = _c(bases, np.uint8)
        out = np.zeros(y - x + 1, dtype=PILEUP)
        self.lib.bso_pileup_block(_p(templates), C.c_size_t(len(templates)), _p(bases), C.c_uint32(x), C.c_uint32(y),
                                  C.byref(self.params), _p(out))
        return out

    def normalise_block(self, templates, bases, misms):
        templates = _c(templates, TEMPLATE)
        bases = _c(bases, np.uint8)
        misms = _c(misms, MISMS)
        cap = int(len(bases) + (misms["size"].sum() if len(misms) else 0) + 64)
        out_t = np.zeros(len(templates), dtype=TEMPLATE)
        out_b = np.zeros(cap, dtype=np.uint8)
        used = C.c_size_t(0)
        rc = self.lib.bso_normalise_block(_p(templates), C.c_size_t(len(templates)), _p(bases), _p(misms),
                                          C.byref(self.params), _p(out_t), _p(out_b), C.c_size_t(cap), C.byref(used))
        if rc:
            raise RuntimeError("bso_normalise_block failed: %d" % rc)
        return out_t, out_b[:used.value].copy()

    def decode_records(self, bam, mapq_thresh=20, max_template_len=1000, keep_unmatched=False, ignore_dup=False):
        """raw BAM records -> (RECORD[], packed bases, misms): the restatement of get_next_align_details()"""
        return _decode_records(self.lib.bso_decode_records, bam, mapq_thresh, max_template_len, keep_unmatched, ignore_dup)

    def read_input(self, bam, target_len, ctg_codes=None, mapq_thresh=20, max_template_len=1000, keep_unmatched=False,
                   ignore_duplicates=False, keep_duplicates=False, run_chain=False):
        """raw BAM records -> (BLOCK[], TEMPLATE[], bases, misms, gt_vcf[]): the restatement of read_input(); with
        run_chain the blocks also go through the restated process_template_vector / call_genotypes_ML"""
        self.lib.bso_set_params(C.byref(self.params))
        return _read_input(self.lib.bso_read_input, bam, target_len, ctg_codes, mapq_thresh, max_template_len,
                           keep_unmatched, ignore_duplicates, keep_duplicates, run_chain)

    def print_block(self, vcf, refcodes, x, rid=0, ctg_end=0xffffffff, vcf_ids=None, all_positions=False, region=None, dbsnp=None):
        """the restatement of the reference's writer over one block of gt_vcf[] -> (BCF record bytes, number of records);
        region = (start, stop) of ctg->curr_reg, dbsnp = dbsnp_arrays(...) of the contig's index entries"""
        return _print_block(self.lib.bso_print_block, vcf, refcodes, x, rid, ctg_end, vcf_ids, all_positions, region, dbsnp,
                            self.lib.bso_print_block_ann)

    def stats_block(self, vcf, refcodes, x, ctg_end=0xffffffff, all_positions=False, region=None, dbsnp=None, gc=None, start_pos=1,
                    stats=None, state=None):
        """the writer's --report-file statistics of one block (restatement of src/print_vcf.c:382-526), added to `stats`
        (a SITE_STATS record array of length 1; made when None).  state = [prev_cpg_x, prev_cpg_flt] carried between blocks."""
        vcf = _c(vcf, GT_VCF)
        refcodes = _c(refcodes, np.uint8)
        if stats is None:
            stats = np.zeros(1, dtype=SITE_STATS)
        if state is None:
            state = np.zeros(2, dtype=np.uint32)
        r0, r1 = region if region is not None else (0, 0)
        pos, flags, off, names = dbsnp if dbsnp is not None else dbsnp_arrays([])
        db = BsoDbsnp(len(pos), pos.ctypes.data, flags.ctypes.data, off.ctypes.data, names.ctypes.data)
        gcv = _c(gc, np.uint8) if gc is not None else None
        self.lib.bso_stats_block(_p(vcf), C.c_uint32(len(vcf)), _p(refcodes), C.c_uint32(x), C.c_uint32(ctg_end), C.c_int(1 if all_positions else 0),
                                 C.c_uint32(r0), C.c_uint32(r1), C.byref(db) if len(pos) else None, _p(gcv) if gcv is not None else None,
                                 C.c_int(len(gcv) if gcv is not None else 0), C.c_uint32(start_pos), _p(stats), _p(state))
        return stats, state

    def profile_enable(self, on=True):
        """--report-file side channels (process-wide in the library); process_block then wants codes for [x, y + 1]"""
        self.lib.bso_profile_enable(C.c_int(1 if on else 0))
        self._profile_on = bool(on)

    def profile_reset(self):
        self.lib.bso_profile_reset()

    def profile_read(self):
        pr = BsoProfile()
        self.lib.bso_profile_read(C.byref(pr))
        conv = np.ctypeslib.as_array(pr.conv_cts).reshape(PROFILE_MAX, 4)
        return profile_dict(pr.used, conv, list(pr.base_filter), list(pr.filter_cts), list(pr.filter_bases))

    def process_block(self, templates, bases, misms, refcodes, y):
        """refcodes: codes for positions [x, y] where x = max(first-2, 1)."""
        templates = _c(templates, TEMPLATE)
        bases = _c(bases, np.uint8)
        misms = _c(misms, MISMS)
        refcodes = _c(refcodes, np.uint8)
        first = int(templates[0]["forward_position"]) or int(templates[0]["reverse_position"])
        x = first - 2 if first > 2 else 1
        sz = y - x + 1
        assert len(refcodes) >= sz + (1 if getattr(self, "_profile_on", False) else 0)
        pile = np.zeros(sz, dtype=PILEUP)
        vcf = np.zeros(sz, dtype=GT_VCF)
        xo = C.c_uint32(0)
        rc = self.lib.bso_process_block(_p(templates), C.c_size_t(len(templates)), _p(bases), _p(misms), _p(refcodes),
                                        C.c_uint32(y), C.byref(self.params), C.byref(xo), _p(pile), _p(vcf))
        if rc:
            raise RuntimeError("bso_process_block failed: %d" % rc)
        return xo.value, pile, vcf


_REF_SINGLETON = None
_REF_LIBNAME = "libbsref.so"


def reference_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libbsref.so"))


def dropin_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libbsref_gpu.so"))


class Reference:
    """The reference's own compiled hot path (one instance per library: it owns global thread state)."""
    LIBNAME = "libbsref.so"
    _instances = {}

    def __new__(cls, *a, **k):
        inst = Reference._instances.get(cls.LIBNAME)
        if inst is None:
            inst = super().__new__(cls)
            inst._started = False
            Reference._instances[cls.LIBNAME] = inst
        return inst

    def __init__(self, under_conv=0.01, over_conv=0.05, ref_bias=2.0, min_qual=20, calc_threads=1,
                 left_trim=(0, 0), right_trim=(0, 0)):
        lt = (C.c_uint32 * 2)(*left_trim)
        rt = (C.c_uint32 * 2)(*right_trim)
        if not self._started:
            path = os.path.join(HERE, "_ref", self.LIBNAME)
            if not os.path.exists(path):
                raise FileNotFoundError(path)
            self.lib = C.CDLL(path)
            self.lib.bsref_fisher.restype = C.c_double
            self.lib.bsref_last_call_seconds.restype = C.c_double
            assert self.lib.bsref_sizeof(0) == PILEUP.itemsize
            assert self.lib.bsref_sizeof(1) == GT_METH.itemsize
            assert self.lib.bsref_sizeof(2) == GT_VCF.itemsize
            assert self.lib.bsref_sizeof(3) == TEMPLATE.itemsize
            rc = self.lib.bsref_init(C.c_double(under_conv), C.c_double(over_conv), C.c_double(ref_bias),
                                     C.c_int(min_qual), C.c_int(max(calc_threads - 1, 0)), lt, rt)
            assert rc == 0
            self._started = True
        else:
            self.lib.bsref_set_params(C.c_double(under_conv), C.c_double(over_conv), C.c_double(ref_bias),
                                      C.c_int(min_qual), lt, rt)

    def calc_gt_prob(self, counts, qual, rf):
        g = np.zeros(1, dtype=GT_METH)
        c = np.ascontiguousarray(counts, dtype=np.uint64)
        q = np.ascontiguousarray(qual, dtype=np.int32)
        self.lib.bsref_calc_gt_prob(_p(c), _p(q), C.c_int(int(rf)), _p(g))
        return g[0]

    def calc_gt_prob_batch(self, counts, qual, rf):
        c = np.ascontiguousarray(counts, dtype=np.uint64)
        q = np.ascontiguousarray(qual, dtype=np.int32)
        r = np.ascontiguousarray(rf, dtype=np.uint8)
        g = np.zeros(len(r), dtype=GT_METH)
        self.lib.bsref_calc_gt_prob_batch(_p(c), _p(q), _p(r), C.c_size_t(len(r)), _p(g))
        return g

    def fisher(self, tab):
        t = np.ascontiguousarray(tab, dtype=np.int32)
        return float(self.lib.bsref_fisher(_p(t)))

    def call_sites(self, pileup, ref, nthreads=1):
        """pileup[] -> gt_meth[] through the reference's calc_gt_prob()/fisher() (see ref_harness.c)."""
        pileup = _c(pileup, PILEUP)
        ref = _c(ref, np.uint8)
        n = len(pileup)
        out = np.zeros(n, dtype=GT_METH)
        skip = np.zeros(n, dtype=np.uint8)
        rc = self.lib.bsref_call_sites_mt(_p(pileup), _p(ref), C.c_size_t(n), _p(out), _p(skip), C.c_int(nthreads))
        assert rc == 0
        return out, skip

    def lfact_table(self):
        t = np.zeros(256, dtype=np.float64)
        self.lib.bsref_lfact_table(_p(t))
        return t

    def last_call_seconds(self):
        """wall time of the last call_genotypes_ML / process_template_vector until every site was ready"""
        return float(self.lib.bsref_last_call_seconds())

    def call_block(self, templates, bases, refcodes, x, y):
        """Normalised templates -> (pileup[], gt_vcf[]) via call_genotypes_ML.  refcodes covers [x, y+2]."""
        templates = _c(templates, TEMPLATE)
        bases = _c(bases, np.uint8)
        sz = y - x + 1
        rc_ = np.zeros(sz + 2, dtype=np.uint8)
        rc_[:min(len(refcodes), sz + 2)] = refcodes[:sz + 2]
        pile = np.zeros(sz, dtype=PILEUP)
        vcf = np.zeros(sz, dtype=GT_VCF)
        rc = self.lib.bsref_call_block(_p(templates), C.c_size_t(len(templates)), _p(bases), _p(rc_),
                                       C.c_uint32(x), C.c_uint32(y), _p(pile), _p(vcf))
        if rc:
            raise RuntimeError("bsref_call_block failed: %d" % rc)
        return pile, vcf

    def print_block(self, vcf, refcodes, x, rid=0, ctg_end=0xffffffff, vcf_ids=None, all_positions=False, region=None, dbsnp=None):
        """one block of gt_vcf[] through the reference's print_vcf_entry / flush_vcf_entries (src/print_vcf.c) as the print
        thread runs them; refcodes covers [x, x + len(vcf) + 1].  Returns (BCF record bytes, number of records)."""
        return _print_block(self.lib.bsref_print_block, vcf, refcodes, x, rid, ctg_end, vcf_ids, all_positions, region, dbsnp,
                            self.lib.bsref_print_block_ann, flat_db=True)

    def writer_stats(self, on=True, gc=None, start_pos=1):
        """blocks printed from now on count into the reference's bs_stats (src/print_vcf.c:382-526); gc = the contig's GC bins"""
        self._gc_keep = _c(gc, np.uint8) if gc is not None else None
        self.lib.bsref_writer_stats(C.c_int(1 if on else 0), _p(self._gc_keep) if self._gc_keep is not None else None,
                                    C.c_int(len(self._gc_keep) if self._gc_keep is not None else 0), C.c_uint32(start_pos))

    def writer_stats_reset(self):
        self.lib.bsref_writer_stats_reset()

    def writer_stats_read(self):
        """(SITE_STATS record array of length 1, per-contig counters u64[6][2] of the block printed last)"""
        st = np.zeros(1, dtype=SITE_STATS)
        self.lib.bsref_writer_stats_read(_p(st))
        cs = np.zeros(12, dtype=np.uint64)
        self.lib.bsref_writer_ctg_stats_read(_p(cs))
        return st, cs.reshape(6, 2)

    def stats_enable(self, on=True):
        """give the reference a bs_stats (what --report-file does): meth_profile() and the tallies become live"""
        self.lib.bsref_stats_enable(C.c_int(1 if on else 0))

    def stats_reset(self):
        self.lib.bsref_stats_reset()

    def stats_read(self):
        """-> profile_dict of the reference's bs_stats"""
        conv = np.zeros((1 << 16, 4), dtype=np.uint64)
        bf = np.zeros(5, dtype=np.uint64)
        fc = np.zeros(15, dtype=np.uint64)
        fb = np.zeros(15, dtype=np.uint64)
        self.lib.bsref_stats_read.restype = C.c_uint32
        used = self.lib.bsref_stats_read(_p(conv), C.c_size_t(len(conv)), _p(bf), _p(fc), _p(fb))
        return profile_dict(used, conv, bf, fc, fb)

    def decode_records(self, bam, mapq_thresh=20, max_template_len=1000, keep_unmatched=False, ignore_dup=False):
        """raw BAM records through the reference's get_next_align_details() (src/input_sam.c:222)"""
        return _decode_records(self.lib.bsref_decode_records, bam, mapq_thresh, max_template_len, keep_unmatched, ignore_dup)

    def read_input(self, bam, target_len, ctg_codes=None, mapq_thresh=20, max_template_len=1000, keep_unmatched=False,
                   ignore_duplicates=False, keep_duplicates=False, run_chain=False):
        """raw BAM records through the reference's read_input() (src/get_template_vector.c:49); with run_chain every
        block continues through process_template_vector() and call_genotypes_ML() as in the reference binary"""
        return _read_input(self.lib.bsref_read_input, bam, target_len, ctg_codes, mapq_thresh, max_template_len,
                           keep_unmatched, ignore_duplicates, keep_duplicates, run_chain)

    def process_block(self, templates, bases, misms, ctg_codes, y):
        """Raw templates -> (x, pileup[], gt_vcf[], ref[], normalised templates, normalised bases)."""
        templates = _c(templates, TEMPLATE)
        bases = _c(bases, np.uint8)
        misms = _c(misms, MISMS)
        ctg_codes = _c(ctg_codes, np.uint8)
        first = int(templates[0]["forward_position"]) or int(templates[0]["reverse_position"])
        x = first - 2 if first > 2 else 1
        sz = y - x + 1
        pile = np.zeros(sz, dtype=PILEUP)
        vcf = np.zeros(sz, dtype=GT_VCF)
        ref = np.zeros(sz, dtype=np.uint8)
        cap = int(len(bases) + (misms["size"].sum() if len(misms) else 0) + 64)
        nt = np.zeros(len(templates), dtype=TEMPLATE)
        nb = np.zeros(cap, dtype=np.uint8)
        xo = C.c_uint32(0)
        rc = self.lib.bsref_process_block(_p(templates), C.c_size_t(len(templates)), _p(bases), _p(misms),
                                          _p(ctg_codes), C.c_uint32(len(ctg_codes)), C.c_uint32(y), C.byref(xo),
                                          _p(pile), _p(vcf), _p(ref), _p(nt), _p(nb), C.c_size_t(cap))
        if rc:
            raise RuntimeError("bsref_process_block failed: %d" % rc)
        used = int(nt["read_len"].sum())
        return xo.value, pile, vcf, ref, nt, nb[:used].copy()


class ReferenceWithGpuDropin(Reference):
    """The reference's process_template_vector etc. with src/call_genotypes.c replaced by the product's drop-in
    (bs_call_b200/csrc/bsgpu_dropin.c over libbsgpu.so): oracle/_ref/libbsref_gpu.so.  Needs a GPU."""
    LIBNAME = "libbsref_gpu.so"


def seam_available(name):
    return os.path.exists(os.path.join(HERE, "_ref", "libbsref_%s.so" % name))


class ReferenceWithSeamB(Reference):
    """The reference with src/process_template.c and src/call_genotypes.c replaced by the product's
    bs_call_b200/csrc/bsgpu_seam_template.c (process_template_vector on the device): oracle/_ref/libbsref_seamB.so.
    read_input, get_next_align_details and the rest are the reference's own objects.  Needs a GPU."""
    LIBNAME = "libbsref_seamB.so"


class ReferenceWithSeamC(Reference):
    """The reference with src/get_template_vector.c, src/process_template.c and src/call_genotypes.c replaced by the product's
    bs_call_b200/csrc/bsgpu_seam_reader.c (read_input feeding a streaming session of the device library):
    oracle/_ref/libbsref_seamC.so.  get_next_align_details' record source (sam_read1), the sequence store, the print-thread
    protocol and bcf_write are the harness's stand-ins for htslib / src/process.c.  Needs a GPU."""
    LIBNAME = "libbsref_seamC.so"

    def seam_read_input(self, bam, target_len, ctg_codes, mapq_thresh=20, max_template_len=1000, keep_unmatched=False,
                        ignore_duplicates=False, keep_duplicates=False, vcf_ids=None, all_positions=False):
        """the product's read_input() under the harness's print thread -> (blocks, gt_vcf[], BCF bytes, BCF records); which
        of gt_vcf[] (seam C) or records (seam D) comes back is decided by BSGPU_SEAM_RECORDS in the environment"""
        bam = _c(bam, np.uint8)
        target_len = _c(target_len, np.uint32)
        codes = [_c(c, np.uint8) for c in ctg_codes]
        ptrs = (C.c_void_p * len(codes))(*[c.ctypes.data for c in codes])
        ids = _c(VCF_IDS if vcf_ids is None else vcf_ids, np.int32)
        nrec_max = len(bam) // 36 + 8
        blocks = np.zeros(nrec_max, dtype=BLOCK)
        vcf = np.zeros(int(target_len.sum()) + 8, dtype=GT_VCF)
        out = np.zeros(int(target_len.sum()) * 160 + 4096, dtype=np.uint8)
        nb, nv, ob, orec = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        rc = self.lib.bsref_seam_read_input(_p(bam), C.c_size_t(len(bam)), C.c_int(len(codes)), _p(target_len), ptrs,
                                            C.c_int(mapq_thresh), C.c_uint32(max_template_len), C.c_int(int(keep_unmatched)),
                                            C.c_int(int(ignore_duplicates)), C.c_int(int(keep_duplicates)), _p(ids), C.c_int(int(all_positions)),
                                            _p(blocks), C.c_size_t(len(blocks)), C.byref(nb), _p(vcf), C.c_size_t(len(vcf)), C.byref(nv),
                                            _p(out), C.c_size_t(len(out)), C.byref(ob), C.byref(orec))
        if rc:
            raise RuntimeError("bsref_seam_read_input failed: %d" % rc)
        return blocks[:nb.value].copy(), vcf[:nv.value].copy(), out[:ob.value].copy(), orec.value
