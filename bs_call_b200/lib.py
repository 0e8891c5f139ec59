"""ctypes binding of the C ABI in include/bsgpu.h (libbsgpu.so, built in-tree by bs_call_b200/csrc/Makefile).

This is the host-side mirror used by the tests and by bench.py: numpy arrays in the reference's record layouts
go in and come out, exactly what a C host would hand to the same entry points.  There is no fallback of any
kind: a missing library or a missing sm_100 device raises.
"""
import ctypes as C
import os

import numpy as np

from .records import PILEUP, GT_METH, GT_VCF, SEG, TEMPLATE, MISMS, RECORD, BLOCK

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BSGPU_LIB_PATH") or os.path.join(_HERE, "libbsgpu.so")      # (the override is for A/B builds of the kernels)

BSGPU_OK = 1
BSGPU_FAIL = -1

# every symbol include/bsgpu.h declares (tests/test_abi.py checks the header and this list against the .so)
EXPORTS = [
    "bsgpu_default_params", "bsgpu_init", "bsgpu_destroy", "bsgpu_last_error", "bsgpu_get_stats", "bsgpu_version",
    "bsgpu_sync", "bsgpu_guard_read", "bsgpu_debug_redzones", "bsgpu_bam_on_contig", "bsgpu_host_alloc", "bsgpu_host_free",
    "bsgpu_call_sites", "bsgpu_pileup_block", "bsgpu_call_block", "bsgpu_stage_bound", "bsgpu_stage_templates",
    "bsgpu_process_block", "bsgpu_profile_enable", "bsgpu_profile_read", "bsgpu_build_blocks_tally",
    "bsgpu_site_stats_enable", "bsgpu_set_contig_gc", "bsgpu_site_stats_read",
    "bsgpu_call_bam_bcf", "bsgpu_default_bcf_params", "bsgpu_set_contig_annotation", "bsgpu_bcf_block", "bsgpu_bcf_block_dev", "bsgpu_call_block_bcf", "bsgpu_call_sites_bcf",
    "bsgpu_default_reader_params", "bsgpu_decode_records", "bsgpu_build_blocks", "bsgpu_call_bam",
    "bsgpu_bam_open", "bsgpu_bam_set_contig", "bsgpu_bam_feed", "bsgpu_bam_reserve", "bsgpu_bam_commit", "bsgpu_bam_finish", "bsgpu_bam_cut", "bsgpu_bam_rewind", "bsgpu_bam_drain",
    "bsgpu_bam_release", "bsgpu_bam_progress", "bsgpu_bam_close",
    "bsgpu_call_sites_dev", "bsgpu_call_sites_vcf_dev", "bsgpu_pileup_block_dev", "bsgpu_call_block_dev",
    "bsgpu_synth_sites_dev", "bsgpu_synth_block_nseg", "bsgpu_synth_block_dev",
    "bsgpu_synth_bam_bytes", "bsgpu_synth_bam_dev", "bsgpu_synth_ref_dev",
    "bsgpu_math_probe", "bsgpu_wire_pack", "bsgpu_wire_expand",
]


class Params(C.Structure):
    _fields_ = [("under_conv", C.c_double), ("over_conv", C.c_double), ("ref_bias", C.c_double),
                ("left_trim", C.c_uint32 * 2), ("right_trim", C.c_uint32 * 2), ("min_qual", C.c_uint8),
                ("device", C.c_int32)]


class ReaderParams(C.Structure):
    _fields_ = [("max_template_len", C.c_uint32), ("mapq_thresh", C.c_uint8), ("keep_unmatched", C.c_uint8),
                ("ignore_duplicates", C.c_uint8), ("keep_duplicates", C.c_uint8)]


def reader_params(mapq_thresh=20, max_template_len=1000, keep_unmatched=False, ignore_duplicates=False, keep_duplicates=False):
    return ReaderParams(int(max_template_len), int(mapq_thresh), int(bool(keep_unmatched)), int(bool(ignore_duplicates)),
                        int(bool(keep_duplicates)))


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("sites", C.c_uint64), ("sites_called", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("qsum_overflow", C.c_uint64),
                ("bam_decode_s", C.c_double), ("bam_build_s", C.c_double), ("bam_call_s", C.c_double),
                ("near_tie_sites", C.c_uint64), ("exact_tie_sites", C.c_uint64), ("near_qual_sites", C.c_uint64), ("near_fs_sites", C.c_uint64),
                ("long_segments", C.c_uint64), ("wire_sites", C.c_uint64), ("wire_refetched_chunks", C.c_uint64)]


PROFILE_MAX = 1024


class Profile(C.Structure):
    """bsgpu_profile (include/bsgpu.h): the --report-file side channels of the path"""
    _fields_ = [("conv_cts", (C.c_uint64 * 4) * PROFILE_MAX), ("used", C.c_uint32), ("pad_", C.c_uint32),
                ("base_filter", C.c_uint64 * 5), ("filter_cts", C.c_uint64 * 15), ("filter_bases", C.c_uint64 * 15)]


BCF_MAX_RECORD = 384


class BamResult(C.Structure):
    """bsgpu_bam_result (include/bsgpu.h): the results of one batch of a streaming session, lent out in pinned memory"""
    _fields_ = [("id", C.c_uint64), ("blocks", C.c_void_p), ("nblocks", C.c_size_t), ("data", C.c_void_p), ("nbytes", C.c_size_t),
                ("nrec", C.c_size_t), ("bytes_in", C.c_size_t), ("records_in", C.c_size_t), ("finished", C.c_int), ("pad_", C.c_int)]


class BamProgress(C.Structure):
    _fields_ = [("bytes_fed", C.c_uint64), ("bytes_done", C.c_uint64), ("records_done", C.c_uint64), ("batches", C.c_uint64),
                ("carry_bytes", C.c_uint64), ("empty_batches", C.c_uint64), ("pinned_bytes", C.c_uint64)]


class BcfParams(C.Structure):
    """bsgpu_bcf_params (include/bsgpu.h): header dictionary ids, contig id and end, -A"""
    _fields_ = [("ids", C.c_int32 * 16), ("rid", C.c_int32), ("ctg_end", C.c_uint32), ("all_positions", C.c_uint8), ("pad_", C.c_uint8 * 3),
                ("reg_start", C.c_uint32), ("reg_stop", C.c_uint32), ("dbsnp", C.c_void_p)]


class Dbsnp(C.Structure):
    """bsgpu_dbsnp (include/bsgpu.h): what dbSNP_lookup_name() answers for the known positions of one contig"""
    _fields_ = [("n", C.c_uint32), ("pos", C.c_void_p), ("flags", C.c_void_p), ("name_off", C.c_void_p), ("names", C.c_void_p)]


def dbsnp(pos, flags, name_off, names):
    """-> Dbsnp over numpy arrays (kept alive on the returned object)"""
    a = [np.ascontiguousarray(pos, dtype=np.uint32), np.ascontiguousarray(flags, dtype=np.uint8),
         np.ascontiguousarray(name_off, dtype=np.uint32), np.ascontiguousarray(names, dtype=np.uint8)]
    d = Dbsnp(len(a[0]), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data)
    d._keep = a
    return d


def bcf_params(ids=None, rid=0, ctg_end=0xffffffff, all_positions=False, region=None, dbsnp=None):
    p = BcfParams()
    load().bsgpu_default_bcf_params(C.byref(p))
    if ids is not None:
        p.ids = (C.c_int32 * 16)(*[int(v) for v in ids])
    p.rid, p.ctg_end, p.all_positions = int(rid), int(ctg_end), 1 if all_positions else 0
    if region is not None:
        p.reg_start, p.reg_stop = int(region[0]), int(region[1])
    if dbsnp is not None:
        p.dbsnp = C.addressof(dbsnp)
        p._keep = dbsnp
    return p


class BsGpuError(RuntimeError):
    pass


_lib = None


def load():
    """Load libbsgpu.so; raises if it has not been built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BsGpuError("libbsgpu.so is missing at %s: build it with bs_call_b200/csrc/Makefile "
                         "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.bsgpu_last_error.restype = C.c_char_p
    lib.bsgpu_host_alloc.restype = C.c_void_p
    lib.bsgpu_host_alloc.argtypes = [C.c_size_t]
    lib.bsgpu_host_free.argtypes = [C.c_void_p]
    lib.bsgpu_stage_bound.restype = C.c_size_t
    lib.bsgpu_synth_bam_bytes.restype = C.c_size_t
    lib.bsgpu_synth_bam_bytes.argtypes = [C.c_size_t, C.c_uint32]
    lib.bsgpu_synth_block_nseg.restype = C.c_size_t
    lib.bsgpu_synth_block_nseg.argtypes = [C.c_uint32, C.c_uint32, C.c_double]
    lib.bsgpu_init.argtypes = [C.POINTER(Params), C.POINTER(C.c_void_p)]
    lib.bsgpu_destroy.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(int(a))         # raw device / host address


class HostBuffer:
    """Page-locked host array (bsgpu_host_alloc) viewed as a numpy array of `dtype`."""

    def __init__(self, n, dtype):
        lib = load()
        self.dtype = np.dtype(dtype)
        self.nbytes = max(int(n) * self.dtype.itemsize, 16)
        self.addr = lib.bsgpu_host_alloc(C.c_size_t(self.nbytes))
        if not self.addr:
            raise BsGpuError(lib.bsgpu_last_error().decode())
        buf = (C.c_uint8 * self.nbytes).from_address(self.addr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(n))

    def free(self):
        if self.addr:
            self.array = None
            load().bsgpu_host_free(C.c_void_p(self.addr))
            self.addr = None


class BsGpu:
    """One context = one device.  Methods map one-to-one onto the C entry points."""

    def __init__(self, under_conv=0.01, over_conv=0.05, ref_bias=2.0, min_qual=20, left_trim=(0, 0),
                 right_trim=(0, 0), device=0):
        self.lib = load()
        self.params = Params(under_conv, over_conv, ref_bias, (C.c_uint32 * 2)(*left_trim),
                             (C.c_uint32 * 2)(*right_trim), min_qual, device)
        self.ctx = C.c_void_p()
        self._check(self.lib.bsgpu_init(C.byref(self.params), C.byref(self.ctx)))

    def _check(self, rc):
        if rc != BSGPU_OK:
            raise BsGpuError(self.lib.bsgpu_last_error().decode())

    def close(self):
        if self.ctx:
            self.lib.bsgpu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.lib.bsgpu_sync(self.ctx))

    def guard_read(self, reset=True):
        """-> (kinds, ids) of the sites inside a guard band since the last reset (kind 1 call, 2 QUAL / GQ, 3 FS)"""
        ids = np.zeros(65536, dtype=np.uint64)
        n = C.c_size_t(0)
        self._check(self.lib.bsgpu_guard_read(self.ctx, _ptr(ids), C.c_size_t(len(ids)), C.byref(n), C.c_int(1 if reset else 0)))
        ids = ids[:n.value]
        return (ids >> np.uint64(56)).astype(np.int64), (ids & np.uint64((1 << 56) - 1)).astype(np.int64)

    def debug_redzones(self):
        """-> (ok, zones checked, zones written into); only meaningful with BSGPU_REDZONE=1 set before the library was loaded"""
        a, b = C.c_ulonglong(0), C.c_ulonglong(0)
        rc = self.lib.bsgpu_debug_redzones(C.byref(a), C.byref(b))
        return rc == 1, a.value, b.value

    def stats(self):
        s = Stats()
        self._check(self.lib.bsgpu_get_stats(self.ctx, C.byref(s)))
        return {k: (float(getattr(s, k)) if t is C.c_double else int(getattr(s, k))) for k, t in Stats._fields_}

    # ---- host-buffer entry points -------------------------------------------------------------
    def call_sites(self, pileup, ref, out=None, skip=None):
        pileup = np.ascontiguousarray(pileup, dtype=PILEUP)
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        n = len(pileup)
        assert len(ref) == n
        out = np.empty(n, dtype=GT_METH) if out is None else out
        skip = np.empty(n, dtype=np.uint8) if skip is None else skip
        self._check(self.lib.bsgpu_call_sites(self.ctx, _ptr(pileup), _ptr(ref), C.c_size_t(n), _ptr(out), _ptr(skip)))
        return out, skip

    def pileup_block(self, segs, bases, x, sz, out=None):
        segs = np.ascontiguousarray(segs, dtype=SEG)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        out = np.empty(sz, dtype=PILEUP) if out is None else out
        self._check(self.lib.bsgpu_pileup_block(self.ctx, _ptr(segs), C.c_size_t(len(segs)), _ptr(bases),
                                                C.c_size_t(len(bases)), C.c_uint32(x), C.c_uint32(sz), _ptr(out)))
        return out

    def call_block(self, segs, bases, ref, x, sz, out=None):
        segs = np.ascontiguousarray(segs, dtype=SEG)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        assert len(ref) >= sz
        out = np.empty(sz, dtype=GT_VCF) if out is None else out
        self._check(self.lib.bsgpu_call_block(self.ctx, _ptr(segs), C.c_size_t(len(segs)), _ptr(bases),
                                              C.c_size_t(len(bases)), _ptr(ref), C.c_uint32(x), C.c_uint32(sz), _ptr(out)))
        return out

    def process_block(self, templates, bases, misms, ref, y, out=None):
        templates = np.ascontiguousarray(templates, dtype=TEMPLATE)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        misms = np.ascontiguousarray(misms, dtype=MISMS)
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        first = int(templates[0]["forward_position"]) or int(templates[0]["reverse_position"])
        x = first - 2 if first > 2 else 1
        sz = y - x + 1
        assert len(ref) >= sz + (1 if getattr(self, "_profile_on", False) else 0)
        out = np.empty(sz, dtype=GT_VCF) if out is None else out
        xo = C.c_uint32(0)
        self._check(self.lib.bsgpu_process_block(self.ctx, _ptr(templates), C.c_size_t(len(templates)), _ptr(bases),
                                                 C.c_size_t(len(bases)), _ptr(misms), C.c_size_t(len(misms)), _ptr(ref),
                                                 C.c_uint32(y), C.byref(xo), _ptr(out)))
        return xo.value, out

    # ---- writer side ------------------------------------------------------------------------------
    def bcf_block(self, vcf, ref, x, params=None, out=None):
        """gt_vcf[] of one block + reference codes of [x, x + len(vcf) + 1] -> (BCF record bytes, number of records)"""
        vcf = np.ascontiguousarray(vcf, dtype=GT_VCF)
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        sz = len(vcf)
        assert len(ref) >= sz + 2
        p = params or bcf_params()
        out = np.empty(sz * BCF_MAX_RECORD + 64, dtype=np.uint8) if out is None else out
        nb, nr = C.c_size_t(0), C.c_size_t(0)
        self._check(self.lib.bsgpu_bcf_block(self.ctx, _ptr(vcf), _ptr(ref), C.c_uint32(x), C.c_uint32(sz), C.byref(p), _ptr(out),
                                             C.c_size_t(len(out)), C.byref(nb), C.byref(nr)))
        return out[:nb.value], nr.value

    def bcf_block_dev(self, d_vcf, d_ref, x, sz, d_out, out_cap, params=None, stream=0):
        """device-resident gt_vcf[] -> records in device memory; returns (bytes, records) after waiting for the stream"""
        p = params or bcf_params()
        nb, nr = C.c_size_t(0), C.c_size_t(0)
        self._check(self.lib.bsgpu_bcf_block_dev(self.ctx, C.c_void_p(d_vcf), C.c_void_p(d_ref), C.c_uint32(x), C.c_uint32(sz), C.byref(p),
                                                 C.c_void_p(d_out), C.c_size_t(out_cap), C.byref(nb), C.byref(nr), C.c_void_p(stream)))
        return nb.value, nr.value

    def call_block_bcf(self, segs, bases, ref, x, sz, params=None, out=None):
        """sorted segments + reference codes of [x, x + sz + 1] -> (BCF record bytes, number of records)"""
        segs = np.ascontiguousarray(segs, dtype=SEG)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        assert len(ref) >= sz + 2
        p = params or bcf_params()
        out = np.empty(sz * BCF_MAX_RECORD + 64, dtype=np.uint8) if out is None else out
        nb, nr = C.c_size_t(0), C.c_size_t(0)
        self._check(self.lib.bsgpu_call_block_bcf(self.ctx, _ptr(segs), C.c_size_t(len(segs)), _ptr(bases), C.c_size_t(len(bases)), _ptr(ref),
                                                  C.c_uint32(x), C.c_uint32(sz), C.byref(p), _ptr(out), C.c_size_t(len(out)), C.byref(nb), C.byref(nr)))
        return out[:nb.value], nr.value

    def call_sites_bcf(self, pileup, ref, x, params=None, out=None):
        """pileup[] of consecutive sites from position x (one block) + reference codes of [x, x + n + 1] -> (BCF bytes, records)"""
        pileup = np.ascontiguousarray(pileup, dtype=PILEUP)
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        n = len(pileup)
        assert len(ref) >= n + 2
        p = params or bcf_params()
        out = np.empty(n * BCF_MAX_RECORD + 64, dtype=np.uint8) if out is None else out
        nb, nr = C.c_size_t(0), C.c_size_t(0)
        self._check(self.lib.bsgpu_call_sites_bcf(self.ctx, _ptr(pileup), _ptr(ref), C.c_size_t(n), C.c_uint32(x), C.byref(p), _ptr(out),
                                                  C.c_size_t(len(out)), C.byref(nb), C.byref(nr)))
        return out[:nb.value], nr.value

    # ---- --report-file side channels -------------------------------------------------------------
    def set_contig_annotation(self, tid, region=None, dbsnp=None):
        """region (start, stop) and / or Dbsnp of contig `tid` for call_bam_bcf and sessions (bsgpu_set_contig_annotation)"""
        r0, r1 = region if region is not None else (0, 0)
        self._check(self.lib.bsgpu_set_contig_annotation(self.ctx, C.c_int(int(tid)), C.c_uint32(int(r0)), C.c_uint32(int(r1)),
                                                         C.byref(dbsnp) if dbsnp is not None else None))

    def profile_enable(self, on=True):
        """gather the non-CpG conversion profile and the base / read tallies in process_block and call_bam
        (process_block then wants reference codes for [x, y + 1])"""
        self._check(self.lib.bsgpu_profile_enable(self.ctx, C.c_int(1 if on else 0)))
        self._profile_on = bool(on)

    def profile_read(self, reset=False):
        """-> dict(used, conv[used, 4], base_filter[5], filter_cts[15], filter_bases[15]); conv[i] = counters of original read
        position i - 1, the filter arrays are indexed by gt_filter_reason"""
        pr = Profile()
        self._check(self.lib.bsgpu_profile_read(self.ctx, C.byref(pr), C.c_int(1 if reset else 0)))
        conv = np.ctypeslib.as_array(pr.conv_cts).reshape(PROFILE_MAX, 4)
        return dict(used=int(pr.used), conv=conv[:pr.used].copy(), base_filter=np.array(list(pr.base_filter), dtype=np.uint64),
                    filter_cts=np.array(list(pr.filter_cts), dtype=np.uint64), filter_bases=np.array(list(pr.filter_bases), dtype=np.uint64))

    def site_stats_enable(self, on=True, n_contigs=0):
        """gather the writer's --report-file statistics (src/print_vcf.c:382-526) in every entry point that writes BCF records"""
        self._check(self.lib.bsgpu_site_stats_enable(self.ctx, C.c_int(1 if on else 0), C.c_int(int(n_contigs))))

    def set_contig_gc(self, rid, gc, start_pos=1):
        """GC percentage of every 100-base bin of contig `rid` from start_pos on (ctg_stats->gc)"""
        gc = np.ascontiguousarray(gc, dtype=np.uint8)
        self._check(self.lib.bsgpu_set_contig_gc(self.ctx, C.c_int(int(rid)), C.c_uint32(int(start_pos)), _ptr(gc), C.c_uint32(len(gc))))

    def site_stats_read(self, n_contigs=0, reset=False):
        """-> (SITE_STATS record array of length 1, CTG_SITE_STATS[n_contigs])"""
        from .records import SITE_STATS, CTG_SITE_STATS
        st = np.zeros(1, dtype=SITE_STATS)
        ctg = np.zeros(max(int(n_contigs), 1), dtype=CTG_SITE_STATS)
        self._check(self.lib.bsgpu_site_stats_read(self.ctx, _ptr(st), _ptr(ctg) if n_contigs else None, C.c_int(int(n_contigs)), C.c_int(1 if reset else 0)))
        return st, ctg[:int(n_contigs)]

    # ---- reader side ----------------------------------------------------------------------------
    def decode_records(self, bam, rp=None, want_reads=True):
        """raw BAM records -> (RECORD[], packed bases, events); the decoded arrays also stay resident in the context"""
        bam = np.ascontiguousarray(bam, dtype=np.uint8)
        rp = rp or reader_params()
        cap = len(bam) // 36 + 8
        rec = np.zeros(cap, dtype=RECORD)
        bases = np.zeros(len(bam) + 64, dtype=np.uint8) if want_reads else None
        misms = np.zeros(len(bam) // 4 + 64, dtype=MISMS) if want_reads else None
        n, nb, nm = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        self._check(self.lib.bsgpu_decode_records(self.ctx, _ptr(bam), C.c_size_t(len(bam)), C.byref(rp), _ptr(rec), C.c_size_t(cap),
                                                  C.byref(n), _ptr(bases), C.c_size_t(len(bases) if want_reads else 0), C.byref(nb),
                                                  _ptr(misms), C.c_size_t(len(misms) if want_reads else 0), C.byref(nm)))
        if not want_reads:
            return rec[:n.value], None, None
        return rec[:n.value], bases[:nb.value], misms[:nm.value]

    def call_bam(self, bam, target_len, ctg_codes, rp=None, vcf=None):
        """raw BAM records + per-contig reference codes -> (BLOCK[], gt_vcf[]); the whole path on the device except the
        block builder"""
        bam = np.ascontiguousarray(bam, dtype=np.uint8)
        target_len = np.ascontiguousarray(target_len, dtype=np.uint32)
        rp = rp or reader_params()
        codes = [np.ascontiguousarray(c, dtype=np.uint8) for c in ctg_codes]
        ptrs = (C.c_void_p * len(codes))(*[c.ctypes.data for c in codes])
        bcap = len(bam) // 36 + 8
        blocks = np.zeros(bcap, dtype=BLOCK)
        if vcf is None:
            vcf = np.zeros(int(target_len.sum()) + 8, dtype=GT_VCF)
        nb, nv = C.c_size_t(0), C.c_size_t(0)
        self._check(self.lib.bsgpu_call_bam(self.ctx, _ptr(bam), C.c_size_t(len(bam)), C.c_int(len(codes)), _ptr(target_len), ptrs,
                                            C.byref(rp), _ptr(blocks), C.c_size_t(bcap), C.byref(nb), _ptr(vcf), C.c_size_t(len(vcf)),
                                            C.byref(nv)))
        return blocks[:nb.value], vcf[:nv.value]

    def bam_session(self, target_len, ctg_codes, rp=None, bcf=None, vcf_rid=None, batch_bytes=0):
        """streaming session over this context (bsgpu_bam_open); bcf: None -> gt_vcf[] results, True / BcfParams -> BCF records"""
        return BamSession(self, target_len, ctg_codes, rp, bcf, vcf_rid, batch_bytes)

    # ---- device-pointer entry points (addresses as ints, e.g. torch.Tensor.data_ptr()) --------
    def call_bam_bcf(self, bam, target_len, ctg_codes, rp=None, params=None, vcf_rid=None, out=None):
        """raw BAM records + per-contig reference codes -> (BLOCK[], BCF record bytes, number of records)"""
        bam = np.ascontiguousarray(bam, dtype=np.uint8)
        target_len = np.ascontiguousarray(target_len, dtype=np.uint32)
        rp = rp or reader_params()
        p = params or bcf_params()
        codes = [np.ascontiguousarray(c, dtype=np.uint8) for c in ctg_codes]
        ptrs = (C.c_void_p * len(codes))(*[c.ctypes.data for c in codes])
        rid = None if vcf_rid is None else np.ascontiguousarray(vcf_rid, dtype=np.int32)
        blocks = np.zeros(len(bam) // 36 + 8, dtype=BLOCK)
        out = np.empty(int(target_len.sum()) * 160 + 4096, dtype=np.uint8) if out is None else out
        nbk, nb, nr = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        self._check(self.lib.bsgpu_call_bam_bcf(self.ctx, _ptr(bam), C.c_size_t(len(bam)), C.c_int(len(codes)), _ptr(target_len), ptrs,
                                                C.byref(rp), C.byref(p), _ptr(rid), _ptr(blocks), C.c_size_t(len(blocks)), C.byref(nbk),
                                                _ptr(out), C.c_size_t(len(out)), C.byref(nb), C.byref(nr)))
        return blocks[:nbk.value], out[:nb.value], nr.value

    def call_sites_dev(self, d_pileup, d_ref, n, d_out, d_skip, stream=0):
        self._check(self.lib.bsgpu_call_sites_dev(self.ctx, _ptr(d_pileup), _ptr(d_ref), C.c_size_t(n), _ptr(d_out),
                                                  _ptr(d_skip), C.c_void_p(stream)))

    def call_sites_vcf_dev(self, d_pileup, d_ref, n, d_vcf, stream=0):
        self._check(self.lib.bsgpu_call_sites_vcf_dev(self.ctx, _ptr(d_pileup), _ptr(d_ref), C.c_size_t(n), _ptr(d_vcf),
                                                      C.c_void_p(stream)))

    def pileup_block_dev(self, d_segs, nseg, d_bases, x, sz, d_out, stream=0):
        self._check(self.lib.bsgpu_pileup_block_dev(self.ctx, _ptr(d_segs), C.c_size_t(nseg), _ptr(d_bases), C.c_uint32(x),
                                                    C.c_uint32(sz), _ptr(d_out), C.c_void_p(stream)))

    def call_block_dev(self, d_segs, nseg, d_bases, d_ref, x, sz, d_out, stream=0):
        self._check(self.lib.bsgpu_call_block_dev(self.ctx, _ptr(d_segs), C.c_size_t(nseg), _ptr(d_bases), _ptr(d_ref),
                                                  C.c_uint32(x), C.c_uint32(sz), _ptr(d_out), C.c_void_p(stream)))

    def synth_sites_dev(self, seed, first_site, n, mean_depth, d_pileup, d_ref, stream=0):
        self._check(self.lib.bsgpu_synth_sites_dev(self.ctx, C.c_uint64(seed), C.c_uint64(first_site), C.c_size_t(n),
                                                   C.c_double(mean_depth), _ptr(d_pileup), _ptr(d_ref), C.c_void_p(stream)))

    def synth_block_nseg(self, sz, read_len, depth):
        return int(self.lib.bsgpu_synth_block_nseg(C.c_uint32(sz), C.c_uint32(read_len), C.c_double(depth)))

    def synth_block_dev(self, seed, x, sz, read_len, depth, d_segs, seg_cap, d_bases, base_cap, d_ref, stream=0):
        ns, nb = C.c_size_t(0), C.c_size_t(0)
        self._check(self.lib.bsgpu_synth_block_dev(self.ctx, C.c_uint64(seed), C.c_uint32(x), C.c_uint32(sz),
                                                   C.c_uint32(read_len), C.c_double(depth), _ptr(d_segs), C.c_size_t(seg_cap),
                                                   _ptr(d_bases), C.c_size_t(base_cap), _ptr(d_ref), C.byref(ns), C.byref(nb),
                                                   C.c_void_p(stream)))
        return ns.value, nb.value

    def synth_bam_bytes(self, ntemplates, read_len):
        return int(self.lib.bsgpu_synth_bam_bytes(C.c_size_t(ntemplates), C.c_uint32(read_len)))

    def synth_bam_dev(self, seed, ntemplates, read_len, d_pos_f, d_pos_r, d_src, d_rank, d_out, stream=0):
        self._check(self.lib.bsgpu_synth_bam_dev(self.ctx, C.c_uint64(seed), C.c_size_t(ntemplates), C.c_uint32(read_len), _ptr(d_pos_f),
                                                 _ptr(d_pos_r), _ptr(d_src), _ptr(d_rank), _ptr(d_out), C.c_void_p(stream)))

    def synth_ref_dev(self, seed, x, sz, d_ref, stream=0):
        self._check(self.lib.bsgpu_synth_ref_dev(self.ctx, C.c_uint64(seed), C.c_uint32(x), C.c_uint32(sz), _ptr(d_ref), C.c_void_p(stream)))

    # ---- host staging --------------------------------------------------------------------------
    def stage_templates(self, templates, bases, x, y):
        templates = np.ascontiguousarray(templates, dtype=TEMPLATE)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        bound = int(self.lib.bsgpu_stage_bound(_ptr(templates), C.c_size_t(len(templates))))
        segs = np.zeros(max(bound, 1), dtype=SEG)
        ns = C.c_size_t(0)
        self._check(self.lib.bsgpu_stage_templates(_ptr(templates), C.c_size_t(len(templates)), _ptr(bases), C.c_uint32(x),
                                                   C.c_uint32(y), _ptr(segs), C.byref(ns)))
        return segs[:ns.value]


class BamSession:
    """bsgpu_bam_open / _feed / _finish / _drain / _release / _close: the reader side in bounded memory.  Results come back
    as (BLOCK[] view, data view, nrec, BamResult); the views alias the session's pinned memory until release(id)."""

    def __init__(self, gpu, target_len, ctg_codes, rp=None, bcf=None, vcf_rid=None, batch_bytes=0):
        self.gpu, self.lib = gpu, gpu.lib
        self.target_len = np.ascontiguousarray(target_len, dtype=np.uint32)
        self.codes = [None if c is None else np.ascontiguousarray(c, dtype=np.uint8) for c in ctg_codes]      # kept alive: the session reads them
        ptrs = (C.c_void_p * len(self.codes))(*[None if c is None else c.ctypes.data for c in self.codes])
        self.rp = rp or reader_params()
        self.bcf = bcf_params() if bcf is True else bcf
        self.rid = None if vcf_rid is None else np.ascontiguousarray(vcf_rid, dtype=np.int32)
        self.s = C.c_void_p()
        self.lib.bsgpu_bam_open.argtypes = None
        gpu._check(self.lib.bsgpu_bam_open(gpu.ctx, C.c_int(len(self.codes)), _ptr(self.target_len), ptrs, C.byref(self.rp),
                                           C.byref(self.bcf) if self.bcf is not None else None, _ptr(self.rid), C.c_size_t(int(batch_bytes)),
                                           C.byref(self.s)))

    def feed(self, data, nowait=False):
        """returns the number of bytes taken (all of them unless nowait)"""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        acc = C.c_size_t(0)
        self.gpu._check(self.lib.bsgpu_bam_feed(self.s, _ptr(data), C.c_size_t(len(data)), C.byref(acc) if nowait else None))
        return acc.value if nowait else len(data)

    def reserve(self, wait=True):
        """-> numpy view of the staging area where the next bytes of the stream go (commit(n) afterwards)"""
        ptr, avail = C.c_void_p(), C.c_size_t(0)
        self.gpu._check(self.lib.bsgpu_bam_reserve(self.s, C.byref(ptr), C.byref(avail), C.c_int(1 if wait else 0)))
        if not avail.value:
            return np.zeros(0, dtype=np.uint8)
        return np.frombuffer((C.c_uint8 * avail.value).from_address(ptr.value), dtype=np.uint8)

    def commit(self, n):
        self.gpu._check(self.lib.bsgpu_bam_commit(self.s, C.c_size_t(int(n))))

    def finish(self):
        self.gpu._check(self.lib.bsgpu_bam_finish(self.s))

    def cut(self):
        self.gpu._check(self.lib.bsgpu_bam_cut(self.s))

    def rewind(self):
        self.gpu._check(self.lib.bsgpu_bam_rewind(self.s))

    def drain(self, wait=True):
        """-> None when no result is ready (check .finished), else (blocks, data, nrec, result)"""
        r = BamResult()
        self.gpu._check(self.lib.bsgpu_bam_drain(self.s, C.c_int(1 if wait else 0), C.byref(r)))
        self.finished = bool(r.finished)
        if not r.id:
            return None
        blocks = np.frombuffer((C.c_uint8 * (r.nblocks * BLOCK.itemsize)).from_address(r.blocks), dtype=BLOCK) if r.nblocks else np.zeros(0, dtype=BLOCK)
        if self.bcf is not None:
            data = np.frombuffer((C.c_uint8 * r.nbytes).from_address(r.data), dtype=np.uint8) if r.nbytes else np.zeros(0, dtype=np.uint8)
        else:
            data = np.frombuffer((C.c_uint8 * r.nbytes).from_address(r.data), dtype=GT_VCF) if r.nbytes else np.zeros(0, dtype=GT_VCF)
        return blocks, data, int(r.nrec), r

    def release(self, r):
        self.gpu._check(self.lib.bsgpu_bam_release(self.s, C.c_uint64(r.id if isinstance(r, BamResult) else int(r))))

    def progress(self):
        p = BamProgress()
        self.gpu._check(self.lib.bsgpu_bam_progress(self.s, C.byref(p)))
        return {k: int(getattr(p, k)) for k, _ in BamProgress._fields_}

    def close(self):
        if self.s:
            self.lib.bsgpu_bam_close(self.s)
            self.s = C.c_void_p()

    def run(self, bam, slice_bytes=1 << 20, nowait=True, keep=True):
        """convenience for tests: feeds `bam` in slices and collects the results as (blocks, data copy, nrec) per batch.
        nowait: one thread, non-blocking feeds, draining whenever a feed comes up short.  Otherwise two threads as in the
        reference: this one feeds (blocking), a second one drains."""
        bam = np.ascontiguousarray(bam, dtype=np.uint8)
        out = []

        def collect(got):
            b, d, n, r = got
            out.append((b.copy(), d.copy() if keep else None, n))
            self.release(r)

        def drain_all():
            while True:
                got = self.drain(wait=True)
                if got is not None:
                    collect(got)
                if self.finished:
                    return

        if nowait:
            at = 0
            while at < len(bam):
                m = min(slice_bytes, len(bam) - at)
                took = self.feed(bam[at:at + m], nowait=True)
                at += took
                if took < m:
                    got = self.drain(wait=True)
                    if got is not None:
                        collect(got)
            self.finish()
            drain_all()
            return out
        import threading
        err = []

        def printer():
            try:
                while True:
                    got = self.drain(wait=True)
                    if got is not None:
                        collect(got)
                    elif not self.finished:
                        import time
                        time.sleep(0.0005)          # nothing in flight yet: the feeder has not filled a batch
                    if self.finished:
                        return
            except Exception as e:        # noqa: BLE001 -- handed to the feeding thread
                err.append(e)

        t = threading.Thread(target=printer)
        t.start()
        try:
            for at in range(0, len(bam), slice_bytes):
                self.feed(bam[at:at + slice_bytes])
            self.finish()
        finally:
            t.join()
        if err:
            raise err[0]
        return out


def math_probe(x):
    """(log, exp) of the kernels' table-driven routines, evaluated on the host (bsgpu_math_probe)."""
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    lo, ex = np.zeros_like(x), np.zeros_like(x)
    lib.bsgpu_math_probe(_ptr(x), C.c_size_t(len(x)), _ptr(lo), _ptr(ex))
    return lo, ex


WIRE_BYTES = 120


def wire_pack(records, skip=None):
    """records (GT_METH with skip[], or GT_VCF) -> (wire bytes [n, 120], ok) on the host (bsgpu_wire_pack)"""
    lib = load()
    records = np.ascontiguousarray(records)
    n, rb = len(records), records.dtype.itemsize
    wire = np.zeros((n, WIRE_BYTES), dtype=np.uint8)
    rc = lib.bsgpu_wire_pack(_ptr(records), C.c_size_t(n), C.c_size_t(rb), _ptr(skip) if skip is not None else None, _ptr(wire))
    return wire, rc == BSGPU_OK


def wire_expand(wire, dtype, threads=1):
    """wire bytes -> (records of `dtype`, skip[] or None) on the host, the pass the library runs on its own threads (bsgpu_wire_expand)"""
    lib = load()
    wire = np.ascontiguousarray(wire, dtype=np.uint8)
    n = wire.size // WIRE_BYTES
    rb = np.dtype(dtype).itemsize
    out = np.full(n, 0, dtype=dtype)
    out.view(np.uint8)[:] = 0xa5
    skip = np.full(n, 0xa5, dtype=np.uint8) if rb == 200 else None
    rc = lib.bsgpu_wire_expand(_ptr(wire), C.c_size_t(n), C.c_size_t(rb), _ptr(out), _ptr(skip) if skip is not None else None, C.c_int(threads))
    if rc != BSGPU_OK:
        raise RuntimeError(last_error(lib))
    return out, skip


def build_blocks(bam, rec, rp=None, tally=False):
    """bsgpu_build_blocks: descriptors of decode_records -> (BLOCK[], TEMPLATE[]).  Pure host function of the ABI.
    tally: also return read_input's filter_cts[15] / filter_bases[15] (bsgpu_build_blocks_tally)."""
    lib = load()
    bam = np.ascontiguousarray(bam, dtype=np.uint8)
    rec = np.ascontiguousarray(rec, dtype=RECORD)
    rp = rp or reader_params()
    blocks = np.zeros(len(rec) + 8, dtype=BLOCK)
    tmpl = np.zeros(len(rec) + 8, dtype=TEMPLATE)
    nb, nt = C.c_size_t(0), C.c_size_t(0)
    if tally:
        fc, fb = np.zeros(15, dtype=np.uint64), np.zeros(15, dtype=np.uint64)
        rc = lib.bsgpu_build_blocks_tally(_ptr(bam), C.c_size_t(len(bam)), _ptr(rec), C.c_size_t(len(rec)), C.byref(rp), _ptr(blocks),
                                          C.c_size_t(len(blocks)), C.byref(nb), _ptr(tmpl), C.c_size_t(len(tmpl)), C.byref(nt), _ptr(fc), _ptr(fb))
    else:
        rc = lib.bsgpu_build_blocks(_ptr(bam), C.c_size_t(len(bam)), _ptr(rec), C.c_size_t(len(rec)), C.byref(rp), _ptr(blocks),
                                    C.c_size_t(len(blocks)), C.byref(nb), _ptr(tmpl), C.c_size_t(len(tmpl)), C.byref(nt))
    if rc != BSGPU_OK:
        raise BsGpuError(lib.bsgpu_last_error().decode())
    if tally:
        return blocks[:nb.value], tmpl[:nt.value], fc, fb
    return blocks[:nb.value], tmpl[:nt.value]


def stage_templates_host(templates, bases, x, y):
    """bsgpu_stage_templates without a device context (pure host function of the ABI)."""
    lib = load()
    templates = np.ascontiguousarray(templates, dtype=TEMPLATE)
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    bound = int(lib.bsgpu_stage_bound(_ptr(templates), C.c_size_t(len(templates))))
    segs = np.zeros(max(bound, 1), dtype=SEG)
    ns = C.c_size_t(0)
    if lib.bsgpu_stage_templates(_ptr(templates), C.c_size_t(len(templates)), _ptr(bases), C.c_uint32(x), C.c_uint32(y),
                                 _ptr(segs), C.byref(ns)) != BSGPU_OK:
        raise BsGpuError(lib.bsgpu_last_error().decode())
    return segs[:ns.value]
