"""CPU-side checks (run with -m "not gpu"): oracle against the committed golden fixtures (captured from the
reference's own code by tests/golden/make_golden.py), the C-ABI surface of libbsgpu.so, and the host staging logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bs_call_b200 import lib as bslib
from bs_call_b200 import records
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BLOCKS = ["block_pe_plain", "block_pe_indel_clip_trim", "block_se_deep", "block_mixed"]


def test_record_layouts_match_oracle():
    from oracle import bindings
    for name in ("PILEUP", "GT_METH", "GT_VCF", "TEMPLATE", "MISMS"):
        assert getattr(records, name) == getattr(bindings, name), name


def test_oracle_sites_golden(oracle):
    g = util.load_golden("sites_v1")
    out, skip = oracle.call_sites(g["pileup"], g["ref"], nthreads=2)
    n = util.assert_gt_meth_close(out, skip, g["gt_meth"], g["skip"])
    assert n > 7000
    # same machine family as the one that made the fixtures -> normally bit-identical; tolerance covers libm versions
    for i in range(len(g["kat_rf"])):
        got = oracle.calc_gt_prob(g["kat_counts"][i], g["kat_qual"][i], g["kat_rf"][i])
        want = g["kat_out"][i]
        if not util.near_tie(want["gt_prob"][None])[0]:
            assert got["max_gt"] == want["max_gt"]
        np.testing.assert_allclose(got["gt_prob"], want["gt_prob"], rtol=util.PROB_RTOL, atol=util.PROB_ATOL)
    for t, p in zip(g["fisher_tabs"], g["fisher_p"]):
        assert oracle.fisher(t) == pytest.approx(p, rel=1e-10, abs=1e-300)


@pytest.mark.parametrize("name", BLOCKS)
def test_oracle_block_golden(name):
    from oracle.bindings import Oracle
    g = util.load_golden(name)
    o = Oracle(left_trim=tuple(g["left_trim"]), right_trim=tuple(g["right_trim"]))
    nt, nb = o.normalise_block(g["templates"], g["bases"], g["misms"])
    assert nb.tobytes() == g["norm_bases"].tobytes()
    for f in ("forward_position", "reverse_position", "read_len", "read_off", "present"):
        assert (nt[f] == g["norm_templates"][f]).all(), f
    x, pile, vcf = o.process_block(g["templates"], g["bases"], g["misms"], g["ref"], int(g["y"]))
    assert x == int(g["x"])
    util.assert_pileup_equal(pile, g["pileup"])
    util.assert_vcf_close(vcf, g["vcf"])
    # pileup straight from the golden normalised templates
    pile2 = o.pileup_block(g["norm_templates"], g["norm_bases"], x, int(g["y"]))
    util.assert_pileup_equal(pile2, g["pileup"])


@pytest.mark.parametrize("name", BLOCKS)
def test_oracle_writer_matches_golden(name, oracle):
    """the restated writer against the BCF records captured from the reference's compiled src/print_vcf.c"""
    g = util.load_golden(name)
    w = util.load_golden("writer_v1")
    for allp in (0, 1):
        b, n = oracle.print_block(g["vcf"], w[name + "__ref"], int(g["x"]), rid=2, all_positions=bool(allp))
        assert n == int(w["%s__n%d" % (name, allp)]) and n > 300
        assert b.tobytes() == w["%s__all%d" % (name, allp)].tobytes()
    recs = util.split_bcf(b)
    assert len(recs) == n and max(len(r) for r in recs) <= 384


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bsgpu.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bsgpu_[a-z0-9_]+)\s*\(", body))
    assert declared == set(bslib.EXPORTS), (declared ^ set(bslib.EXPORTS))
    so = C.CDLL(bslib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(so, name), "libbsgpu.so does not export %s" % name
    assert so.bsgpu_version() >= 100


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(bslib.BsGpuError):
        bslib.BsGpu()


@pytest.mark.parametrize("name", BLOCKS)
def test_stage_templates_reproduces_pileup(name, oracle):
    """Segments staged by the ABI's host function, accumulated naively in numpy, give the golden pileup: checks the
    mate walk / strand-index flip / window clip logic without a GPU."""
    g = util.load_golden(name)
    x, y = int(g["x"]), int(g["y"])
    segs = bslib.stage_templates_host(g["norm_templates"], g["norm_bases"], x, y)
    assert (segs["len"] <= 256).all() and (segs["len"] > 0).all()
    sz = y - x + 1
    cls = np.array([[0, 1, 2, 3], [0, 5, 2, 7], [4, 1, 6, 3]])
    counts = np.zeros((sz, 2, 8), dtype=np.uint32)
    qsum = np.zeros((sz, 8), dtype=np.float64)
    mq2 = np.zeros(sz, dtype=np.float64)
    nb = g["norm_bases"]
    for s in segs:
        b = nb[s["off"]:s["off"] + s["len"]]
        q = b >> 2
        ok = (q >= 20) & (q != 63)
        site = np.arange(s["pos"] - x, s["pos"] - x + s["len"])
        assert site.max() < sz
        c = cls[(s["flags"] >> 1) & 3][b & 3]
        np.add.at(counts, (site[ok], s["flags"] & 1, c[ok]), 1)
        np.add.at(qsum, (site[ok], c[ok]), q[ok])
        np.add.at(mq2, site[ok], float(s["mapq"]) ** 2)
    want = g["pileup"]
    assert (counts == want["counts"]).all()
    assert (qsum.astype(np.float32) == want["quality"]).all()
    assert (mq2.astype(np.float32) == want["mapq2"]).all()
    assert (counts.sum(axis=(1, 2)) == want["n"]).all()


def test_stage_rejects_bad_input():
    t = np.zeros(1, dtype=records.TEMPLATE)
    t["present"][0, 0] = 1
    t["read_len"][0, 0] = 4
    t["forward_position"] = 5
    t["bs_strand"] = 3
    with pytest.raises(bslib.BsGpuError):
        bslib.stage_templates_host(t, np.full(4, 37 << 2, dtype=np.uint8), 1, 100)
    t["bs_strand"] = 1
    with pytest.raises(bslib.BsGpuError):      # starts before the window
        bslib.stage_templates_host(t, np.full(4, 37 << 2, dtype=np.uint8), 10, 100)
    segs = bslib.stage_templates_host(t, np.full(4, 37 << 2, dtype=np.uint8), 1, 6)     # clipped at y
    assert len(segs) == 1 and segs[0]["len"] == 2


def test_fast_math_accuracy():
    """The kernels' table-driven log/exp (bsgpu_math.cuh), evaluated by the host build of the same source, against
    long double.  Domains are the ones the genotype model produces."""
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(1e-6, 3.0, 400000), np.exp(rng.uniform(np.log(1e-6), np.log(3.0), 400000)),
                        1.0 + np.exp(rng.uniform(np.log(1e-16), np.log(9.0), 400000)), [1.0, 2.0, 0.5, 1.375, 0.6875]])
    lo, _ = bslib.math_probe(x)
    want = np.log(x.astype(np.longdouble))
    err = np.abs(lo.astype(np.longdouble) - want)
    assert lo[-5] == 0.0                                     # log(1) is exactly 0
    assert float(err.max()) < 2.5e-15                        # absolute, at |log| ~ 14
    big = np.abs(lo) > 0.01
    assert float((err / np.spacing(np.abs(lo)))[big].max()) < 2.0      # ulps
    m = (x >= 1.0) & (want > 0)
    assert float((err[m] / want[m]).max()) < 4e-16           # log-sum-exp arguments keep relative accuracy down to 1+eps
    xe = np.concatenate([rng.uniform(-45.5, 0.0, 800000), [0.0, -45.0]])
    _, ex = bslib.math_probe(xe)
    we = np.exp(xe.astype(np.longdouble))
    assert float(np.abs((ex.astype(np.longdouble) - we) / we).max()) < 2.3e-16      # about 1 ulp
    assert ex[-2] == 1.0


def test_wire_records_round_trip():
    """bsgpu_wire_pack / bsgpu_wire_expand (host side of the compact records the host-buffer entry points can send over PCIe,
    bs_call_b200/csrc/bsgpu_wire.h): every byte of gt_meth / gt_vcf records comes back, on one thread and through the pool,
    and a field too wide for the wire is refused."""
    from bs_call_b200 import lib as bslib
    from bs_call_b200.records import GT_METH, GT_VCF
    rng = np.random.default_rng(1)
    n = 300_001
    for dt in (GT_METH, GT_VCF):
        r = np.zeros(n, dtype=dt)
        g = r if dt is GT_METH else r["gtm"]
        g["counts"] = rng.integers(0, 65536, (n, 8))
        g["qual"] = rng.integers(0, 256, (n, 8))
        g["gt_prob"] = rng.standard_normal((n, 10))
        g["fisher_strand"] = rng.standard_normal(n)
        g["mq"] = rng.integers(0, 256, n)
        g["aq"] = rng.integers(0, 256, n)
        g["max_gt"] = rng.integers(0, 10, n)
        sk = rng.integers(0, 2, n).astype(np.uint8)
        if dt is GT_VCF:
            r["ready"] = 1
            r["skip"] = sk
        w, ok = bslib.wire_pack(r, sk if dt is GT_METH else None)
        assert ok and w.shape == (n, bslib.WIRE_BYTES)
        for threads in (1, 3):
            o, s2 = bslib.wire_expand(w, dt, threads)
            assert o.tobytes() == r.tobytes()
            assert s2 is None or (s2 == sk).all()
        for field, val in (("counts", 65536), ("qual", 256), ("mq", 256), ("aq", -1)):
            bad = r.copy()
            gb = bad if dt is GT_METH else bad["gtm"]
            if gb[field].ndim == 2:
                gb[field][7, 3] = val
            else:
                gb[field][7] = val
            assert not bslib.wire_pack(bad, sk if dt is GT_METH else None)[1], field
