"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI of libbsgpu.so, against
(a) the golden fixtures captured from the reference's own code and (b) the oracle on seeded inputs.

Bit-exact for counts / qualities / indices / flags; PROB_RTOL for the log10 posteriors (tests/util.py).
"""
import numpy as np
import pytest

from bs_call_b200 import lib as bslib
from bs_call_b200.records import GT_METH, GT_VCF, PILEUP, SEG
from tests import blockgen, util

pytestmark = pytest.mark.gpu

BLOCKS = ["block_pe_plain", "block_pe_indel_clip_trim", "block_se_deep", "block_mixed"]


@pytest.fixture(scope="module")
def gpu():
    g = bslib.BsGpu()
    yield g
    g.close()


def test_native_library_is_the_path(gpu):
    before = gpu.stats()["kernel_launches"]
    p = np.zeros(4, dtype=PILEUP)
    gpu.call_sites(p, np.zeros(4, dtype=np.uint8))
    assert gpu.stats()["kernel_launches"] == before + 1


def test_call_sites_golden(gpu):
    g = util.load_golden("sites_v1")
    out, skip = gpu.call_sites(g["pileup"], g["ref"])
    n = util.assert_gt_meth_close(out, skip, g["gt_meth"], g["skip"])
    assert n > 7000


def test_model_kats_golden(gpu):
    """calc_gt_prob KATs: build a one-strand pileup record whose per-class mean quality is the KAT's qual."""
    g = util.load_golden("sites_v1")
    n = len(g["kat_rf"])
    p = np.zeros(n, dtype=PILEUP)
    c = g["kat_counts"].astype(np.uint32)
    p["counts"][:, 0, :] = c
    p["n"] = c.sum(axis=1)
    p["quality"] = (c * g["kat_qual"]).astype(np.float32)       # exact below 2^24: 20000 * 43 < 2^24
    p["mapq2"] = 3600.0 * p["n"]
    out, skip = gpu.call_sites(p, g["kat_rf"])
    m = p["n"] > 0
    want = g["kat_out"][m]
    got = out[m]
    assert (got["qual"] == g["kat_qual"][m]).all()
    tie = util.near_tie(want["gt_prob"])
    assert (got["max_gt"][~tie] == want["max_gt"][~tie]).all()
    np.testing.assert_allclose(got["gt_prob"], want["gt_prob"], rtol=util.PROB_RTOL, atol=util.PROB_ATOL)
    assert (skip[~m] == 1).all() and (skip[m] == 0).all()


@pytest.mark.parametrize("n", [0, 1, 2, 127, 128, 129, 255, 1000, 4097])
def test_call_sites_ragged_sizes(gpu, oracle, n):
    rng = np.random.default_rng(n)
    p, ref = blockgen.random_pileups(rng, n, depth=12)
    out, skip = gpu.call_sites(p, ref)
    wout, wskip = oracle.call_sites(p, ref)
    assert len(out) == n
    if n:
        util.assert_gt_meth_close(out, skip, wout, wskip)


def test_call_sites_random_vs_oracle(gpu, oracle):
    rng = np.random.default_rng(2024)
    p, ref = blockgen.random_pileups(rng, 30000, depth=30, het_frac=0.2)
    out, skip = gpu.call_sites(p, ref)
    wout, wskip = oracle.call_sites(p, ref, nthreads=4)
    n = util.assert_gt_meth_close(out, skip, wout, wskip)
    assert n > 25000
    het = np.isin(wout["max_gt"], [1, 2, 3, 5, 6, 8]) & (wskip == 0)
    assert het.sum() > 500


def test_writer_integer_fields_identical(gpu, oracle):
    """GT, QUAL/GQ, FS, QD and the FILTER bits as the reference's writer derives them from gt_meth (tests/util.py:
    writer_fields) -- computed from the GPU records and from the oracle's: identical on 1.2 M sites of the config 2
    stream plus 20 k het-enriched random sites.  The doubles differ by ulps; nothing of that reaches the VCF."""
    p, r = oracle.synth_sites(20261018, 7_000_000, 1_200_000, 30.0, nthreads=8)
    rng = np.random.default_rng(5)
    p2, r2 = blockgen.random_pileups(rng, 20000, depth=18, het_frac=0.5)
    p = np.concatenate([p, p2])
    r = np.concatenate([r, r2])
    out, skip = gpu.call_sites(p, r)
    wout, wskip = oracle.call_sites(p, r, nthreads=8)
    a, b = util.writer_fields(out, skip), util.writer_fields(wout, wskip)
    for f in a:
        bad = np.nonzero(a[f] != b[f])[0]
        assert len(bad) == 0, "%s differs at %d of %d sites, first %d: %r vs %r" % (f, len(bad), len(a[f]), bad[0], a[f][bad[0]], b[f][bad[0]])
    assert (a["flt"] == 0).sum() > 100000 and (a["flt"] != 0).sum() > 1000 and len(np.unique(a["gt"])) == 10


def test_call_sites_deep_counts_lgamma(gpu, oracle):
    """depth ~2000: strand tables with margins >= 256 take the lgamma branch of the Fisher test"""
    rng = np.random.default_rng(5)
    p, ref = blockgen.random_pileups(rng, 300, depth=2000, het_frac=0.7)
    out, skip = gpu.call_sites(p, ref)
    wout, wskip = oracle.call_sites(p, ref)
    util.assert_gt_meth_close(out, skip, wout, wskip)


def test_other_parameters(oracle):
    from oracle.bindings import Oracle
    g = bslib.BsGpu(under_conv=0.03, over_conv=0.1, ref_bias=1.0, min_qual=10)
    o = Oracle(under_conv=0.03, over_conv=0.1, ref_bias=1.0, min_qual=10)
    rng = np.random.default_rng(8)
    p, ref = blockgen.random_pileups(rng, 5000, depth=20, het_frac=0.2)
    out, skip = g.call_sites(p, ref)
    wout, wskip = o.call_sites(p, ref)
    util.assert_gt_meth_close(out, skip, wout, wskip)
    g.close()


@pytest.mark.parametrize("name", BLOCKS)
def test_pileup_block_golden(gpu, name):
    g = util.load_golden(name)
    x, y = int(g["x"]), int(g["y"])
    segs = gpu.stage_templates(g["norm_templates"], g["norm_bases"], x, y)
    pile = gpu.pileup_block(segs, g["norm_bases"], x, y - x + 1)
    util.assert_pileup_equal(pile, g["pileup"])
    # segment order must not matter (the device bins them itself)
    rng = np.random.default_rng(1)
    pile2 = gpu.pileup_block(segs[rng.permutation(len(segs))], g["norm_bases"], x, y - x + 1)
    util.assert_pileup_equal(pile2, g["pileup"])
    assert gpu.stats()["qsum_overflow"] == 0


@pytest.mark.parametrize("name", BLOCKS)
def test_call_block_golden(gpu, name):
    g = util.load_golden(name)
    x, y = int(g["x"]), int(g["y"])
    segs = gpu.stage_templates(g["norm_templates"], g["norm_bases"], x, y)
    vcf = gpu.call_block(segs, g["norm_bases"], g["ref"], x, y - x + 1)
    util.assert_vcf_close(vcf, g["vcf"])


def test_block_edge_cases(gpu, oracle):
    # empty block: no segments at all -> every site skipped
    vcf = gpu.call_block(np.zeros(0, dtype=SEG), np.zeros(0, dtype=np.uint8), np.ones(300, dtype=np.uint8), 10, 300)
    assert (vcf["skip"] == 1).all() and (vcf["ready"] == 1).all()
    pile = gpu.pileup_block(np.zeros(0, dtype=SEG), np.zeros(0, dtype=np.uint8), 10, 300)
    assert not pile.tobytes().strip(b"\0")
    # one read, window of one tile minus one / exactly one / plus one site
    for sz in (255, 256, 257, 513):
        bases = np.full(100, (37 << 2) | 1, dtype=np.uint8)
        segs = np.zeros(1, dtype=SEG)
        segs[0] = (sz - 60 + 5, 0, 100, 60, 1 | (1 << 1), 0)        # runs past the window end: clipped
        pile = gpu.pileup_block(segs, bases, 5, sz)
        want = np.zeros(sz, dtype=PILEUP)
        want["n"][sz - 60:] = 1
        want["counts"][sz - 60:, 1, 5] = 1
        want["quality"][sz - 60:, 5] = 37
        want["mapq2"][sz - 60:] = 3600
        util.assert_pileup_equal(pile, want)


def test_large_block_vs_oracle(gpu, oracle):
    """a 120 kb block at 30x through the fused kernel against the oracle's pileup + model"""
    rng = np.random.default_rng(77)
    ref = blockgen.random_reference(rng, 130000, n_runs=5)
    T, B, M, y = blockgen.make_block(rng, ref, 500, 120000, depth=30, read_len=150, paired=True, frag_mean=300)
    nt, nb = oracle.normalise_block(T, B, M)
    first = int(T[0]["forward_position"]) or int(T[0]["reverse_position"])
    x = first - 2
    want_pile = oracle.pileup_block(nt, nb, x, y)
    segs = gpu.stage_templates(nt, nb, x, y)
    sz = y - x + 1
    pile = gpu.pileup_block(segs, nb, x, sz)
    util.assert_pileup_equal(pile, want_pile)
    refw = ref[x - 1:x - 1 + sz]
    vcf = gpu.call_block(segs, nb, refw, x, sz)
    wout, wskip = oracle.call_sites(want_pile, refw, nthreads=4)
    n = util.assert_gt_meth_close(vcf["gtm"], vcf["skip"], wout, wskip)
    assert n > 100000


def test_deep_panel_widening(gpu, oracle):
    """500x single-end panel (config 4): more than 255 hits per site exercises the packed-counter widening"""
    rng = np.random.default_rng(78)
    ref = blockgen.random_reference(rng, 3000)
    T, B, M, y = blockgen.make_block(rng, ref, 100, 2400, depth=700, read_len=150, paired=False, snp_rate=0.01)
    nt, nb = oracle.normalise_block(T, B, M)
    x = int(T[0]["forward_position"]) - 2
    sz = y - x + 1
    want_pile = oracle.pileup_block(nt, nb, x, y)
    assert want_pile["n"].max() > 600
    segs = gpu.stage_templates(nt, nb, x, y)
    util.assert_pileup_equal(gpu.pileup_block(segs, nb, x, sz), want_pile)
    refw = ref[x - 1:x - 1 + sz]
    vcf = gpu.call_block(segs, nb, refw, x, sz)
    wout, wskip = oracle.call_sites(want_pile, refw, nthreads=4)
    util.assert_gt_meth_close(vcf["gtm"], vcf["skip"], wout, wskip)


def test_synthetic_generators_and_device_entry_points(gpu, oracle):
    """device-resident path used by bench.py: generate in HBM, run the _dev entry points, check a sample on the host"""
    import torch
    n = 200000
    d_p = torch.empty(n * 104, dtype=torch.uint8, device="cuda")
    d_r = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_o = torch.empty(n * 200, dtype=torch.uint8, device="cuda")
    d_s = torch.empty(n, dtype=torch.uint8, device="cuda")
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream
    gpu.synth_sites_dev(20261018, 0, n, 30.0, d_p.data_ptr(), d_r.data_ptr(), st)
    gpu.call_sites_dev(d_p.data_ptr(), d_r.data_ptr(), n, d_o.data_ptr(), d_s.data_ptr(), st)
    torch.cuda.synchronize()
    p = d_p.cpu().numpy().view(PILEUP)
    r = d_r.cpu().numpy()
    out = d_o.cpu().numpy().view(GT_METH)
    skip = d_s.cpu().numpy()
    assert 25 < p["n"].mean() < 32 and 0.02 < (p["n"] == 0).mean() < 0.04
    # the host twin of the generator (oracle/bs_oracle.c, used by the CPU arm of bench.py) draws the same records
    hp, hr = oracle.synth_sites(20261018, 0, n, 30.0, nthreads=4)
    same = (hp.view(np.uint8).reshape(n, 104) == p.view(np.uint8).reshape(n, 104)).all(axis=1)
    assert same.mean() > 0.9999 and (hr == r).all()
    wout, wskip = oracle.call_sites(p, r, nthreads=4)
    util.assert_gt_meth_close(out, skip, wout, wskip)
    # same seed -> same records; vcf-layout variant agrees with the gt_meth variant
    d_p2 = torch.empty_like(d_p)
    d_r2 = torch.empty_like(d_r)
    gpu.synth_sites_dev(20261018, 0, n, 30.0, d_p2.data_ptr(), d_r2.data_ptr(), st)
    d_v = torch.empty(n * 208, dtype=torch.uint8, device="cuda")
    gpu.call_sites_vcf_dev(d_p2.data_ptr(), d_r2.data_ptr(), n, d_v.data_ptr(), st)
    torch.cuda.synchronize()
    assert torch.equal(d_p, d_p2) and torch.equal(d_r, d_r2)
    v = d_v.cpu().numpy().view(GT_VCF)
    assert v["gtm"].tobytes() == out.tobytes() and (v["skip"] == skip).all() and (v["ready"] == 1).all()


def test_synthetic_block_fused_vs_oracle(gpu, oracle):
    import torch
    x, sz, L, depth = 1000, 300000, 150, 30.0
    ns = gpu.synth_block_nseg(sz, L, depth)
    d_seg = torch.empty(ns * 16, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(ns * L, dtype=torch.uint8, device="cuda")
    d_r = torch.empty(sz, dtype=torch.uint8, device="cuda")
    d_v = torch.empty(sz * 208, dtype=torch.uint8, device="cuda")
    d_p = torch.empty(sz * 104, dtype=torch.uint8, device="cuda")
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream
    nseg, nb = gpu.synth_block_dev(7, x, sz, L, depth, d_seg.data_ptr(), ns, d_b.data_ptr(), ns * L, d_r.data_ptr(), st)
    assert nseg == ns
    gpu.call_block_dev(d_seg.data_ptr(), nseg, d_b.data_ptr(), d_r.data_ptr(), x, sz, d_v.data_ptr(), st)
    gpu.pileup_block_dev(d_seg.data_ptr(), nseg, d_b.data_ptr(), x, sz, d_p.data_ptr(), st)
    torch.cuda.synchronize()
    segs = d_seg.cpu().numpy().view(SEG)
    bases = d_b.cpu().numpy()
    ref = d_r.cpu().numpy()
    # oracle pileup over the same reads presented as single-mate templates
    from bs_call_b200.records import TEMPLATE
    T = np.zeros(len(segs), dtype=TEMPLATE)
    T["forward_position"] = segs["pos"]
    T["read_off"][:, 0] = segs["off"]
    T["read_len"][:, 0] = segs["len"]
    T["present"][:, 0] = 1
    T["mapq"][:, 0] = segs["mapq"]
    T["orientation"] = segs["flags"] & 1
    T["bs_strand"] = (segs["flags"] >> 1) & 3
    want_pile = oracle.pileup_block(T, bases, x, x + sz - 1)
    pile = d_p.cpu().numpy().view(PILEUP)
    util.assert_pileup_equal(pile, want_pile)
    assert 27 < pile["n"][1000:-1000].mean() < 31
    wout, wskip = oracle.call_sites(want_pile, ref, nthreads=4)
    v = d_v.cpu().numpy().view(GT_VCF)
    n = util.assert_gt_meth_close(v["gtm"], v["skip"], wout, wskip)
    assert n > 0.99 * sz - 400


@pytest.mark.parametrize("name", BLOCKS)
def test_process_block_golden(name):
    """raw templates -> gt_vcf[] with trimming, soft clips, mate overlap and indel normalisation done on the device,
    against what the reference's process_template_vector -> call_genotypes_ML produced"""
    g = util.load_golden(name)
    gpu = bslib.BsGpu(left_trim=tuple(int(v) for v in g["left_trim"]), right_trim=tuple(int(v) for v in g["right_trim"]))
    x, vcf = gpu.process_block(g["templates"], g["bases"], g["misms"], g["ref"], int(g["y"]))
    assert x == int(g["x"])
    util.assert_vcf_close(vcf, g["vcf"])
    gpu.close()


@pytest.mark.parametrize("case", range(4))
def test_process_block_random_vs_oracle(case):
    from oracle.bindings import Oracle
    kw = [dict(depth=30, read_len=150, paired=True, frag_mean=220, indel_frac=0.3, clip_frac=0.3),
          dict(depth=10, read_len=100, paired=True, frag_mean=120, frag_sd=40, indel_frac=0.6, clip_frac=0.4),
          dict(depth=80, read_len=75, paired=False, indel_frac=0.3, clip_frac=0.3, nonconv_frac=0.3),
          dict(depth=25, read_len=250, paired=True, frag_mean=300, single_mate_frac=0.2, indel_frac=0.4, n_frac=0.05)][case]
    trims = [((0, 0), (0, 0)), ((5, 5), (3, 3)), ((10, 0), (0, 10)), ((1, 2), (3, 4))][case]
    rng = np.random.default_rng(300 + case)
    ref = blockgen.random_reference(rng, 40000, n_runs=3)
    T, B, M, y = blockgen.make_block(rng, ref, 300, 36000, **kw)
    o = Oracle(left_trim=trims[0], right_trim=trims[1])
    gpu = bslib.BsGpu(left_trim=trims[0], right_trim=trims[1])
    first = int(T[0]["forward_position"]) or int(T[0]["reverse_position"])
    x = first - 2 if first > 2 else 1
    refw = ref[x - 1:x - 1 + (y - x + 1)]
    xo, pile, want = o.process_block(T, B, M, refw, y)
    xg, vcf = gpu.process_block(T, B, M, refw, y)
    assert xg == xo == x
    n = util.assert_vcf_close(vcf, want)
    assert n > 30000
    gpu.close()


def test_process_block_rejects_bad_cigar(gpu):
    from bs_call_b200.records import MISMS, TEMPLATE
    t = np.zeros(1, dtype=TEMPLATE)
    t["forward_position"] = 10
    t["present"][0, 0] = 1
    t["read_len"][0, 0] = 20
    t["reference_span"][0, 0] = 20
    t["mm_n"][0, 0] = 1
    m = np.zeros(1, dtype=MISMS)
    m[0] = (3, 0, 25)                                     # soft clip longer than the read: fatal in the reference
    with pytest.raises(bslib.BsGpuError):
        gpu.process_block(t, np.full(20, 37 << 2, dtype=np.uint8), m, np.ones(64, dtype=np.uint8), 30)


def test_pileup_very_deep_widens_more_than_once(oracle):
    """> 1500 reads over one tile: the 16-bit packed counters are widened repeatedly"""
    rng = np.random.default_rng(79)
    ref = blockgen.random_reference(rng, 1200)
    T, B, M, y = blockgen.make_block(rng, ref, 100, 500, depth=2500, read_len=100, paired=False, snp_rate=0.01, lowmapq_frac=0.0)
    nt, nb = oracle.normalise_block(T, B, M)
    x = int(T[0]["forward_position"]) - 2
    sz = y - x + 1
    want = oracle.pileup_block(nt, nb, x, y)
    assert want["n"].max() > 1600
    g = bslib.BsGpu()
    segs = g.stage_templates(nt, nb, x, y)
    util.assert_pileup_equal(g.pileup_block(segs, nb, x, sz), want)
    assert g.stats()["qsum_overflow"] == 0
    refw = ref[x - 1:x - 1 + sz]
    vcf = g.call_block(segs, nb, refw, x, sz)
    wout, wskip = oracle.call_sites(want, refw, nthreads=4)
    util.assert_gt_meth_close(vcf["gtm"], vcf["skip"], wout, wskip)
    g.close()


def test_fused_variant_agrees(oracle):
    """the single fused pileup+model kernel (BSGPU_FUSED=1) stays correct"""
    import os
    g0 = util.load_golden("block_mixed")
    x, y = int(g0["x"]), int(g0["y"])
    for var in ("BSGPU_FUSED",):
        os.environ[var] = "1"
        g = bslib.BsGpu()
        del os.environ[var]
        segs = g.stage_templates(g0["norm_templates"], g0["norm_bases"], x, y)
        util.assert_pileup_equal(g.pileup_block(segs, g0["norm_bases"], x, y - x + 1), g0["pileup"])
        util.assert_vcf_close(g.call_block(segs, g0["norm_bases"], g0["ref"], x, y - x + 1), g0["vcf"])
        g.close()


@pytest.mark.parametrize("mode", ["1", "2"])
def test_call_sites_wire_records(gpu, mode, monkeypatch):
    """BSGPU_WIRE: results cross PCIe as 120-byte wire records and are rebuilt by host threads (bsgpu_wire.h) -- the caller's
    arrays must hold the same bytes as with full records, including a chunk that has to be fetched again because a count
    does not fit the wire's 16 bits."""
    rng = np.random.default_rng(77)
    n = 2 * (1 << 18) + 1234                            # three chunks of the default size
    p = np.zeros(n, dtype=PILEUP)
    c = rng.poisson(3.0, (n, 2, 8)).astype(np.uint32)
    c[rng.random(n) < 0.05] = 0
    p["counts"] = c
    p["n"] = c.sum(axis=(1, 2))
    p["quality"] = (c.sum(axis=1) * rng.integers(20, 41, (n, 8))).astype(np.float32)
    p["mapq2"] = (p["n"] * 3600).astype(np.float32)
    p["counts"][(1 << 18) + 5, 0, 4] = 70000          # second chunk: too wide for the wire
    p["n"][(1 << 18) + 5] += 70000
    ref = rng.integers(0, 5, n).astype(np.uint8)
    want, want_skip = gpu.call_sites(p, ref)
    monkeypatch.setenv("BSGPU_WIRE", mode)
    other = bslib.BsGpu(device=0)
    try:
        out = np.full(n, 0, dtype=want.dtype)
        out.view(np.uint8)[:] = 0x5a
        skip = np.full(n, 0x5a, dtype=np.uint8)
        other.call_sites(p, ref, out=out, skip=skip)
        st = other.stats()
    finally:
        other.close()
    assert out.tobytes() == want.tobytes()
    assert (skip == want_skip).all()
    assert st["wire_refetched_chunks"] == 1
    assert 0 < st["wire_sites"] <= n - (1 << 18)
