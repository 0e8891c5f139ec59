"""Size-independent properties at BASELINE.json's full sizes (-m gpu), where the oracle cannot follow:

  config 2  1e9 synthetic per-site count vectors through the likelihood kernel, in slabs: two runs are byte-identical,
            skip <=> n == 0, max_gt is the first strict maximum of gt_prob[], the ten posteriors sum to one, counts and
            qualities are the summarised inputs
  config 3  a 50 M-site 30x window through the pileup kernel: two runs byte-identical, additive over a split of the
            segments, and the counted bases of the window are conserved (sum of n == bytes that count in the input)

torch is used for device memory and for the reductions that check the properties; the kernels under test are reached
through the C ABI like everywhere else."""
import numpy as np
import pytest

from bs_call_b200 import lib as bslib

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

SEED = 20261018


@pytest.fixture(scope="module")
def gpu():
    g = bslib.BsGpu()
    yield g
    g.close()


def test_likelihood_kernel_properties_at_1e9_sites(gpu):
    total, slab = 1_000_000_000, 50_000_000
    d_p = torch.empty(slab * 104 + 16, dtype=torch.uint8, device="cuda")
    d_r = torch.empty(slab + 16, dtype=torch.uint8, device="cuda")
    d_o = torch.empty(slab * 200 + 16, dtype=torch.uint8, device="cuda")
    d_o2 = torch.empty(slab * 200 + 16, dtype=torch.uint8, device="cuda")
    d_s = torch.empty(slab + 16, dtype=torch.uint8, device="cuda")
    called = 0
    for first in range(0, total, slab):
        gpu.synth_sites_dev(SEED, first, slab, 30.0, d_p.data_ptr(), d_r.data_ptr())
        gpu.call_sites_dev(d_p.data_ptr(), d_r.data_ptr(), slab, d_o.data_ptr(), d_s.data_ptr())
        gpu.sync()
        pile = d_p[:slab * 104].view(torch.int32).view(slab, 26)
        out64 = d_o[:slab * 200].view(torch.int64).view(slab, 25)
        n = pile[:, 16]
        skip = d_s[:slab]
        assert bool(((n == 0) == (skip == 1)).all())
        live = n > 0
        called += int(live.sum())
        # counts[j] = counts[0][j] + counts[1][j]
        assert bool((out64[:, :8][live] == (pile[:, :8] + pile[:, 8:16]).to(torch.int64)[live]).all())
        prob = d_o[:slab * 200].view(torch.float64).view(slab, 25)[:, 12:22]
        best = out64[:, 24] & 0xff
        assert bool((torch.argmax(prob, dim=1)[live] == best[live]).all())          # argmax returns the first maximum
        tot = torch.pow(10.0, prob[live]).sum(dim=1)
        assert float((tot - 1.0).abs().max()) < 1e-9
        assert bool((out64[~live] == 0).all())                                      # skipped sites are zero records
        if first % (5 * slab) == 0:                                                 # determinism on every fifth slab
            gpu.call_sites_dev(d_p.data_ptr(), d_r.data_ptr(), slab, d_o2.data_ptr(), d_s.data_ptr())
            gpu.sync()
            assert torch.equal(d_o[:slab * 200], d_o2[:slab * 200])
    assert 0.96 * total < called < 0.98 * total          # 3 % empty sites in the stream


def test_pileup_kernel_properties_at_50m_sites(gpu):
    sz, L, depth, x = 50_000_000, 150, 30.0, 1000
    ns = gpu.synth_block_nseg(sz, L, depth)
    d_seg = torch.empty(ns * 16 + 16, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(ns * L + 16, dtype=torch.uint8, device="cuda")
    d_r = torch.empty(sz + 16, dtype=torch.uint8, device="cuda")
    gpu.synth_block_dev(SEED, x, sz, L, depth, d_seg.data_ptr(), ns, d_b.data_ptr(), ns * L, d_r.data_ptr())
    outs = [torch.empty(sz * 104 + 16, dtype=torch.uint8, device="cuda") for _ in range(3)]
    gpu.pileup_block_dev(d_seg.data_ptr(), ns, d_b.data_ptr(), x, sz, outs[0].data_ptr())
    gpu.pileup_block_dev(d_seg.data_ptr(), ns, d_b.data_ptr(), x, sz, outs[1].data_ptr())
    gpu.sync()
    assert torch.equal(outs[0][:sz * 104], outs[1][:sz * 104])                      # deterministic
    # additive: even-numbered segments + odd-numbered segments = all segments (integer counts; the float sums hold integers)
    segs = d_seg[:ns * 16].view(torch.int32).view(ns, 4)
    half = [segs[0::2].contiguous(), segs[1::2].contiguous()]
    acc = None
    for h in half:
        gpu.pileup_block_dev(h.data_ptr(), h.shape[0], d_b.data_ptr(), x, sz, outs[2].data_ptr())
        gpu.sync()
        rec = outs[2][:sz * 104].view(torch.int32).view(sz, 26)
        ints = rec[:, :17].to(torch.int64)
        flts = rec[:, 17:].view(torch.float32).to(torch.float64)
        acc = (ints, flts) if acc is None else (acc[0] + ints, acc[1] + flts)
    rec = outs[0][:sz * 104].view(torch.int32).view(sz, 26)
    assert bool((rec[:, :17].to(torch.int64) == acc[0]).all())
    assert bool((rec[:, 17:].view(torch.float32).to(torch.float64) == acc[1]).all())
    # conservation: every byte of every segment that counts (min_qual <= q != 63) lands in exactly one site's n
    lens = (segs[:, 2] & 0xffff).to(torch.int64)
    offs = segs[:, 1].to(torch.int64) & 0xffffffff
    assert bool((offs == torch.arange(ns, device="cuda") * L).all())               # the generator's layout: read i at i * L
    q = (d_b[:ns * L].view(ns, L) >> 2).to(torch.int32)
    inside = torch.arange(L, device="cuda")[None, :] < lens[:, None]
    counted = ((q >= 20) & (q != 63) & inside).sum()
    assert int(rec[:, 16].to(torch.int64).sum()) == int(counted)
    assert gpu.stats()["qsum_overflow"] == 0
