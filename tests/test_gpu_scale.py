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


def test_record_stream_properties_at_8m_sites(gpu):
    """8 M sites of the config-2 stream all the way to BCF records (bsgpu_call_sites_bcf, 31 chunks): the stream parses
    into whole records; positions strictly increase; the sites that have a record, their QUAL, genotype filter and allele
    counts are what an independent numpy derivation from the gt_meth records of the same sites gives
    (tests/util.py:writer_fields, src/print_vcf.c:139-217); two runs are byte-identical"""
    import struct
    from tests import util
    from bs_call_b200.records import PILEUP
    n = 8_000_000
    d_p = torch.empty(n * 104 + 16, dtype=torch.uint8, device="cuda")
    d_r = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    gpu.synth_sites_dev(SEED, 12345, n, 30.0, d_p.data_ptr(), d_r.data_ptr(), 0)
    torch.cuda.synchronize()
    pile = d_p[:n * 104].cpu().numpy().view(PILEUP)
    ref = d_r[:n].cpu().numpy()
    refw = np.concatenate([ref, [1, 1]]).astype(np.uint8)
    x0 = 1000
    b1, n1 = gpu.call_sites_bcf(pile, refw, x0)
    b1 = np.asarray(b1).copy()
    b2, n2 = gpu.call_sites_bcf(pile, refw, x0)
    assert n1 == n2 and b1.tobytes() == np.asarray(b2).tobytes()
    # walk the records
    buf = memoryview(b1.tobytes())
    pos = np.empty(n1, dtype=np.int64); qual = np.empty(n1, dtype=np.float32); nal = np.empty(n1, dtype=np.int32); nfmt = np.empty(n1, dtype=np.int32)
    at, k = 0, 0
    unpack = struct.Struct("<IIiiifII").unpack_from
    while at < len(buf):
        ls, li, rid, p, rlen, q, w6, w7 = unpack(buf, at)
        assert ls >= 24 + 12 and li > 40 and rid == 0 and rlen == 1 and (w6 & 0xffff) == 1 and (w7 & 0xffffff) == 1
        pos[k], qual[k], nal[k], nfmt[k] = p, q, w6 >> 16, w7 >> 24
        at += 8 + ls + li
        k += 1
    assert at == len(buf) and k == n1
    assert (np.diff(pos) > 0).all() and pos[0] >= x0 - 1 and pos[-1] <= x0 + n - 2
    # the independent derivation
    gtm, skip = gpu.call_sites(pile, ref)
    f = util.writer_fields(gtm, skip)
    called = np.nonzero(skip == 0)[0]
    gt, rf = f["gt"], ref[called]
    emitted = ~(((gt == 0) & (rf == 1)) | ((gt == 9) & (rf == 4)))          # hom-ref A / T is skipped without -A
    want_pos = called[emitted] + x0 - 1
    assert len(want_pos) == n1 and (want_pos == pos).all()
    assert (qual.astype(np.int64) == f["phred"][emitted]).all()
    a0 = np.array([1, 1, 1, 1, 2, 2, 2, 3, 3, 4])[gt[emitted]]
    a1 = np.array([1, 2, 3, 4, 2, 3, 4, 3, 4, 4])[gt[emitted]]
    r = rf[emitted]
    n_alt = (a0 != r).astype(np.int32) + ((a1 != r) & (a1 != a0)).astype(np.int32)
    assert (nal == 1 + n_alt).all()
    het = np.isin(gt[emitted], [1, 2, 3, 5, 6, 8])
    assert (nfmt == 12 + het).all()          # 11 + AMQ (a called site has a non-empty count slot) + FS for heterozygotes
