"""Guard bands (-m gpu): a call, a QUAL or an FS that hangs on the last bits of a likelihood is REPORTED by the device
(bsgpu_stats.near_tie_sites / exact_tie_sites / near_qual_sites / near_fs_sites, bsgpu_guard_read), never waived by the
tests: max_gt must equal the oracle's at every site the device did not flag.  src/genotype_model.c:231-239 takes the
first strict maximum of ten sums; for several genotype pairs the reference adds the same terms in a different order, so
low-depth sites with three or more alleles tie up to one ulp -- which way such a tie falls depends on libm's last bit."""
import itertools

import numpy as np
import pytest

from bs_call_b200 import lib as bslib
from bs_call_b200.records import PILEUP
from tests import blockgen, util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    g = bslib.BsGpu()
    yield g
    g.close()


def _low_depth_sites():
    """every multiset of up to four reads over the 16 (strand index, class) cells x three qualities x every reference code:
    72 660 sites, 9 % of which tie at the top in the oracle (exactly, or by an ulp)"""
    rows = [(comb, q, rf) for depth in (1, 2, 3, 4) for comb in itertools.combinations_with_replacement(range(16), depth)
            for q in (20, 30, 37) for rf in range(5)]
    p = np.zeros(len(rows), dtype=PILEUP)
    ref = np.zeros(len(rows), dtype=np.uint8)
    for i, (comb, q, rf) in enumerate(rows):
        for c in comb:
            p["counts"][i, c // 8, c % 8] += 1
            p["quality"][i, c % 8] += q
        p["n"][i] = len(comb)
        p["mapq2"][i] = 3600 * len(comb)
        ref[i] = rf
    return p, ref


def test_every_site_that_may_flip_is_flagged(gpu, oracle):
    p, ref = _low_depth_sites()
    want, wskip = oracle.call_sites(p, ref, nthreads=4)
    gpu.guard_read(reset=True)
    s0 = gpu.stats()
    got, skip = gpu.call_sites(p, ref)
    kinds, ids = gpu.guard_read(reset=True)
    s1 = gpu.stats()
    flagged = ids[kinds == 1]
    assert len(set(flagged.tolist())) == len(flagged) == (s1["near_tie_sites"] - s0["near_tie_sites"]) + (s1["exact_tie_sites"] - s0["exact_tie_sites"])
    rep = {}
    util.assert_gt_meth_close(got, skip, want, wskip, flagged=flagged, report=rep)
    # the flags are not a blanket: they sit exactly where the oracle's own two best posteriors are (nearly) equal
    tie = util.near_tie(want["gt_prob"])
    fl = np.zeros(len(p), dtype=bool)
    fl[flagged] = True
    assert (fl == tie).all(), "flagged %d, tied in the oracle %d, both %d" % (fl.sum(), tie.sum(), (fl & tie).sum())
    assert 0.05 < fl.mean() < 0.15
    print("guard: %d of %d low-depth sites flagged, %d of them called differently from the oracle" % (rep["flagged"], len(p), rep["flagged_differ"]))


def test_ordinary_depth_has_no_unflagged_difference(gpu, oracle):
    rng = np.random.default_rng(77)
    p, ref = blockgen.random_pileups(rng, 100000, depth=30, het_frac=0.05)
    want, wskip = oracle.call_sites(p, ref, nthreads=4)
    gpu.guard_read(reset=True)
    got, skip = gpu.call_sites(p, ref)
    kinds, ids = gpu.guard_read(reset=True)
    rep = {}
    util.assert_gt_meth_close(got, skip, want, wskip, flagged=ids[kinds == 1], report=rep)
    assert rep["flagged"] < 100            # a 30x site rarely ties: the band is not a licence
    print("guard: %d of 100000 30x sites flagged, %d differ" % (rep["flagged"], rep["flagged_differ"]))


def test_writer_bands_cover_every_integer_that_differs(gpu, oracle):
    """QUAL / GQ and FS are truncations of doubles: the device reports the records whose value sits within its error band of an
    integer; everywhere else the bytes must be the oracle writer's"""
    rng = np.random.default_rng(78)
    v = util.random_gt_vcf(rng, 200000)
    # plant posteriors and strand biases that land exactly on an integer QUAL / FS
    g = v["gtm"]
    k = np.arange(0, len(v), 97)
    g["gt_prob"][k, g["max_gt"][k]] = np.log10(1.0 - 10.0 ** (-(rng.integers(1, 60, size=len(k))) / 10.0))
    g["fisher_strand"][k] = -(rng.integers(0, 90, size=len(k)) + 0.5) / 10.0
    ref = rng.integers(1, 5, size=len(v) + 2).astype(np.uint8)
    gpu.guard_read(reset=True)
    s0 = gpu.stats()
    gb, gn = gpu.bcf_block(v, ref, 1000)
    kinds, ids = gpu.guard_read(reset=True)
    s1 = gpu.stats()
    wb, wn = oracle.print_block(v, ref, 1000)
    assert gn == wn
    got, want = util.split_bcf(gb), util.split_bcf(wb)
    pos_flagged = set((ids[kinds >= 2]).tolist())
    assert s1["near_qual_sites"] - s0["near_qual_sites"] == (kinds == 2).sum() and s1["near_fs_sites"] - s0["near_fs_sites"] == (kinds == 3).sum()
    assert len(pos_flagged) >= len(k) // 4            # the planted ones are seen
    bad = 0
    for a, b in zip(got, want):
        if a != b:
            pos = int(np.frombuffer(a[12:16], dtype="<i4")[0]) + 1
            assert pos in pos_flagged, "record at %d differs from the oracle writer's and was not flagged" % pos
            bad += 1
    print("guard: %d records flagged (QUAL %d, FS %d), %d differ from the oracle writer" % (len(pos_flagged), (kinds == 2).sum(), (kinds == 3).sum(), bad))
