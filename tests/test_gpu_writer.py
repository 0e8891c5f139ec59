"""Writer-side parity (-m gpu): the BCF records the device builds from gt_vcf[] (k_bcf_calls / k_bcf_measure / k_bcf_offsets /
k_bcf_emit behind bsgpu_bcf_block, bsgpu_call_block_bcf, bsgpu_call_sites_bcf) against the oracle's restatement of the
reference's print_vcf_entry / flush_vcf_entries / _print_vcf_entry, which tests/test_oracle_vs_reference.py holds byte-identical to
the reference's compiled src/print_vcf.c, and against goldens captured from that (tests/golden/writer_v1.npz).
Integer / byte work: bit-exact.  QUAL / GQ go through exp and log of a posterior; the device's table-driven versions and
libm agree on the integer everywhere in these sets (asserted)."""
import numpy as np
import pytest

from bs_call_b200 import lib as bslib
from tests import blockgen, util

pytestmark = pytest.mark.gpu

BLOCKS = ["block_pe_plain", "block_pe_indel_clip_trim", "block_se_deep", "block_mixed"]


@pytest.fixture(scope="module")
def gpu():
    g = bslib.BsGpu()
    yield g
    g.close()


def same_bcf(got, want, what):
    gb, gn = got
    wb, wn = want
    if gb.tobytes() == wb.tobytes() and gn == wn:
        return
    a, b = util.split_bcf(np.asarray(gb)), util.split_bcf(np.asarray(wb))
    for k, (ra, rb) in enumerate(zip(a, b)):
        assert ra == rb, "%s: record %d of %d differs\n got  %s\n want %s" % (what, k, len(b), ra.hex(), rb.hex())
    raise AssertionError("%s: %d records vs %d" % (what, len(a), len(b)))


@pytest.mark.parametrize("name", BLOCKS)
def test_goldens_from_reference(gpu, name):
    """records of the reference's own writer for the reference's own gt_vcf[] of the block goldens"""
    g = util.load_golden(name)
    w = util.load_golden("writer_v1")
    for allp in (0, 1):
        want = (w["%s__all%d" % (name, allp)], int(w["%s__n%d" % (name, allp)]))
        got = gpu.bcf_block(g["vcf"], w[name + "__ref"], int(g["x"]), bslib.bcf_params(all_positions=bool(allp), rid=2))
        same_bcf(got, want, "%s all_positions=%d" % (name, allp))


@pytest.mark.parametrize("seed", range(6))
def test_random_records_match_oracle(gpu, oracle, seed):
    """records built to reach every branch of the writer, N runs in the reference window, blocks of every small size and
    sizes around the CTA width, header ids of one, two and four bytes, a contig end inside the block"""
    rng = np.random.default_rng(900 + seed)
    ids = [list(range(16)), [3, 200, 5, 40000, 7, 100000, 9, 11, 127, 128, 13, 15, 17, 19, 21, 23]][seed % 2]
    for sz in list(range(1, 9)) + [40, 127, 128, 129, 333, 2000, 20011]:
        vcf = util.random_gt_vcf(rng, sz, skip_frac=[0.0, 0.25, 0.6][seed % 3])
        refw = rng.integers(1, 5, size=sz + 2).astype(np.uint8)
        refw[rng.random(sz + 2) < [0.0, 0.03, 0.3][(seed // 2) % 3]] = 0
        x = int(rng.integers(1, 1000))
        ctg_end = x + sz - 1 - int(rng.integers(0, 3))
        for allp in (False, True):
            want = oracle.print_block(vcf, refw, x, rid=3, ctg_end=ctg_end, vcf_ids=ids, all_positions=allp)
            got = gpu.bcf_block(vcf, refw, x, bslib.bcf_params(ids=ids, rid=3, ctg_end=ctg_end, all_positions=allp))
            same_bcf(got, want, "seed %d size %d all_positions %r" % (seed, sz, allp))


def test_block_path_to_records(gpu, oracle):
    """segments -> records with nothing but the records coming back, against the oracle's writer over the device's own
    gt_vcf[] of the same block (so the comparison is exact whatever the last bits of the posteriors are)"""
    for name in BLOCKS:
        g = util.load_golden(name)
        x, y = int(g["x"]), int(g["y"])
        sz = y - x + 1
        segs = bslib.stage_templates_host(g["norm_templates"], g["norm_bases"], x, y)
        refw = util.load_golden("writer_v1")[name + "__ref"]
        vcf = gpu.call_block(segs, g["norm_bases"], refw[:sz], x, sz)
        want = oracle.print_block(vcf, refw, x, rid=2, all_positions=False)
        before = gpu.stats()
        got = gpu.call_block_bcf(segs, g["norm_bases"], refw, x, sz, bslib.bcf_params(rid=2))
        after = gpu.stats()
        same_bcf(got, want, name)
        assert after["d2h_bytes"] - before["d2h_bytes"] < 0.3 * sz * 208          # only the records come home
        # and those are the reference's records for this block, up to sites whose posteriors differ in the last bits
        w = util.load_golden("writer_v1")
        ref_recs = util.split_bcf(w["%s__all0" % name])
        got_recs = util.split_bcf(np.asarray(got[0]))
        assert len(ref_recs) == len(got_recs)
        same = sum(a[:32] == b[:32] for a, b in zip(ref_recs, got_recs))          # fixed fields: POS, QUAL, allele / field counts
        assert same == len(ref_recs)


@pytest.mark.parametrize("n", [1, 5, 1000, (1 << 20) - 1, (1 << 20) + 3, 3 * (1 << 20) + 12345])
def test_sites_to_records_pipeline(gpu, oracle, n):
    """count vectors -> records in chunks (the writer runs one chunk behind the model): the chunk seams must not show"""
    pile, ref = oracle.synth_sites(7, 1000, n, nthreads=8)
    refw = np.concatenate([ref, [1, 2]]).astype(np.uint8)
    gtm, skip = gpu.call_sites(pile, ref)
    from bs_call_b200.records import GT_VCF
    vcf = np.zeros(n, dtype=GT_VCF)
    vcf["gtm"] = gtm; vcf["skip"] = skip; vcf["ready"] = 1
    want = oracle.print_block(vcf, refw, 5, rid=0)
    got = gpu.call_sites_bcf(pile, refw, 5)
    same_bcf(got, want, "n = %d" % n)
    assert got[1] > 0 or n < 5


def test_output_too_small_is_an_error(gpu):
    g = util.load_golden("block_pe_plain")
    refw = util.load_golden("writer_v1")["block_pe_plain__ref"]
    with pytest.raises(bslib.BsGpuError):
        gpu.bcf_block(g["vcf"], refw, int(g["x"]), out=np.empty(1000, dtype=np.uint8))
