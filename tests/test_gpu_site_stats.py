"""The writer's --report-file statistics on the device (-m gpu): k_bcf_stats behind bsgpu_site_stats_enable / _read against the
oracle's restatement of src/print_vcf.c:382-526 (oracle/bs_oracle_stats.c), which tests/test_oracle_vs_reference.py holds
identical to the compiled print_vcf.c with a live bs_stats.  Counters: bit-exact.  The two methylation posteriors are sums of
doubles gathered in device order: 1e-9 relative (observed ~1e-13)."""
import numpy as np
import pytest

from bs_call_b200 import lib as bslib
from oracle.bindings import SITE_STATS, site_stats_equal
from tests import blockgen, util
from tests.test_oracle_vs_reference import cpg_rich, random_dbsnp

pytestmark = pytest.mark.gpu


@pytest.fixture()
def gpu():
    g = bslib.BsGpu()
    yield g
    g.close()


def test_random_blocks(gpu, oracle):
    """records that reach every branch, rich in called CpGs on reference CpGs and elsewhere; dbSNP, regions, -A, a contig end
    inside the block, GC bins that start before the block and end inside it, sizes around the CTA width"""
    rng = np.random.default_rng(4100)
    want = np.zeros(1, dtype=SITE_STATS)
    state = np.zeros(2, dtype=np.uint32)
    gpu.site_stats_enable(True, n_contigs=4)
    x = 1
    per_rid = np.zeros((4, 12), dtype=np.uint64)
    for sz in (1, 2, 5, 127, 128, 129, 700, 4000, 33333):
        vcf = cpg_rich(rng, util.random_gt_vcf(rng, sz, skip_frac=0.2, deep_frac=0.02))
        refw = rng.integers(1, 5, size=sz + 2).astype(np.uint8)
        refw[rng.random(sz + 2) < 0.02] = 0
        for i in np.flatnonzero(rng.random(sz) < 0.3):
            refw[i], refw[i + 1] = 2, 3
        x += int(rng.integers(1, 500))
        start_pos = int(rng.integers(1, x + 1))
        gc = rng.integers(0, 120, size=(x + sz - start_pos) // 100 + 1 - int(rng.integers(0, 2))).astype(np.uint8)
        ctg_end = x + sz - 1 - int(rng.integers(0, 3))
        db = random_dbsnp(rng, max(1, x - 5), x + sz + 5)
        rid = int(rng.integers(0, 4))
        if len(gc):
            gpu.set_contig_gc(rid, gc, start_pos)
        else:
            gpu.set_contig_gc(rid, np.zeros(0, dtype=np.uint8), start_pos)
        for region in (None, (x + sz // 4, x + (3 * sz) // 4)):
            for allp in (False, True):
                for d in (None, db):
                    before = want.copy()
                    oracle.stats_block(vcf, refw, x, ctg_end=ctg_end, all_positions=allp, region=region, dbsnp=d,
                                       gc=gc if len(gc) else None, start_pos=start_pos, stats=want, state=state)
                    for k, f in enumerate(("snps", "multi", "dbSNP_sites", "dbSNP_var", "CpG_ref", "CpG_nonref")):
                        per_rid[rid, 2 * k:2 * k + 2] += want[0][f] - before[0][f]
                    p = bslib.bcf_params(rid=rid, ctg_end=ctg_end, all_positions=allp, region=region, dbsnp=bslib.dbsnp(*d) if d is not None else None)
                    gpu.bcf_block(vcf, refw, x, p)
        x += sz
    got, ctg = gpu.site_stats_read(n_contigs=4)
    site_stats_equal(got[0], want[0], rtol=1e-9, what="random blocks")
    flat = np.stack([np.concatenate([ctg[r][f] for f in ("snps", "multi", "dbSNP_sites", "dbSNP_var", "CpG_ref", "CpG_nonref")]) for r in range(4)])
    assert np.array_equal(flat, per_rid)
    assert want[0]["CpG_ref"][0] > 100 and want[0]["CpG_nonref"][0] > 100 and want[0]["CpG_ref_meth"].sum() > 100
    assert want[0]["cov"]["gc_pcent"].sum() > 10000 and want[0]["mut_counts"].sum() > 1000 and want[0]["fs_stats"][64:].sum() > 100
    # reset
    gpu.site_stats_read(reset=True)
    again, _ = gpu.site_stats_read()
    assert not again.view(np.uint8).any()


@pytest.mark.parametrize("n", [1000, (1 << 18) + 1, 2 * (1 << 18) + 3, 5 * (1 << 18) + 777])
def test_chunked_sites_pipeline(gpu, oracle, n):
    """count vectors -> records in chunks: the statistics of a '-' strand CpG on the first site of a chunk need the site
    before it, whose record has left the ring by then (the carry word of k_bcf_stats)"""
    pile, ref = oracle.synth_sites(11, 5000, n, nthreads=8)
    refw = np.concatenate([ref, [1, 2]]).astype(np.uint8)
    gtm, skip = gpu.call_sites(pile, ref)
    from bs_call_b200.records import GT_VCF
    vcf = np.zeros(n, dtype=GT_VCF)
    vcf["gtm"] = gtm; vcf["skip"] = skip; vcf["ready"] = 1
    gc = (np.arange(n // 100 + 2) % 101).astype(np.uint8)
    want, _ = oracle.stats_block(vcf, refw, 5, gc=gc, start_pos=1)
    gpu.site_stats_enable(True)
    gpu.set_contig_gc(0, gc, 1)
    gpu.call_sites_bcf(pile, refw, 5)
    got, _ = gpu.site_stats_read()
    site_stats_equal(got[0], want[0], rtol=1e-9, what="n = %d" % n)
    assert want[0]["snps"][0] + want[0]["multi"][0] > 100


def test_bam_to_records_with_statistics(oracle):
    """the whole path (BAM records -> BCF records) with the statistics on: many contigs, many windows per call, a session"""
    from tests import bamgen
    bam, n, tl, refs = bamgen.make_stream(91, n_contigs=3, dup=0.1, contig_len=9000)
    rng = np.random.default_rng(92)
    gcs = [rng.integers(0, 101, size=int(tl[t]) // 100 + 1).astype(np.uint8) for t in range(3)]
    g = bslib.BsGpu()
    try:
        blocks, vcf = g.call_bam(bam, tl, refs)
        want = np.zeros(1, dtype=SITE_STATS)
        state = np.zeros(2, dtype=np.uint32)
        for b in blocks:
            x, y, tid = int(b["x"]), int(b["y"]), int(b["tid"])
            v = vcf[int(b["vcf_off"]):int(b["vcf_off"]) + y - x + 1]
            oracle.stats_block(v, blockgen.window_codes(refs[tid], x, y + 2), x, ctg_end=int(tl[tid]), gc=gcs[tid], start_pos=1, stats=want, state=state)
        g.site_stats_enable(True, n_contigs=3)
        for t in range(3):
            g.set_contig_gc(t, gcs[t], 1)
        g.call_bam_bcf(bam, tl, refs)
        got, ctg = g.site_stats_read(n_contigs=3, reset=True)
        site_stats_equal(got[0], want[0], rtol=1e-9, what="call_bam_bcf")
        assert int(ctg["snps"][:, 0].sum()) == int(want[0]["snps"][0]) and (ctg["snps"][:, 0] > 0).all()
        s = g.bam_session(tl, refs, bcf=True, batch_bytes=30000)
        try:
            s.run(bam, slice_bytes=7777)
        finally:
            s.close()
        got, _ = g.site_stats_read(n_contigs=0)
        site_stats_equal(got[0], want[0], rtol=1e-9, what="session")
        assert want[0]["snps"][0] + want[0]["multi"][0] > 300
    finally:
        g.close()
