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


@pytest.fixture(autouse=True, params=["split", "fused"])
def writer_mode(request, monkeypatch):
    """every test of the module runs with the three-kernel writer (default) and with the single-pass kernel k_bcf_fused"""
    if request.param == "fused":
        monkeypatch.setenv("BSGPU_WRITER", "fused")
    else:
        monkeypatch.delenv("BSGPU_WRITER", raising=False)
    return request.param


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


@pytest.mark.parametrize("n", [1, 5, 1000, (1 << 18) - 1, (1 << 18) + 1, (1 << 18) + 2, 2 * (1 << 18) + 3, (1 << 20) + 3, 9 * (1 << 18) + 12345])
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


def _reader_opts(g):
    return dict(mapq_thresh=int(g["mapq_thresh"]), max_template_len=int(g["max_template_len"]), keep_unmatched=bool(g["keep_unmatched"]),
                ignore_duplicates=bool(g["ignore_duplicates"]), keep_duplicates=bool(g["keep_duplicates"]))


def oracle_bcf_of_stream(oracle, bam, tl, refs, opts, all_positions=False):
    """the oracle's chain read_input -> process_template_vector -> call_genotypes_ML, then its writer block by block"""
    bk, _, _, _, vcf = oracle.read_input(bam, tl, refs, run_chain=True, **opts)
    parts, total = [], 0
    for b in bk:
        x, y, tid = int(b["x"]), int(b["y"]), int(b["tid"])
        sz = y - x + 1
        v = vcf[int(b["vcf_off"]):int(b["vcf_off"]) + sz]
        rb, n = oracle.print_block(v, blockgen.window_codes(refs[tid], x, y + 2), x, rid=tid, ctg_end=int(tl[tid]), all_positions=all_positions)
        parts.append(rb)
        total += n
    return bk, (np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)), total


@pytest.mark.parametrize("name", ["reader_pe", "reader_mixed"])
def test_bam_to_records_goldens(gpu, name):
    """raw BAM records -> BCF records, everything between on the device except the block builder, against the records of
    the reference's own chain read_input -> process_template_vector -> call_genotypes_ML -> print_vcf_entry"""
    g = util.load_golden(name)
    w = util.load_golden("writer_v1")
    refs = [g["ref%d" % i] for i in range(len(g["target_len"]))]
    blocks, out, n = gpu.call_bam_bcf(g["bam"], g["target_len"], refs, bslib.reader_params(**_reader_opts(g)))
    assert len(blocks) == len(g["blocks"])
    got, want = util.split_bcf(np.asarray(out)), util.split_bcf(w[name + "__bcf"])
    assert n == len(got) == len(want) == int(w[name + "__nrec"])
    # fixed fields (CHROM, POS, QUAL, allele and field counts) and everything that is not a float: identical; the GL floats
    # carry the last bits of the device's posteriors
    assert [r[:32] for r in got] == [r[:32] for r in want]
    same = sum(a == b for a, b in zip(got, want))
    assert same >= 0.98 * len(want), (same, len(want))


@pytest.mark.parametrize("seed", [2, 5, 9, 14])
def test_bam_to_records_streams(gpu, oracle, monkeypatch, seed):
    """streams with duplicates, junk and gaps through bsgpu_call_bam_bcf (windows that hold several blocks, blocks that
    touch, chunked upload, parallel builder) against the oracle's writer over the DEVICE's gt_vcf[] of the same stream"""
    bam, n, tl, refs = bamgen_stream(seed)
    opts = dict(mapq_thresh=20, max_template_len=1000, keep_unmatched=seed % 2 == 1, ignore_duplicates=False, keep_duplicates=False)
    monkeypatch.setenv("BSGPU_READER_CHUNK_MIN_BYTES", "1")
    monkeypatch.setenv("BSGPU_BUILDER_MIN_RECORDS", "1")
    monkeypatch.setenv("BSGPU_BUILDER_THREADS", "3")
    allp = seed % 4 == 1
    blocks, vcf = gpu.call_bam(bam, tl, refs, bslib.reader_params(**opts))
    parts, total = [], 0
    for b in blocks:
        x, y, tid = int(b["x"]), int(b["y"]), int(b["tid"])
        v = vcf[int(b["vcf_off"]):int(b["vcf_off"]) + y - x + 1]
        rb, k = oracle.print_block(v, blockgen.window_codes(refs[tid], x, y + 2), x, rid=tid + 7, ctg_end=int(tl[tid]), all_positions=allp)
        parts.append(rb)
        total += k
    want = (np.concatenate(parts), total)
    b2, out, nrec = gpu.call_bam_bcf(bam, tl, refs, bslib.reader_params(**opts), bslib.bcf_params(all_positions=allp),
                                     vcf_rid=[t + 7 for t in range(len(tl))])
    assert len(b2) == len(blocks) and total > 500
    same_bcf((np.asarray(out), nrec), want, "seed %d" % seed)


def bamgen_stream(seed):
    from tests import bamgen
    return bamgen.make_stream(seed, dup=0.2, junk=0.1)


def test_edge_cases(gpu, oracle):
    """empty inputs, a stream without a single kept record, buffers that are too small in the middle of a pipeline"""
    from bs_call_b200.records import GT_VCF, PILEUP
    # nothing in, nothing out
    b, n = gpu.bcf_block(np.zeros(0, dtype=GT_VCF), np.zeros(2, dtype=np.uint8), 1)
    assert len(b) == 0 and n == 0
    b, n = gpu.call_sites_bcf(np.zeros(0, dtype=PILEUP), np.zeros(2, dtype=np.uint8), 1)
    assert len(b) == 0 and n == 0
    # a block of skipped sites only
    v = np.zeros(300, dtype=GT_VCF); v["skip"] = 1; v["ready"] = 1
    b, n = gpu.bcf_block(v, np.ones(302, dtype=np.uint8), 10)
    assert len(b) == 0 and n == 0
    # every record of the stream filtered (unmapped): no blocks, no records, no error
    from tests import bamgen
    bam, _, tl, refs = bamgen.make_stream(4)
    bad = bam.copy()
    at = 0
    while at < len(bad):
        bs = int(bad[at:at + 4].view("<i4")[0])
        bad[at + 4 + 14] |= 4                       # BAM_FUNMAP
        at += 4 + bs
    blocks, out, nrec = gpu.call_bam_bcf(bad, tl, refs)
    assert len(blocks) == 0 and len(out) == 0 and nrec == 0
    # output buffers too small: reported, and the context keeps working afterwards
    pile, ref = oracle.synth_sites(9, 0, 2500000, nthreads=8)
    refw = np.concatenate([ref, [1, 1]]).astype(np.uint8)
    with pytest.raises(bslib.BsGpuError):
        gpu.call_sites_bcf(pile, refw, 1, out=np.empty(40 << 20, dtype=np.uint8))          # room for about one chunk of three
    with pytest.raises(bslib.BsGpuError):
        gpu.call_bam_bcf(bam, tl, refs, out=np.empty(20000, dtype=np.uint8))
    b1, n1 = gpu.call_sites_bcf(pile[:70000], refw[:70002], 1)
    gtm, skip = gpu.call_sites(pile[:70000], ref[:70000])
    v = np.zeros(70000, dtype=GT_VCF); v["gtm"] = gtm; v["skip"] = skip; v["ready"] = 1
    same_bcf((b1, n1), oracle.print_block(v, refw[:70002], 1), "after the failed calls")


def test_records_and_side_channels_together(gpu, oracle):
    """--report-file side channels switched on while the records are produced: both as when each runs alone"""
    from tests import bamgen
    bam, _, tl, refs = bamgen.make_stream(11, dup=0.2, junk=0.1)
    opts = dict(mapq_thresh=20, max_template_len=1000, keep_unmatched=False, ignore_duplicates=False, keep_duplicates=False)
    _, plain, nplain = gpu.call_bam_bcf(bam, tl, refs, bslib.reader_params(**opts))
    plain = np.asarray(plain).copy()
    oracle.profile_enable(True); oracle.profile_reset()
    try:
        oracle.read_input(bam, tl, refs, run_chain=True, **opts)
        want = oracle.profile_read()
    finally:
        oracle.profile_enable(False)
    g2 = bslib.BsGpu()
    try:
        g2.profile_enable(True)
        _, rec, nrec = g2.call_bam_bcf(bam, tl, refs, bslib.reader_params(**opts))
        assert nrec == nplain and np.asarray(rec).tobytes() == plain.tobytes()
        util.same_profile(g2.profile_read(), want, "records + side channels")
    finally:
        g2.close()


@pytest.mark.parametrize("seed", range(4))
def test_dbsnp_ids_and_regions(gpu, oracle, seed):
    """-D and -C on the device (src/print_vcf.c:133, 139, 154-157, 163-167): dbSNP IDs in the record, "always written" sites,
    clipping to the region instead of the contig end; the oracle's annotated writer is held to the compiled reference's in
    tests/test_oracle_vs_reference.py"""
    from tests.test_oracle_vs_reference import random_dbsnp
    rng = np.random.default_rng(2900 + seed)
    for sz in (1, 7, 129, 3000, 70001):
        vcf = util.random_gt_vcf(rng, sz, skip_frac=[0.0, 0.3][seed % 2])
        refw = rng.integers(1, 5, size=sz + 2).astype(np.uint8)
        refw[rng.random(sz + 2) < 0.02] = 0
        x = int(rng.integers(1, 100000))
        ctg_end = x + sz - 1 - int(rng.integers(0, 3))
        db = random_dbsnp(rng, max(1, x - 5), x + sz + 5)
        for region in (None, (x + sz // 4, x + (3 * sz) // 4), (x + 2, x + sz + 100)):
            for allp in (False, True):
                for d in (None, db):
                    want = oracle.print_block(vcf, refw, x, rid=2, ctg_end=ctg_end, all_positions=allp, region=region, dbsnp=d)
                    p = bslib.bcf_params(rid=2, ctg_end=ctg_end, all_positions=allp, region=region, dbsnp=bslib.dbsnp(*d) if d is not None else None)
                    got = gpu.bcf_block(vcf, refw, x, p)
                    same_bcf(got, want, "seed %d size %d %r" % (seed, sz, (region, allp, d is not None)))


def test_bam_to_records_with_contig_annotation(oracle, monkeypatch):
    """bsgpu_set_contig_annotation: per-contig dbSNP entries and regions on the many-contig path (bsgpu_call_bam_bcf, and a
    streaming session on the same context), against the oracle's annotated writer block by block"""
    from tests import bamgen
    from tests.test_oracle_vs_reference import random_dbsnp
    bam, n, tl, refs = bamgen.make_stream(77, n_contigs=3, dup=0.1, contig_len=9000)
    rng = np.random.default_rng(78)
    dbs = [random_dbsnp(rng, 1, int(tl[t]), frac=0.1) for t in range(3)]
    regions = [None, (1500, 7000), (1, int(tl[2]) + 50)]
    g = bslib.BsGpu()
    try:
        blocks, vcf = g.call_bam(bam, tl, refs)
        keep = []
        for t in range(3):
            d = bslib.dbsnp(*dbs[t]) if t != 1 else None
            keep.append(d)
            g.set_contig_annotation(t, region=regions[t], dbsnp=d)
        parts, total = [], 0
        for b in blocks:
            x, y, tid = int(b["x"]), int(b["y"]), int(b["tid"])
            v = vcf[int(b["vcf_off"]):int(b["vcf_off"]) + y - x + 1]
            rb, k = oracle.print_block(v, blockgen.window_codes(refs[tid], x, y + 2), x, rid=tid, ctg_end=int(tl[tid]),
                                       region=regions[tid], dbsnp=dbs[tid] if tid != 1 else None)
            parts.append(rb)
            total += k
        want = (np.concatenate(parts), total)
        _, out, nrec = g.call_bam_bcf(bam, tl, refs)
        same_bcf((np.asarray(out).copy(), nrec), want, "call_bam_bcf")
        s = g.bam_session(tl, refs, bcf=True, batch_bytes=30000)
        try:
            res = s.run(bam, slice_bytes=7777)
        finally:
            s.close()
        same_bcf((np.concatenate([d for _, d, _ in res]), sum(k for _, _, k in res)), want, "session")
        assert total > 300
    finally:
        g.close()
