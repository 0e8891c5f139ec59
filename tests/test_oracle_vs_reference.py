"""Pins the restatement (oracle/bs_oracle.c) against the reference's own compiled objects
(oracle/_ref/libbsref.so = unmodified /root/reference sources behind oracle/ref_harness.c).

Both run on the same CPU with the same libm, so every field -- including the doubles -- must be
bit-identical.  Skipped where the prebuilt reference library is absent.
"""
import os

import numpy as np
import pytest

from tests import blockgen, util

pytestmark = pytest.mark.reference


def _same_records(a, b, what):
    assert a.dtype == b.dtype and a.shape == b.shape
    if a.tobytes() != b.tobytes():
        for name in a.dtype.names:
            if name.startswith("pad"):
                continue
            if a[name].tobytes() != b[name].tobytes():
                bad = np.nonzero(np.any((a[name] != b[name]).reshape(len(a), -1), axis=1))[0]
                raise AssertionError("%s: field %s differs at %d sites, first %d: %r vs %r"
                                     % (what, name, len(bad), bad[0], a[name][bad[0]], b[name][bad[0]]))


def test_kat_from_survey(oracle, reference):
    # captured from the reference with defaults
    c = [0, 14, 0, 0, 0, 10, 0, 5]
    q = [35 if v else 0 for v in c]
    want = [-111.48773751377976, -8.8528005438194999, -111.48773751377976, -93.107739053342456,
            -6.5076274832880412e-06, -8.8528005438194999, -4.8244462420618976, -111.48773751377976,
            -93.107739053342456, -91.602818028647008]
    for impl in (oracle, reference):
        g = impl.calc_gt_prob(c, q, 2)
        assert g["max_gt"] == 4
        np.testing.assert_allclose(g["gt_prob"], want, rtol=1e-13)
        assert abs(impl.fisher([12, 3, 4, 11]) - 0.0092205703133985545) < 1e-16


def test_tables(oracle, reference):
    assert oracle.lfact_table().tobytes() == reference.lfact_table().tobytes()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_calc_gt_prob_random(oracle, reference, seed):
    rng = np.random.default_rng(seed)
    n = 4000
    counts = np.zeros((n, 8), dtype=np.uint64)
    for i in range(n):
        k = rng.integers(0, 6)
        cls = rng.choice(8, size=k, replace=False) if k else []
        for c in cls:
            counts[i, c] = int(rng.integers(1, 60)) if rng.random() < 0.95 else int(rng.integers(60, 5000))
    qual = np.where(counts > 0, rng.integers(1, 44, size=(n, 8)), 0).astype(np.int32)
    rf = rng.integers(0, 5, size=n).astype(np.uint8)
    want = reference.calc_gt_prob_batch(counts, qual, rf)
    for i in range(n):
        got = oracle.calc_gt_prob(counts[i], qual[i], rf[i])
        assert got["max_gt"] == want[i]["max_gt"], (i, counts[i], qual[i], rf[i])
        assert got["gt_prob"].tobytes() == want[i]["gt_prob"].tobytes(), (i, counts[i], qual[i], rf[i])


def test_calc_gt_prob_other_params(reference):
    from oracle.bindings import Oracle, Reference
    o = Oracle(under_conv=0.02, over_conv=0.1, ref_bias=1.0)
    r = Reference(under_conv=0.02, over_conv=0.1, ref_bias=1.0)
    try:
        rng = np.random.default_rng(7)
        for _ in range(500):
            c = rng.integers(0, 30, size=8) * (rng.random(8) < 0.5)
            q = np.where(c > 0, rng.integers(1, 44, size=8), 0)
            rf = int(rng.integers(0, 5))
            a, b = o.calc_gt_prob(c, q, rf), r.calc_gt_prob(c, q, rf)
            assert a["max_gt"] == b["max_gt"] and a["gt_prob"].tobytes() == b["gt_prob"].tobytes()
    finally:
        Reference()  # restore defaults for the other tests


def test_fisher_random(oracle, reference):
    rng = np.random.default_rng(11)
    tabs = [rng.integers(0, 40, size=4) for _ in range(3000)]
    tabs += [rng.integers(0, 400, size=4) for _ in range(1000)]       # margins >= 256 -> lgamma path
    tabs += [np.array(t) for t in ([0, 0, 0, 0], [5, 0, 0, 5], [0, 7, 7, 0], [1, 0, 0, 0], [300, 2, 1, 280])]
    for t in tabs:
        a, b = oracle.fisher(t), reference.fisher(t)
        assert a == b or (np.isnan(a) and np.isnan(b)), (t, a, b)


CASES = [
    dict(depth=30, read_len=100, paired=True),
    dict(depth=30, read_len=100, paired=True, indel_frac=0.3, clip_frac=0.3),
    dict(depth=12, read_len=75, paired=True, frag_mean=110, frag_sd=30, indel_frac=0.5, clip_frac=0.2),
    dict(depth=60, read_len=50, paired=False, indel_frac=0.2, clip_frac=0.2, nonconv_frac=0.3),
    dict(depth=20, read_len=100, paired=True, single_mate_frac=0.3, indel_frac=0.2, n_frac=0.05),
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("trims", [((0, 0), (0, 0)), ((5, 3), (2, 4))])
def test_process_block(reference, case, trims):
    from oracle.bindings import Oracle, Reference
    lt, rt = trims
    rng = np.random.default_rng(100 + case)
    ref = blockgen.random_reference(rng, 6000, n_runs=2)
    T, B, M, y = blockgen.make_block(rng, ref, 200, 4800, **CASES[case])
    o = Oracle(left_trim=lt, right_trim=rt)
    r = Reference(left_trim=lt, right_trim=rt)
    try:
        x, pile_r, vcf_r, ref_r, nt_r, nb_r = r.process_block(T, B, M, ref, y)
    finally:
        Reference()
    nt_o, nb_o = o.normalise_block(T, B, M)
    # the normalised reads
    assert nb_o.tobytes() == nb_r.tobytes()
    for f in ("forward_position", "reverse_position", "read_len", "read_off", "present"):
        assert (nt_o[f] == nt_r[f]).all(), f
    # ref window the reference decoded (last contig base reads as N, src/get_sequence.c:41)
    sz = y - x + 1
    np.testing.assert_array_equal(ref_r, ref[x - 1:x - 1 + sz])
    xo, pile_o, vcf_o = o.process_block(T, B, M, ref_r, y)
    assert xo == x
    _same_records(pile_o, pile_r, "pileup")
    _same_records(vcf_o, vcf_r, "gt_vcf")
    assert (vcf_r["skip"] == 0).sum() > 1000
    # the reference gives the same answer through call_genotypes_ML alone on the normalised block
    pile_c, vcf_c = r.call_block(nt_r, nb_r, np.concatenate([ref_r, [0, 0]]).astype(np.uint8), x, y)
    _same_records(pile_c, pile_r, "pileup via call_block")
    _same_records(vcf_c, vcf_r, "gt_vcf via call_block")


same_profile = util.same_profile


PROFILE_RUNS = blockgen.PROFILE_RUNS


@pytest.mark.parametrize("run", range(len(PROFILE_RUNS)))
@pytest.mark.parametrize("trims", [((0, 0), (0, 0)), ((5, 3), (2, 4))])
def test_profile(reference, run, trims):
    """--report-file side channels of the path (non-CpG conversion profile, base / read tallies): the restatement
    against the reference's meth_profile() / process_template_vector() with stats switched on, block after block"""
    from oracle.bindings import Oracle, Reference
    lt, rt = trims
    rng = np.random.default_rng(700 + run)
    ref = blockgen.random_reference(rng, 4000, n_runs=2)
    o = Oracle(left_trim=lt, right_trim=rt, min_qual=25 if run == 1 else 20)
    r = Reference(left_trim=lt, right_trim=rt, min_qual=25 if run == 1 else 20)
    r.stats_enable(True); r.stats_reset()
    o.profile_enable(True); o.profile_reset()
    try:
        # last: a pile of reads at positions 1 and 2, where the FSM starts one code late (src/meth_profile.c:65)
        blocks = [(150 + 40 * b, 2650 + 40 * b, case) for b, case in enumerate(PROFILE_RUNS[run])]
        blocks.append((1, 3, dict(depth=4000, read_len=50, paired=run != 0, frag_mean=70, frag_sd=10, single_mate_frac=0.3 if run else 0.0)))
        for b, (start, end, case) in enumerate(blocks):
            T, B, M, y = blockgen.make_block(rng, ref, start, end, **case)
            x, pile_r, vcf_r, ref_r, nt_r, nb_r = r.process_block(T, B, M, ref, y)
            refw = blockgen.window_codes(ref, x, y + 1)
            xo, pile_o, vcf_o = o.process_block(T, B, M, refw, y)
            _same_records(vcf_o, vcf_r, "gt_vcf")
            pr, po = r.stats_read(), o.profile_read()
            same_profile(po, pr, "run %d block %d" % (run, b))
        assert pr["conv"].sum() > 1000 and pr["base_filter"][0] > 0
    finally:
        r.stats_enable(False); o.profile_enable(False)
        Reference()


def test_profile_growth_drops_top_entry(reference):
    """fresh profile, lone mates in either slot: whether the count of a slot-1 mate's first byte survives depends on
    which templates came before it (gt_vector_reserve clears from the old `used` upwards, gt/src/gt_vector.c:34-37)"""
    from oracle.bindings import Oracle, Reference
    rng = np.random.default_rng(4242)
    ref = blockgen.random_reference(rng, 3000)
    o, r = Oracle(), Reference()
    r.stats_enable(True); o.profile_enable(True)
    try:
        for b in range(40):
            r.stats_reset(); o.profile_reset()
            start = 20 + 60 * b
            T, B, M, y = blockgen.make_block(rng, ref, start, start + 40, depth=30, read_len=int(rng.integers(30, 60)), paired=True,
                                             single_mate_frac=1.0, clip_frac=0.3, conv=0.5)
            x = r.process_block(T, B, M, ref, y)[0]
            o.process_block(T, B, M, blockgen.window_codes(ref, x, y + 1), y)
            same_profile(o.profile_read(), r.stats_read(), "block %d" % b)
    finally:
        r.stats_enable(False); o.profile_enable(False)


def test_deep_block_uses_lgamma(reference):
    """500x single-end panel (config 4): strand tables with margins >= 256."""
    from oracle.bindings import Oracle
    rng = np.random.default_rng(5)
    ref = blockgen.random_reference(rng, 1500)
    T, B, M, y = blockgen.make_block(rng, ref, 100, 900, depth=500, read_len=150, paired=False, snp_rate=0.02)
    x, pile_r, vcf_r, ref_r, nt_r, nb_r = reference.process_block(T, B, M, ref, y)
    xo, pile_o, vcf_o = Oracle().process_block(T, B, M, ref_r, y)
    _same_records(pile_o, pile_r, "pileup")
    _same_records(vcf_o, vcf_r, "gt_vcf")
    het = np.isin(vcf_r["gtm"]["max_gt"], [1, 2, 3, 5, 6, 8]) & (vcf_r["skip"] == 0)
    assert het.sum() > 0 and pile_r["n"].max() >= 256


# ---- reader side: record decode (src/input_sam.c) and read_input (src/get_template_vector.c) ----------------------
KEPT_ONLY = ("bs_strand", "align_length", "reference_span", "read_off", "read_len", "mm_off", "mm_n")


def _reader_opts(seed):
    rng = np.random.default_rng(1000 + seed)
    return dict(mapq_thresh=int(rng.integers(0, 30)), max_template_len=int(rng.integers(300, 1200)),
                keep_unmatched=seed % 7 == 5, ignore_duplicates=bool(rng.random() < 0.3), keep_duplicates=seed % 5 == 3)


@pytest.mark.parametrize("seed", range(8))
def test_decode_records(oracle, reference, seed):
    """every field get_next_align_details() produces, the packed reads and the CIGAR events: identical"""
    from tests import bamgen
    bam, n, _, _ = bamgen.make_stream(seed)
    o = _reader_opts(seed)
    args = (bam, o["mapq_thresh"], o["max_template_len"], o["keep_unmatched"], o["ignore_duplicates"])
    r, rb, rm = reference.decode_records(*args)
    w, wb, wm = oracle.decode_records(*args)
    assert len(r) == len(w) == n
    kept = r["ret"] == 0
    assert kept.sum() > 0 and (~kept).sum() > 0
    for f in r.dtype.names:
        a, b = (r[f][kept], w[f][kept]) if f in KEPT_ONLY else (r[f], w[f])
        assert (a == b).all(), f
    assert rb.tobytes() == wb.tobytes() and rm.tobytes() == wm.tobytes()
    assert set(np.unique(r["bs_strand"][kept])) == {0, 1, 2}


@pytest.mark.parametrize("seed", range(4))
def test_decode_exotic_records(oracle, reference, seed):
    """= / X / H / N / P / B operators, arbitrary flag words, unknown tag types, qualities up to 254"""
    from tests import bamgen
    bam, n = bamgen.exotic_stream(seed)
    for ku in (False, True):
        r, rb, rm = reference.decode_records(bam, 10, 700, ku, seed % 2 == 0)
        w, wb, wm = oracle.decode_records(bam, 10, 700, ku, seed % 2 == 0)
        kept = r["ret"] == 0
        assert len(r) == n and kept.sum() > 50
        for f in r.dtype.names:
            a, b = (r[f][kept], w[f][kept]) if f in KEPT_ONLY else (r[f], w[f])
            assert (a == b).all(), f
        assert rb.tobytes() == wb.tobytes() and rm.tobytes() == wm.tobytes()


@pytest.mark.parametrize("seed", range(12))
def test_read_input(oracle, reference, seed):
    """blocks (contig, window, template ranges) and every template with its mates' bytes and events: identical;
    every fourth seed also runs the chain read_input -> process_template_vector -> call_genotypes_ML on both sides"""
    from tests import bamgen, util
    bam, n, tl, refs = bamgen.make_stream(seed)
    o = _reader_opts(seed)
    chain = seed % 4 == 0
    # with --report-file's counters live on both sides: read_input's own tallies always, the conversion profile and the
    # normalisation tallies when the chain runs
    reference.stats_enable(True); reference.stats_reset()
    oracle.profile_enable(True); oracle.profile_reset()
    try:
        rbk, rt, rb, rm, rv = reference.read_input(bam, tl, refs, run_chain=chain, **o)
        wbk, wt, wb, wm, wv = oracle.read_input(bam, tl, refs, run_chain=chain, **o)
        pr, po = reference.stats_read(), oracle.profile_read()
    finally:
        reference.stats_enable(False); oracle.profile_enable(False)
    same_profile(po, pr, "seed %d" % seed, recycled_vectors=True)
    assert not chain or po["filter_cts"][0] == wt["present"].sum()
    assert pr["filter_cts"].sum() > 0 and (not chain or pr["conv"].sum() > 0)
    assert len(rbk) == len(wbk) and len(rbk) > 0
    for f in ("tid", "x", "y", "first_template", "n_templates", "vcf_off"):
        assert (rbk[f] == wbk[f]).all(), f
    assert bamgen.template_keys(rt, rb, rm) == bamgen.template_keys(wt, wb, wm)
    if chain:
        assert len(rv) == len(wv) and (rv["skip"] == 0).sum() > 0
        util.assert_gt_meth_close(wv["gtm"], wv["skip"], rv["gtm"], rv["skip"], exact_doubles=True)


# ---- writer side: print_vcf_entry / flush_vcf_entries / _print_vcf_entry (src/print_vcf.c) ------------------------------
def _same_bcf(got, want, what):
    if got[0].tobytes() != want[0].tobytes() or got[1] != want[1]:
        a, b = util.split_bcf(got[0]), util.split_bcf(want[0])
        for k, (ra, rb) in enumerate(zip(a, b)):
            assert ra == rb, "%s: record %d differs\n%s\n%s" % (what, k, ra.hex(), rb.hex())
        raise AssertionError("%s: %d records vs %d" % (what, len(a), len(b)))


@pytest.mark.parametrize("case", range(len(CASES)))
def test_print_block_on_called_blocks(oracle, reference, case):
    """the BCF records the reference's writer emits for a block (its own compiled print_vcf.c behind a capturing
    bcf_write) against the restatement, byte for byte, on blocks called by the reference itself"""
    rng = np.random.default_rng(300 + case)
    ref = blockgen.random_reference(rng, 6000, n_runs=6)
    T, B, M, y = blockgen.make_block(rng, ref, 200, 4800, **CASES[case])
    x, pile_r, vcf_r, ref_r, nt_r, nb_r = reference.process_block(T, B, M, ref, y)
    refw = blockgen.window_codes(ref, x, y + 2)
    for allp in (False, True):
        for ctg_end in (0xffffffff, x + len(vcf_r) // 2):
            want = reference.print_block(vcf_r, refw, x, rid=case, ctg_end=ctg_end, all_positions=allp)
            got = oracle.print_block(vcf_r, refw, x, rid=case, ctg_end=ctg_end, all_positions=allp)
            _same_bcf(got, want, "case %d all_positions %r" % (case, allp))
    assert want[1] > 100


@pytest.mark.parametrize("seed", range(6))
def test_print_block_on_random_records(oracle, reference, seed):
    """records built to reach every branch of the writer (any call on any reference base incl. N, two ALT alleles, deep
    counts, all filter combinations, certain and hopeless posteriors), N runs in the reference window, blocks of every
    small size, header ids that need one, two and four bytes"""
    rng = np.random.default_rng(900 + seed)
    ids = [list(range(16)), [3, 200, 5, 40000, 7, 100000, 9, 11, 127, 128, 13, 15, 17, 19, 21, 23]][seed % 2]
    sizes = list(range(1, 9)) + [40, 333, 2000]
    for sz in sizes:
        vcf = util.random_gt_vcf(rng, sz, skip_frac=[0.0, 0.25, 0.6][seed % 3])
        refw = rng.integers(1, 5, size=sz + 2).astype(np.uint8)
        refw[rng.random(sz + 2) < [0.0, 0.03, 0.3][(seed // 2) % 3]] = 0
        x = int(rng.integers(1, 1000))
        ctg_end = x + sz - 1 - int(rng.integers(0, 3))
        for allp in (False, True):
            want = reference.print_block(vcf, refw, x, rid=3, ctg_end=ctg_end, vcf_ids=ids, all_positions=allp)
            got = oracle.print_block(vcf, refw, x, rid=3, ctg_end=ctg_end, vcf_ids=ids, all_positions=allp)
            _same_bcf(got, want, "seed %d size %d all_positions %r" % (seed, sz, allp))


def random_dbsnp(rng, lo, hi, frac=0.15):
    """dbSNP answers for a random subset of [lo, hi): known / always-written flags, IDs of 3..14 bytes, some with the trailing
    NUL an odd number of digits leaves behind (src/dbSNP.c:337-341)"""
    from oracle.bindings import dbsnp_arrays
    n = max(1, int((hi - lo) * frac))
    ents = []
    for p in rng.choice(np.arange(lo, hi), size=min(n, hi - lo), replace=False):
        nm = b"rs%d" % int(rng.integers(1, 10 ** int(rng.integers(1, 11))))
        if rng.random() < 0.3:
            nm += b"\0"
        ents.append((int(p), 3 if rng.random() < 0.4 else 1, nm))
    return dbsnp_arrays(ents)


@pytest.mark.parametrize("seed", range(4))
def test_print_block_with_dbsnp_and_regions(oracle, reference, seed):
    """-D and -C as the writer sees them (src/print_vcf.c:133, 139, 154-157, 163-167): IDs, "always written" homozygous
    reference A / T sites, clipping to ctg->curr_reg instead of the contig end -- restatement against the compiled writer fed
    through its own dbSNP_lookup_name() interface"""
    rng = np.random.default_rng(1900 + seed)
    for sz in (1, 7, 300, 5000):
        vcf = util.random_gt_vcf(rng, sz, skip_frac=[0.0, 0.3][seed % 2])
        refw = rng.integers(1, 5, size=sz + 2).astype(np.uint8)
        refw[rng.random(sz + 2) < 0.02] = 0
        x = int(rng.integers(1, 1000))
        ctg_end = x + sz - 1 - int(rng.integers(0, 3))
        db = random_dbsnp(rng, max(1, x - 5), x + sz + 5)
        regions = [None, (x + sz // 4, x + (3 * sz) // 4), (1, x + sz // 2), (x + 2, x + sz + 100)]
        for region in regions:
            for allp in (False, True):
                for d in (None, db):
                    kw = dict(rid=2, ctg_end=ctg_end, all_positions=allp, region=region, dbsnp=d)
                    want = reference.print_block(vcf, refw, x, **kw)
                    got = oracle.print_block(vcf, refw, x, **kw)
                    _same_bcf(got, want, "seed %d size %d %r" % (seed, sz, (region, allp, d is not None)))


def cpg_rich(rng, vcf, frac=0.2):
    """make a share of adjacent site pairs a called CpG (CC then GG) with informative counts, so the CpG branches of the
    writer's statistics see traffic"""
    g = vcf["gtm"]
    n = len(vcf)
    for i in np.flatnonzero(rng.random(max(n - 1, 0)) < frac):
        for j, gt in ((i, 4), (i + 1, 7)):
            lp = g["gt_prob"][j].copy()
            b = int(np.argmax(lp))
            lp[b], lp[gt] = lp[gt], lp[b]
            if lp[gt] < lp.max():
                lp[gt] = lp.max() + 1e-3
            g["gt_prob"][j] = np.minimum(lp, 0.0)
            g["max_gt"][j] = gt
            if rng.random() < 0.9:
                g["counts"][j, 4:8] = rng.integers(0, [3, 40, 300][int(rng.integers(0, 3))], size=4)
                vcf["skip"][j] = 0 if g["counts"][j].sum() else 1
    return vcf


@pytest.mark.parametrize("seed", range(4))
def test_writer_statistics(oracle, reference, seed):
    """the --report-file statistics of the writer (src/print_vcf.c:382-526): the compiled print_vcf.c with a live bs_stats
    against the restatement, over random records rich in CpGs, with dbSNP, regions, -A and GC bins; counters identical, the
    methylation posteriors (sums of doubles in the same order) identical too"""
    from oracle.bindings import SITE_STATS, site_stats_equal
    rng = np.random.default_rng(2900 + seed)
    reference.writer_stats_reset()
    st = np.zeros(1, dtype=SITE_STATS)
    state = np.zeros(2, dtype=np.uint32)
    x = 1
    try:
        for sz in (1, 2, 5, 60, 700, 4000):
            vcf = cpg_rich(rng, util.random_gt_vcf(rng, sz, skip_frac=[0.0, 0.2][seed % 2], deep_frac=0.02))
            refw = rng.integers(1, 5, size=sz + 2).astype(np.uint8)
            refw[rng.random(sz + 2) < 0.02] = 0
            for i in np.flatnonzero(rng.random(sz) < 0.3):      # reference CpGs
                refw[i], refw[i + 1] = 2, 3
            x += int(rng.integers(1, 500))
            start_pos = int(rng.integers(1, x + 1))
            gc = rng.integers(0, 120, size=(x + sz - start_pos) // 100 + int(rng.integers(0, 2))).astype(np.uint8)
            ctg_end = x + sz - 1 - int(rng.integers(0, 3))
            db = random_dbsnp(rng, max(1, x - 5), x + sz + 5)
            for region in (None, (x + sz // 4, x + (3 * sz) // 4)):
                for allp in (False, True):
                    for d in (None, db):
                        reference.writer_stats(True, gc=gc if len(gc) else None, start_pos=start_pos)
                        kw = dict(ctg_end=ctg_end, all_positions=allp, region=region, dbsnp=d)
                        reference.print_block(vcf, refw, x, rid=1, **kw)
                        oracle.stats_block(vcf, refw, x, gc=gc if len(gc) else None, start_pos=start_pos, stats=st, state=state, **kw)
            x += sz
        want, _ = reference.writer_stats_read()
    finally:
        reference.writer_stats(False)
    site_stats_equal(st[0], want[0], rtol=1e-13, what="seed %d" % seed)
    assert want[0]["snps"][0] > 1000 and want[0]["CpG_ref"][0] + want[0]["CpG_nonref"][0] > 50 and want[0]["multi"][0] > 50
    assert want[0]["CpG_ref_meth"].sum() > 10 and want[0]["cov"]["gc_pcent"].sum() > 1000 and want[0]["dbSNP_sites"][0] > 20


@pytest.mark.parametrize("seed", [5007, 5080])
def test_lone_mate_before_its_block_aborts_the_reference(oracle, seed):
    """-k with -d: a lone mate is kept with the position its absent partner claimed (src/get_template_vector.c:247-268), which
    may lie before the block; read_input hands the block on, call_genotypes_ML asserts `x1 >= x` on it (src/call_genotypes.c:186;
    the reference is built with its asserts) and the process dies.  Found by tests/fuzz_cpu.py.  The restatement's read_input
    succeeds on the same stream with the same blocks as the product's host builder, and its chain refuses it; the compiled
    reference is run in a child process, because it takes the process down."""
    import subprocess
    import sys
    from bs_call_b200 import lib
    from tests import bamgen
    from tests.test_cpu_reader import descriptors_from_oracle
    rng = np.random.default_rng(seed)
    bam, n, tl, refs = bamgen.make_stream(seed, dup=float(rng.choice([0.0, 0.1, 0.3])), junk=float(rng.choice([0.0, 0.1, 0.3])))
    o = dict(mapq_thresh=int(rng.integers(0, 40)), max_template_len=int(rng.integers(200, 1500)), keep_unmatched=bool(rng.random() < 0.3),
             ignore_duplicates=bool(rng.random() < 0.3), keep_duplicates=bool(rng.random() < 0.3))
    assert o["keep_unmatched"] and o["keep_duplicates"]
    wbk = oracle.read_input(bam, tl, refs, **o)[0]
    with pytest.raises(RuntimeError):
        oracle.read_input(bam, tl, refs, run_chain=True, **o)
    orec, ob, om = oracle.decode_records(bam, o["mapq_thresh"], o["max_template_len"], o["keep_unmatched"], o["ignore_duplicates"])
    hb, ht = lib.build_blocks(bam, descriptors_from_oracle(orec, ob, bam), lib.reader_params(keep_unmatched=True, keep_duplicates=True))
    assert len(hb) == len(wbk) and all((hb[f] == wbk[f]).all() for f in ("tid", "x", "y", "n_templates"))
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from oracle.bindings import Reference; from tests import bamgen; "
            "rng = np.random.default_rng(%d); "
            "bam, n, tl, refs = bamgen.make_stream(%d, dup=float(rng.choice([0.0, 0.1, 0.3])), junk=float(rng.choice([0.0, 0.1, 0.3]))); "
            "Reference(calc_threads=1).read_input(bam, tl, refs, run_chain=True, **%r); print('survived')"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), seed, seed, o))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "survived" not in r.stdout and "x1 >= x" in r.stderr, (r.returncode, r.stderr[-300:])


def test_block_at_the_very_start_of_a_contig_is_where_the_restatement_leaves_the_reference(oracle, reference):
    """Known, deliberate deviation (found by running test_print_block_on_random_records over several hundred more seeds): the
    reference's writer keeps its five-site window in file statics (gt_store, store_x; src/print_vcf.c:529-533) and clears it at a
    block start only when the block begins at position 5 or later (`l = x - store_x`, :563-570).  A block that begins at
    position 2 therefore sees a call of whatever block was printed BEFORE it -- another contig -- in the context fields of its
    first record (from position 3 on the stale calls have left the window before a site is printed), and one that begins at
    position 1 makes the writer print the stale site `x - 2` = 2^32 - 1, after which `x <= old_x` (:120-121) discards every site
    of the contig.  Only contigs whose first read begins within their first four bases (x = first read - 2) are affected.  The restatement -- and the device writer pinned to it -- starts every block from a clean window: it writes the
    contig, with the context a first block has."""
    rng = np.random.default_rng(2010)
    stale = util.random_gt_vcf(rng, 50, skip_frac=0.0)
    refs = rng.integers(1, 5, size=52).astype(np.uint8)
    vcf = util.random_gt_vcf(rng, 60, skip_frac=0.0)
    refw = rng.integers(1, 5, size=62).astype(np.uint8)
    for x, what in ((1, "nothing"), (2, "context"), (3, "same"), (5, "same")):
        reference.print_block(stale, refs, 700, ctg_end=10000)                 # the block "before": leaves its last calls behind
        want = reference.print_block(vcf, refw, x, ctg_end=10000)
        got = oracle.print_block(vcf, refw, x, ctg_end=10000)
        assert got[1] > 20
        if what == "nothing":
            assert want[1] == 0
        elif what == "context":
            a, b = util.split_bcf(got[0]), util.split_bcf(want[0])
            assert len(a) == len(b) and a[0] != b[0] and a[1:] == b[1:]        # only the first record's context fields differ
        else:
            assert got[0].tobytes() == want[0].tobytes()
