"""--report-file side channels on the device (-m gpu): the non-CpG conversion profile of meth_profile()
(src/meth_profile.c:48-76) and the base / read tallies of process_template_vector (src/process_template.c:52-63), gathered
by the normalisation kernel while it rewrites the reads, plus read_input's per-reason tallies from the host block builder.
Checked against the goldens the reference left in bs_stats (tests/golden/profile_v1.npz) and against the oracle on seeded
blocks and streams.  All integer counts: bit-exact."""
import numpy as np
import pytest

from bs_call_b200 import lib as bslib
from tests import bamgen, blockgen, util

pytestmark = pytest.mark.gpu

BLOCKS = ["block_pe_plain", "block_pe_indel_clip_trim", "block_se_deep", "block_mixed"]


def _reader_opts(g):
    return dict(mapq_thresh=int(g["mapq_thresh"]), max_template_len=int(g["max_template_len"]), keep_unmatched=bool(g["keep_unmatched"]),
                ignore_duplicates=bool(g["ignore_duplicates"]), keep_duplicates=bool(g["keep_duplicates"]))


@pytest.mark.parametrize("name", BLOCKS)
def test_block_goldens(name):
    g = util.load_golden(name)
    want, refw = util.golden_profile(name)
    gpu = bslib.BsGpu(left_trim=tuple(int(v) for v in g["left_trim"]), right_trim=tuple(int(v) for v in g["right_trim"]))
    try:
        gpu.profile_enable(True)
        before = gpu.stats()["kernel_launches"]
        x, vcf = gpu.process_block(g["templates"], g["bases"], g["misms"], refw, int(g["y"]))
        assert gpu.stats()["kernel_launches"] >= before + 3          # normalise + resolve + the block kernels
        util.assert_vcf_close(vcf, g["vcf"])                          # the calls are what they are without the profile
        util.same_profile(gpu.profile_read(), want, name)
        # reading does not clear; reset does
        util.same_profile(gpu.profile_read(reset=True), want, name)
        z = gpu.profile_read()
        assert z["used"] == 0 and z["filter_cts"].sum() == 0 and z["base_filter"].sum() == 0
    finally:
        gpu.close()


@pytest.mark.parametrize("name", ["reader_pe", "reader_mixed"])
def test_reader_goldens(name):
    g = util.load_golden(name)
    want, _ = util.golden_profile(name)
    o = _reader_opts(g)
    gpu = bslib.BsGpu()
    try:
        gpu.profile_enable(True)
        refs = [g["ref%d" % i] for i in range(len(g["target_len"]))]
        blocks, vcf = gpu.call_bam(g["bam"], g["target_len"], refs, bslib.reader_params(**o))
        assert len(blocks) == len(g["blocks"])          # (the calls themselves are checked in tests/test_gpu_reader.py)
        got = gpu.profile_read()
        util.same_profile(got, want, name, recycled_vectors=True)
        assert got["filter_cts"][0] == (g["templates"]["read_len"] > 0).sum()      # mates that exist (the snapshot also marks recycled empty vectors)
    finally:
        gpu.close()


def test_runs_of_blocks_match_oracle():
    """several blocks in a row through one context (the profile vector grows with the longest read seen so far, so the
    order matters), with and without -L/-R trimming and a non-default -Q"""
    from oracle.bindings import Oracle
    for run, cases in enumerate(blockgen.PROFILE_RUNS):
        for lt, rt in (((0, 0), (0, 0)), ((5, 3), (2, 4))):
            rng = np.random.default_rng(700 + run)
            ref = blockgen.random_reference(rng, 4000, n_runs=2)
            mq = 25 if run == 1 else 20
            o = Oracle(left_trim=lt, right_trim=rt, min_qual=mq)
            gpu = bslib.BsGpu(left_trim=lt, right_trim=rt, min_qual=mq)
            o.profile_enable(True); o.profile_reset()
            try:
                gpu.profile_enable(True)
                blocks = [(150 + 40 * b, 2650 + 40 * b, case) for b, case in enumerate(cases)]
                blocks.append((1, 3, dict(depth=4000, read_len=50, paired=run != 0, frag_mean=70, frag_sd=10, single_mate_frac=0.3 if run else 0.0)))
                for b, (start, end, case) in enumerate(blocks):
                    T, B, M, y = blockgen.make_block(rng, ref, start, end, **case)
                    first = int(T[0]["forward_position"]) or int(T[0]["reverse_position"])
                    x = first - 2 if first > 2 else 1
                    refw = blockgen.window_codes(ref, x, y + 1)
                    xo, pile, want_vcf = o.process_block(T, B, M, refw, y)
                    xg, vcf = gpu.process_block(T, B, M, refw, y)
                    assert xo == xg == x
                    util.assert_vcf_close(vcf, want_vcf)
                    util.same_profile(gpu.profile_read(), o.profile_read(), "run %d block %d trims %r" % (run, b, (lt, rt)))
            finally:
                o.profile_enable(False)
                gpu.close()


def test_growth_drops_top_entry(oracle):
    """fresh profile, lone mates in either slot: a count on the entry that is the profile's `used` at that moment is lost
    in the reference (gt_vector_reserve clears from the old `used` upwards); k_profile_resolve reproduces which"""
    rng = np.random.default_rng(4242)
    ref = blockgen.random_reference(rng, 3000)
    gpu = bslib.BsGpu()
    oracle.profile_enable(True)
    try:
        gpu.profile_enable(True)
        for b in range(40):
            oracle.profile_reset()
            start = 20 + 60 * b
            T, B, M, y = blockgen.make_block(rng, ref, start, start + 40, depth=30, read_len=int(rng.integers(30, 60)), paired=True,
                                             single_mate_frac=1.0, clip_frac=0.3, conv=0.5)
            first = int(T[0]["forward_position"]) or int(T[0]["reverse_position"])
            x = first - 2 if first > 2 else 1
            refw = blockgen.window_codes(ref, x, y + 1)
            oracle.process_block(T, B, M, refw, y)
            gpu.process_block(T, B, M, refw, y)
            util.same_profile(gpu.profile_read(reset=True), oracle.profile_read(), "block %d" % b)
    finally:
        oracle.profile_enable(False)
        gpu.close()


@pytest.mark.parametrize("seed", [3, 8, 21])
def test_streams_match_oracle(oracle, monkeypatch, seed):
    """raw BAM records -> profile through bsgpu_call_bam (windows that span several blocks, chunked upload, parallel
    builder) against the oracle's chain read_input -> process_template_vector"""
    bam, n, tl, refs = bamgen.make_stream(seed, dup=0.2, junk=0.15)
    opts = dict(mapq_thresh=20, max_template_len=1000, keep_unmatched=seed % 2 == 1, ignore_duplicates=False, keep_duplicates=seed == 21)
    oracle.profile_enable(True); oracle.profile_reset()
    try:
        wbk, wt, wb, wm, wv = oracle.read_input(bam, tl, refs, run_chain=True, **opts)
        want = oracle.profile_read()
    finally:
        oracle.profile_enable(False)
    monkeypatch.setenv("BSGPU_READER_CHUNK_MIN_BYTES", "1")
    monkeypatch.setenv("BSGPU_BUILDER_MIN_RECORDS", "1")
    monkeypatch.setenv("BSGPU_BUILDER_THREADS", "4")
    gpu = bslib.BsGpu()
    try:
        gpu.profile_enable(True)
        blocks, vcf = gpu.call_bam(bam, tl, refs, bslib.reader_params(**opts))
        assert len(blocks) == len(wbk)
        util.same_profile(gpu.profile_read(), want, "seed %d" % seed)
        # a second pass over the same stream doubles every count and leaves `used` alone
        gpu.call_bam(bam, tl, refs, bslib.reader_params(**opts))
        twice = gpu.profile_read()
        assert twice["used"] == want["used"]
        np.testing.assert_array_equal(twice["filter_bases"], 2 * want["filter_bases"])
        assert (twice["conv"] >= 2 * want["conv"]).all()          # entries the first pass lost to growth are kept the second time
    finally:
        gpu.close()


def test_many_templates_in_one_window(oracle):
    """more templates than one scan chunk of k_profile_resolve (4096) in a single window, lone reverse mates first, so the
    running maximum has to be carried across chunks"""
    rng = np.random.default_rng(99)
    ref = blockgen.random_reference(rng, 30000)
    T, B, M, y = blockgen.make_block(rng, ref, 100, 29000, depth=90, read_len=90, paired=True, single_mate_frac=0.6, clip_frac=0.2, indel_frac=0.1)
    assert len(T) > 3 * 4096
    first = int(T[0]["forward_position"]) or int(T[0]["reverse_position"])
    x = first - 2 if first > 2 else 1
    refw = blockgen.window_codes(ref, x, y + 1)
    oracle.profile_enable(True); oracle.profile_reset()
    gpu = bslib.BsGpu()
    try:
        oracle.process_block(T, B, M, refw, y)
        gpu.profile_enable(True)
        gpu.process_block(T, B, M, refw, y)
        util.same_profile(gpu.profile_read(), oracle.profile_read(), "big window")
    finally:
        oracle.profile_enable(False)
        gpu.close()


def test_profile_is_off_by_default():
    g = util.load_golden("block_pe_plain")
    gpu = bslib.BsGpu()
    try:
        with pytest.raises(bslib.BsGpuError):
            gpu.profile_read()
        before = gpu.stats()["kernel_launches"]
        gpu.process_block(g["templates"], g["bases"], g["misms"], g["ref"], int(g["y"]))
        plain = gpu.stats()["kernel_launches"] - before
        gpu.profile_enable(True)
        want, refw = util.golden_profile("block_pe_plain")
        before = gpu.stats()["kernel_launches"]
        gpu.process_block(g["templates"], g["bases"], g["misms"], refw, int(g["y"]))
        assert gpu.stats()["kernel_launches"] - before == plain + 1      # the profile costs one more launch (resolve)
        gpu.profile_enable(False)
        before = gpu.stats()["kernel_launches"]
        gpu.process_block(g["templates"], g["bases"], g["misms"], g["ref"], int(g["y"]))
        assert gpu.stats()["kernel_launches"] - before == plain
        util.same_profile(gpu.profile_read(), want, "after switching off")
    finally:
        gpu.close()
