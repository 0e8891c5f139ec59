"""CPU checks of the reader side that need no GPU: the product's host block builder (bsgpu_build_blocks, a pure host
function of the ABI) against the oracle's restatement of read_input(), and the committed reader golden (captured
from the reference's own read_input / process_template_vector / call_genotypes_ML) against the oracle."""
import numpy as np
import pytest

from bs_call_b200 import lib
from bs_call_b200.records import RECORD
from tests import bamgen, util


def descriptors_from_oracle(orec, obases, bam):
    """the oracle's per-record view widened to the product's descriptor (adds the contig id and the two qualities the
    duplicate tie-break reads)"""
    r = np.zeros(len(orec), dtype=RECORD)
    for f in orec.dtype.names:
        r[f] = orec[f]
    at = 0
    for i in range(len(orec)):
        bs = int(bam[at:at + 4].view("<i4")[0])
        r["tid"][i] = int(bam[at + 4:at + 8].view("<i4")[0])
        at += 4 + bs
        if orec["ret"][i] == 0:
            for k in range(min(2, int(orec["read_len"][i]))):
                r["q01"][i, k] = obases[int(orec["read_off"][i]) + k] >> 2
    return r


@pytest.mark.parametrize("seed", range(10))
def test_build_blocks_matches_oracle(oracle, seed):
    bam, n, tl, _ = bamgen.make_stream(seed, dup=0.2)
    kd, ku = seed % 5 == 3, seed % 7 == 5
    orec, ob, om = oracle.decode_records(bam, 20, 1000, ku, False)
    rec = descriptors_from_oracle(orec, ob, bam)
    blocks, tm = lib.build_blocks(bam, rec, lib.reader_params(keep_unmatched=ku, keep_duplicates=kd))
    wbk, wt, wb, wm, _ = oracle.read_input(bam, tl, None, keep_unmatched=ku, keep_duplicates=kd)
    assert len(blocks) == len(wbk) > 0
    for f in ("tid", "x", "y", "first_template", "n_templates"):
        assert (blocks[f] == wbk[f]).all(), f
    assert bamgen.template_keys(tm, ob, om) == bamgen.template_keys(wt, wb, wm)


def test_build_blocks_rejects_bad_stream():
    bam, n, _, _ = bamgen.make_stream(1)
    rec = np.zeros(n, dtype=RECORD)
    with pytest.raises(lib.BsGpuError):
        lib.build_blocks(bam[:-7], rec)


def test_empty_stream(oracle):
    blocks, tm = lib.build_blocks(np.zeros(0, dtype=np.uint8), np.zeros(0, dtype=RECORD))
    assert len(blocks) == 0 and len(tm) == 0
    r, b, m = oracle.decode_records(np.zeros(0, dtype=np.uint8))
    assert len(r) == 0


@pytest.mark.parametrize("name", ["reader_pe", "reader_mixed"])
def test_oracle_matches_reader_golden(oracle, name):
    g = util.load_golden(name)
    o = dict(mapq_thresh=int(g["mapq_thresh"]), max_template_len=int(g["max_template_len"]), keep_unmatched=bool(g["keep_unmatched"]),
             ignore_duplicates=bool(g["ignore_duplicates"]), keep_duplicates=bool(g["keep_duplicates"]))
    refs = [g["ref%d" % i] for i in range(len(g["target_len"]))]
    rec, rb, rm = oracle.decode_records(g["bam"], o["mapq_thresh"], o["max_template_len"], o["keep_unmatched"], o["ignore_duplicates"])
    kept = g["rec"]["ret"] == 0
    for f in g["rec"].dtype.names:
        a, b = (g["rec"][f][kept], rec[f][kept]) if f in ("bs_strand", "align_length", "reference_span", "read_off", "read_len", "mm_off", "mm_n") else (g["rec"][f], rec[f])
        assert (a == b).all(), f
    assert rb.tobytes() == g["rec_bases"].tobytes()
    wbk, wt, wb, wm, wv = oracle.read_input(g["bam"], g["target_len"], refs, run_chain=True, **o)
    for f in ("tid", "x", "y", "first_template", "n_templates", "vcf_off"):
        assert (wbk[f] == g["blocks"][f]).all(), f
    assert bamgen.template_keys(wt, wb, wm) == bamgen.template_keys(g["templates"], g["bases"], g["misms"])
    util.assert_gt_meth_close(wv["gtm"], wv["skip"], g["vcf"]["gtm"], g["vcf"]["skip"], exact_doubles=True)


BLOCKS = ["block_pe_plain", "block_pe_indel_clip_trim", "block_se_deep", "block_mixed"]


@pytest.mark.parametrize("name", BLOCKS + ["reader_pe", "reader_mixed"])
def test_oracle_matches_profile_golden(name):
    """--report-file side channels: the restatement against what the reference's meth_profile() / process_template_vector()
    / read_input() left in bs_stats for the committed goldens"""
    from oracle.bindings import Oracle
    g = util.load_golden(name)
    want, refw = util.golden_profile(name)
    if name in BLOCKS:
        o = Oracle(left_trim=tuple(int(v) for v in g["left_trim"]), right_trim=tuple(int(v) for v in g["right_trim"]))
    else:
        o = Oracle()
    o.profile_enable(True); o.profile_reset()
    try:
        if name in BLOCKS:
            o.process_block(g["templates"], g["bases"], g["misms"], refw, int(g["y"]))
        else:
            opts = dict(mapq_thresh=int(g["mapq_thresh"]), max_template_len=int(g["max_template_len"]), keep_unmatched=bool(g["keep_unmatched"]),
                        ignore_duplicates=bool(g["ignore_duplicates"]), keep_duplicates=bool(g["keep_duplicates"]))
            o.read_input(g["bam"], g["target_len"], [g["ref%d" % i] for i in range(len(g["target_len"]))], run_chain=True, **opts)
        got = o.profile_read()
    finally:
        o.profile_enable(False)
    util.same_profile(got, want, name, recycled_vectors=name not in BLOCKS)
    assert want["conv"].sum() > 5000


def test_parallel_builder_equals_sequential(oracle, monkeypatch):
    """the stream is cut at records where read_input is certain to start a new block and the pieces are built on
    separate host threads: same blocks, same templates as one thread"""
    bam, n, tl, _ = bamgen.make_stream(77, n_contigs=3, dup=0.2, contig_len=9000)
    orec, ob, om = oracle.decode_records(bam, 20, 1000, False, False)
    rec = descriptors_from_oracle(orec, ob, bam)
    monkeypatch.setenv("BSGPU_BUILDER_THREADS", "1")
    b1, t1 = lib.build_blocks(bam, rec)
    monkeypatch.setenv("BSGPU_BUILDER_THREADS", "7")
    monkeypatch.setenv("BSGPU_BUILDER_MIN_RECORDS", "1")
    b7, t7 = lib.build_blocks(bam, rec)
    assert len(b1) > 3
    assert b1.tobytes() == b7.tobytes() and t1.tobytes() == t7.tobytes()
    wbk, wt, wb, wm, _ = oracle.read_input(bam, tl, None)
    assert (b7["y"] == wbk["y"]).all() and bamgen.template_keys(t7, ob, om) == bamgen.template_keys(wt, wb, wm)


@pytest.mark.parametrize("blind", [False, True])
def test_parallel_framer_equals_sequential(oracle, monkeypatch, blind):
    """several host threads walk the block_size chain from guessed record boundaries; the stitch only takes a piece over
    from an offset the true chain lands on.  With BSGPU_FRAMER_BLIND the guesses are raw byte offsets (wrong almost
    every time) and the result must still be the sequential chain."""
    bam, n, tl, _ = bamgen.make_stream(78, n_contigs=2, contig_len=8000)
    orec, ob, om = oracle.decode_records(bam, 20, 1000, False, False)
    rec = descriptors_from_oracle(orec, ob, bam)
    monkeypatch.setenv("BSGPU_FRAMER_THREADS", "1")
    b1, t1 = lib.build_blocks(bam, rec)
    monkeypatch.setenv("BSGPU_FRAMER_THREADS", "6")
    monkeypatch.setenv("BSGPU_FRAMER_MIN_BYTES", "1")
    if blind:
        monkeypatch.setenv("BSGPU_FRAMER_BLIND", "1")
    b6, t6 = lib.build_blocks(bam, rec)
    assert b1.tobytes() == b6.tobytes() and t1.tobytes() == t6.tobytes()
    # a truncated stream is rejected by the parallel framer too
    with pytest.raises(lib.BsGpuError):
        lib.build_blocks(bam[:-9], rec)
    with pytest.raises(lib.BsGpuError):
        lib.build_blocks(bam, rec[:-1])


@pytest.mark.parametrize("seed,code", [(2002, "positions"), (2017, "duplicate read name")])
def test_streams_the_reference_aborts_on_are_errors(oracle, seed, code):
    """read_input() asserts that a mate agrees with its waiting partner about both positions
    (src/get_template_vector.c:239) and treats a read name that is already waiting as fatal (:327-328): the oracle and
    the product's builder return an error on exactly those streams (found by fuzzing against the compiled reference,
    which aborts the process there)"""
    rng = np.random.default_rng(seed)
    bam, n, tl, refs = bamgen.make_stream(seed, dup=float(rng.choice([0.0, 0.1, 0.3])), junk=float(rng.choice([0.0, 0.1, 0.3])))
    o = dict(mapq_thresh=int(rng.integers(0, 40)), max_template_len=int(rng.integers(200, 1500)), keep_unmatched=bool(rng.random() < 0.3),
             ignore_duplicates=bool(rng.random() < 0.3), keep_duplicates=bool(rng.random() < 0.3))
    with pytest.raises(RuntimeError):
        oracle.read_input(bam, tl, refs, **o)
    orec, ob, om = oracle.decode_records(bam, o["mapq_thresh"], o["max_template_len"], o["keep_unmatched"], o["ignore_duplicates"])
    rec = descriptors_from_oracle(orec, ob, bam)
    with pytest.raises(lib.BsGpuError, match=code):
        lib.build_blocks(bam, rec, lib.reader_params(**o))


def test_three_phase_certain_start_scan_equals_the_sequential_one():
    """The device marks certain block starts (bs_call_b200/csrc/bsgpu_reader.cu: k_certain_tile / k_certain_combine) tile by
    tile: phase 1 summarises every tile as if the stream continued into it (first / last kept contig, "a contig change inside",
    running end since that change), phase 2 chains the summaries from the carried state, phase 3 replays every tile from its
    incoming state.  Emulated here in Python over random key streams with tiny tiles (every boundary case: empty tiles, tiles
    of dropped records only, contig changes on tile edges) against the sequential scan (certain_scan_seq)."""
    NO = 0xffffffff
    rng = np.random.default_rng(4)

    def seq(keys, st):
        tid, m = st
        out = []
        for i, (t, mn, e) in enumerate(keys):
            if t == NO:
                continue
            if t != tid:
                tid, m = t, 0
                out.append(i)
            elif mn and mn > m + 1:
                out.append(i)
            m = max(m, e)
        return out, (tid, m)

    def three(keys, st, T):
        nt = (len(keys) + T - 1) // T
        agg = []
        for t in range(nt):
            tile = [k for k in keys[t * T:(t + 1) * T] if k[0] != NO]
            if not tile:
                agg.append((NO, NO, 0, 0))
                continue
            pt, m, flag = tile[0][0], 0, 0
            for (td, mn, e) in tile:
                if td != pt:
                    flag, m, pt = 1, 0, td
                m = max(m, e)
            agg.append((tile[0][0], tile[-1][0], flag, m))
        tid, m = st
        tin = []
        for (first, last, flag, am) in agg:
            tin.append((tid, m))
            if first == NO:
                continue
            f = flag or first != tid
            m = am if f else max(m, am)
            tid = last
        out = []
        for t in range(nt):
            o, _ = seq(keys[t * T:(t + 1) * T], tin[t])
            out += [t * T + i for i in o]
        return out, (tid, m)

    for trial in range(300):
        n = int(rng.integers(0, 120))
        T = int(rng.choice([1, 2, 3, 4, 8, 16]))
        keys, tid, pos = [], int(rng.integers(0, 3)), 10
        for _ in range(n):
            r = rng.random()
            if r < 0.25:
                keys.append((NO, 0, 0))
                continue
            if r < 0.33:
                tid += 1
                pos = int(rng.integers(1, 50))
            pos += int(rng.integers(0, 40))
            keys.append((tid, pos if rng.random() < 0.7 else 0, pos + int(rng.integers(1, 60))))
        st = (int(rng.choice([-1 & NO, 0, 1, tid])), int(rng.integers(0, 200)))
        # a stream in two chunks, state carried
        cut = int(rng.integers(0, n + 1))
        a1, s1 = seq(keys[:cut], st)
        a2, s2 = seq(keys[cut:], s1)
        b1, t1 = three(keys[:cut], st, T)
        b2, t2 = three(keys[cut:], t1, T)
        assert a1 == b1 and a2 == b2 and s1 == t1 and s2 == t2, (trial, T, keys, st)
