"""The wide seams (-m gpu), tested like the narrow one (tests/test_gpu_dropin.py): the reference's own unmodified objects
linked with the product's replacement for a part of the chain, driven by the reference-side harness, against what the
all-CPU reference produced.

  seam B  bs_call_b200/csrc/bsgpu_seam_template.c replaces src/process_template.c + src/call_genotypes.c:
          process_template_vector (trim, soft clips, mate overlap, indel normalisation, pileup, model) on the device,
          read_input and the print-thread protocol still the reference's  -> oracle/_ref/libbsref_seamB.so
  seams C / D: see the second half of this file
"""
import numpy as np
import pytest

from tests import bamgen, blockgen, util

pytestmark = pytest.mark.gpu

BLOCKS = ["block_pe_plain", "block_pe_indel_clip_trim", "block_se_deep", "block_mixed"]


@pytest.fixture(scope="module")
def seam_b():
    from oracle.bindings import ReferenceWithSeamB, seam_available
    if not seam_available("seamB"):
        pytest.skip("oracle/_ref/libbsref_seamB.so not built (reference tree absent at build time)")
    return ReferenceWithSeamB(calc_threads=1)


@pytest.mark.parametrize("name", BLOCKS)
def test_seam_b_blocks_match_reference_goldens(seam_b, name):
    from oracle.bindings import ReferenceWithSeamB
    g = util.load_golden(name)
    d = ReferenceWithSeamB(left_trim=tuple(int(v) for v in g["left_trim"]), right_trim=tuple(int(v) for v in g["right_trim"]))
    x, y = int(g["x"]), int(g["y"])
    ctg = np.zeros(y + 16, dtype=np.uint8)
    ctg[x - 1:x - 1 + len(g["ref"])] = g["ref"]
    xo, pile, vcf, ref, nt, nb = d.process_block(g["templates"], g["bases"], g["misms"], ctg, y)
    assert xo == x
    util.assert_vcf_close(vcf, g["vcf"])
    ReferenceWithSeamB()


def test_seam_b_many_blocks_back_to_back(seam_b, oracle):
    """successive blocks of different sizes through the double-buffered hand-off (both arrays grow, ref / ref1 swap)"""
    rng = np.random.default_rng(19)
    for i, span in enumerate((3000, 40000, 800, 15000, 60000, 500)):
        ref = blockgen.random_reference(rng, span + 3000, n_runs=1)
        T, B, M, y = blockgen.make_block(rng, ref, 200, 200 + span, depth=20, read_len=100, paired=True, indel_frac=0.1, clip_frac=0.1)
        x, pile, vcf, refw, nt, nb = seam_b.process_block(T, B, M, ref, y)
        xo, wpile, want = oracle.process_block(T, B, M, refw, y)
        assert xo == x
        util.assert_vcf_close(vcf, want)


@pytest.mark.parametrize("name", ["reader_pe", "reader_mixed"])
def test_seam_b_under_the_references_read_input(seam_b, name):
    """the reference's own read_input (compiled) feeding the device's process_template_vector block after block"""
    g = util.load_golden(name)
    refs = [g["ref%d" % i] for i in range(len(g["target_len"]))]
    blocks, tm, bases, misms, vcf = seam_b.read_input(g["bam"], g["target_len"], refs, mapq_thresh=int(g["mapq_thresh"]),
                                                      max_template_len=int(g["max_template_len"]), keep_unmatched=bool(g["keep_unmatched"]),
                                                      ignore_duplicates=bool(g["ignore_duplicates"]), keep_duplicates=bool(g["keep_duplicates"]), run_chain=True)
    assert len(blocks) == len(g["blocks"])
    n = 0
    for b, w in zip(blocks, g["blocks"]):
        assert (b["tid"], b["x"], b["y"], b["n_templates"]) == (w["tid"], w["x"], w["y"], w["n_templates"])
        sz = int(w["y"]) - int(w["x"]) + 1
        n += util.assert_vcf_close(vcf[int(b["vcf_off"]):int(b["vcf_off"]) + sz], g["vcf"][int(w["vcf_off"]):int(w["vcf_off"]) + sz])
    assert n > 5000


# ---- seams C / D: bs_call_b200/csrc/bsgpu_seam_reader.c replaces src/get_template_vector.c + src/process_template.c +
# src/call_genotypes.c: read_input feeds a streaming session record by record; results go to the reference's print-thread
# protocol (C) or, as BCF records, to bcf_write (D; BSGPU_SEAM_RECORDS=1)  -> oracle/_ref/libbsref_seamC.so
@pytest.fixture(scope="module")
def seam_c():
    from oracle.bindings import ReferenceWithSeamC, seam_available
    if not seam_available("seamC"):
        pytest.skip("oracle/_ref/libbsref_seamC.so not built (reference tree absent at build time)")
    return ReferenceWithSeamC(calc_threads=1)


def _golden_kw(g):
    return dict(mapq_thresh=int(g["mapq_thresh"]), max_template_len=int(g["max_template_len"]), keep_unmatched=bool(g["keep_unmatched"]),
                ignore_duplicates=bool(g["ignore_duplicates"]), keep_duplicates=bool(g["keep_duplicates"]))


@pytest.mark.parametrize("name", ["reader_pe", "reader_mixed"])
def test_seam_c_blocks_reach_the_print_thread(seam_c, name, monkeypatch):
    """record by record through the product's read_input; every block arrives at the print thread with the reference's window,
    reference string and gt_vcf records (goldens of the all-CPU chain)"""
    monkeypatch.delenv("BSGPU_SEAM_RECORDS", raising=False)
    monkeypatch.setenv("BSGPU_BATCH_BYTES", "60000")        # several batches, carried records, results released in turn
    g = util.load_golden(name)
    refs = [g["ref%d" % i] for i in range(len(g["target_len"]))]
    blocks, vcf, rec, nrec = seam_c.seam_read_input(g["bam"], g["target_len"], refs, **_golden_kw(g))
    assert nrec == 0 and len(blocks) == len(g["blocks"])
    n = 0
    for b, w in zip(blocks, g["blocks"]):
        assert (b["tid"], b["x"], b["y"]) == (w["tid"], w["x"], w["y"])
        sz = int(w["y"]) - int(w["x"]) + 1
        n += util.assert_vcf_close(vcf[int(b["vcf_off"]):int(b["vcf_off"]) + sz], g["vcf"][int(w["vcf_off"]):int(w["vcf_off"]) + sz])
    assert n > 5000


def test_seam_d_records_reach_bcf_write(seam_c, monkeypatch):
    """the same with BSGPU_SEAM_RECORDS=1: the print thread stays idle, bcf_write receives the records of the reference's writer"""
    monkeypatch.setenv("BSGPU_SEAM_RECORDS", "1")
    monkeypatch.setenv("BSGPU_BATCH_BYTES", "80000")
    g = util.load_golden("reader_pe")
    w = util.load_golden("writer_v1")
    refs = [g["ref%d" % i] for i in range(len(g["target_len"]))]
    blocks, vcf, rec, nrec = seam_c.seam_read_input(g["bam"], g["target_len"], refs, **_golden_kw(g))
    assert len(blocks) == 0
    got, want = util.split_bcf(rec), util.split_bcf(w["reader_pe__bcf"])
    assert nrec == len(want) == len(got) and [r[:32] for r in got] == [r[:32] for r in want]


@pytest.mark.parametrize("seed", [31, 32])
def test_seam_c_oracle_streams(seam_c, oracle, seed, monkeypatch):
    monkeypatch.delenv("BSGPU_SEAM_RECORDS", raising=False)
    monkeypatch.setenv("BSGPU_BATCH_BYTES", "50000")
    bam, n, tl, refs = bamgen.make_stream(500 + seed, n_contigs=3, dup=0.2, junk=0.15, contig_len=8000)
    o = dict(mapq_thresh=15, max_template_len=800, keep_unmatched=seed == 32, ignore_duplicates=False, keep_duplicates=False)
    wbk, wt, wb, wm, wv = oracle.read_input(bam, tl, refs, run_chain=True, **o)
    blocks, vcf, rec, nrec = seam_c.seam_read_input(bam, tl, refs, **o)
    assert len(blocks) == len(wbk) > 3
    for b, w in zip(blocks, wbk):
        assert (b["tid"], b["x"], b["y"]) == (w["tid"], w["x"], w["y"])
        sz = int(w["y"]) - int(w["x"]) + 1
        util.assert_vcf_close(vcf[int(b["vcf_off"]):int(b["vcf_off"]) + sz], wv[int(w["vcf_off"]):int(w["vcf_off"]) + sz])


def test_seam_d_statistics_reach_the_report(seam_c, oracle, monkeypatch):
    """seam D with --report-file: no gt_vcf record reaches the reference's writer, so what it would have added to bs_stats and
    to the contigs' ctg_stats (src/print_vcf.c:382-526) is gathered on the device and folded in by join_calc_threads -- against
    the pinned restatement over the blocks a seam C run of the same stream hands to the print thread"""
    # snps / multi: the restatement is pinned to what oracle/_ref/libbsref.so does (a hom-ref call counts as "multi"); which of the
    # two the seam's own link would do is a property of its string pool (bsgpu_seam_reader.c: homref_is_multi), so say it here
    monkeypatch.setenv("BSGPU_STATS_HOMREF_MULTI", "1")
    from oracle.bindings import SITE_STATS, site_stats_equal
    from tests import blockgen
    bam, n, tl, refs = bamgen.make_stream(640, n_contigs=3, dup=0.15, contig_len=9000)
    o = dict(mapq_thresh=20, max_template_len=1000, keep_unmatched=False, ignore_duplicates=False, keep_duplicates=False)
    monkeypatch.setenv("BSGPU_BATCH_BYTES", "70000")
    monkeypatch.delenv("BSGPU_SEAM_RECORDS", raising=False)
    blocks, vcf, _, _ = seam_c.seam_read_input(bam, tl, refs, **o)
    want = np.zeros(1, dtype=SITE_STATS)
    state = np.zeros(2, dtype=np.uint32)
    for b in blocks:
        x, y, tid = int(b["x"]), int(b["y"]), int(b["tid"])
        codes = np.asarray(refs[tid], dtype=np.uint8) & 7
        nb = (len(codes) + 99) // 100
        pad = np.zeros(nb * 100, dtype=np.uint8)
        pad[:len(codes)] = codes
        known = (pad != 0).reshape(nb, 100).sum(axis=1)
        gcn = ((pad == 2) | (pad == 3)).reshape(nb, 100).sum(axis=1)
        gc = np.where(known > 0, 100 * gcn // np.maximum(known, 1), 255).astype(np.uint8)
        v = vcf[int(b["vcf_off"]):int(b["vcf_off"]) + y - x + 1]
        oracle.stats_block(v, blockgen.window_codes(refs[tid], x, y + 2), x, ctg_end=int(tl[tid]), gc=gc, start_pos=1, stats=want, state=state)
    monkeypatch.setenv("BSGPU_SEAM_RECORDS", "1")
    seam_c.stats_enable(True)
    seam_c.writer_stats_reset()
    try:
        _, _, rec, nrec = seam_c.seam_read_input(bam, tl, refs, **o)
        got, ctg = seam_c.writer_stats_read()
    finally:
        seam_c.stats_enable(False)
    site_stats_equal(got[0], want[0], rtol=1e-9, what="seam D")
    assert nrec > 300 and int(want[0]["snps"][0] + want[0]["multi"][0]) == nrec
    assert np.array_equal(ctg[:, 0], [want[0][f][0] for f in ("snps", "multi", "dbSNP_sites", "dbSNP_var", "CpG_ref", "CpG_nonref")])
    assert want[0]["cov"]["gc_pcent"].sum() > 1000
