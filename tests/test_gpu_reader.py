"""Reader-side parity (-m gpu): BAM record decode on the device (bsgpu_decode_records), the block builder over its
descriptors, and the whole chain raw records -> gt_vcf[] (bsgpu_call_bam), against the goldens captured from the
reference's own get_next_align_details / read_input / process_template_vector / call_genotypes_ML and against the
oracle on seeded streams.  Everything on the reader side is integer / byte work: bit-exact."""
import os

import numpy as np
import pytest

from bs_call_b200 import lib as bslib
from tests import bamgen, util

pytestmark = pytest.mark.gpu

KEPT_ONLY = ("bs_strand", "align_length", "reference_span", "read_len", "mm_n")


@pytest.fixture(scope="module")
def gpu():
    g = bslib.BsGpu()
    yield g
    g.close()


def _rp(o):
    return bslib.reader_params(mapq_thresh=o["mapq_thresh"], max_template_len=o["max_template_len"], keep_unmatched=o["keep_unmatched"],
                               ignore_duplicates=o["ignore_duplicates"], keep_duplicates=o["keep_duplicates"])


def _check_records(rec, bases, misms, want, wbases, wmisms):
    assert len(rec) == len(want)
    kept = want["ret"] == 0
    for f in ("ret", "filtered", "forward_position", "reverse_position", "alignment_flag", "reverse", "orientation", "mapq"):
        assert (rec[f] == want[f]).all(), f
    for f in KEPT_ONLY:
        assert (rec[f][kept] == want[f][kept]).all(), f
    # decoded reads and events, record by record (the two sides lay them out differently)
    for i in np.nonzero(kept)[0]:
        a = bases[int(rec["read_off"][i]):int(rec["read_off"][i]) + int(rec["read_len"][i])]
        b = wbases[int(want["read_off"][i]):int(want["read_off"][i]) + int(want["read_len"][i])]
        assert a.tobytes() == b.tobytes(), "read of record %d" % i
        a = misms[int(rec["mm_off"][i]):int(rec["mm_off"][i]) + int(rec["mm_n"][i])]
        b = wmisms[int(want["mm_off"][i]):int(want["mm_off"][i]) + int(want["mm_n"][i])]
        assert a.tobytes() == b.tobytes(), "events of record %d" % i
        for k in range(min(2, int(rec["read_len"][i]))):
            assert rec["q01"][i, k] == b"".join([bases[int(rec["read_off"][i]) + k:int(rec["read_off"][i]) + k + 1].tobytes()])[0] >> 2


def _golden_opts(g):
    return dict(mapq_thresh=int(g["mapq_thresh"]), max_template_len=int(g["max_template_len"]), keep_unmatched=bool(g["keep_unmatched"]),
                ignore_duplicates=bool(g["ignore_duplicates"]), keep_duplicates=bool(g["keep_duplicates"]))


@pytest.mark.parametrize("name", ["reader_pe", "reader_mixed"])
def test_decode_records_golden(gpu, name):
    g = util.load_golden(name)
    before = gpu.stats()["kernel_launches"]
    rec, bases, misms = gpu.decode_records(g["bam"], _rp(_golden_opts(g)))
    # k_decode_records + k_name_ids, and the three launches of the certain-start scan unless the host scans (BSGPU_HOST_SCAN)
    assert gpu.stats()["kernel_launches"] == before + (2 if os.environ.get("BSGPU_HOST_SCAN") else 5)
    _check_records(rec, bases, misms, g["rec"], g["rec_bases"], g["rec_misms"])


@pytest.mark.parametrize("seed", range(6))
def test_decode_records_oracle(gpu, oracle, seed):
    bam, n, _, _ = bamgen.make_stream(100 + seed)
    ku, ig = seed % 3 == 1, seed % 2 == 1
    want, wb, wm = oracle.decode_records(bam, 20 - seed, 400 + 150 * seed, ku, ig)
    rec, bases, misms = gpu.decode_records(bam, bslib.reader_params(mapq_thresh=20 - seed, max_template_len=400 + 150 * seed,
                                                                  keep_unmatched=ku, ignore_duplicates=ig))
    assert len(rec) == n
    _check_records(rec, bases, misms, want, wb, wm)


@pytest.mark.parametrize("seed", range(3))
def test_decode_exotic_records(gpu, oracle, seed):
    """= / X / H / N / P / B operators, arbitrary flag words, unknown tag types, qualities up to 254"""
    bam, n = bamgen.exotic_stream(40 + seed)
    for ku in (False, True):
        want, wb, wm = oracle.decode_records(bam, 10, 700, ku, seed == 1)
        rec, bases, misms = gpu.decode_records(bam, bslib.reader_params(mapq_thresh=10, max_template_len=700, keep_unmatched=ku,
                                                                      ignore_duplicates=seed == 1))
        assert len(rec) == n
        _check_records(rec, bases, misms, want, wb, wm)


@pytest.mark.parametrize("name", ["reader_pe", "reader_mixed"])
def test_blocks_from_device_descriptors_golden(gpu, name):
    """decode on the device, build blocks on the host from the device's descriptors: the reference's blocks"""
    g = util.load_golden(name)
    o = _golden_opts(g)
    rec, bases, misms = gpu.decode_records(g["bam"], _rp(o))
    blocks, tm = bslib.build_blocks(g["bam"], rec, _rp(o))
    for f in ("tid", "x", "y", "first_template", "n_templates"):
        assert (blocks[f] == g["blocks"][f]).all(), f
    assert bamgen.template_keys(tm, bases, misms) == bamgen.template_keys(g["templates"], g["bases"], g["misms"])


@pytest.mark.parametrize("name", ["reader_pe", "reader_mixed"])
def test_call_bam_golden(gpu, name):
    """raw records in, gt_vcf[] of every block out: what the reference's whole chain produced"""
    g = util.load_golden(name)
    refs = [g["ref%d" % i] for i in range(len(g["target_len"]))]
    blocks, vcf = gpu.call_bam(g["bam"], g["target_len"], refs, _rp(_golden_opts(g)))
    want_b, want_v = g["blocks"], g["vcf"]
    assert len(blocks) == len(want_b)
    ncalled = 0
    for b, w in zip(blocks, want_b):
        assert (b["tid"], b["x"], b["y"], b["n_templates"]) == (w["tid"], w["x"], w["y"], w["n_templates"])
        sz = int(w["y"]) - int(w["x"]) + 1
        got = vcf[int(b["vcf_off"]):int(b["vcf_off"]) + sz]
        ncalled += util.assert_vcf_close(got, want_v[int(w["vcf_off"]):int(w["vcf_off"]) + sz])
    assert ncalled > 5000


@pytest.mark.parametrize("seed", range(4))
def test_call_bam_oracle(gpu, oracle, seed):
    bam, n, tl, refs = bamgen.make_stream(200 + seed, dup=0.2)
    o = dict(mapq_thresh=15, max_template_len=800, keep_unmatched=seed == 2, ignore_duplicates=False, keep_duplicates=seed == 3)
    wbk, wt, wb, wm, wv = oracle.read_input(bam, tl, refs, run_chain=True, **o)
    blocks, vcf = gpu.call_bam(bam, tl, refs, _rp(o))
    assert len(blocks) == len(wbk) > 0
    for b, w in zip(blocks, wbk):
        assert (b["tid"], b["x"], b["y"], b["n_templates"]) == (w["tid"], w["x"], w["y"], w["n_templates"])
        sz = int(w["y"]) - int(w["x"]) + 1
        util.assert_vcf_close(vcf[int(b["vcf_off"]):int(b["vcf_off"]) + sz], wv[int(w["vcf_off"]):int(w["vcf_off"]) + sz])
    # positions between blocks of a contig are uncovered: skip records
    covered = np.zeros(len(vcf), dtype=bool)
    for b in blocks:
        covered[int(b["vcf_off"]):int(b["vcf_off"]) + int(b["y"]) - int(b["x"]) + 1] = True
    assert (vcf["skip"][~covered] == 1).all()


def test_call_bam_chunked_upload_and_decode(gpu, oracle, monkeypatch):
    """the stream uploaded in four byte pieces, decoded chunk by chunk, blocks built up to the last certain block start of
    each chunk: same result as in one piece"""
    bam, n, tl, refs = bamgen.make_stream(301, n_contigs=3, dup=0.2, contig_len=9000)
    o = dict(mapq_thresh=20, max_template_len=1000, keep_unmatched=False, ignore_duplicates=False, keep_duplicates=False)
    wbk, wt, wb, wm, wv = oracle.read_input(bam, tl, refs, run_chain=True, **o)
    monkeypatch.setenv("BSGPU_READER_CHUNK_MIN_BYTES", "1")
    monkeypatch.setenv("BSGPU_BUILDER_THREADS", "3")
    monkeypatch.setenv("BSGPU_BUILDER_MIN_RECORDS", "1")
    blocks, vcf = gpu.call_bam(bam, tl, refs, _rp(o))
    assert len(blocks) == len(wbk) >= 4
    for b, w in zip(blocks, wbk):
        assert (b["tid"], b["x"], b["y"], b["n_templates"], b["first_template"]) == (w["tid"], w["x"], w["y"], w["n_templates"], w["first_template"])
        sz = int(w["y"]) - int(w["x"]) + 1
        util.assert_vcf_close(vcf[int(b["vcf_off"]):int(b["vcf_off"]) + sz], wv[int(w["vcf_off"]):int(w["vcf_off"]) + sz])
    rec, bases, misms = gpu.decode_records(bam, _rp(o))
    want, wb2, wm2 = oracle.decode_records(bam, 20, 1000, False, False)
    _check_records(rec, bases, misms, want, wb2, wm2)


def test_reader_rejects_truncated_stream(gpu):
    bam, n, tl, refs = bamgen.make_stream(7)
    with pytest.raises(bslib.BsGpuError):
        gpu.decode_records(bam[:-5])


def test_reader_empty_stream(gpu):
    rec, bases, misms = gpu.decode_records(np.zeros(0, dtype=np.uint8))
    assert len(rec) == 0


def test_call_bam_pieces_tile_the_contigs(gpu, oracle, monkeypatch):
    """the block builder cut into many pieces on a thread pool, pieces consumed in order by the device stages: the windows
    of successive pieces tile each contig and every block's records are those of the single-piece run"""
    bam, n, tl, refs = bamgen.make_stream(300, n_contigs=3, dup=0.2, contig_len=9000)
    o = dict(mapq_thresh=20, max_template_len=1000, keep_unmatched=False, ignore_duplicates=False, keep_duplicates=False)
    wbk, wt, wb, wm, wv = oracle.read_input(bam, tl, refs, run_chain=True, **o)
    monkeypatch.setenv("BSGPU_BUILDER_THREADS", "5")
    monkeypatch.setenv("BSGPU_BUILDER_MIN_RECORDS", "1")
    blocks, vcf = gpu.call_bam(bam, tl, refs, _rp(o))
    assert len(blocks) == len(wbk) > 6
    for b, w in zip(blocks, wbk):
        assert (b["tid"], b["x"], b["y"], b["n_templates"], b["first_template"]) == (w["tid"], w["x"], w["y"], w["n_templates"], w["first_template"])
        sz = int(w["y"]) - int(w["x"]) + 1
        util.assert_vcf_close(vcf[int(b["vcf_off"]):int(b["vcf_off"]) + sz], wv[int(w["vcf_off"]):int(w["vcf_off"]) + sz])
    covered = np.zeros(len(vcf), dtype=bool)
    for b in blocks:
        covered[int(b["vcf_off"]):int(b["vcf_off"]) + int(b["y"]) - int(b["x"]) + 1] = True
    assert (vcf["skip"][~covered] == 1).all() and (vcf["ready"] == 1).all()


@pytest.mark.parametrize("seed", [1, 6, 13, 27])
def test_device_scan_of_certain_block_starts(monkeypatch, seed):
    """the segmented max-scan on the device (k_certain_starts) marks exactly the records the host scan marks, chunk after
    chunk with the state carried between launches: BSGPU_CHECK_SCAN makes bsgpu_call_bam run both and fail on a difference"""
    bam, n, tl, refs = bamgen.make_stream(seed, n_contigs=3, dup=0.2, junk=0.2, contig_len=6000 + 700 * seed)
    monkeypatch.setenv("BSGPU_CHECK_SCAN", "1")
    monkeypatch.setenv("BSGPU_READER_CHUNK_MIN_BYTES", "1")
    monkeypatch.setenv("BSGPU_BUILDER_MIN_RECORDS", "1")
    g = bslib.BsGpu()
    try:
        for ku in (False, True):
            blocks, vcf = g.call_bam(bam, tl, refs, bslib.reader_params(keep_unmatched=ku))
            assert len(blocks) > 3
    finally:
        g.close()


@pytest.mark.parametrize("name", ["reader_pe", "reader_mixed"])
def test_host_scan_of_certain_block_starts(monkeypatch, name):
    """BSGPU_HOST_SCAN=1: the keys come home and host threads mark the certain block starts (the default before the device's
    scan was spread over the SMs): same blocks as the golden"""
    monkeypatch.setenv("BSGPU_HOST_SCAN", "1")
    monkeypatch.setenv("BSGPU_READER_CHUNK_MIN_BYTES", "1")
    monkeypatch.setenv("BSGPU_BUILDER_MIN_RECORDS", "1")
    g = util.load_golden(name)
    refs = [g["ref%d" % i] for i in range(len(g["target_len"]))]
    gpu = bslib.BsGpu()
    try:
        blocks, vcf = gpu.call_bam(g["bam"], g["target_len"], refs, _rp(_golden_opts(g)))
        assert len(blocks) == len(g["blocks"])
        for f in ("tid", "x", "y", "n_templates"):
            assert (blocks[f] == g["blocks"][f]).all(), f
    finally:
        gpu.close()


def test_exact_offset_table_path(monkeypatch):
    """windows whose mates get exact offsets instead of uniform slots (what a piece with a very long reference span falls
    back to; BSGPU_EXACT_SLOTS forces it): same records"""
    bam, n, tl, refs = bamgen.make_stream(19, n_contigs=2, dup=0.2, junk=0.1)
    g = bslib.BsGpu()
    try:
        b1, v1 = g.call_bam(bam, tl, refs)
        v1 = v1.copy()
    finally:
        g.close()
    monkeypatch.setenv("BSGPU_EXACT_SLOTS", "1")
    import subprocess, sys, os, json
    # the switch is read once per process: run the second pass in a fresh interpreter
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from bs_call_b200 import lib; from tests import bamgen; "
            "bam, n, tl, refs = bamgen.make_stream(19, n_contigs=2, dup=0.2, junk=0.1); g = lib.BsGpu(); b, v = g.call_bam(bam, tl, refs); "
            "import hashlib; print(len(b), hashlib.sha1(b''.join(np.ascontiguousarray(v['gtm'][f]).tobytes() for f in ('counts', 'qual', 'gt_prob', 'fisher_strand', 'mq', 'aq', 'max_gt')) + v['skip'].tobytes()).hexdigest())" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, BSGPU_EXACT_SLOTS="1"))
    assert out.returncode == 0, out.stderr[-500:]
    import hashlib
    nb, digest = out.stdout.split()[-2:]
    want = hashlib.sha1(b"".join(np.ascontiguousarray(v1["gtm"][f]).tobytes() for f in ("counts", "qual", "gt_prob", "fisher_strand", "mq", "aq", "max_gt")) + v1["skip"].tobytes()).hexdigest()
    assert int(nb) == len(b1) and digest == want


def _repeat_on_contigs(bam, copies):
    """the records of a one-contig stream again on contigs 1 .. copies - 1 (same read names: legal, read_input's name table
    is emptied at every block end)"""
    out, at, n = [], 0, len(bam)
    offs = []
    while at < n:
        bs = int(bam[at:at + 4].view("<i4")[0])
        offs.append((at, 4 + bs))
        at += 4 + bs
    for t in range(copies):
        b = bam.copy()
        for o, l in offs:
            b[o + 4:o + 8] = np.frombuffer(np.int32(t).tobytes(), dtype=np.uint8)
            if int(bam[o + 24:o + 28].view("<i4")[0]) >= 0:
                b[o + 24:o + 28] = np.frombuffer(np.int32(t).tobytes(), dtype=np.uint8)
        out.append(b)
    return np.concatenate(out), len(offs) * copies


def test_name_join_overflow_falls_back_to_the_host(gpu, oracle, monkeypatch):
    """every read name on eight kept records (four contigs x two mates): the device's name table keeps five per hash, raises
    its overflow flag, and the host computes the name ids for the affected chunks -- same blocks, same records"""
    bam1, n1, tl1, refs1 = bamgen.make_stream(611, n_contigs=1, dup=0.2, junk=0.1, contig_len=7000, paired=True)
    bam, n = _repeat_on_contigs(bam1, 4)
    tl, refs = np.repeat(tl1, 4), refs1 * 4
    o = dict(mapq_thresh=20, max_template_len=1000, keep_unmatched=False, ignore_duplicates=False, keep_duplicates=False)
    wbk, wt, wb, wm, wv = oracle.read_input(bam, tl, refs, run_chain=True, **o)
    for chunked in (False, True):
        if chunked:
            monkeypatch.setenv("BSGPU_READER_CHUNK_MIN_BYTES", "1")
            monkeypatch.setenv("BSGPU_BUILDER_MIN_RECORDS", "1")
            monkeypatch.setenv("BSGPU_BUILDER_THREADS", "3")
        blocks, vcf = gpu.call_bam(bam, tl, refs, _rp(o))
        assert len(blocks) == len(wbk) > 4
        for b, w in zip(blocks, wbk):
            assert (b["tid"], b["x"], b["y"], b["n_templates"]) == (w["tid"], w["x"], w["y"], w["n_templates"])
            sz = int(w["y"]) - int(w["x"]) + 1
            util.assert_vcf_close(vcf[int(b["vcf_off"]):int(b["vcf_off"]) + sz], wv[int(w["vcf_off"]):int(w["vcf_off"]) + sz])
