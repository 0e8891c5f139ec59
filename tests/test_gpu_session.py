"""Streaming session (-m gpu): bsgpu_bam_open / _feed / _finish / _drain / _release / _close.

A session must give exactly what the one-shot entry points give on the same stream -- the same blocks in the same order,
the same gt_vcf[] records per block, the same BCF bytes -- whatever the batch size and however the stream is sliced on the
way in (also inside records), because it only ever cuts the stream at records where read_input (src/get_template_vector.c:
111-149) starts a new block with blank state.  Integer / byte work and the same device code on both sides: bit-exact."""
import numpy as np
import pytest

from bs_call_b200 import lib as bslib
from tests import bamgen, util

pytestmark = pytest.mark.gpu

FIELDS = ("counts", "qual", "gt_prob", "fisher_strand", "mq", "aq", "max_gt")


@pytest.fixture(scope="module")
def gpu():
    g = bslib.BsGpu()
    yield g
    g.close()


def _same_records(a, b):
    assert len(a) == len(b)
    assert a["skip"].tobytes() == b["skip"].tobytes()
    for f in FIELDS:
        assert np.ascontiguousarray(a["gtm"][f]).tobytes() == np.ascontiguousarray(b["gtm"][f]).tobytes(), f


def _check_vcf_batches(batches, blocks, vcf):
    """batches of a gt_vcf session against the (blocks, vcf) of bsgpu_call_bam on the whole stream"""
    k = 0
    for b, d, n in batches:
        assert n == len(d)
        for blk in b:
            w = blocks[k]
            assert (blk["tid"], blk["x"], blk["y"], blk["n_templates"]) == (w["tid"], w["x"], w["y"], w["n_templates"]), k
            sz = int(w["y"]) - int(w["x"]) + 1
            _same_records(d[int(blk["vcf_off"]):int(blk["vcf_off"]) + sz], vcf[int(w["vcf_off"]):int(w["vcf_off"]) + sz])
            k += 1
    assert k == len(blocks)


@pytest.mark.parametrize("batch,slice_bytes,nowait", [(4096, 777, True), (20000, 1 << 20, True), (9000, 5001, False), (1 << 30, 1234, True), (6000, 1 << 20, False)])
def test_session_vcf_equals_one_shot(gpu, batch, slice_bytes, nowait):
    bam, n, tl, refs = bamgen.make_stream(411, n_contigs=3, dup=0.2, junk=0.1, contig_len=9000)
    blocks, vcf = gpu.call_bam(bam, tl, refs)
    blocks, vcf = blocks.copy(), vcf.copy()
    s = gpu.bam_session(tl, refs, batch_bytes=batch)
    try:
        out = s.run(bam, slice_bytes=slice_bytes, nowait=nowait)
        pr = s.progress()
    finally:
        s.close()
    assert pr["bytes_fed"] == pr["bytes_done"] == len(bam) and pr["records_done"] == n
    if batch < len(bam) // 4:
        assert pr["batches"] > 2
    _check_vcf_batches(out, blocks, vcf)


@pytest.mark.parametrize("batch,slice_bytes", [(4096, 901), (30000, 1 << 16), (1 << 30, 1 << 20)])
def test_session_bcf_equals_one_shot(gpu, batch, slice_bytes):
    bam, n, tl, refs = bamgen.make_stream(412, n_contigs=3, dup=0.1, contig_len=8000)
    rid = np.array([5, 0, 9], dtype=np.int32)
    blocks, rec, nrec = gpu.call_bam_bcf(bam, tl, refs, vcf_rid=rid)
    want = rec.tobytes()
    s = gpu.bam_session(tl, refs, bcf=True, vcf_rid=rid, batch_bytes=batch)
    try:
        out = s.run(bam, slice_bytes=slice_bytes)
    finally:
        s.close()
    got = b"".join(d.tobytes() for _, d, _ in out)
    assert sum(n_ for _, _, n_ in out) == nrec and got == want
    gb = np.concatenate([b for b, _, _ in out])
    for f in ("tid", "x", "y", "n_templates"):
        assert (gb[f] == blocks[f]).all(), f


def test_session_golden_writer_records(gpu):
    """reader golden through a session with batches far smaller than the stream: the records of the reference's writer"""
    g = util.load_golden("reader_pe")
    w = util.load_golden("writer_v1")
    refs = [g["ref%d" % i] for i in range(len(g["target_len"]))]
    rp = bslib.reader_params(mapq_thresh=int(g["mapq_thresh"]), max_template_len=int(g["max_template_len"]))
    s = gpu.bam_session(g["target_len"], refs, rp=rp, bcf=True, batch_bytes=8192)
    try:
        out = s.run(g["bam"], slice_bytes=3333)
    finally:
        s.close()
    got = util.split_bcf(np.frombuffer(b"".join(d.tobytes() for _, d, _ in out), dtype=np.uint8))
    want = util.split_bcf(w["reader_pe__bcf"])
    assert len(got) == len(want) and [r[:32] for r in got] == [r[:32] for r in want]


def test_session_block_larger_than_batch(gpu):
    """one contig, no coverage gap: no certain block start inside any batch, so the staging grows until the stream ends"""
    bam, n, tl, refs = bamgen.make_stream(413, n_contigs=1, contig_len=4000)
    blocks, vcf = gpu.call_bam(bam, tl, refs)
    blocks, vcf = blocks.copy(), vcf.copy()
    s = gpu.bam_session(tl, refs, batch_bytes=4096)
    try:
        out = s.run(bam, slice_bytes=2048)
        pr = s.progress()
    finally:
        s.close()
    _check_vcf_batches(out, blocks, vcf)
    if len(blocks) == 1:
        assert pr["empty_batches"] > 0


def test_session_reserve_commit(gpu):
    bam, n, tl, refs = bamgen.make_stream(414, n_contigs=2, dup=0.2, contig_len=7000)
    blocks, vcf = gpu.call_bam(bam, tl, refs)
    blocks, vcf = blocks.copy(), vcf.copy()
    s = gpu.bam_session(tl, refs, batch_bytes=16384)
    out = []

    def collect(got):
        b, d, k, r = got
        out.append((b.copy(), d.copy(), k))
        s.release(r)

    try:
        at = 0
        while at < len(bam):
            view = s.reserve(wait=False)          # one thread: never wait for room, drain instead
            if not len(view):
                got = s.drain(wait=True)
                if got is not None:
                    collect(got)
                continue
            m = min(len(view), len(bam) - at, 5000)
            view[:m] = bam[at:at + m]
            s.commit(m)
            at += m
        s.finish()
        while True:
            got = s.drain(wait=True)
            if got is not None:
                collect(got)
            if s.finished:
                break
    finally:
        s.close()
    _check_vcf_batches(out, blocks, vcf)


def test_session_truncated_stream_fails_at_finish(gpu):
    bam, n, tl, refs = bamgen.make_stream(415, n_contigs=1, contig_len=3000)
    s = gpu.bam_session(tl, refs, batch_bytes=1 << 20)
    try:
        s.feed(bam[:-7])
        s.finish()
        with pytest.raises(bslib.BsGpuError):
            while True:
                got = s.drain(wait=True)
                if got is not None:
                    s.release(got[3])
                if s.finished:
                    break
    finally:
        s.close()


def test_session_profile_side_channels(gpu):
    """the --report-file tallies accumulate over the batches of a session exactly as over one call"""
    bam, n, tl, refs = bamgen.make_stream(416, n_contigs=2, dup=0.2, junk=0.2, contig_len=8000)
    gpu.profile_enable(True)
    try:
        gpu.profile_read(reset=True)
        gpu.call_bam(bam, tl, refs)
        want = gpu.profile_read(reset=True)
        s = gpu.bam_session(tl, refs, batch_bytes=10000)
        try:
            s.run(bam, slice_bytes=4097, keep=False)
        finally:
            s.close()
        got = gpu.profile_read(reset=True)
    finally:
        gpu.profile_enable(False)
    for k in want:
        assert np.array_equal(np.asarray(got[k]), np.asarray(want[k])), k


def test_session_empty_stream(gpu):
    s = gpu.bam_session([1000], [np.ones(1000, np.uint8)], batch_bytes=4096)
    try:
        s.finish()
        assert s.drain(wait=True) is None and s.finished
    finally:
        s.close()
