"""Multi-rank host logic on CPU: sharding plans, and a world_size-2 gloo run in which each rank calls the (CPU) oracle
on its shard of the site stream and rank 0 merges in coordinate order -- the N>1 plumbing of bench.py without GPUs."""
import os
import subprocess
import sys

import numpy as np

from bs_call_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_balances_hg38_over_8():
    owner = shard.lpt_assign(shard.HG38_CONTIGS, 8)
    load = np.zeros(8)
    for ln, o in zip(shard.HG38_CONTIGS, owner):
        load[o] += ln
    assert load.max() / load.mean() < 1.05          # within a few %
    assert sorted(set(owner)) == list(range(8))


def test_plan_covers_everything_once():
    for n in (1, 2, 4, 8):
        for lengths in (shard.HG38_CONTIGS, [50_000_000], [1000, 7, 5_000_000, 12]):
            p = shard.plan(lengths, n)
            assert len(p) == n
            seen = {c: [] for c in range(len(lengths))}
            for lst in p:
                for r in lst:
                    seen[r.contig].append((r.start, r.stop))
            for c, ln in enumerate(lengths):
                iv = sorted(seen[c])
                assert iv[0][0] == 1 and iv[-1][1] == ln
                for a, b in zip(iv, iv[1:]):
                    assert a[1] + 1 == b[0]
    # one contig on 8 GPUs is split (level 2) -- at block boundaries, and only when they are known
    bounds = list(range(1, 50_000_000, 99_371))
    p = shard.plan([50_000_000], 8, boundaries=[bounds])
    assert all(len(lst) == 1 for lst in p)
    for lst in p:
        assert lst[0].start in bounds
    loads = [lst[0].stop - lst[0].start + 1 for lst in p]
    assert max(loads) / (50_000_000 / 8) < 1.05
    p = shard.plan([50_000_000], 8)                    # no boundaries: never cut through a block
    assert sum(len(lst) for lst in p) == 1


def test_plan_hg38_level2_split_at_8():
    """24 hg38-shaped contigs on 8 ranks with split_over = 0.4: the largest contigs are cut in two at block boundaries and
    the plan balances within a few per cent"""
    lens = shard.HG38_CONTIGS
    bounds = [list(range(1, ln, 100_400)) for ln in lens]
    p = shard.plan(lens, 8, split_over=0.4, boundaries=bounds)
    load = [sum(r.stop - r.start + 1 for r in lst) for lst in p]
    assert max(load) / (sum(lens) / 8) < 1.05
    split = [r for lst in p for r in lst if not (r.start == 1 and r.stop == lens[r.contig])]
    assert split and all(r.start == 1 or r.start in bounds[r.contig] for r in split)


def test_split_respects_block_boundaries():
    bounds = [1, 900, 2500, 2600, 7000, 9100]
    regs = shard.split_contig(0, 10000, 4, bounds)
    for r in regs[1:]:
        assert r.start in bounds
    assert regs[0].start == 1 and regs[-1].stop == 10000


def test_site_range():
    for n in (0, 1, 10, 1_000_000_007):
        for w in (1, 2, 3, 8):
            rs = [shard.site_range(r, w, n) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from bs_call_b200 import shard
from oracle.bindings import Oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
o = Oracle()
N = 60000
first, last = shard.site_range(rank, world, N)
p, r = o.synth_sites(20261018, first, last - first, 30.0)
out, skip = o.call_sites(p, r)
mine = [(shard.Region(0, first + 1, last), (out.tobytes(), skip.tobytes()))]
gathered = [None] * world
dist.all_gather_object(gathered, mine)
dist.barrier()
if rank == 0:
    merged = shard.merge_in_coordinate_order(gathered)
    whole_p, whole_r = o.synth_sites(20261018, 0, N, 30.0)
    wout, wskip = o.call_sites(whole_p, whole_r)
    assert b"".join(m[0] for m in merged) == wout.tobytes()
    assert b"".join(m[1] for m in merged) == wskip.tobytes()
    print("MERGE_OK", world)
dist.destroy_process_group()
'''


def test_two_rank_gloo_shard_and_merge(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "MERGE_OK 2" in res.stdout


BAM_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from bs_call_b200 import shard
from oracle.bindings import Oracle
from tests import bamgen
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
o = Oracle()
bam, n, tl, refs = bamgen.make_stream(55, n_contigs=3, contig_len=4000)
owner = shard.lpt_assign([int(v) for v in tl], world)
mine = shard.split_records_by_contig(bam, owner, world)[rank]
blocks, tm, b, m, vcf = o.read_input(mine, tl, refs, run_chain=True)
out = [(shard.Region(int(k["tid"]), int(k["x"]), int(k["y"])),
        vcf[int(k["vcf_off"]):int(k["vcf_off"]) + int(k["y"]) - int(k["x"]) + 1].tobytes()) for k in blocks]
gathered = [None] * world
dist.all_gather_object(gathered, out)
dist.barrier()
if rank == 0:
    # consecutive blocks of a contig may share one (uncovered) position: merge on (contig, start) without the overlap check
    flat = sorted((it for lst in gathered for it in lst), key=lambda it: (it[0].contig, it[0].start))
    wb, wt, wbb, wm, wv = o.read_input(bam, tl, refs, run_chain=True)
    assert len(flat) == len(wb)
    for (reg, payload), k in zip(flat, wb):
        assert (reg.contig, reg.start, reg.stop) == (int(k["tid"]), int(k["x"]), int(k["y"]))
        assert payload == wv[int(k["vcf_off"]):int(k["vcf_off"]) + int(k["y"]) - int(k["x"]) + 1].tobytes()
    assert sorted(set(owner)) == list(range(world))
    print("BAM_MERGE_OK", world, len(flat))
dist.destroy_process_group()
"""


def test_two_rank_gloo_bam_stream_sharded_by_contig(tmp_path):
    """level-1 sharding of a BAM record stream: each rank runs the whole chain (the CPU oracle stands in for the device)
    on the contigs it owns; merged in coordinate order the blocks are those of the single-process run"""
    script = tmp_path / "bam_worker.py"
    script.write_text(BAM_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29534", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "BAM_MERGE_OK 2" in res.stdout


def test_synthetic_dbsnp_table():
    """bench.py's genome leg annotates every contig (configs[4]: "with dbSNP annotation"): the table is a pure function of
    (seed, contig), sorted, one entry per stride, laid out as bsgpu_dbsnp wants it"""
    from bs_call_b200 import synthgenome as sg
    pos, flags, off, names = sg.contig_dbsnp(7, 3, 1_000_000)
    pos2 = sg.contig_dbsnp(7, 3, 1_000_000)[0]
    assert (pos == pos2).all() and (sg.contig_dbsnp(7, 4, 1_000_000)[0] != pos).any()
    n = len(pos)
    assert n == 999_999 // sg.DBSNP_EVERY and (np.diff(pos.astype(np.int64)) > 0).all() and pos[0] >= 1 and pos[-1] < 1_000_000
    assert set(np.unique(flags)) == {1, 3} and 0.02 < (flags == 3).mean() < 0.1
    assert len(off) == n + 1 and off[-1] == 11 * n and len(names) == 11 * n + 1
    nm = names[:-1].reshape(n, 11)
    assert (nm[:, 0] == ord("r")).all() and (nm[:, 1] == ord("s")).all() and ((nm[:, 2:] >= 48) & (nm[:, 2:] <= 57)).all()
