"""Sharding of the path across the GPUs of one box.

The unit of work is independent by construction: contigs share nothing, and within a contig a *block* (maximal run of
templates with gaps <= 1 bp, src/get_template_vector.c:140-148) shares nothing with other blocks, not even the writer's
5-site context (flush_vcf_entries, src/print_vcf.c:536-546).  So ranks never exchange data on the hot path; the only
collective is the final ordered merge of per-rank results on the host (what the reference does with
`bcftools concat`, src/process_sam_header.c:52-70).

  level 1  contigs -> ranks by longest-processing-time first on their base count
  level 2  one long contig -> windows cut at block boundaries (the `-C` region machinery of the reference)
  stream   the synthetic per-site stream of config 2 -> equal contiguous site ranges
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple


def lpt_assign(weights: Sequence[float], n_ranks: int) -> List[int]:
    """Longest-processing-time-first: returns rank of each item; deterministic (ties by index)."""
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    load = [0.0] * n_ranks
    owner = [0] * len(weights)
    for i in order:
        r = min(range(n_ranks), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += weights[i]
    return owner


@dataclass(frozen=True)
class Region:
    contig: int
    start: int        # 1-based inclusive
    stop: int         # inclusive


def split_contig(contig: int, length: int, n_parts: int, boundaries: Sequence[int] = ()) -> List[Region]:
    """Cut [1, length] into n_parts windows of near-equal size.  If `boundaries` (sorted positions where a new block
    starts) is given, every cut is moved to the nearest block boundary so that no block straddles two windows."""
    cuts = [1]
    for k in range(1, n_parts):
        ideal = 1 + (length * k) // n_parts
        if boundaries:
            import bisect
            j = bisect.bisect_left(boundaries, ideal)
            cand = [b for b in boundaries[max(0, j - 1):j + 1]]
            ideal = min(cand, key=lambda b: abs(b - ideal)) if cand else ideal
        if ideal > cuts[-1]:
            cuts.append(ideal)
    cuts.append(length + 1)
    return [Region(contig, cuts[i], cuts[i + 1] - 1) for i in range(len(cuts) - 1)]


def plan(contig_lengths: Sequence[int], n_ranks: int, split_over: float = 1.5,
         boundaries: Optional[Sequence[Sequence[int]]] = None, weights: Optional[Sequence[float]] = None) -> List[List[Region]]:
    """Per-rank region lists.  A contig heavier than split_over x the ideal per-rank share is first split (level 2).

    boundaries[c] = sorted positions at which a new block starts on contig c (block x; from the BAM index in a real run).
    A contig is only ever cut AT such a position: a block that straddled a cut would be split between two ranks, which
    changes mate pairing, duplicate removal and the pileup at the edge.  Without boundaries for a contig it is never split
    (plan() then stays valid for reads; equal-size cuts are only right for the per-site stream, see site_range()).
    weights[c] = work of contig c (its base count) when that is not proportional to its length."""
    w = [float(x) for x in (weights if weights is not None else contig_lengths)]
    total = float(sum(w))
    share = total / n_ranks
    regions: List[Region] = []
    rw: List[float] = []
    for c, ln in enumerate(contig_lengths):
        bnd = boundaries[c] if boundaries is not None and c < len(boundaries) else None
        parts = max(2, int(round(w[c] / share))) if n_ranks > 1 and w[c] > split_over * share and bnd else 1
        rs = split_contig(c, ln, parts, bnd or ())
        regions.extend(rs)
        rw.extend(w[c] * (r.stop - r.start + 1) / max(ln, 1) for r in rs)
    owner = lpt_assign(rw, n_ranks)
    out: List[List[Region]] = [[] for _ in range(n_ranks)]
    for r, o in zip(regions, owner):
        out[o].append(r)
    for lst in out:
        lst.sort(key=lambda r: (r.contig, r.start))
    return out


def site_range(rank: int, world: int, n: int) -> Tuple[int, int]:
    """contiguous [first, last) slice of a stream of n sites"""
    base, rem = divmod(n, world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def merge_in_coordinate_order(per_rank):
    """per_rank: list (one per rank) of lists of (Region, payload).  Returns payloads in (contig, start) order -- the
    order the single print thread of the reference would have produced."""
    flat = [item for lst in per_rank for item in lst]
    flat.sort(key=lambda it: (it[0].contig, it[0].start))
    for a, b in zip(flat, flat[1:]):
        if a[0].contig == b[0].contig and a[0].stop >= b[0].start:
            raise ValueError("overlapping regions %r %r" % (a[0], b[0]))
    return [p for _, p in flat]


def split_records_by_contig(bam, owner_of_contig: Sequence[int], n_ranks: int):
    """Level 1 for a coordinate-sorted BAM record stream (the bytes after the header of an uncompressed BAM): records of
    contig c go to rank owner_of_contig[c], order preserved, so every rank's stream is itself coordinate sorted and holds
    whole contigs -- each can be handed to bsgpu_call_bam as it is.  Records without a contig (refID < 0) are dropped
    (read_input asserts curr_tid >= 0, src/get_template_vector.c:119).  Returns one uint8 array per rank."""
    import numpy as np
    bam = np.ascontiguousarray(bam, dtype=np.uint8)
    parts = [[] for _ in range(n_ranks)]
    at, n = 0, len(bam)
    run_start, run_owner = 0, None
    while at < n:
        bs = int(bam[at:at + 4].view("<i4")[0])
        tid = int(bam[at + 4:at + 8].view("<i4")[0])
        owner = owner_of_contig[tid] if 0 <= tid < len(owner_of_contig) else None
        if owner != run_owner:
            if run_owner is not None:
                parts[run_owner].append(bam[run_start:at])
            run_start, run_owner = at, owner
        at += 4 + bs
    if run_owner is not None:
        parts[run_owner].append(bam[run_start:at])
    return [np.concatenate(p) if p else np.zeros(0, dtype=np.uint8) for p in parts]


HG38_CONTIGS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
                133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
                58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
