"""Synthetic hg38-shaped WGBS genome for bench.py's genome leg and its tests (BASELINE.json configs[4] shape).

24 contigs with the proportions of hg38 (shard.HG38_CONTIGS divided by `scale`), 30x paired-end 150-bp directional
bisulfite reads with 5 % positional duplicates and a 400-bp coverage gap every 10 000 templates (~100 kb), so that
read_input cuts a block there (src/get_template_vector.c:140-148) -- the block boundaries a BAM index would give a real
run.  Everything about a template is a pure function of (contig, absolute template number), so a region of a contig
generated on its own holds exactly the records the whole contig would hold there: the genome can be sharded before it
is generated, and the records every rank calls are the same at any number of ranks.

The record bytes come from the library's device generator (bsgpu_synth_bam_dev); this module only lays out positions.
"""
import bisect

from . import shard

READ_LEN, DEPTH, FRAG, TPB, GAP = 150, 30.0, 300, 10000, 400
STEP = 2 * READ_LEN / DEPTH                      # a new template every STEP positions -> DEPTH x from two mates


def template_pos(t):
    """1-based start of the forward mate of template t (int, numpy or torch integer tensor)"""
    return 101 + (t * int(STEP)) + (t // TPB) * GAP


def contig_lengths(scale):
    return [max(ln // scale, 400000) for ln in shard.HG38_CONTIGS]


def n_templates(length):
    """templates that fit on a contig of `length` positions (the last mate ends well inside it)"""
    per = STEP + GAP / TPB
    nt = int((length - 2000) / per)
    while nt > 0 and template_pos(nt - 1) + FRAG + READ_LEN + 400 > length:
        nt -= 1
    return max(nt, 0)


def block_starts(length):
    """x of every block of the contig (src/process_template.c:24-28: two positions before its first template)"""
    nt = n_templates(length)
    return [template_pos(b * TPB) - 2 for b in range((nt + TPB - 1) // TPB)]


def region_templates(region, length):
    """[t0, t1): the templates of the blocks whose x lies inside the region (regions are cut at block starts)"""
    starts = block_starts(length)
    nt = n_templates(length)
    b0 = bisect.bisect_left(starts, region.start)
    b1 = bisect.bisect_right(starts, region.stop)
    return min(b0 * TPB, nt), min(b1 * TPB, nt)


def plan(scale, world, split_over=0.4):
    lens = contig_lengths(scale)
    return lens, shard.plan(lens, world, split_over=split_over, boundaries=[block_starts(ln) for ln in lens])


def region_stream_dev(gpu, torch, dev, seed, contig, t0, t1, d_out, stream):
    """record stream of templates [t0, t1) of `contig` into the device buffer d_out (uint8 tensor, at least
    gpu.synth_bam_bytes(t1 - t0, READ_LEN) bytes); refID 0 -- the caller patches the contig number in.  Returns the byte count."""
    nt = t1 - t0
    if nt <= 0:
        return 0
    t = torch.arange(t0, t1, device=dev, dtype=torch.int64)
    dup = ((((t * 2654435761) + contig * 40503) >> 11) % 20 == 0) & (t % TPB != 0)
    src = torch.where(dup, t - 1, t)
    dlen = (FRAG - READ_LEN) + ((src * 2654435761) >> 7) % 41 - 20
    pos_f = template_pos(src)
    pos_r = pos_f + dlen
    order = torch.argsort(torch.cat([pos_f, pos_r]), stable=True)
    rank_of = torch.empty_like(order)
    rank_of[order] = torch.arange(2 * nt, device=dev)
    i32 = lambda a: a.to(torch.int32).contiguous()
    a_f, a_r, a_s, a_k = i32(pos_f), i32(pos_r), i32(src), i32(rank_of)
    gpu.synth_bam_dev(seed + 1000003 * contig, nt, READ_LEN, a_f.data_ptr(), a_r.data_ptr(), a_s.data_ptr(), a_k.data_ptr(), d_out.data_ptr(), stream)
    torch.cuda.synchronize()
    return gpu.synth_bam_bytes(nt, READ_LEN)


def patch_contig(view, contig, record_bytes):
    """writes refID / next refID = contig into every (fixed-size) record of a numpy byte view of the stream"""
    if not len(view):
        return
    v = view.reshape(-1, record_bytes)
    v[:, 4] = contig & 0xff
    v[:, 5] = (contig >> 8) & 0xff
    v[:, 24] = contig & 0xff
    v[:, 25] = (contig >> 8) & 0xff


DBSNP_EVERY, DBSNP_ALWAYS = 100, 20


def contig_dbsnp(seed, contig, length):
    """Synthetic dbSNP annotation of a contig (BASELINE.json configs[4]: "with dbSNP annotation"): one known position in
    every DBSNP_EVERY (at a pseudo-random offset inside its stride), one in DBSNP_ALWAYS of them flagged "always written"
    (flag 3, else 1: what dbSNP_lookup_name() answers, src/print_vcf.c:133), ids "rs" + nine digits.  A pure function of
    (seed, contig), so every rank and the CPU check build the same table.  -> (pos u32[], flags u8[], name_off u32[n + 1],
    names u8[]) as bsgpu_dbsnp / oracle.bindings.dbsnp_arrays lay it out."""
    import numpy as np
    n = max((int(length) - 1) // DBSNP_EVERY, 0)
    k = np.arange(n, dtype=np.uint64)
    h = (k + np.uint64(contig) * np.uint64(0x9E3779B1) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    h ^= h >> np.uint64(29)
    h *= np.uint64(0xBF58476D1CE4E5B9)
    h ^= h >> np.uint64(32)
    pos = (1 + k * np.uint64(DBSNP_EVERY) + h % np.uint64(DBSNP_EVERY)).astype(np.uint32)
    flags = np.where((h >> np.uint64(20)) % np.uint64(DBSNP_ALWAYS) == 0, 3, 1).astype(np.uint8)
    num = ((h >> np.uint64(33)) % np.uint64(10 ** 9)).astype(np.int64)
    names = np.empty((n, 11), dtype=np.uint8)
    names[:, 0], names[:, 1] = ord("r"), ord("s")
    for d in range(9):
        names[:, 10 - d] = 48 + num % 10
        num //= 10
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(11)).astype(np.uint32)
    return pos, flags, off, np.concatenate([names.reshape(-1), np.zeros(1, dtype=np.uint8)])
