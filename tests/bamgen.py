"""Synthetic raw BAM alignment records for the reader-side parity tests (numpy, CPU only).

`records_from_block` turns the flat template records of tests/blockgen.py into the byte stream that follows the
header in an uncompressed BAM file (int32 block_size, 32 bytes of fixed fields, qname, cigar, 4-bit seq, qual, aux),
coordinate sorted, with the bisulfite strand written as the tag of one of the aligners the reference recognises
(src/input_sam.c:150-198).  Options add what the reference's reader has to cope with: positional duplicates, records
that the flag / MAPQ / insert-size / orientation filters drop (src/input_sam.c:234-300), quality values above the
clamp (MAX_QUAL = 43), and unrelated tags of every type in front of the strand tag.
"""
import struct

import numpy as np

INS, DEL, SOFT = 1, 2, 3
CIGAR_M, CIGAR_I, CIGAR_D, CIGAR_S, CIGAR_H = 0, 1, 2, 4, 5

FPAIRED, FPROPER, FUNMAP, FMUNMAP, FREVERSE, FMREVERSE, FREAD1, FREAD2 = 1, 2, 4, 8, 16, 32, 64, 128
FSECONDARY, FQCFAIL, FDUP, FSUPP = 256, 512, 1024, 2048


def _cigar(read_len, events):
    """events (type, read position, size) in read order -> list of (op, len)"""
    ops = []
    pos = 0
    for ty, p, sz in events:
        if p > pos:
            ops.append((CIGAR_M, p - pos))
            pos = p
        if ty == SOFT:
            ops.append((CIGAR_S, sz))
            pos += sz
        elif ty == DEL:          # CIGAR I: bases in the read that the reference lacks
            ops.append((CIGAR_I, sz))
            pos += sz
        elif ty == INS:          # CIGAR D: reference bases the read lacks
            ops.append((CIGAR_D, sz))
    if read_len > pos:
        ops.append((CIGAR_M, read_len - pos))
    return ops


def _strand_tag(rng, strand, flavour):
    """aux bytes naming the bisulfite strand the way aligner `flavour` does; strand 0 -> no tag"""
    if strand == 0:
        return b""
    c2t = strand == 1
    if flavour == 0:      # GEM
        return b"XBA" + (b"C" if c2t else b"G")
    if flavour == 1:      # Bowtie / Bismark
        return b"XGZ" + (b"CT" if c2t else b"GA") + b"\0"
    if flavour == 2:      # Novoalign
        return b"ZBZ" + (b"CT" if c2t else b"GA") + b"\0"
    if flavour == 3:      # BSMAP
        return b"ZSZ" + (b"+" if c2t else b"-") + (b"+" if rng.random() < 0.5 else b"-") + b"\0"
    return b"YDZ" + (b"f" if c2t else b"r") + b"\0"      # bwa-meth


def _other_tags(rng):
    out = b""
    if rng.random() < 0.7:
        out += b"NMi" + struct.pack("<i", int(rng.integers(0, 9)))
    if rng.random() < 0.5:
        out += b"MDZ" + bytes(rng.integers(48, 58, size=int(rng.integers(1, 12))).astype(np.uint8)) + b"\0"
    if rng.random() < 0.3:
        out += b"ASC" + bytes([int(rng.integers(0, 200))])
    if rng.random() < 0.3:
        out += b"XSs" + struct.pack("<h", int(rng.integers(-300, 300)))
    if rng.random() < 0.2:
        out += b"ZPf" + struct.pack("<f", float(rng.random()))
    if rng.random() < 0.15:
        n = int(rng.integers(0, 5))
        out += b"ZQBS" + struct.pack("<I", n) + bytes(rng.integers(0, 255, size=2 * n).astype(np.uint8))
    if rng.random() < 0.1:
        out += b"ZHH" + b"1AE3" + b"\0"
    return out


def exotic_cigar(rng, cigar):
    """the same alignment written with the operators aligners rarely use: = / X for M, hard clips at the ends, and now
    and then an N (skipped region: ignored by the reference) or a P (padding: treated like a soft clip) in the middle"""
    out = []
    for op, ln in cigar:
        if op == CIGAR_M and ln > 3 and rng.random() < 0.5:
            a = int(rng.integers(1, ln - 1))
            out += [(7, a), (8, 1), (7, ln - a - 1)] if ln - a - 1 > 0 else [(7, a), (8, ln - a)]
        else:
            out.append((op, ln))
        if rng.random() < 0.05:
            out.append((3, int(rng.integers(1, 50))))
        if rng.random() < 0.03:
            out.append((6, int(rng.integers(1, 4))))
    if rng.random() < 0.3:
        out.insert(0, (CIGAR_H, int(rng.integers(1, 30))))
    if rng.random() < 0.3:
        out.append((CIGAR_H, int(rng.integers(1, 30))))
    if rng.random() < 0.02:
        out.append((9, 3))          # 'B': no case in the reference's switch
    return out


def exotic_stream(seed, n=600):
    """records for the decode-only tests: unusual CIGAR operators, every flag combination, empty tag lists"""
    rng = np.random.default_rng(seed)
    recs = []
    pos = 10
    for i in range(n):
        pos += int(rng.integers(0, 30))
        l = int(rng.integers(1, 90))
        packed = (rng.integers(0, 4, size=l) | (rng.integers(1, 44, size=l) << 2)).astype(np.uint8)
        packed[rng.random(l) < 0.05] = 0
        quals = np.where(rng.random(l) < 0.1, rng.integers(44, 255, size=l), packed >> 2)
        cig = exotic_cigar(rng, [(CIGAR_S, 2), (CIGAR_M, max(l - 2, 1))] if l > 4 and rng.random() < 0.3 else [(CIGAR_M, l)])
        flag = int(rng.integers(0, 4096)) if rng.random() < 0.5 else int(rng.choice([0, 16, 99, 147, 83, 163, 65, 129, 73, 137]))
        mtid = 0 if rng.random() < 0.9 else 1
        mpos = pos + int(rng.integers(-300, 300))
        tl = int(rng.integers(-1500, 1500))
        aux = b"" if rng.random() < 0.2 else _other_tags(rng) + _strand_tag(rng, int(rng.integers(0, 3)), int(rng.integers(0, 5))) + _other_tags(rng)
        if rng.random() < 0.05:
            aux += b"Q?x"             # a tag type the reference's switch has no case for
        recs.append(pack_record(0, pos, int(rng.integers(0, 61)), flag, mtid, max(mpos, 0), tl, ("x%d" % i).encode(), cig, packed, quals, aux))
    return np.frombuffer(b"".join(recs), dtype=np.uint8).copy(), n


def pack_record(tid, pos0, mapq, flag, mtid, mpos0, tlen, qname, cigar, packed, quals, aux):
    """one BAM alignment record (with its leading block_size)"""
    l_seq = len(packed)
    name = qname + b"\0"
    nib = np.where(packed == 0, 15, (1 << (packed & 3))).astype(np.uint8)
    if l_seq & 1:
        nib = np.concatenate([nib, [0]]).astype(np.uint8)
    seq = ((nib[0::2] << 4) | nib[1::2]).astype(np.uint8).tobytes()
    cig = b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in cigar)
    body = struct.pack("<iiBBHHHiiii", tid, pos0, len(name), mapq, 4680, len(cigar), flag, l_seq, mtid, mpos0, tlen)
    body += name + cig + seq + bytes(quals.astype(np.uint8)) + aux
    return struct.pack("<i", len(body)) + body


def records_from_block(rng, T, B, M, tid=0, name_prefix="r", flavour=None, dup_frac=0.0, junk_frac=0.0,
                       qual_over=0.0, single_flag_paired=False):
    """-> list of (sort key, record bytes); the caller concatenates blocks / contigs and sorts."""
    recs = []
    serial = 0
    for i in range(len(T)):
        t = T[i]
        copies = 1 + (1 if rng.random() < dup_frac else 0) + (1 if rng.random() < dup_frac * 0.3 else 0)
        for cp in range(copies):
            qname = ("%s%d_%d_%d" % (name_prefix, tid, i, cp)).encode()
            fl = int(rng.integers(0, 5)) if flavour is None else flavour
            strand = int(t["bs_strand"])
            present = [bool(t["present"][k]) and int(t["read_len"][k]) > 0 for k in (0, 1)]
            paired = present[0] and present[1]
            fpos, rpos = int(t["forward_position"]), int(t["reverse_position"])
            tl = 0
            if paired:
                tl = rpos + int(t["reference_span"][1]) - fpos
            for k in (0, 1):
                if not present[k]:
                    continue
                packed = B[int(t["read_off"][k]):int(t["read_off"][k]) + int(t["read_len"][k])].copy()
                ev = [tuple(int(v) for v in M[int(t["mm_off"][k]) + z]) for z in range(int(t["mm_n"][k]))]
                quals = (packed >> 2).astype(np.int64)
                if cp:          # a duplicate: same coordinates, its own qualities and MAPQ
                    quals = np.where(packed == 0, quals, rng.integers(5, 44, size=len(packed)))
                    packed = np.where(packed == 0, 0, (packed & 3) | (quals << 2)).astype(np.uint8)
                over = rng.random(len(packed)) < qual_over
                quals = np.where(over, rng.integers(44, 94, size=len(packed)), quals)
                quals = np.where(packed == 0, rng.integers(0, 41, size=len(packed)), quals)
                mapq = int(t["mapq"][k]) if cp == 0 else int(rng.integers(15, 61))
                pos = fpos if k == 0 else rpos
                mpos = (rpos if k == 0 else fpos) if paired else 0
                flag = FREVERSE if k == 1 else 0
                if paired:
                    flag |= FPAIRED | FPROPER | (FMREVERSE if k == 0 else 0)
                    first = (k == 0) == (int(t["orientation"]) == 0)      # orientation FORWARD: the forward read is read 1
                    flag |= FREAD1 if first else FREAD2
                elif single_flag_paired:
                    flag |= FPAIRED | FMUNMAP | FREAD1
                aux = _other_tags(rng) + _strand_tag(rng, strand, fl) + (_other_tags(rng) if rng.random() < 0.3 else b"")
                rec = pack_record(tid, pos - 1, mapq, flag, tid if paired else -1, mpos - 1 if paired else -1,
                                  (tl if k == 0 else -tl), qname, _cigar(len(packed), ev), packed, quals, aux)
                recs.append(((tid, pos - 1, serial), rec))
                serial += 1
        # records the filters must drop
        if rng.random() < junk_frac:
            k = int(rng.integers(0, 2))
            if not t["present"][k] or not t["read_len"][k]:
                continue
            packed = B[int(t["read_off"][k]):int(t["read_off"][k]) + int(t["read_len"][k])].copy()
            quals = (packed >> 2).astype(np.int64)
            pos = int(t["forward_position"]) if k == 0 else int(t["reverse_position"])
            kind = int(rng.integers(0, 8))
            flag = (FREVERSE if k else 0) | FPAIRED | FPROPER | FREAD1 | (0 if k else FMREVERSE)
            mapq, mtid, mpos, tl = 60, tid, pos + 50, 200
            if k:
                mpos, tl = max(pos - 50, 1), -200
            if kind == 0:
                flag |= FSECONDARY
            elif kind == 1:
                flag |= FQCFAIL
            elif kind == 2:
                flag |= FDUP
            elif kind == 3:
                flag |= FSUPP
            elif kind == 4:
                mapq = int(rng.integers(0, 20))
            elif kind == 5:
                tl = 5000 if not k else -5000
            elif kind == 6:
                mtid = tid + 1
            else:             # wrong orientation: forward read to the right of its mate
                mpos = max(pos - 40, 1) if not k else pos + 40
            qname = ("junk%d_%d_%d" % (tid, i, kind)).encode()
            rec = pack_record(tid, pos - 1, mapq, flag, mtid, mpos - 1, tl, qname, [(CIGAR_M, len(packed))], packed, quals,
                              _strand_tag(rng, int(t["bs_strand"]), 0))
            recs.append(((tid, pos - 1, serial), rec))
            serial += 1
    return recs


def concat_sorted(rec_lists):
    """coordinate sort (stable) and concatenate -> (bytes as uint8 array, number of records)"""
    allr = [r for lst in rec_lists for r in lst]
    # stable on (tid, pos); the serial only keeps mates of equal position in generation order within one list
    order = sorted(range(len(allr)), key=lambda i: (allr[i][0][0], allr[i][0][1]))
    buf = b"".join(allr[i][1] for i in order)
    return np.frombuffer(buf, dtype=np.uint8).copy(), len(allr)


def make_stream(seed, n_contigs=None, dup=0.15, junk=0.1, depth=None, read_len=None, paired=None, contig_len=None):
    """A small multi-contig, multi-block coordinate-sorted record stream.
    -> (bam uint8[], number of records, target_len[], [reference codes per contig])"""
    from tests import blockgen
    rng = np.random.default_rng(seed)
    nctg = int(rng.integers(1, 4)) if n_contigs is None else n_contigs
    lists, refs, tl = [], [], []
    for tid in range(nctg):
        L = int(rng.integers(3000, 9000)) if contig_len is None else contig_len
        ref = blockgen.random_reference(rng, L, n_runs=2)
        refs.append(ref)
        tl.append(L)
        a = 50
        while a < L - 900:
            b = min(a + int(rng.integers(300, 2500)), L - 600)
            T, B, M, _ = blockgen.make_block(
                rng, ref, a, b, depth=int(rng.integers(4, 40)) if depth is None else depth,
                read_len=int(rng.integers(30, 120)) if read_len is None else read_len,
                paired=(rng.random() < 0.75) if paired is None else paired, indel_frac=0.1, clip_frac=0.1,
                single_mate_frac=0.15, frag_mean=int(rng.integers(60, 300)), nonconv_frac=0.05)
            lists.append(records_from_block(rng, T, B, M, tid=tid, name_prefix="s%d_%d_" % (seed, a), dup_frac=dup,
                                            junk_frac=junk, qual_over=0.02, single_flag_paired=rng.random() < 0.2))
            a = b + int(rng.integers(1, 400))
    bam, n = concat_sorted(lists)
    return bam, n, np.array(tl, dtype=np.uint32), refs


def template_keys(tm, bases, misms):
    """canonical, layout-independent view of a template list: positions, orientation, strand and per mate
    (length, span, mapq, packed bytes, events); a mate without bytes only keeps its mapq"""
    out = []
    for t in tm:
        mates = []
        for k in (0, 1):
            rl = int(t["read_len"][k]) if t["present"][k] else 0
            if rl == 0:
                mates.append((0, int(t["mapq"][k])))
                continue
            o = int(t["read_off"][k])
            ev = tuple(tuple(int(v) for v in misms[int(t["mm_off"][k]) + z]) for z in range(int(t["mm_n"][k])))
            mates.append((rl, int(t["reference_span"][k]), int(t["mapq"][k]), bytes(bases[o:o + rl]), ev))
        out.append((int(t["forward_position"]), int(t["reverse_position"]), int(t["orientation"]), int(t["bs_strand"]), tuple(mates)))
    return out
