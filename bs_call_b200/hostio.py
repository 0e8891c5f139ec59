"""Files for whole-program runs of bs_call (tests and the `full_binary` leg of bench.py): a BAM file from a stream of raw
alignment records, a FASTA file with its .fai index from per-contig reference codes, and a reader for the BCF files the
binaries write.  Formats: SAM specification sections 4.1 (BGZF) and 4.2 (BAM), VCF/BCF specification v4.3 section 6.3.
Pure host-side helpers; nothing here is on the compute path."""
import struct
import zlib

import numpy as np

_BASES = np.frombuffer(b"NACGT", dtype=np.uint8)
_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _bgzf_block(data, level):
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    head = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25)
    return head + comp + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


def write_bgzf(path, payload, level=1, block=0xff00):
    """payload: bytes-like (or a list of them, written back to back)"""
    parts = payload if isinstance(payload, (list, tuple)) else [payload]
    with open(path, "wb") as f:
        pend = b""
        for p in parts:
            mv = memoryview(p).cast("B")
            pos = 0
            if pend:
                take = min(block - len(pend), len(mv))
                pend += bytes(mv[:take])
                pos = take
                if len(pend) == block:
                    f.write(_bgzf_block(pend, level))
                    pend = b""
            while len(mv) - pos >= block:
                f.write(_bgzf_block(mv[pos:pos + block], level))
                pos += block
            if pos < len(mv):
                pend += bytes(mv[pos:])
        if pend:
            f.write(_bgzf_block(pend, level))
        f.write(_BGZF_EOF)


def bam_header(names, lengths, extra_text=""):
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, int(l)) for n, l in zip(names, lengths)) + extra_text
    tb = text.encode()
    out = [b"BAM\x01", struct.pack("<i", len(tb)), tb, struct.pack("<i", len(names))]
    for n, l in zip(names, lengths):
        nb = n.encode() + b"\0"
        out += [struct.pack("<i", len(nb)), nb, struct.pack("<i", int(l))]
    return b"".join(out)


def write_bam(path, names, lengths, records, level=1, extra_text=""):
    """records: the byte stream that follows the header in an uncompressed BAM file (uint8 array or bytes)"""
    write_bgzf(path, [bam_header(names, lengths, extra_text), np.ascontiguousarray(records).tobytes() if isinstance(records, np.ndarray) else records], level)


def write_fasta(path, names, refs, width=60):
    """refs: per contig codes 0 = N, 1..4 = A C G T for positions 1..len; writes path and path.fai"""
    off = 0
    fai = []
    with open(path, "wb") as f:
        for n, r in zip(names, refs):
            hdr = (">%s\n" % n).encode()
            f.write(hdr)
            off += len(hdr)
            seq = _BASES[np.asarray(r, dtype=np.uint8)]
            L = len(seq)
            full, rest = divmod(L, width)
            body = np.empty(L + full + (1 if rest else 0), dtype=np.uint8)
            if full:
                blk = body[:full * (width + 1)].reshape(full, width + 1)
                blk[:, :width] = seq[:full * width].reshape(full, width)
                blk[:, width] = 10
            if rest:
                body[full * (width + 1):-1] = seq[full * width:]
                body[-1] = 10
            f.write(body.tobytes())
            fai.append("%s\t%d\t%d\t%d\t%d\n" % (n, L, off, width, width + 1))
            off += len(body)
    with open(path + ".fai", "w") as f:
        f.write("".join(fai))


def read_bgzf(path):
    """whole uncompressed payload of a BGZF (or plain) file"""
    raw = open(path, "rb").read()
    if raw[:2] != b"\x1f\x8b":
        return raw
    out = []
    pos = 0
    while pos < len(raw):
        xlen = struct.unpack_from("<H", raw, pos + 10)[0]
        bsize = None
        q = pos + 12
        while q < pos + 12 + xlen:
            si1, si2, sl = raw[q], raw[q + 1], struct.unpack_from("<H", raw, q + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", raw, q + 4)[0]
            q += 4 + sl
        cdata = raw[pos + 12 + xlen:pos + bsize + 1 - 8]
        out.append(zlib.decompress(cdata, -15))
        pos += bsize + 1
    return b"".join(out)


def read_bcf(path):
    """-> (header text, record bytes as uint8 array): the records as they lie in the file (two length words, 24 fixed
    bytes, shared, indiv each), which is also the layout of the device writer's output"""
    data = read_bgzf(path)
    assert data[:5] == b"BCF\x02\x02", "not a BCF2.2 file"
    l_text = struct.unpack_from("<I", data, 5)[0]
    text = data[9:9 + l_text].rstrip(b"\0").decode()
    return text, np.frombuffer(data, dtype=np.uint8, offset=9 + l_text).copy()


def write_dbsnp_index(path, contigs, prefixes=("rs",), header="name=dbSNP_synthetic", bins_per_block=4096):
    """A dbSNP index file in the format src/dbSNP.c:27-304 reads (written by the reference's dbSNP_idx tool,
    src/dbSNP_output.c:139-300): magic, offsets, per contig a chain of zlib-compressed blocks of 64-position bins, and a
    compressed directory at the end.  contigs: {name: [(position (1-based), prefix index, digits (str of decimal digits),
    always_written (bool)), ...]}; at most 64 entries fall into a bin by construction (one per position)."""
    magic = 0xd7278434
    body = bytearray()
    body += struct.pack("<II", magic, 0)
    body += b"\0" * 24                                # directory offset, buffer size, compressed directory size: patched below
    directory = []
    bufsize = 4096
    for name, ents in contigs.items():
        ents = sorted(ents, key=lambda e: e[0])
        if not ents:
            continue
        bins = {}
        for pos, pfx, digits, always in ents:
            bins.setdefault(pos >> 6, []).append((pos & 63, pfx, digits, always))
        order = sorted(bins)
        min_bin, max_bin = order[0], order[-1]
        offset = len(body)
        curr = min_bin
        for s0 in range(0, len(order), bins_per_block):
            blk = bytearray()
            for b in order[s0:s0 + bins_per_block]:
                inc = b - curr
                curr = b
                if inc < 64:
                    blk.append(inc << 2)
                elif inc < 256:
                    blk += bytes([1, inc])
                elif inc < 65536:
                    blk += bytes([2]) + struct.pack("<H", inc)
                else:
                    blk += bytes([3]) + struct.pack("<I", inc)
                items = bins[b]
                for k, (ix, pfx, digits, always) in enumerate(items):
                    if pfx < 3:
                        blk.append(((pfx + 1) << 6) | ix)              # prefixes 0..2 ride in the entry byte
                    else:
                        blk.append(ix)
                        blk += struct.pack(">H", pfx)
                    d = [int(c) for c in digits]
                    for i in range(0, len(d) - 1, 2):
                        blk.append(0x21 + 10 * d[i] + d[i + 1])
                    if len(d) & 1:
                        blk.append(0x85 + d[-1])
                    blk.append((2 if always else 0) | (1 if k == len(items) - 1 else 0))
            bufsize = max(bufsize, len(blk) + 64)
            comp = zlib.compress(bytes(blk))
            body += struct.pack("<Q", len(comp)) + comp
        body += struct.pack("<Q", 0)
        directory.append((min_bin, max_bin, offset, name))
    d = bytearray(b"\0\0") + struct.pack("<HI", len(prefixes), len(directory))
    d += b"track " + header.encode() + b"\0"
    for p in prefixes:
        d += p.encode() + b"\0"
    for mn, mx, off, name in directory:
        d += struct.pack("<IIQ", mn, mx, off) + name.encode() + b"\0"
    d += b"\0" * 8                                    # the reader wants bytes behind the last name (src/dbSNP.c:113)
    bufsize = max(bufsize, len(d) + 64)
    comp = zlib.compress(bytes(d))
    struct.pack_into("<QQQ", body, 8, len(body), bufsize, len(comp))
    body += comp + struct.pack("<I", magic)
    with open(path, "wb") as f:
        f.write(bytes(body))
