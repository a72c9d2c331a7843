"""TEST INFRASTRUCTURE: a minimal BAM + BAI writer (the image has no samtools / pysam) so that the BGZF/BAM/BAI
reader in vapor_b200/seqio.py can be checked against the SAM text of the same records."""
import struct
import zlib

_SEQ_ENC = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
_CIG_OPS = "MIDNSHP=X"


def _bgzf_block(data: bytes) -> bytes:
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    cdata = comp.compress(data) + comp.flush()
    bsize = len(cdata) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + cdata +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14: return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17: return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20: return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23: return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26: return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def _parse_cigar(cigar):
    import re
    return [(int(n), _CIG_OPS.index(op)) for n, op in re.findall(r"(\d+)([MIDNSHP=X])", cigar)]


def write_bam(path, refs, records, block_records=40, with_bai=True):
    """refs: [(name, length)]; records: [(qname, tid, pos1, cigar, seq)] sorted by (tid, pos)."""
    header_text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)
    hdr = b"BAM\x01" + struct.pack("<i", len(header_text)) + header_text.encode() + struct.pack("<i", len(refs))
    for n, l in refs:
        hdr += struct.pack("<i", len(n) + 1) + n.encode() + b"\x00" + struct.pack("<i", l)
    blocks = [hdr]                        # uncompressed payload of each BGZF block
    rec_loc = []                          # (block index, offset in block, length) per record
    cur = b""
    for i, (qname, tid, pos1, cigar, seq) in enumerate(records):
        cig = _parse_cigar(cigar)
        span = sum(n for n, op in cig if op in (0, 2, 3, 7, 8)) or 1
        pos0 = pos1 - 1
        b = reg2bin(pos0, pos0 + span)
        sq = bytearray((len(seq) + 1) // 2)
        for j, c in enumerate(seq):
            sq[j // 2] |= _SEQ_ENC.get(c, 15) << (4 if j % 2 == 0 else 0)
        aux = b""
        cig_field = cig
        if len(cig) > 65535:              # SAM spec 4.2.2: <l_seq>S<ref_len>N in the field, the real CIGAR in CG:B,I
            cig_field = [(len(seq), 4), (span, 3)]
            aux = b"CGBI" + struct.pack("<I", len(cig)) + b"".join(struct.pack("<I", (n << 4) | op) for n, op in cig)
        body = struct.pack("<iiBBHHHiiii", tid, pos0, len(qname) + 1, 60, b, len(cig_field), 0, len(seq), -1, -1, 0)
        body += (qname.encode() + b"\x00" + b"".join(struct.pack("<I", (n << 4) | op) for n, op in cig_field) + bytes(sq) +
                 b"\xff" * len(seq) + b"NMC\x03" + aux)
        rec = struct.pack("<i", len(body)) + body
        if len(cur) and (i % block_records == 0 or len(cur) + len(rec) > 60000):
            blocks.append(cur); cur = b""
        rec_loc.append((len(blocks), len(cur), len(rec), tid, pos0, pos0 + span, b))
        cur += rec
    if cur:
        blocks.append(cur)
    coff, offs = 0, []
    with open(path, "wb") as f:
        for blk in blocks:
            offs.append(coff)
            z = _bgzf_block(blk)
            f.write(z); coff += len(z)
        f.write(_bgzf_block(b""))
    offs.append(coff)
    if not with_bai:
        return
    # BAI: per reference, bins -> chunks (one chunk per record, merged when adjacent) and the 16 kb linear index
    n_ref = len(refs)
    bins = [dict() for _ in range(n_ref)]
    lin = [dict() for _ in range(n_ref)]
    for (bi, off, ln, tid, beg, end, b) in rec_loc:
        vbeg = (offs[bi] << 16) | off
        vend = (offs[bi] << 16) | (off + ln) if off + ln < len(blocks[bi]) else (offs[bi + 1] << 16)
        ch = bins[tid].setdefault(b, [])
        if ch and ch[-1][1] == vbeg:
            ch[-1] = (ch[-1][0], vend)
        else:
            ch.append((vbeg, vend))
        for w in range(beg >> 14, ((end - 1) >> 14) + 1):
            if w not in lin[tid] or vbeg < lin[tid][w]:
                lin[tid][w] = vbeg
    with open(path + ".bai", "wb") as f:
        f.write(b"BAI\x01" + struct.pack("<i", n_ref))
        for t in range(n_ref):
            f.write(struct.pack("<i", len(bins[t])))
            for b, chunks in bins[t].items():
                f.write(struct.pack("<Ii", b, len(chunks)))
                for c in chunks:
                    f.write(struct.pack("<QQ", *c))
            n_intv = (max(lin[t]) + 1) if lin[t] else 0
            f.write(struct.pack("<i", n_intv))
            last = 0
            for w in range(n_intv):
                last = lin[t].get(w, last)
                f.write(struct.pack("<Q", last))
