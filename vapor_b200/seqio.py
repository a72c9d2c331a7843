"""In-process sequence access for the host side of the scoring path.

The reference shells out twice or more per SV: ``samtools faidx ref chr:a-b`` inside
``ref_seq_readin`` (vapor_vali/Simple_function.pyx:1203-1217) and ``samtools view bam chr:a-b`` inside
``chop_pacbio_read_by_pos`` (:339-354).  With the scoring on the GPU those subprocesses dominate the
wall time, and this image has no samtools at all, so the two queries are answered in-process:

* ``FastaFile.fetch(chrom, start, end)`` -- random access through the ``.fai`` index, same coordinates
  (1-based, inclusive) and the same clipping as ``samtools faidx``;
* ``AlignmentFile.fetch(chrom, start, end)`` -- the records ``samtools view file chrom:start-end`` prints
  (every record whose alignment overlaps the region, in file order; no FLAG or MAPQ filter, as in the
  reference), from SAM text, or from BAM through BGZF + the ``.bai`` index (linear scan without one).

Only the fields the reference touches are kept: QNAME, POS, CIGAR, SEQ (``pbam[0]``, ``pbam[3]``, ``pbam[5]``,
``pbam[9]``).  When a real ``samtools`` is on PATH and ``VAPOR_SAMTOOLS=1`` is set, the subprocess route of the
reference is used instead (identical strings, only slower).
"""
from __future__ import annotations

import bisect
import gzip
import os
import re
import shutil
import struct
import subprocess
import threading
import zlib
from typing import Dict, Iterator, List, NamedTuple, Optional, Tuple


def use_samtools() -> bool:
    return os.environ.get("VAPOR_SAMTOOLS", "0") == "1" and shutil.which("samtools") is not None


# ----------------------------------------------------------------------------------------------------
# FASTA
# ----------------------------------------------------------------------------------------------------
class FastaFile:
    """``samtools faidx`` without the subprocess: needs ``<path>.fai`` (built on the fly when missing)."""

    def __init__(self, path: str):
        self.path = path
        self.index: Dict[str, Tuple[int, int, int, int]] = {}     # name -> (length, offset, line_bases, line_width)
        self.order: List[str] = []
        fai = path + ".fai"
        if not os.path.exists(fai):
            build_fai(path)
        with open(fai) as f:
            for line in f:
                p = line.rstrip("\n").split("\t")
                if len(p) < 5:
                    p = line.split()
                if len(p) >= 5:
                    self.index[p[0]] = (int(p[1]), int(p[2]), int(p[3]), int(p[4]))
                    self.order.append(p[0])
        self._fd = os.open(path, os.O_RDONLY)           # positional reads (os.pread): safe from several threads

    def close(self):
        os.close(self._fd)

    def fetch(self, chrom: str, start: int, end: int) -> str:
        """Bases ``start..end`` (1-based, inclusive) of ``chrom``; clipped to the contig like samtools
        (start < 1 -> 1, end > length -> length); unknown contig or empty interval -> ``''``."""
        ent = self.index.get(chrom)
        if ent is None:
            return ""
        length, offset, lb, lw = ent
        start = max(int(start), 1)
        end = min(int(end), length)
        if end < start:
            return ""
        s0, e0 = start - 1, end                                   # 0-based half open
        b0 = offset + (s0 // lb) * lw + s0 % lb
        b1 = offset + ((e0 - 1) // lb) * lw + (e0 - 1) % lb + 1
        raw = os.pread(self._fd, b1 - b0, b0)
        return raw.replace(b"\n", b"").replace(b"\r", b"").decode("latin-1")


def build_fai(path: str) -> str:
    """Write ``<path>.fai`` (same five columns as ``samtools faidx``)."""
    out = []
    with open(path, "rb") as f:
        name, length, offset, lb, lw = None, 0, 0, 0, 0
        pos = 0
        for line in f:
            if line.startswith(b">"):
                if name is not None:
                    out.append((name, length, offset, lb, lw))
                name = line[1:].split()[0].decode("latin-1")
                length, lb, lw = 0, 0, 0
                offset = pos + len(line)
            else:
                bases = len(line.rstrip(b"\r\n"))
                if lb == 0 and bases:
                    lb, lw = bases, len(line)
                length += bases
            pos += len(line)
        if name is not None:
            out.append((name, length, offset, lb, lw))
    with open(path + ".fai", "w") as f:
        for r in out:
            f.write("\t".join(str(v) for v in r) + "\n")
    return path + ".fai"


# ----------------------------------------------------------------------------------------------------
# alignments
# ----------------------------------------------------------------------------------------------------
class AlnRecord(NamedTuple):
    qname: str
    pos: int          # 1-based leftmost position (SAM POS)
    cigar: str
    seq: str


_CIG = re.compile(r"(\d+)([MIDNSHP=X])")


def _ref_span(cigar: str) -> int:
    """Reference bases an alignment covers (M, D, N, =, X), at least 1 (samtools treats an empty CIGAR as length 1)."""
    n = 0
    for m in _CIG.finditer(cigar):
        if m.group(2) in "MDN=X":
            n += int(m.group(1))
    return n if n > 0 else 1


class _SamText:
    """Whole-file index of a (possibly gzipped) SAM text file: per contig, records in file order."""

    def __init__(self, path: str):
        self.by_chrom: Dict[str, List[Tuple[int, int, AlnRecord]]] = {}
        self.max_span: Dict[str, int] = {}
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "rt") as f:
            for line in f:
                if line.startswith("@"):
                    continue
                p = line.rstrip("\n").split("\t")
                if len(p) < 10:
                    p = line.split()
                    if len(p) < 10:
                        continue
                if p[2] == "*":
                    continue
                pos = int(p[3])
                end = pos + _ref_span(p[5]) - 1
                self.by_chrom.setdefault(p[2], []).append((pos, end, AlnRecord(p[0], pos, p[5], p[9])))
        self.sorted: Dict[str, bool] = {}
        self.starts: Dict[str, List[int]] = {}
        for c, recs in self.by_chrom.items():
            st = [r[0] for r in recs]
            self.sorted[c] = all(st[i] <= st[i + 1] for i in range(len(st) - 1))
            self.starts[c] = st
            self.max_span[c] = max((r[1] - r[0] + 1 for r in recs), default=1)

    def fetch(self, chrom: str, start: int, end: int) -> Iterator[AlnRecord]:
        recs = self.by_chrom.get(chrom)
        if not recs:
            return
        if self.sorted[chrom]:
            st = self.starts[chrom]
            lo = bisect.bisect_left(st, start - self.max_span[chrom] + 1)
            hi = bisect.bisect_right(st, end)
            it = recs[lo:hi]
        else:
            it = recs
        for pos, rend, rec in it:
            if pos <= end and rend >= start:
                yield rec


_SEQ_DEC = "=ACMGRSVTWYHKDBN"
_CIG_OPS = "MIDNSHP=XB"


class _BgzfReader:
    """Random access into a BGZF file by virtual offset (coffset << 16 | uoffset)."""

    def __init__(self, path: str):
        self.fh = open(path, "rb")
        self.block_start = -1
        self.block_len = 0
        self.data = b""
        self.upos = 0

    def close(self):
        self.fh.close()

    def _load(self, coffset: int) -> bool:
        self.fh.seek(coffset)
        hdr = self.fh.read(18)
        if len(hdr) < 18:
            self.data = b""; self.block_start = coffset; self.block_len = 0
            return False
        if hdr[:4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF block")
        xlen = struct.unpack("<H", hdr[10:12])[0]
        extra = hdr[12:18] + self.fh.read(xlen - 6)
        bsize = None
        i = 0
        while i + 4 <= len(extra):
            si1, si2, slen = extra[i], extra[i + 1], struct.unpack("<H", extra[i + 2:i + 4])[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack("<H", extra[i + 4:i + 6])[0]
            i += 4 + slen
        if bsize is None:
            raise ValueError("BGZF block without BC field")
        cdata = self.fh.read(bsize - xlen - 19)
        self.fh.read(8)
        self.data = zlib.decompress(cdata, -15)
        self.block_start = coffset
        self.block_len = bsize + 1
        return True

    def seek(self, voffset: int):
        co, uo = voffset >> 16, voffset & 0xFFFF
        if co != self.block_start:
            self._load(co)
        self.upos = uo

    def tell(self) -> int:
        return (self.block_start << 16) | self.upos

    def read(self, n: int) -> bytes:
        out = []
        while n > 0:
            if self.upos >= len(self.data):
                if not self._load(self.block_start + self.block_len):
                    break
                self.upos = 0
                continue                   # an empty block (the EOF marker) just falls through to the next load
            take = self.data[self.upos:self.upos + n]
            out.append(take)
            self.upos += len(take)
            n -= len(take)
        return b"".join(out)


_AUX_SIZE = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}


def _cg_tag(rec: bytes, p: int):
    """The uint32 array of the CG:B,I auxiliary field of a BAM record (aux fields start at ``p``), or None."""
    n = len(rec)
    while p + 3 <= n:
        tag, ty = rec[p:p + 2], chr(rec[p + 2])
        p += 3
        if ty in _AUX_SIZE:
            p += _AUX_SIZE[ty]
        elif ty in "ZH":
            while p < n and rec[p]:
                p += 1
            p += 1
        elif ty == "B":
            if p + 5 > n:
                return None
            sub, cnt = chr(rec[p]), struct.unpack_from("<I", rec, p + 1)[0]
            es = 1 if sub in "cC" else (2 if sub in "sS" else 4)
            if tag == b"CG" and sub == "I" and p + 5 + 4 * cnt <= n:
                return struct.unpack_from(f"<{cnt}I", rec, p + 5)
            p += 5 + es * cnt
        else:
            return None
    return None


def _reg2bins(beg: int, end: int) -> List[int]:
    end -= 1
    bins = [0]
    for shift, base in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        bins.extend(range(base + (beg >> shift), base + (end >> shift) + 1))
    return bins


class _Bam:
    def __init__(self, path: str):
        self.path = path
        self.bg = _BgzfReader(path)
        self.bg.seek(0)
        if self.bg.read(4) != b"BAM\x01":
            raise ValueError(f"{path}: not a BAM file")
        l_text = struct.unpack("<i", self.bg.read(4))[0]
        self.bg.read(l_text)
        n_ref = struct.unpack("<i", self.bg.read(4))[0]
        self.refs: List[str] = []
        self.ref_len: List[int] = []
        for _ in range(n_ref):
            l_name = struct.unpack("<i", self.bg.read(4))[0]
            self.refs.append(self.bg.read(l_name)[:-1].decode("latin-1"))
            self.ref_len.append(struct.unpack("<i", self.bg.read(4))[0])
        self.first_rec = self.bg.tell()
        self.tid = {n: i for i, n in enumerate(self.refs)}
        self.bai = None
        for cand in (path + ".bai", os.path.splitext(path)[0] + ".bai"):
            if os.path.exists(cand):
                self.bai = self._read_bai(cand)
                break

    @staticmethod
    def _read_bai(path: str):
        with open(path, "rb") as f:
            raw = f.read()
        if raw[:4] != b"BAI\x01":
            raise ValueError(f"{path}: not a BAI index")
        p = 4
        n_ref = struct.unpack_from("<i", raw, p)[0]; p += 4
        out = []
        for _ in range(n_ref):
            n_bin = struct.unpack_from("<i", raw, p)[0]; p += 4
            bins = {}
            for _ in range(n_bin):
                b, n_chunk = struct.unpack_from("<Ii", raw, p); p += 8
                chunks = [struct.unpack_from("<QQ", raw, p + 16 * i) for i in range(n_chunk)]
                p += 16 * n_chunk
                bins[b] = chunks
            n_intv = struct.unpack_from("<i", raw, p)[0]; p += 4
            ioff = list(struct.unpack_from(f"<{n_intv}Q", raw, p)); p += 8 * n_intv
            out.append((bins, ioff))
        return out

    def _records_from(self, voffset: int) -> Iterator[Tuple[int, int, int, AlnRecord]]:
        self.bg.seek(voffset)
        while True:
            head = self.bg.read(4)
            if len(head) < 4:
                return
            bs = struct.unpack("<i", head)[0]
            rec = self.bg.read(bs)
            if len(rec) < bs:
                return
            tid, pos, l_rn, _mapq, _bin, n_cig, _flag, l_seq = struct.unpack_from("<iiBBHHHi", rec, 0)
            p = 32
            qname = rec[p:p + l_rn - 1].decode("latin-1"); p += l_rn
            cig = struct.unpack_from(f"<{n_cig}I", rec, p); p += 4 * n_cig
            span = 0
            parts = []
            for c in cig:
                ln, op = c >> 4, c & 15
                parts.append(f"{ln}{_CIG_OPS[op]}")
                if op in (0, 2, 3, 7, 8):
                    span += ln
            nb = (l_seq + 1) // 2
            sq = rec[p:p + nb]
            if n_cig == 2 and (cig[0] & 15) == 4 and (cig[0] >> 4) == l_seq and (cig[1] & 15) == 3:
                # more than 65535 operations: the field holds <l_seq>S<ref_len>N, the real CIGAR sits in the CG:B,I tag
                # (samtools view, which the reference shells out to, prints the real one)
                real = _cg_tag(rec, p + nb + l_seq)
                if real is not None:
                    span, parts = 0, []
                    for c in real:
                        ln, op = c >> 4, c & 15
                        parts.append(f"{ln}{_CIG_OPS[op]}")
                        if op in (0, 2, 3, 7, 8):
                            span += ln
            seq = "".join(_SEQ_DEC[b >> 4] + _SEQ_DEC[b & 15] for b in sq)[:l_seq]
            yield tid, pos + 1, pos + max(span, 1), AlnRecord(qname, pos + 1, "".join(parts) or "*", seq or "*")

    def fetch(self, chrom: str, start: int, end: int) -> Iterator[AlnRecord]:
        tid = self.tid.get(chrom)
        if tid is None:
            return
        beg0, end0 = max(start - 1, 0), max(end, 1)
        voff = self.first_rec
        if self.bai is not None and tid < len(self.bai):
            bins, ioff = self.bai[tid]
            lin = ioff[min(beg0 >> 14, len(ioff) - 1)] if ioff else 0
            cands = [c[0] for b in _reg2bins(beg0, end0) for c in bins.get(b, []) if c[1] > lin]
            if not cands:
                return
            voff = max(min(cands), lin) if lin else min(cands)
        for rtid, pos, rend, rec in self._records_from(voff):
            if rtid != tid:
                if rtid > tid or rtid < 0:
                    return
                continue
            if pos > end:
                return
            if rend >= start:
                yield rec


class AlignmentFile:
    """``samtools view <file> chrom:start-end`` without the subprocess."""

    def __init__(self, path: str):
        self.path = path
        low = path.lower()
        if low.endswith(".sam") or low.endswith(".sam.gz"):
            self._impl = _SamText(path)
        else:
            with open(path, "rb") as f:
                magic = f.read(4)
            self._impl = _Bam(path) if magic[:2] == b"\x1f\x8b" else _SamText(path)

    def fetch(self, chrom: str, start: int, end: int) -> Iterator[AlnRecord]:
        return self._impl.fetch(chrom, int(start), int(end))


_fasta_cache: Dict[str, FastaFile] = {}
_aln_cache: Dict[object, AlignmentFile] = {}


def fasta(path: str) -> FastaFile:
    if path not in _fasta_cache:
        _fasta_cache[path] = FastaFile(path)
    return _fasta_cache[path]


def alignments(path: str) -> AlignmentFile:
    """Cached reader.  SAM text is an immutable in-memory index shared by all threads; a BAM reader keeps a file
    position and a decompressed block, so every thread gets its own."""
    a = _aln_cache.get(path)
    if a is None:
        a = _aln_cache.setdefault(path, AlignmentFile(path))
    if isinstance(a._impl, _SamText):
        return a
    key = (path, threading.get_ident())
    if key not in _aln_cache:
        _aln_cache[key] = AlignmentFile(path)
    return _aln_cache[key]


# ----------------------------------------------------------------------------------------------------
# native route (csrc/hostio.cpp through _hostio): the default.  VAPOR_HOSTIO=python keeps everything above in charge
# (the pure-Python readers are what the native ones are tested against).
# ----------------------------------------------------------------------------------------------------
def native_enabled() -> bool:
    return os.environ.get("VAPOR_HOSTIO", "native") != "python" and not use_samtools()


_native_fa: Dict[str, object] = {}
_native_aln: Dict[str, object] = {}
_region_cache: Dict[tuple, str] = {}          # prefetched samtools-faidx answers
_reads_cache: Dict[tuple, list] = {}           # prefetched chop + minimize answers


def native_fasta(path: str):
    from . import _hostio
    f = _native_fa.get(path)
    if f is None:
        f = _native_fa[path] = _hostio.FastaIndex(path)
    return f


def native_aln(path: str):
    from . import _hostio
    a = _native_aln.get(path)
    if a is None:
        a = _native_aln[path] = _hostio.AlnFile(path)
    return a


def chop_reads(files, chrom: str, start: int, end: int, flank: int, max_reads: int = 20) -> list:
    """``[[read, miss_bp, qname], ...]`` of one window over ``files`` in order: chop_pacbio_read_by_pos per file
    (Simple_function.pyx:339-354), concatenated, then minimize_pacbio_read_list when ``max_reads`` > 0 (:1091-1102).
    Native route only (callers check ``native_enabled()``)."""
    key = (tuple(files), chrom, int(start), int(end), int(flank), int(max_reads))
    hit = _reads_cache.get(key)
    if hit is not None:
        return [list(x) for x in hit]
    from . import _hostio
    out, _ = _hostio.chop_many([native_aln(f) for f in files], [(chrom, int(start), int(end), int(flank))], max_reads=max_reads, threads=1)
    return out[0]


def prefetch(ref: str, regions, files, windows, max_reads: int = 20, threads: int = 0) -> None:
    """Answer many queries ahead of the drivers in two native calls on several host threads: ``regions`` =
    ``(chrom, start, end)`` of ``samtools faidx ref``, ``windows`` = ``(chrom, start, end, flank)`` of chop + minimize
    over ``files``.  The answers wait in caches that ``faidx`` / ``chop_reads`` consult first; ``clear_prefetch`` drops them."""
    if not native_enabled():
        return
    from . import _hostio
    regions = [r for r in dict.fromkeys((c, int(a), int(b)) for c, a, b in regions) if (ref,) + r not in _region_cache]
    if regions:
        for r, s in zip(regions, native_fasta(ref).fetch_many(regions, threads)):
            _region_cache[(ref,) + r] = s
    files = tuple(files)
    windows = [w for w in dict.fromkeys((c, int(a), int(b), int(f)) for c, a, b, f in windows) if (files,) + w + (max_reads,) not in _reads_cache]
    if windows and files:
        lists, _ = _hostio.chop_many([native_aln(f) for f in files], windows, max_reads=max_reads, threads=threads)
        for w, l in zip(windows, lists):
            _reads_cache[(files,) + w + (max_reads,)] = l


def clear_prefetch() -> None:
    _region_cache.clear()
    _reads_cache.clear()


def faidx(ref: str, chrom: str, start: int, end: int) -> str:
    """The sequence ``ref_seq_readin`` assembles from ``samtools faidx ref chrom:start-end``
    (Simple_function.pyx:1206-1213): header dropped, lines joined."""
    if use_samtools():
        out = subprocess.run(["samtools", "faidx", ref, f"{chrom}:{int(start)}-{int(end)}"], capture_output=True, text=True).stdout
        seq = ""
        for line in out.split("\n")[1:]:
            f = line.strip().split()
            if not f:
                break
            seq += f[0]
        return seq
    if native_enabled():
        hit = _region_cache.get((ref, chrom, int(start), int(end)))
        return hit if hit is not None else native_fasta(ref).fetch(chrom, int(start), int(end))
    return fasta(ref).fetch(chrom, start, end)


def view(bam: str, chrom: str, start: int, end: int) -> Iterator[AlnRecord]:
    """Records of ``samtools view bam chrom:start-end`` (Simple_function.pyx:340)."""
    if use_samtools():
        out = subprocess.run(["samtools", "view", bam, f"{chrom}:{int(start)}-{int(end)}"], capture_output=True, text=True).stdout
        for line in out.split("\n"):
            p = line.strip().split()
            if len(p) >= 10 and p[0] != "@":
                yield AlnRecord(p[0], int(p[3]), p[5], p[9])
        return
    yield from alignments(bam).fetch(chrom, start, end)
