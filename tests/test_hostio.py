"""Native region extraction (csrc/hostio.cpp through vapor_b200/_hostio.py) against
 (a) the pure-Python readers / chop rules of vapor_b200 (pinned to the reference by test_cli_host.py and test_seqio.py),
 (b) a BAM transcribed byte by byte from the example of the SAM specification (section 1.1), decoded here with Python's
     own gzip module and a struct parser written from the specification -- independent of tests/bam_writer.py,
 (c) BAMs with and without a .bai index, several files per window, several threads, CIGARs with more than 65535
     operations (CG:B,I tag)."""
import gzip
import os
import struct
import zlib

import numpy as np
import pytest

from vapor_b200 import Simple_function as SF
from vapor_b200 import _hostio as H
from vapor_b200 import seqio

import bam_writer

HERE = os.path.dirname(os.path.abspath(__file__))
CASE = os.path.join(HERE, "golden", "cli_case")
SAM = os.path.join(CASE, "reads.sam.gz")
REF = os.path.join(CASE, "ref.fa")


@pytest.fixture()
def python_io(monkeypatch):
    """Switch vapor_b200 to its pure-Python readers for the duration of a call."""
    def run(fn, *a, **k):
        monkeypatch.setenv("VAPOR_HOSTIO", "python")
        try:
            return fn(*a, **k)
        finally:
            monkeypatch.delenv("VAPOR_HOSTIO")
    return run


def _windows(rng, chrom, length, n):
    out = []
    for _ in range(n):
        s = int(rng.integers(1, length - 100)); f = int(rng.choice([50, 137, 333, 500]))
        out.append((chrom, s, s + int(rng.integers(2 * f - 20, 2 * f + 4000)), f))
    return out


def test_fasta_native_equals_python():
    pf = seqio.FastaFile(REF)
    nf = H.FastaIndex(REF)
    chrom = pf.order[0]; L = pf.index[chrom][0]
    rng = np.random.default_rng(1)
    regs = [(chrom, int(s), int(s + d)) for s, d in zip(rng.integers(-50, L, 300), rng.integers(-5, 3000, 300))]
    regs += [("nope", 1, 10), (chrom, L - 5, L + 100), (chrom, 1, L), (chrom, 0, 0)]
    exp = [pf.fetch(*r) for r in regs]
    assert [nf.fetch(*r) for r in regs] == exp
    assert nf.fetch_many(regs, threads=3) == exp and nf.fetch_many(regs, threads=1) == exp
    nf.close()


def test_fai_written_by_native_equals_python(tmp_path):
    rng = np.random.default_rng(3)
    for name, nl in (("a.fa", "\n"), ("b.fa", "\r\n")):
        p = tmp_path / name
        with open(p, "w", newline="") as f:
            for c, (n, w) in enumerate(((1234, 60), (77, 10), (4, 80))):
                f.write(f">c{c} desc{nl}")
                s = "".join(rng.choice(list("ACGTNacgt"), size=n))
                for i in range(0, n, w):
                    f.write(s[i:i + w] + nl)
        seqio.build_fai(str(p))
        py = open(str(p) + ".fai").read()
        os.unlink(str(p) + ".fai")
        nf = H.FastaIndex(str(p))                       # writes the .fai
        assert open(str(p) + ".fai").read() == py
        pf = seqio.FastaFile(str(p))
        for c in pf.order:
            n = pf.index[c][0]
            for a, b in ((1, n), (3, 61), (60, 61), (n - 2, n + 9)):
                assert nf.fetch(c, a, b) == pf.fetch(c, a, b)
        nf.close()


def test_chop_native_equals_python_on_sam(python_io):
    pf = seqio.FastaFile(REF)
    chrom = pf.order[0]; L = pf.index[chrom][0]
    wins = _windows(np.random.default_rng(2), chrom, L, 250)
    al = H.AlnFile(SAM)
    exp0 = [python_io(SF.chop_pacbio_read_by_pos, SAM, c, s, e, f) for c, s, e, f in wins]
    exp = [SF.minimize_pacbio_read_list(x) for x in exp0]
    assert max(len(x) for x in exp0) > 20 > min(len(x) for x in exp0)
    for th in (1, 4):
        got, seen = H.chop_many([al], wins, max_reads=20, threads=th)
        assert got == exp and seen > sum(len(x) for x in exp0)
    assert H.chop_many([al], wins, max_reads=0, threads=3)[0] == exp0
    # the drop-in functions route through the native code by default and give the same lists
    for c, s, e, f in wins[:40]:
        assert SF.chop_pacbio_read_by_pos(SAM, c, s, e, f) == python_io(SF.chop_pacbio_read_by_pos, SAM, c, s, e, f)
    sv = [chrom, wins[0][1] + 500, wins[0][1] + 900]
    assert SF.simple_chop_pacbio_read_simple_short(SAM, sv, 333) == python_io(SF.simple_chop_pacbio_read_simple_short, SAM, sv, 333)
    assert SF.simple_del_chop_pacbio_read_simple_short(SAM, sv, 333) == python_io(SF.simple_del_chop_pacbio_read_simple_short, SAM, sv, 333)
    al.close()


def _sam_records(n=None):
    recs = []
    with gzip.open(SAM, "rt") as f:
        for line in f:
            if not line.startswith("@"):
                p = line.rstrip("\n").split("\t")
                recs.append((p[0], 0, int(p[3]), p[5], p[9]))
    return recs[:n] if n else recs


def test_chop_native_on_bam_with_and_without_index(tmp_path, python_io):
    recs = _sam_records()
    chrom_len = int(open(REF + ".fai").read().split()[1])
    wins = _windows(np.random.default_rng(4), "chr1", chrom_len, 80) + [("chr1", 1, 400, 50), ("chrZ", 5, 900, 50)]
    sam = H.AlnFile(SAM)
    exp, _ = H.chop_many([sam], wins, max_reads=20, threads=2)
    for with_bai in (True, False):
        path = str(tmp_path / f"r{int(with_bai)}.bam")
        bam_writer.write_bam(path, [("chr1", chrom_len)], recs, with_bai=with_bai)
        b = H.AlnFile(path)
        for th in (1, 3):
            got, _ = H.chop_many([b], wins, max_reads=20, threads=th)
            assert got == exp, (with_bai, th)
        pyl = [SF.minimize_pacbio_read_list(python_io(SF.chop_pacbio_read_by_pos, path, c, s, e, f)) for c, s, e, f in wins[:25]]
        assert pyl == exp[:25]
        b.close()
    # two files for one window (bam_in_decide may return several): lists are concatenated in file order, then cut to 20
    half = len(recs) // 2
    pa, pb = str(tmp_path / "a.bam"), str(tmp_path / "b.bam")
    bam_writer.write_bam(pa, [("chr1", chrom_len)], recs[:half])
    bam_writer.write_bam(pb, [("chr1", chrom_len)], recs[half:])
    fa, fb = H.AlnFile(pa), H.AlnFile(pb)
    got, _ = H.chop_many([fa, fb], wins[:30], max_reads=20, threads=2)
    for w, g in zip(wins[:30], got):
        x = python_io(SF.chop_pacbio_read_by_pos, pa, *w) + python_io(SF.chop_pacbio_read_by_pos, pb, *w)
        assert g == SF.minimize_pacbio_read_list(x)
    sam.close(); fa.close(); fb.close()


# ---- the example of the SAM specification (section 1.1), transcribed to the BAM layout of section 4.2 by hand ----
SPEC_HEADER = ("42414d01" "2a000000"                                    # magic, l_text = 42
               "40484409564e3a312e3609534f3a636f6f7264696e6174650a"     # @HD VN:1.6 SO:coordinate
               "40535109534e3a726566094c4e3a34350a"                     # @SQ SN:ref LN:45
               "01000000" "04000000" "72656600" "2d000000")             # n_ref = 1, l_name = 4, "ref\0", l_ref = 45
SPEC_RECORDS = [
    # r001 99 ref 7 30 8M2I4M1D3M = 37 39 TTAGATAAAGGATACTG *
    "53000000" "00000000" "06000000" "05" "1e" "4912" "0500" "6300" "11000000" "00000000" "24000000" "27000000" "7230303100"
    "80000000" "21000000" "40000000" "12000000" "30000000" "881418111441812840" + "ff" * 17,
    # r002 0 ref 9 30 3S6M1P1I4M * 0 0 AAAAGATAAGGATA *
    "4e000000" "00000000" "08000000" "05" "1e" "4912" "0500" "0000" "0e000000" "ffffffff" "ffffffff" "00000000" "7230303200"
    "34000000" "60000000" "16000000" "11000000" "40000000" "11114181144181" + "ff" * 14,
    # r003 0 ref 9 30 5S6M * 0 0 GCCTAAGCTAA * SA:Z:ref,29,-,6H5M,17,0;
    "55000000" "00000000" "08000000" "05" "1e" "4912" "0200" "0000" "0b000000" "ffffffff" "ffffffff" "00000000" "7230303300"
    "54000000" "60000000" "422811428110" + "ff" * 11 + "53415a7265662c32392c2d2c3648354d2c31372c303b00",
    # r004 0 ref 16 30 6M14N5M * 0 0 ATAGCTTCAGC *
    "42000000" "00000000" "0f000000" "05" "1e" "4912" "0300" "0000" "0b000000" "ffffffff" "ffffffff" "00000000" "7230303400"
    "60000000" "e3000000" "50000000" "181428821420" + "ff" * 11,
    # r003 2064 ref 29 17 6H5M * 0 0 TAGGC * SA:Z:ref,9,+,5S6M,30,1;
    "4b000000" "00000000" "1c000000" "05" "11" "4912" "0200" "1008" "05000000" "ffffffff" "ffffffff" "00000000" "7230303300"
    "65000000" "50000000" "814420" + "ff" * 5 + "53415a7265662c392c2b2c3553364d2c33302c313b00",
    # r001 147 ref 37 30 9M = 7 -39 CAGCGGCAT * NM:i:1
    "3b000000" "00000000" "24000000" "05" "1e" "4912" "0100" "9300" "09000000" "00000000" "06000000" "d9ffffff" "7230303100"
    "90000000" "2142442180" + "ff" * 9 + "4e4d4301",
]
SPEC_SAM = [("r001", 7, "8M2I4M1D3M", "TTAGATAAAGGATACTG"), ("r002", 9, "3S6M1P1I4M", "AAAAGATAAGGATA"),
            ("r003", 9, "5S6M", "GCCTAAGCTAA"), ("r004", 16, "6M14N5M", "ATAGCTTCAGC"), ("r003", 29, "6H5M", "TAGGC"),
            ("r001", 37, "9M", "CAGCGGCAT")]


def _bgzf(data: bytes) -> bytes:
    """One BGZF block (SAM specification section 4.1): a gzip member with the BC extra subfield."""
    c = zlib.compressobj(9, zlib.DEFLATED, -15)
    cd = c.compress(data) + c.flush()
    return (bytes.fromhex("1f8b08040000000000ff0600424302") + b"\x00" + struct.pack("<H", len(cd) + 25) + cd +
            struct.pack("<II", zlib.crc32(data), len(data)))


def _spec_bam(path):
    raw = bytes.fromhex(SPEC_HEADER + "".join(SPEC_RECORDS))
    cut1, cut2 = len(bytes.fromhex(SPEC_HEADER)) + 40, len(raw) - 70            # block boundaries inside records
    with open(path, "wb") as f:
        f.write(_bgzf(raw[:cut1]) + _bgzf(raw[cut1:cut2]) + _bgzf(raw[cut2:]) + _bgzf(b""))
    return raw


def test_spec_example_bam(tmp_path, python_io):
    path = str(tmp_path / "spec.bam")
    raw = _spec_bam(path)
    # independent decoders: Python's gzip reads the BGZF layer; a struct walk written from section 4.2 reads the records
    assert gzip.decompress(open(path, "rb").read()) == raw
    p = 4
    l_text = struct.unpack_from("<i", raw, p)[0]; p += 4 + l_text
    assert raw[8:8 + l_text].decode().splitlines()[1] == "@SQ\tSN:ref\tLN:45"
    n_ref, l_name = struct.unpack_from("<ii", raw, p); p += 8
    assert (n_ref, raw[p:p + l_name]) == (1, b"ref\0"); p += l_name + 4
    parsed = []
    while p < len(raw):
        bs, tid, pos, l_rn, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiiBBHHHi", raw, p)
        q = p + 36
        qname = raw[q:q + l_rn - 1].decode(); q += l_rn
        cig = "".join(f"{c >> 4}{'MIDNSHP=X'[c & 15]}" for c in struct.unpack_from(f"<{n_cig}I", raw, q)); q += 4 * n_cig
        seq = "".join("=ACMGRSVTWYHKDBN"[(raw[q + i // 2] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
        parsed.append((qname, pos + 1, cig, seq))
        p += 4 + bs
    assert parsed == SPEC_SAM
    # region queries: what `samtools view spec.bam ref:a-b` prints, worked out from the alignment spans of the example
    # (r001 7-22, r002 9-18, r003 9-14, r004 16-40, r003 29-33, r001 37-45)
    expect = {(1, 6): [], (7, 7): [0], (15, 15): [0, 1], (23, 28): [3], (30, 30): [3, 4], (41, 45): [5], (19, 22): [0, 3],
              (1, 45): [0, 1, 2, 3, 4, 5], (14, 16): [0, 1, 2, 3]}
    pb = python_io(seqio.AlignmentFile, path)
    nb = H.AlnFile(path)
    for (a, b), idx in expect.items():
        got = [(r.qname, r.pos, r.cigar, r.seq) for r in pb.fetch("ref", a, b)]
        assert got == [SPEC_SAM[i] for i in idx], (a, b)
        # the native reader sees the same records: with a 0-length window cut nothing is kept, so count what it saw
        _, seen = H.chop_many([nb], [("ref", a, b, 50)], max_reads=0, threads=1)
        assert seen == len(idx), (a, b)
    # the reference's CIGAR walk on these records (Simple_function.pyx:309-337; P, N, H advance nothing)
    assert H.cigar2alignstart("3S6M1P1I4M", 9, 12) == SF.cigar2alignstart_by_pos("3S6M1P1I4M", 9, 12, 20) == [6, 0]
    assert H.cigar2alignstart("6M14N5M", 16, 25) == SF.cigar2alignstart_by_pos("6M14N5M", 16, 25, 30) == [9, 0]
    assert H.cigar2alignstart("6H5M", 29, 31) == SF.cigar2alignstart_by_pos("6H5M", 29, 31, 33) == [2, 0]
    assert H.cigar2alignstart("*", 10, 20) == [0, -10]
    # chop: window 9..12 with flank 4 keeps r001 (POS 7 <= 9: walk 8M -> read offset 2, miss 0; 17 - 2 > 3 -> 3 bases) and
    # r002 / r003 (POS 9: soft clips skipped)
    wins = [("ref", 9, 12, 4), ("ref", 7, 10, 2), ("ref", 16, 20, 4), ("ref", 30, 33, 10)]
    got, _ = H.chop_many([nb], wins, max_reads=0, threads=1)
    exp = [python_io(SF.chop_pacbio_read_by_pos, path, *w) for w in wins]
    assert got == exp
    assert got[0] == [["AGA", 0, "r001"], ["AGA", 0, "r002"], ["AGC", 0, "r003"]]
    nb.close()


def test_cigar_walk_native_equals_reference_loop():
    rng = np.random.default_rng(8)
    for _ in range(400):
        n_ops = int(rng.integers(1, 60))
        ops = rng.choice(list("MIDNSHP=X"), size=n_ops, p=[0.4, 0.15, 0.15, 0.02, 0.05, 0.02, 0.01, 0.1, 0.1])
        cigar = "".join(f"{int(rng.integers(1, 40))}{o}" for o in ops)
        a0 = int(rng.integers(1, 5000))
        for start in (a0 - 10, a0, a0 + 1, a0 + int(rng.integers(0, 600)), a0 + 100000):
            rr, ar, last = SF._cigar_walk(cigar, a0, start)
            sd = ar - start
            exp = [rr - sd, 0] if (last != "" and last in "M=") else [rr, sd]
            assert H.cigar2alignstart(cigar, a0, start) == exp, (cigar, a0, start)
    assert H.cigar2alignstart("12M3B4M", 5, 9) == SF.cigar2alignstart_by_pos("12M3B4M", 5, 9, 30)     # B is not in the reference's pattern


def test_long_cigar_cg_tag(tmp_path, python_io):
    """More than 65535 CIGAR operations: BAM keeps <l_seq>S<ref_len>N in the field and the real CIGAR in CG:B,I
    (SAM specification 4.2.2); samtools view -- what the reference reads -- prints the real one.  Both readers restore it."""
    rng = np.random.default_rng(12)
    n_pairs = 33000                                            # 66000 operations: 1M1I repeated
    cigar = "1M1I" * n_pairs
    seq = "".join(rng.choice(list("ACGT"), size=2 * n_pairs))
    other = ("short", 0, 150, "300M", "".join(rng.choice(list("ACGT"), size=300)))
    path = str(tmp_path / "long.bam")
    bam_writer.write_bam(path, [("c", 100000)], [("long", 0, 100, cigar, seq), other])
    pb = python_io(seqio.AlignmentFile, path)
    r = [x for x in pb.fetch("c", 20000, 20010)]
    assert [x.qname for x in r] == ["long"] and r[0].cigar == cigar
    wins = [("c", 20000, 21000, 500), ("c", 140, 400, 100)]
    nb = H.AlnFile(path)
    got, _ = H.chop_many([nb], wins, max_reads=0, threads=1)
    exp = [python_io(SF.chop_pacbio_read_by_pos, path, *w) for w in wins]
    assert got == exp and len(got[0]) == 1 and got[0][0][2] == "long" and len(got[0][0][0]) == 1000
    nb.close()


def test_prefetch_cache_gives_the_same_answers(python_io):
    pf = seqio.FastaFile(REF)
    chrom = pf.order[0]; L = pf.index[chrom][0]
    wins = _windows(np.random.default_rng(6), chrom, L, 30)
    regs = [(c, s, e) for c, s, e, _ in wins]
    seqio.clear_prefetch()
    seqio.prefetch(REF, regs, [SAM], wins, threads=2)
    assert len(seqio._region_cache) == len(set(regs)) and len(seqio._reads_cache) == len(set(wins))
    for (c, s, e, f) in wins:
        assert seqio.faidx(REF, c, s, e) == pf.fetch(c, s, e)
        got = seqio.chop_reads([SAM], c, s, e, f)
        assert got == SF.minimize_pacbio_read_list(python_io(SF.chop_pacbio_read_by_pos, SAM, c, s, e, f))
        n = len(got)
        got[0:1] = []                                           # callers may edit the list they get: the cache must not change
        assert len(seqio.chop_reads([SAM], c, s, e, f)) == n
    seqio.clear_prefetch()
    assert not seqio._region_cache and not seqio._reads_cache


def test_integration_md_hostio_stub_runs(python_io):
    """The host-I/O ctypes stub INTEGRATION.md shows a maintainer of the reference, executed as written (only the library
    path is made absolute), returns what the Python restatement of the reference's functions returns."""
    import re
    from vapor_b200 import _native
    md = open(os.path.join(os.path.dirname(HERE), "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# vapor_vali/_b200_io\.py.*?)```", md, re.S).group(1)
    block = block.replace('C.CDLL("libvapor_b200.so")', f'C.CDLL({_native.LIB_PATH!r})')
    ns = {}
    exec(compile(block, "INTEGRATION.md:_b200_io.py", "exec"), ns)
    fa, al = ns["open_fasta"](REF), ns["open_alignments"](SAM)
    pf = seqio.FastaFile(REF)
    chrom = pf.order[0]
    assert ns["ref_seq_readin"](fa, chrom, 12000, 13200) == pf.fetch(chrom, 12000, 13200)
    wins = [(chrom, 11500, 13143, 500), (chrom, 24214, 25980, 500), (chrom, 36064, 37064, 500)]
    got = ns["chop_and_minimize"]([al], wins)
    exp = [SF.minimize_pacbio_read_list(python_io(SF.chop_pacbio_read_by_pos, SAM, *w)) for w in wins]
    assert got == exp and all(len(x) > 3 for x in got)
