"""In-process FASTA / SAM / BAM access (vapor_b200/seqio.py) -- the stand-in for the reference's
``samtools faidx`` and ``samtools view`` subprocesses (Simple_function.pyx:1206, :340)."""
import gzip
import os

import numpy as np
import pytest

from vapor_b200 import seqio

import bam_writer

HERE = os.path.dirname(os.path.abspath(__file__))
CASE = os.path.join(HERE, "golden", "cli_case")


def test_fasta_fetch_matches_plain_slicing(tmp_path):
    rng = np.random.default_rng(3)
    seqs = {"chrA": "".join(rng.choice(list("ACGTN"), size=1234)), "chrB": "".join(rng.choice(list("acgt"), size=77)), "c3": "ACGT"}
    fa = tmp_path / "t.fa"
    with open(fa, "w") as f:
        for (name, s), w in zip(seqs.items(), (60, 10, 80)):
            f.write(f">{name} some description\n")
            for i in range(0, len(s), w):
                f.write(s[i:i + w] + "\n")
    ff = seqio.FastaFile(str(fa))
    assert os.path.exists(str(fa) + ".fai")
    for name, s in seqs.items():
        for a, b in [(1, len(s)), (1, 1), (5, 64), (60, 61), (61, 120), (len(s), len(s)), (len(s) - 3, len(s) + 50), (-5, 10), (0, 3)]:
            exp = s[max(a, 1) - 1:min(b, len(s))]
            assert ff.fetch(name, a, b) == exp, (name, a, b)
    assert ff.fetch("nope", 1, 10) == "" and ff.fetch("chrA", 50, 40) == ""
    assert seqio.faidx(str(fa), "chrA", 11, 20) == seqs["chrA"][10:20]


def _sam_records():
    recs = []
    with gzip.open(os.path.join(CASE, "reads.sam.gz"), "rt") as f:
        for line in f:
            if line.startswith("@"):
                continue
            p = line.rstrip("\n").split("\t")
            recs.append((p[0], 0, int(p[3]), p[5], p[9]))
    return recs


def test_bam_reader_matches_sam_text(tmp_path):
    recs = _sam_records()[:900]
    chrom_len = int(open(os.path.join(CASE, "ref.fa.fai")).read().split()[1])
    sam = seqio.AlignmentFile(os.path.join(CASE, "reads.sam.gz"))
    sam_names = {r[0] for r in recs}
    rng = np.random.default_rng(5)
    hi = max(r[2] for r in recs)
    regions = [(1, 500), (hi - 10, hi + 5000), (13000, 13000)] + [tuple(sorted(rng.integers(1, hi + 2000, size=2))) for _ in range(40)]
    for with_bai in (True, False):
        bam = str(tmp_path / f"r_{int(with_bai)}.bam")
        bam_writer.write_bam(bam, [("chr1", chrom_len)], recs, with_bai=with_bai)
        bf = seqio.AlignmentFile(bam)
        assert (bf._impl.bai is not None) == with_bai
        for a, b in regions:
            exp = [(r.qname, r.pos, r.cigar, r.seq) for r in sam.fetch("chr1", int(a), int(b)) if r.qname in sam_names]
            got = [(r.qname, r.pos, r.cigar, r.seq) for r in bf.fetch("chr1", int(a), int(b))]
            assert got == exp, (with_bai, a, b, len(got), len(exp))
        assert list(bf.fetch("chrZ", 1, 100)) == []


def test_chop_reads_same_from_bam_and_sam(tmp_path):
    """chop_pacbio_read_by_pos (Simple_function.pyx:339-354) gives the same kernel inputs whichever container holds the reads."""
    from vapor_b200 import Simple_function as SF
    recs = _sam_records()
    chrom_len = int(open(os.path.join(CASE, "ref.fa.fai")).read().split()[1])
    bam = str(tmp_path / "all.bam")
    bam_writer.write_bam(bam, [("chr1", chrom_len)], recs)
    for s, e in [(12000 - 500, 12643 + 500), (24714 - 500, 25480 + 500), (36564 - 500, 36564 + 500)]:
        a = SF.chop_pacbio_read_by_pos(os.path.join(CASE, "reads.sam.gz"), "chr1", s, e, 500)
        b = SF.chop_pacbio_read_by_pos(bam, "chr1", s, e, 500)
        assert a == b and len(a) > 3


def test_cigar_table_walk_equals_plain_walk():
    """The table-driven cigar2alignstart_by_pos equals the reference's op-by-op loop (Simple_function.pyx:309-337) on
    random CIGARs, including X / N / H / P operations, windows before the alignment and beyond its end."""
    from vapor_b200 import Simple_function as SF
    rng = np.random.default_rng(8)
    for _ in range(300):
        n_ops = int(rng.integers(1, 60))
        ops = rng.choice(list("MIDNSHP=X"), size=n_ops, p=[0.4, 0.15, 0.15, 0.02, 0.05, 0.02, 0.01, 0.1, 0.1])
        cigar = "".join(f"{int(rng.integers(1, 40))}{o}" for o in ops)
        a0 = int(rng.integers(1, 5000))
        for start in (a0 - 10, a0, a0 + 1, a0 + int(rng.integers(0, 600)), a0 + 100000):
            rr, ar, last = SF._cigar_walk(cigar, a0, start)
            sd = ar - start
            exp = [rr - sd, 0] if (last != "" and last in "M=") else [rr, sd]
            assert SF.cigar2alignstart_by_pos(cigar, a0, start, start + 1000) == exp, (cigar, a0, start)
    assert SF.cigar2alignstart_by_pos("*", 10, 20, 30) == [0, -10]


def test_sam_text_multi_contig_and_unsorted(tmp_path):
    """Records of several contigs, not coordinate-sorted: fetch still returns exactly the overlapping ones, in file order."""
    sam = tmp_path / "u.sam"
    recs = [("a", "c2", 500, "100M"), ("b", "c1", 900, "50M20D50M"), ("c", "c1", 100, "10S80M"), ("d", "c1", 950, "*"),
            ("e", "c2", 10, "30M"), ("f", "c1", 880, "30M")]
    with open(sam, "w") as f:
        f.write("@HD\tVN:1.6\n")
        for q, c, p, cg in recs:
            f.write(f"{q}\t0\t{c}\t{p}\t60\t{cg}\t*\t0\t0\t{'ACGT' * 30}\t*\n")
    af = seqio.AlignmentFile(str(sam))
    assert [r.qname for r in af.fetch("c1", 905, 960)] == ["b", "d", "f"]
    assert [r.qname for r in af.fetch("c1", 1, 99)] == []
    assert [r.qname for r in af.fetch("c1", 179, 179)] == ["c"]            # soft clip does not extend the span: 100..179
    assert [r.qname for r in af.fetch("c1", 180, 180)] == []
    assert [r.qname for r in af.fetch("c2", 1, 10000)] == ["a", "e"]
