"""Seeded synthetic *file-level* inputs for the command line: reference FASTA (+ .fai), SV calls as the
5-column BED and as a VCF with complex events in the README's INFO syntax (README.md:79-82), and simulated
PacBio-CLR-like reads aligned to the reference as SAM text with truth CIGARs.

The reference repo plants SVs with ``simulate/generateVariantChromosomes.py`` (non-overlapping events, a
3 kb buffer, :141) but has no read simulator and needs Biopython, which this image lacks; this module follows
the same placement rules and adds the reads: ~15 % error split insertion : deletion : substitution =
50 : 30 : 20, log-normal lengths, half of the reads of a heterozygous event from each haplotype.

Used by the end-to-end CLI tests (``tests/golden/make_cli_golden.py`` runs the unmodified reference CLI on
these files once, here, to produce the committed golden tables) and by ``tools/cli_bench.py``.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[_a] = _b


def _revcomp(a: np.ndarray) -> np.ndarray:
    return _COMP[a[::-1]]


@dataclass
class PlantedSV:
    svid: str
    svtype: str                 # DEL, DUP, INV, INS, DEL_INV, DUP_INV, DISDUP, OTHER
    chrom: str
    start: int                  # BED / VCF coordinates as written to the call files
    end: int
    genotype: int               # 1 het, 2 hom-alt
    segments: list = field(default_factory=list)     # alternative haplotype of the event region
    ins_seq: str = ""
    extra: dict = field(default_factory=dict)


def _hap_arrays(ref: np.ndarray, segments: Sequence[tuple]) -> Tuple[np.ndarray, np.ndarray]:
    """Haplotype bases and, per base, the reference coordinate it aligns to (-1 = inserted)."""
    seqs, pos = [], []
    for seg in segments:
        kind = seg[0]
        if kind == "ref":
            seqs.append(ref[seg[1]:seg[2]]); pos.append(np.arange(seg[1], seg[2], dtype=np.int64))
        elif kind == "inv":                       # aligned through, bases reverse-complemented
            seqs.append(_revcomp(ref[seg[1]:seg[2]])); pos.append(np.arange(seg[1], seg[2], dtype=np.int64))
        elif kind == "ins":
            s = seg[1] if isinstance(seg[1], np.ndarray) else np.frombuffer(seg[1].encode(), dtype=np.uint8)
            seqs.append(s); pos.append(np.full(len(s), -1, dtype=np.int64))
        elif kind == "del":
            continue
        else:
            raise ValueError(kind)
    return np.concatenate(seqs), np.concatenate(pos)


def _rle_cigar(ops: np.ndarray, lens: np.ndarray) -> str:
    keep = lens > 0
    ops, lens = ops[keep], lens[keep]
    if len(ops) == 0:
        return "*"
    brk = np.concatenate([[True], ops[1:] != ops[:-1]])
    idx = np.nonzero(brk)[0]
    tot = np.add.reduceat(lens, idx)
    sym = "MIDS"
    return "".join(f"{int(n)}{sym[int(o)]}" for n, o in zip(tot, ops[idx]))


def simulate_read(rng: np.random.Generator, hap: np.ndarray, hpos: np.ndarray, a: int, b: int, err: float,
                  mix=(0.5, 0.3, 0.2)) -> Optional[Tuple[int, str, str]]:
    """A noisy copy of hap[a:b] with its truth alignment: (POS 1-based, CIGAR, SEQ), or None when it aligns nowhere."""
    base, rp = hap[a:b], hpos[a:b]
    n = len(base)
    p_ins, p_del, p_sub = (err * m for m in mix)
    u = rng.random(n)
    deleted = u < p_del
    sub = (u >= p_del) & (u < p_del + p_sub)
    ins_after = rng.random(n) < p_ins
    aligned = rp >= 0
    al_idx = np.nonzero(aligned)[0]
    if len(al_idx) == 0:
        return None
    first, last = al_idx[0], al_idx[-1]
    # bases
    lut = np.zeros(256, dtype=np.uint8); lut[_ACGT] = np.arange(4, dtype=np.uint8)
    code = lut[base]
    code = np.where(sub, (code + rng.integers(1, 4, size=n, dtype=np.uint8)) & 3, code)
    cnt = (~deleted).astype(np.int64) + ins_after.astype(np.int64)
    owner = np.repeat(np.arange(n), cnt)
    csum = np.cumsum(cnt)
    firstslot = np.arange(int(csum[-1]) if n else 0) - (csum[owner] - cnt[owner])
    is_copy = (firstslot == 0) & (~deleted[owner])
    out_code = np.where(is_copy, code[owner], rng.integers(0, 4, size=len(owner), dtype=np.uint8))
    seq = _ACGT[out_code].tobytes().decode()
    # cigar: per hap base three slots -- reference gap before it (D), the base itself, an inserted base after it
    gap = np.zeros(n, dtype=np.int64)
    prev = np.full(n, -1, dtype=np.int64)
    prev[al_idx[1:]] = rp[al_idx[:-1]]
    gap[al_idx[1:]] = rp[al_idx[1:]] - prev[al_idx[1:]] - 1
    gap = np.maximum(gap, 0)
    base_op = np.where(aligned, np.where(deleted, 2, 0), 1)              # M / D for aligned bases, I for inserted ones
    base_len = np.where(aligned, 1, (~deleted).astype(np.int64))         # a deleted inserted base leaves nothing
    outside = (np.arange(n) < first) | (np.arange(n) > last)             # unaligned ends are soft clips
    base_op = np.where(outside & ~aligned, 3, base_op)
    ins_op = np.where(outside | (np.arange(n) == last), 3, 1)            # insertions after the last aligned base clip too
    ins_op = np.where(np.arange(n) < first, 3, ins_op)
    ops = np.stack([np.full(n, 2), base_op, ins_op], axis=1).ravel()
    lens = np.stack([gap, base_len, ins_after.astype(np.int64)], axis=1).ravel()
    # a leading/trailing D is not a valid alignment edge: start at the first aligned base that survived
    cigar = _rle_cigar(ops, lens)
    import re
    m = re.match(r"^((?:\d+S)?)(\d+)D", cigar)
    pos = int(rp[first]) + 1
    if m:                                                                # first aligned base was deleted: shift POS
        pos += int(m.group(2))
        cigar = m.group(1) + cigar[m.end():]
    cigar = re.sub(r"(\d+)D((?:\d+S)?)$", r"\2", cigar)
    return pos, cigar, seq


@dataclass
class Dataset:
    out_dir: str
    ref_fa: str
    sam: str
    bed: str
    vcf: str
    svs: List[PlantedSV]
    chrom_len: int


def make_dataset(out_dir: str, seed: int = 20261018, n_simple: int = 8, n_complex: int = 4,
                 size_range: Tuple[int, int] = (50, 1500), coverage: float = 30.0, read_len_mean: float = 7000.0,
                 err: float = 0.15, chrom: str = "chr1", spacing: int = 3000, line_width: int = 60,
                 simple_types: Sequence[str] = ("DEL", "DUP", "INV", "INS"),
                 complex_types: Sequence[str] = ("DEL_INV", "DUP_INV", "DISDUP", "OTHER"),
                 ins_with_seq_every: int = 2, het_frac: float = 0.5, ins_len_override: int = 0,
                 kind_size: Optional[Dict[str, Tuple[int, int]]] = None) -> Dataset:
    """Write ref.fa(.fai), reads.sam, svs.bed, svs.vcf, truth.json under ``out_dir``.
    ``kind_size``: per event kind, its own size range (e.g. a >= 10 kb ``OTHER_LONG`` event that sends the complex
    driver down its junction-window fallback, Simple_function.pyx:1537-1555)."""
    rng = np.random.default_rng(seed)
    os.makedirs(out_dir, exist_ok=True)
    n_sv = n_simple + n_complex
    lens = rng.integers(size_range[0], size_range[1] + 1, size=n_sv)
    kinds = [simple_types[i % len(simple_types)] for i in range(n_simple)] + \
            [complex_types[i % len(complex_types)] for i in range(n_complex)]
    for i, k_ in enumerate(kinds):
        if kind_size and k_ in kind_size:
            lens[i] = int(rng.integers(kind_size[k_][0], kind_size[k_][1] + 1))
    # each event owns a stretch: spacing + up to 3 blocks of its size + spacing
    lead = 12000                                                  # room for reads to start upstream of the first window
    starts = []
    cur = lead
    for L in lens:
        starts.append(cur)
        cur += 3 * int(L) + spacing + 1200
    chrom_len = cur + lead
    ref = _ACGT[rng.integers(0, 4, size=chrom_len, dtype=np.uint8)]
    svs: List[PlantedSV] = []
    alt_segments: List[tuple] = []                                # whole-chromosome alternative haplotype per genotype
    cursor = 0
    per_sv_segments = []
    for i, (kind, L, s) in enumerate(zip(kinds, lens, starts)):
        L = int(L); e = s + L
        gt = 1 if rng.random() < het_frac else 2
        svid = f"SV_{i + 1}"
        if kind == "DEL":
            seg = [("del", s, e)]; sv = PlantedSV(svid, kind, chrom, s, e, gt)
            span_end = e
        elif kind == "DUP":
            seg = [("ref", s, e), ("ins", ref[s:e].copy())]; sv = PlantedSV(svid, kind, chrom, s, e, gt)
            span_end = e
        elif kind == "INV":
            seg = [("inv", s, e)]; sv = PlantedSV(svid, kind, chrom, s, e, gt)
            span_end = e
        elif kind == "INS":
            ins = _ACGT[rng.integers(0, 4, size=(ins_len_override or L), dtype=np.uint8)]
            seg = [("ins", ins)]
            sv = PlantedSV(svid, kind, chrom, s, s, gt, ins_seq=ins.tobytes().decode(),
                           extra={"with_seq": bool((i // len(simple_types)) % ins_with_seq_every == 0)})
            span_end = s
        elif kind == "DEL_INV":                                    # a deleted, b inverted (adjacent)
            m_ = s + L
            e = m_ + max(60, L // 2)
            seg = [("del", s, m_), ("inv", m_, e)]
            sv = PlantedSV(svid, kind, chrom, s, e, gt, extra={"del": [s, m_], "inv": [m_, e]})
            span_end = e
        elif kind == "DUP_INV":                                    # a b a^ : inverted copy of a inserted after b
            p = e + max(80, L // 2)
            seg = [("ref", s, p), ("ins", _revcomp(ref[s:e]))]
            sv = PlantedSV(svid, kind, chrom, s, e, gt, extra={"insert_point": p})
            span_end = p
        elif kind == "DISDUP":                                     # a b a : copy of a inserted after b
            p = e + max(80, L // 2)
            seg = [("ref", s, p), ("ins", ref[s:e].copy())]
            sv = PlantedSV(svid, kind, chrom, s, e, gt, extra={"insert_point": p})
            span_end = p
        elif kind == "DEL_DUP_INV":                                # README.md:81  ab/ab_a/bb^ : haplotype A = a, haplotype B = b b^
            m_ = s + L
            e = m_ + max(80, L // 2)
            seg = [("del", s, m_), ("ref", m_, e), ("ins", _revcomp(ref[m_:e]))]
            seg_a = [("ref", s, m_), ("del", m_, e)]
            sv = PlantedSV(svid, kind, chrom, s, e, 2, extra={"bps": [s, m_, e], "ref": "ab", "alt": "a/bb^"})
            span_end = e
        elif kind == "OTHER3":                                     # three blocks, three alternative alleles: ac / ab^c / cba
            m1 = s + L
            m2 = m1 + max(80, L // 2)
            e = m2 + max(80, L // 3)
            seg = [("ref", s, m1), ("inv", m1, m2), ("ref", m2, e)]                      # haplotype B: a b^ c
            seg_a = [("ref", s, m1), ("del", m1, m2), ("ref", m2, e)]                    # haplotype A: a c
            sv = PlantedSV(svid, kind, chrom, s, e, 2, extra={"bps": [s, m1, m2, e], "ref": "abc", "alt": "ac/ab^c/cba"})
            span_end = e
        elif kind in ("OTHER", "OTHER_LONG"):                      # ab -> ba : blocks swapped
            m_ = s + L
            e = m_ + max(80, L // 2)
            seg = [("del", s, m_), ("ref", m_, e), ("ins", ref[s:m_].copy())]
            sv = PlantedSV(svid, kind, chrom, s, e, gt, extra={"bps": [s, m_, e], "ref": "ab", "alt": "ba"})
            span_end = e
        elif kind == "OTHER2":                                     # two different alternative alleles: a (b deleted) and ba
            m_ = s + L
            e = m_ + max(80, L // 2)
            seg = [("del", s, m_), ("ref", m_, e), ("ins", ref[s:m_].copy())]          # haplotype B: ba
            seg_a = [("ref", s, m_), ("del", m_, e)]                                     # haplotype A: a
            sv = PlantedSV(svid, kind, chrom, s, e, 2, extra={"bps": [s, m_, e], "ref": "ab", "alt": "a/ba"})
            span_end = e
        else:
            raise ValueError(kind)
        svs.append(sv)
        per_sv_segments.append((s, span_end, seg, seg_a if kind in ("OTHER2", "DEL_DUP_INV", "OTHER3") else None))
    # haplotypes: A carries hom-alt events only, B carries every event
    def build(which):
        segs, cur = [], 0
        for sv, (s, span_end, seg, seg_a) in zip(svs, per_sv_segments):
            carry = sv.genotype == 2 or which == "B"
            segs.append(("ref", cur, s))
            if carry:
                segs += seg_a if (seg_a is not None and which == "A") else seg
                cur = span_end
            else:
                cur = s
        segs.append(("ref", cur, chrom_len))
        return _hap_arrays(ref, segs)
    haps = [build("A"), build("B")]

    # ---- reads -------------------------------------------------------------------------------------------
    recs = []
    rid = 0
    for h, (hap, hpos) in enumerate(haps):
        target = coverage / 2.0 * len(hap)
        made = 0.0
        while made < target:
            ln = int(min(max(rng.lognormal(np.log(read_len_mean), 0.35), 1500), 20000))
            a = int(rng.integers(0, max(1, len(hap) - ln)))
            r = simulate_read(rng, hap, hpos, a, min(len(hap), a + ln), err)
            made += ln
            if r is None:
                continue
            rid += 1
            recs.append((r[0], f"read{rid}_h{h}", r[1], r[2]))
    recs.sort(key=lambda t: t[0])

    ref_fa = os.path.join(out_dir, "ref.fa")
    with open(ref_fa, "w") as f:
        f.write(f">{chrom}\n")
        s = ref.tobytes().decode()
        for i in range(0, len(s), line_width):
            f.write(s[i:i + line_width] + "\n")
    from .seqio import build_fai
    build_fai(ref_fa)
    sam = os.path.join(out_dir, "reads.sam")
    with open(sam, "w") as f:
        f.write("@HD\tVN:1.6\tSO:coordinate\n")
        f.write(f"@SQ\tSN:{chrom}\tLN:{chrom_len}\n")
        for pos, name, cigar, seq in recs:
            f.write(f"{name}\t0\t{chrom}\t{pos}\t60\t{cigar}\t*\t0\t0\t{seq}\t*\n")

    # ---- call files ---------------------------------------------------------------------------------------
    bed = os.path.join(out_dir, "svs.bed")
    with open(bed, "w") as f:
        for sv in svs:
            if sv.svtype in ("DEL", "DUP", "INV"):
                f.write(f"{sv.chrom}\t{sv.start}\t{sv.end}\t{sv.svid}\t{sv.svtype}\n")
            elif sv.svtype == "INS":
                t = f"INS_{sv.ins_seq}" if sv.extra.get("with_seq") else f"INS_{len(sv.ins_seq)}"
                f.write(f"{sv.chrom}\t{sv.start}\t{sv.end}\t{sv.svid}\t{t}\n")
    vcf = os.path.join(out_dir, "svs.vcf")
    with open(vcf, "w") as f:
        f.write("##fileformat=VCFv4.2\n")
        f.write(f"##contig=<ID={chrom},length={chrom_len}>\n")
        f.write('##INFO=<ID=SVTYPE,Number=1,Type=String,Description="Type of structural variant">\n')
        f.write('##INFO=<ID=END,Number=1,Type=Integer,Description="End position">\n')
        f.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n")
        for sv in svs:
            gt = "0/1" if sv.genotype == 1 else "1/1"
            c = sv.chrom
            if sv.svtype in ("DEL", "INV"):
                info = f"SVTYPE={sv.svtype};END={sv.end}"
            elif sv.svtype == "DUP":
                info = f"SVTYPE=TANDUP;END={sv.end}"
            elif sv.svtype == "INS":
                info = f"SVTYPE=INS;END={sv.end};SVLEN={len(sv.ins_seq)}" + (f";SEQ={sv.ins_seq}" if sv.extra.get("with_seq") else "")
            elif sv.svtype == "DEL_INV":
                d, v = sv.extra["del"], sv.extra["inv"]
                info = (f"SVTYPE=del_inv;END={sv.end};del={c}:{d[0]}-{d[1]};inv={c}:{v[0]}-{v[1]};"
                        f"Other=ab/ab_ab/b^_{c}:{d[0]}:{d[1]}:{v[1]}")
            elif sv.svtype == "DUP_INV":
                p = sv.extra["insert_point"]
                info = f"SVTYPE=dup_inv;END={sv.end};insert_point={c}:{p};Other=ab/ab_ab/aba^_{c}:{sv.start}:{sv.end}:{p}"
            elif sv.svtype == "DISDUP":
                p = sv.extra["insert_point"]
                info = f"SVTYPE=disdup;END={sv.end};insert_point={c}:{p};Other=ab/ab_ab/aba_{c}:{sv.start}:{sv.end}:{p}"
            elif sv.svtype == "OTHER2":
                b0, b1, b2 = sv.extra["bps"]
                info = f"SVTYPE=cannot_classify_for_now;END={sv.end};Other=ab/ab_a/ba_{c}:{b0}:{b1}:{b2}"
            elif sv.svtype == "DEL_DUP_INV":
                b0, b1, b2 = sv.extra["bps"]
                info = (f"SVTYPE=del_dup_inv;END={sv.end};del={c}:{b0}-{b1};dup_inv={c}:{b1}-{b2};insert_point={c}:{b2};"
                        f"Other=ab/ab_a/bb^_{c}:{b0}:{b1}:{b2}")
            elif sv.svtype == "OTHER3":
                b0, b1, b2, b3 = sv.extra["bps"]
                info = f"SVTYPE=cannot_classify_for_now;END={sv.end};Other=abc/abc_ac/ab^c/cba_{c}:{b0}:{b1}:{b2}:{b3}"
            else:
                b0, b1, b2 = sv.extra["bps"]
                info = f"SVTYPE=cannot_classify_for_now;END={sv.end};Other=ab/ab_ab/ba_{c}:{b0}:{b1}:{b2}"
            f.write(f"{c}\t{sv.start}\t{sv.svid}\tN\t<{sv.svtype}>\t60\tPASS\t{info}\tGT\t{gt}\n")
    with open(os.path.join(out_dir, "truth.json"), "w") as f:
        json.dump([{"svid": s.svid, "type": s.svtype, "start": s.start, "end": s.end, "genotype": s.genotype, **{k: v for k, v in s.extra.items()}}
                   for s in svs], f, indent=1)
    return Dataset(out_dir, ref_fa, sam, bed, vcf, svs, chrom_len)
