"""Seeded synthetic workloads for the scoring path (BASELINE.json configs 2-5).

There is no read simulator in the reference (only the SV planter under ``simulate/``), and no
samtools/pysam/Biopython in the image, so this module makes the *kernel inputs* directly: for
every SV it builds the reference/alternative structure strings the reference's L2 drivers would
build (SURVEY.md section 8a recipe table; vapor_vali/Simple_function.pyx:1701-1917) and simulates
PacBio-CLR-like reads (about 15 % error, insertion : deletion : substitution = 50 : 30 : 20) cut
to the read window the way ``chop_pacbio_read_by_pos`` cuts them (:339-354), at most 20 reads per SV
(``minimize_pacbio_read_list``, :1091-1102) -- what 30x coverage of ~10 kb reads leaves after that cap.

SV placement follows ``simulate/``: sizes uniform in the requested range
(``selectVariantChromosomes.py:53``), every SV on its own stretch of uniform-random sequence
(the 3 kb buffer of ``generateVariantChromosomes.py:141`` means windows never share sequence),
12 % of events with a 1-10 bp micro-indel at a breakpoint (``generateVariantChromosomes.py:264``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .engine import MODE_ABS, MODE_ABS_AND_W10, MODE_REDEF, MODE_W10, PackedBatch

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP_ASCII = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP_ASCII[_a] = _b

SV_TYPES = ("DEL", "TANDUP", "INV", "INS")


def random_dna(rng: np.random.Generator, n: int) -> np.ndarray:
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def revcomp(a: np.ndarray) -> np.ndarray:
    return _COMP_ASCII[a[::-1]]


def flank_length_calculate(svlen: int) -> int:
    """Simple_function.pyx:794-802: the SV length below 500 bp, else 500."""
    return svlen if svlen < 500 else 500


def simulate_reads(rng: np.random.Generator, hap: np.ndarray, seg_start: np.ndarray, seg_len: np.ndarray,
                   n_out: np.ndarray, err: float = 0.15, mix=(0.5, 0.3, 0.2)) -> Tuple[np.ndarray, np.ndarray]:
    """Noisy copies of ``hap[seg_start[i] : seg_start[i]+seg_len[i]]``, each cut to exactly
    ``n_out[i]`` bases.  Returns (concatenated reads, offsets[len+1]).  Vectorised over all segments."""
    p_ins, p_del, p_sub = (err * m for m in mix)
    seg_len = np.asarray(seg_len, dtype=np.int64)
    n_seg = len(seg_len)
    off_in = np.zeros(n_seg + 1, dtype=np.int64)
    np.cumsum(seg_len, out=off_in[1:])
    total = int(off_in[-1])
    src = (np.repeat(np.asarray(seg_start, dtype=np.int64) - off_in[:-1], seg_len) + np.arange(total))
    base = hap[src]
    u = rng.random(total, dtype=np.float32)
    deleted = u < p_del
    sub = (u >= p_del) & (u < p_del + p_sub)
    ins = rng.random(total, dtype=np.float32) < p_ins
    # substitution: a different base
    idx = np.zeros(256, dtype=np.uint8)
    idx[_ACGT] = np.arange(4, dtype=np.uint8)
    code = idx[base]
    code = np.where(sub, (code + rng.integers(1, 4, size=total, dtype=np.uint8)) & 3, code)
    cnt = (~deleted).astype(np.int32) + ins.astype(np.int32)
    csum = np.cumsum(cnt, dtype=np.int64)
    tot_out = int(csum[-1]) if total else 0
    owner = np.repeat(np.arange(total, dtype=np.int64), cnt)
    first = np.arange(tot_out, dtype=np.int64) - (csum[owner] - cnt[owner])     # 0 = first base of its group
    is_copy = (first == 0) & (~deleted[owner])
    out_code = np.where(is_copy, code[owner], rng.integers(0, 4, size=tot_out, dtype=np.uint8))
    out = _ACGT[out_code]
    # cut every segment's output to n_out bases
    seg_out_start = np.concatenate([[0], csum[off_in[1:-1] - 1]]) if n_seg > 1 else np.zeros(1, dtype=np.int64)
    seg_out_len = np.diff(np.concatenate([seg_out_start, [tot_out]]))
    n_out = np.asarray(n_out, dtype=np.int64)
    if (seg_out_len < n_out).any():
        raise ValueError("haplotype segment too short for the requested read length")
    off_out = np.zeros(n_seg + 1, dtype=np.int64)
    np.cumsum(n_out, out=off_out[1:])
    take = np.repeat(seg_out_start - off_out[:-1], n_out) + np.arange(int(off_out[-1]))
    return out[take], off_out


@dataclass
class SVCase:
    """One SV as the L2 driver sees it: structures, reads, k, scoring mode."""
    svtype: str
    svlen: int
    ref_seq: np.ndarray
    alt_seq: np.ndarray
    mode: int
    k: int
    read_window: int                 # end - start of the read window (reads hold read_window - miss_bp bases)
    hap_ref: np.ndarray              # haplotypes from the window start onwards (read templates)
    hap_alt: np.ndarray
    genotype: int                    # 1 = het, 2 = hom-alt, 0 = hom-ref (false call)


def make_sv_case(rng: np.random.Generator, svtype: str, svlen: int, genotype: int = 1, k: int = 10,
                 micro_indel: bool = False, lowercase_frac: float = 0.0) -> SVCase:
    """Build the structures of one simple SV the way the reference drivers do."""
    f = flank_length_calculate(svlen)
    margin = 3 * svlen + 4 * f + 400                     # sequence downstream so reads can run on
    G = random_dna(rng, f + svlen + f + margin)
    if lowercase_frac > 0:                               # soft-masked stretch, as samtools faidx returns it
        a = int(rng.integers(0, len(G) // 2))
        b = a + int(lowercase_frac * len(G))
        G[a:b] = G[a:b] + 32
    s, e = f, f + svlen                                   # 0-based SV block [s, e)
    if svtype == "DEL":
        # ref = faidx[s-f, e+f] inclusive -> svlen + 2f + 1 bases (Simple_function.pyx:1709, 1206)
        ref = G[0:e + f + 1]
        alt = np.concatenate([ref[:f], ref[len(ref) - f:]])                     # :1712
        hap_alt = np.concatenate([G[:s], G[e:]])
        if micro_indel:
            hap_alt = np.concatenate([G[:s], random_dna(rng, int(rng.integers(1, 11))), G[e:]])
        return SVCase(svtype, svlen, ref, alt, MODE_ABS_AND_W10, k, 2 * f, G, hap_alt, genotype)
    if svtype == "INV":
        ref = G[0:e + f + 1]
        alt = np.concatenate([ref[:f], revcomp(ref[f:len(ref) - f]), ref[len(ref) - f:]])   # :1905
        hap_alt = np.concatenate([G[:s], revcomp(G[s:e]), G[e:]])
        return SVCase(svtype, svlen, ref, alt, MODE_ABS, k, svlen + 2 * f, G, hap_alt, genotype)
    if svtype == "TANDUP":
        ref = G[0:e + f + 1]
        mid = ref[f:len(ref) - f]
        alt = np.concatenate([ref[:f], mid, mid, ref[len(ref) - f:]])           # :1756
        hap_alt = np.concatenate([G[:e], G[s:e], G[e:]])
        return SVCase(svtype, svlen, ref, alt, MODE_REDEF, k, 2 * svlen + 2 * f, G, hap_alt, genotype)
    if svtype == "INS":
        ins = random_dna(rng, svlen)
        p = f                                            # insertion point
        ref = G[0:p + f + svlen + 1] if svlen < 5000 else G[0:p + f + 1]        # :1868-1871
        alt = np.concatenate([G[0:p + 1], ins, G[p:p + f + 1]])                 # :1874 (both windows inclusive)
        hap_alt = np.concatenate([G[:p], ins, G[p:]])
        return SVCase(svtype, svlen, ref, alt, MODE_ABS, k, svlen + 2 * f, G, hap_alt, genotype)
    raise ValueError(svtype)


@dataclass
class Workload:
    batch: PackedBatch
    sv_type: List[str]
    sv_len: np.ndarray
    sv_genotype: np.ndarray
    reads_per_sv: np.ndarray
    cells: int                        # sum over reads of n*(m_ref+m_alt), each distinct plot once
    meta: Dict[str, object] = field(default_factory=dict)


def _make_one_sv(args):
    """Everything about SV ``i`` derives from ``default_rng([seed, i])``: any subset of a workload can be
    regenerated on its own (the reference arm of bench.py scores a prefix of the same SV list)."""
    (seed, i, types, size_range, reads_per_sv, err, het_frac, homref_frac, max_miss, lowercase_every, k_choices) = args
    rng = np.random.default_rng([seed, i])
    st = types[i % len(types)]
    ln = int(rng.integers(size_range[0], size_range[1] + 1))
    u = rng.random()
    gt = 0 if u < homref_frac else (1 if u < homref_frac + het_frac else 2)
    lc = 0.3 if (lowercase_every and i % lowercase_every == 0) else 0.0
    case = make_sv_case(rng, st, ln, gt, k=int(k_choices[i % len(k_choices)]),
                        micro_indel=bool(rng.random() < 0.12), lowercase_frac=lc)
    if gt == 2:
        from_alt = np.ones(reads_per_sv, dtype=bool)
    elif gt == 1:
        from_alt = rng.random(reads_per_sv) < 0.5
    else:
        from_alt = np.zeros(reads_per_sv, dtype=bool)
    miss = rng.integers(0, max_miss + 1, size=reads_per_sv) if max_miss else np.zeros(reads_per_sv, dtype=np.int64)
    want = case.read_window - miss
    o_alt = len(case.hap_ref)
    hap_len = np.where(from_alt, len(case.hap_alt), len(case.hap_ref))
    need = np.minimum((want * 1.12).astype(np.int64) + 60, hap_len - miss)
    start = np.where(from_alt, o_alt, 0) + miss
    reads, roff = simulate_reads(rng, np.concatenate([case.hap_ref, case.hap_alt]), start, need, want, err=err)
    lens = np.diff(roff)
    n = np.maximum(lens - case.k + 1, 0)
    m = np.maximum(len(case.ref_seq) - miss - case.k + 1, 0) + np.maximum(len(case.alt_seq) - miss - case.k + 1, 0)
    return (st, ln, gt, case.ref_seq, case.alt_seq, reads, lens, miss.astype(np.int32), case.k, case.mode,
            int((n * m).sum()))


def make_workload(n_sv: int, seed: int = 20261018, types: Sequence[str] = SV_TYPES,
                  size_range: Tuple[int, int] = (50, 5000), reads_per_sv: int = 20, err: float = 0.15,
                  het_frac: float = 0.5, homref_frac: float = 0.1, max_miss: int = 0,
                  lowercase_every: int = 0, k_choices: Sequence[int] = (10,), workers: int = 0,
                  first_sv: int = 0) -> Workload:
    """SVs ``first_sv .. first_sv+n_sv-1`` of the seeded SV list (types cycled, sizes uniform in
    ``size_range``), ``reads_per_sv`` reads each.  ``workers`` > 1 generates in that many processes."""
    types = tuple(types); k_choices = tuple(k_choices)
    jobs = [(seed, i, types, tuple(size_range), reads_per_sv, err, het_frac, homref_frac, max_miss,
             lowercase_every, k_choices) for i in range(first_sv, first_sv + n_sv)]
    if workers and workers > 1 and n_sv >= 4 * workers:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            outs = pool.map(_make_one_sv, jobs, chunksize=max(1, n_sv // (workers * 8)))
    else:
        outs = [_make_one_sv(j) for j in jobs]
    seq_parts: List[np.ndarray] = []
    seq_lens: List[np.ndarray] = []
    t_read, t_ref, t_alt, t_miss, t_k, t_mode = [], [], [], [], [], []
    sv_type: List[str] = []
    sv_len = np.zeros(n_sv, dtype=np.int64)
    sv_gt = np.zeros(n_sv, dtype=np.int8)
    cells = 0
    n_seq = 0
    for j, (st, ln, gt, ref, alt, reads, lens, miss, k, mode, c) in enumerate(outs):
        sv_type.append(st); sv_len[j] = ln; sv_gt[j] = gt; cells += c
        seq_parts += [ref, alt, reads]
        seq_lens.append(np.concatenate([[len(ref), len(alt)], lens]).astype(np.int64))
        ids = np.arange(n_seq + 2, n_seq + 2 + reads_per_sv, dtype=np.int32)
        t_read.append(ids)
        t_ref.append(np.full(reads_per_sv, n_seq, np.int32)); t_alt.append(np.full(reads_per_sv, n_seq + 1, np.int32))
        t_miss.append(miss)
        t_k.append(np.full(reads_per_sv, k, np.uint8)); t_mode.append(np.full(reads_per_sv, mode, np.uint8))
        n_seq += 2 + reads_per_sv
    seq_off = np.zeros(n_seq + 1, dtype=np.int64)
    if seq_lens:
        np.cumsum(np.concatenate(seq_lens), out=seq_off[1:])
    sv_off = np.arange(n_sv + 1, dtype=np.int64) * reads_per_sv
    cat = lambda parts, dt: (np.concatenate(parts) if parts else np.zeros(0, dt))
    batch = PackedBatch(cat(seq_parts, np.uint8), seq_off, cat(t_read, np.int32), cat(t_ref, np.int32),
                        cat(t_alt, np.int32), cat(t_miss, np.int32), cat(t_k, np.uint8), cat(t_mode, np.uint8),
                        sv_off).validate()
    return Workload(batch, sv_type, sv_len, sv_gt, np.full(n_sv, reads_per_sv, np.int32), cells,
                    {"seed": seed, "types": list(types), "size_range": list(size_range), "err": err,
                     "reads_per_sv": reads_per_sv, "first_sv": first_sv})
