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


# ----------------------------------------------------------------------------------------------------------
# complex events (BASELINE config 3): the structures the reference's complex-event drivers build
# ----------------------------------------------------------------------------------------------------------
# kind -> (number of blocks, alternative alleles in the reference's block-letter notation, `^` = inverted).
# The driver recipes: vapor_del_inv_Vapor (Simple_function.pyx:1557-1593), vapor_dup_inv_VapoR (:1595-1669),
# vapor_simple_disdup_Vapor (:1786-1854), vapor_CANNOT_CLASSIFY_VapoR (:1490-1555: every distinct alternative
# haplotype is scored against the reference structure with the SAME reads and all scores are pooled; REDEF when a
# block letter repeats in the allele, else ABS, :1519-1523), junction fallback windows of 2 x flank (:1537-1555).
COMPLEX_KINDS = {
    "DEL_INV":      (2, ("b^",)),                 # ab -> b^        README.md:80   mode ABS  (:1577-1581)
    "DUP_INV":      (2, ("aba^",)),               # ab -> aba^      README.md:82   mode REDEF (:1630-1634)
    "DUP_INV_L":    (2, ("b^ab",)),               # insert point left of the block (:1611)
    "DISDUP":       (2, ("aba",)),                # ab -> aba       README.md:79   mode REDEF (:1815-1819)
    "DISDUP_L":     (2, ("bab",)),
    "DEL_DUP_INV":  (2, ("a", "bb^")),            # ab/ab_a/bb^     README.md:81   two alternative haplotypes
    "CPX_3BLOCK":   (3, ("ac^b",)),               # one allele, three blocks
    "CPX_2ALLELE":  (3, ("ac", "ab^c")),          # two alternative haplotypes over three blocks
    "CPX_3ALLELE":  (3, ("ac", "abbc", "c^b^a^")),  # three alternative haplotypes (one with a repeated block)
    "JUNCTION":     (2, ("b^",)),                 # >= 10 kb event: 1 kb junction windows, mode W10 (:1537-1555, 1671-1690)
}
COMPLEX_MIX = ("DEL_INV", "DUP_INV", "DISDUP", "DEL_DUP_INV", "CPX_2ALLELE", "DUP_INV_L", "DISDUP_L", "CPX_3BLOCK",
               "DEL_INV", "DUP_INV", "DISDUP", "DEL_DUP_INV", "CPX_3ALLELE", "JUNCTION")


def _letters(allele: str) -> List[str]:
    out: List[str] = []
    for ch in allele:
        if ch == "^":
            out[-1] += ch
        else:
            out.append(ch)
    return out


@dataclass
class ComplexShape:
    """Everything about a complex event that does not need sequence: enough to know its cost."""
    kind: str
    blocks: Tuple[int, ...]
    alleles: Tuple[str, ...]
    flank: int
    k: int
    genotype: int
    len_ref: int
    len_alt: Tuple[int, ...]
    modes: Tuple[int, ...]
    read_window: int


def draw_complex_shape(rng: np.random.Generator, kind: str, size_range: Tuple[int, int], k: int,
                       het_frac: float, homref_frac: float) -> ComplexShape:
    nblk, alleles = COMPLEX_KINDS[kind]
    total = int(rng.integers(size_range[0], size_range[1] + 1))
    cuts = np.sort(rng.integers(max(1, total // 8), max(2, total - total // 8), size=nblk - 1))
    edges = [0] + [int(c) for c in cuts] + [total]
    blocks = tuple(max(20, edges[i + 1] - edges[i]) for i in range(nblk))
    u = rng.random()
    gt = 0 if u < homref_frac else (1 if u < homref_frac + het_frac else 2)
    f = flank_length_calculate(sum(blocks))
    if kind == "JUNCTION":
        f = 500
        # ref = [s-f, s+f] (2f+1 bases), alt = ref[:f] + RC([e-f, e]) (2f+1 bases), reads cut to 2f
        return ComplexShape(kind, blocks, alleles, f, k, gt, 2 * f + 1, (2 * f + 1,), (MODE_W10,), 2 * f)
    L = dict(zip("abc", blocks))
    len_ref = sum(blocks) + 2 * f + 1
    len_alt, modes = [], []
    for al in alleles:
        let = _letters(al)
        len_alt.append(2 * f + sum(L[x[0]] + 1 for x in let))            # every block is fetched inclusive (+1 base)
        repeated = max(al.count(c) for c in al if c != "^") > 1
        modes.append(MODE_REDEF if repeated else MODE_ABS)
    if kind in ("DUP_INV", "DUP_INV_L", "DISDUP", "DISDUP_L"):
        dup = L["a"] if kind in ("DUP_INV", "DISDUP") else L["b"]
        read_window = sum(blocks) + dup + 2 * f                            # :1093 / :1808: event + the extra copy
    elif kind == "DEL_INV":
        read_window = len_alt[0]                                           # :1575: start + len(alt) - 2 flank (+ flanks)
    else:
        read_window = sum(blocks) + 2 * f                                  # :1137: the whole event +- flank
    return ComplexShape(kind, blocks, alleles, f, k, gt, len_ref, tuple(len_alt), tuple(modes), read_window)


def complex_cost(sh: ComplexShape, reads_per_sv: int) -> int:
    n = max(sh.read_window - sh.k + 1, 0)
    m = sum(max(sh.len_ref - sh.k + 1, 0) + max(la - sh.k + 1, 0) for la in sh.len_alt)
    return reads_per_sv * n * m


def _make_complex_sv(rng, sh: ComplexShape, reads_per_sv: int, err: float):
    """Sequences and reads of one complex event: (ref, [alt...], reads, lens, hap sources)."""
    f = sh.flank
    span = sum(sh.blocks)
    margin = 2 * span + 4 * f + 400
    G = random_dna(rng, f + span + f + margin)
    if sh.kind == "JUNCTION":
        s, e = f, f + span
        ref = G[s - f:s + f + 1]
        alt = np.concatenate([ref[:f], revcomp(G[e - f:e + 1])])
        alts = [alt]
        haps = [np.concatenate([G[:s], revcomp(G[s:e]), G[e:]])]
    else:
        starts = np.concatenate([[0], np.cumsum(sh.blocks)]) + f
        blk = {c: (int(starts[i]), int(starts[i + 1])) for i, c in enumerate("abc"[:len(sh.blocks)])}
        ref = G[0:f + span + f + 1]
        alts, haps = [], []
        for al in sh.alleles:
            parts, hparts = [ref[:f]], [G[:f]]
            for x in _letters(al):
                a, b = blk[x[0]]
                piece = G[a:b + 1]                                      # ref_seq_readin is inclusive at both ends (:1206)
                hp = G[a:b]
                parts.append(revcomp(piece) if "^" in x else piece)
                hparts.append(revcomp(hp) if "^" in x else hp)
            parts.append(ref[len(ref) - f:])
            hparts.append(G[f + span:])
            alts.append(np.concatenate(parts))
            haps.append(np.concatenate(hparts))
    # reads: hom-ref -> all from the reference haplotype; het -> half from the first alternative haplotype;
    # hom-alt -> spread over the alternative haplotypes (a 1/2 call when there are several)
    n_alt = len(haps)
    if sh.genotype == 0:
        src = np.zeros(reads_per_sv, dtype=np.int64)
    elif sh.genotype == 1:
        src = np.where(rng.random(reads_per_sv) < 0.5, 1, 0).astype(np.int64)
    else:
        src = 1 + (np.arange(reads_per_sv) % n_alt)
    hap_all = [G] + haps
    offs = np.concatenate([[0], np.cumsum([len(h) for h in hap_all])])
    want = np.full(reads_per_sv, sh.read_window, dtype=np.int64)
    hap_len = np.array([len(hap_all[i]) for i in src])
    need = np.minimum((want * 1.12).astype(np.int64) + 60, hap_len)
    # a haplotype shorter than the read window (deletion alleles): pad the window from the reference margin instead
    short = need < (want * 1.05).astype(np.int64) + 20
    src = np.where(short, 0, src)
    hap_len = np.array([len(hap_all[i]) for i in src])
    need = np.minimum((want * 1.12).astype(np.int64) + 60, hap_len)
    reads, roff = simulate_reads(rng, np.concatenate(hap_all), offs[src], need, want, err=err)
    return ref, alts, reads, np.diff(roff)


@dataclass
class Workload:
    batch: PackedBatch
    sv_type: List[str]
    sv_len: np.ndarray
    sv_genotype: np.ndarray
    reads_per_sv: np.ndarray
    cells: int                        # sum over reads of n*(m_ref+m_alt), each distinct plot once
    meta: Dict[str, object] = field(default_factory=dict)


def simple_cost(svtype: str, svlen: int, k: int, reads_per_sv: int) -> int:
    """Recurrence cells of one simple SV with miss_bp = 0, from its type and length alone (SURVEY 8a recipe table)."""
    f = flank_length_calculate(svlen)
    if svtype == "DEL":
        lr, la, rw = svlen + 2 * f + 1, 2 * f, 2 * f
    elif svtype == "INV":
        lr, la, rw = svlen + 2 * f + 1, svlen + 2 * f + 1, svlen + 2 * f
    elif svtype == "TANDUP":
        lr, la, rw = svlen + 2 * f + 1, 2 * (svlen + 1) + 2 * f, 2 * svlen + 2 * f
    elif svtype == "INS":
        lr, la, rw = (2 * f + svlen + 1 if svlen < 5000 else 2 * f + 1), svlen + 2 * f + 2, svlen + 2 * f
    else:
        raise ValueError(svtype)
    return reads_per_sv * max(rw - k + 1, 0) * (max(lr - k + 1, 0) + max(la - k + 1, 0))


LARGE_MODES = {"INV": MODE_ABS, "TANDUP": MODE_REDEF, "DEL": MODE_W10, "INS": MODE_W10}


def _sv_shape(seed, i, recipe, types, size_range, k_choices, het_frac, homref_frac):
    """The draws that fix SV ``i``'s shape, and the generator positioned right after them."""
    rng = np.random.default_rng([seed, i])
    k = int(k_choices[i % len(k_choices)])
    if recipe == "complex":
        kind = types[i % len(types)]
        return rng, draw_complex_shape(rng, kind, size_range, k, het_frac, homref_frac)
    st = types[i % len(types)]
    ln = int(rng.integers(size_range[0], size_range[1] + 1))
    return rng, (st, ln, k)


def workload_costs(n_sv: int, seed: int, recipe: str = "simple", types: Sequence[str] = SV_TYPES,
                   size_range: Tuple[int, int] = (50, 5000), reads_per_sv: int = 20, k_choices: Sequence[int] = (10,),
                   het_frac: float = 0.5, homref_frac: float = 0.1, first_sv: int = 0, with_tasks: bool = False):
    """Recurrence cells of SVs ``first_sv .. first_sv+n_sv-1`` of the seeded list WITHOUT generating any sequence
    (``max_miss`` = 0): what ``multi.partition_svs`` needs to deal a long SV list to the GPUs before any rank
    builds its share.  ``with_tasks``: also the number of tasks (reads x alternative haplotypes) of every SV."""
    out = np.zeros(n_sv, dtype=np.int64)
    ntask = np.full(n_sv, reads_per_sv, dtype=np.int64)
    if recipe == "complex" and tuple(types) == SV_TYPES:
        types = COMPLEX_MIX
    types = tuple(types); k_choices = tuple(k_choices)
    for j in range(n_sv):
        _, sh = _sv_shape(seed, first_sv + j, recipe, types, tuple(size_range), k_choices, het_frac, homref_frac)
        out[j] = complex_cost(sh, reads_per_sv) if recipe == "complex" else simple_cost(sh[0], sh[1], sh[2], reads_per_sv)
        if recipe == "complex":
            ntask[j] = reads_per_sv * len(sh.alleles)
    return (out, ntask) if with_tasks else out


def _make_one_sv(args):
    """Everything about SV ``i`` derives from ``default_rng([seed, i])``: any subset of a workload can be
    regenerated on its own (the reference arm of bench.py scores a sample of the same SV list).
    Returns (type, length, genotype, ref, [alt...], reads, read lengths, miss, k, [mode...], cells)."""
    (seed, i, types, size_range, reads_per_sv, err, het_frac, homref_frac, max_miss, lowercase_every, k_choices, recipe) = args
    rng, sh = _sv_shape(seed, i, recipe, types, size_range, k_choices, het_frac, homref_frac)
    if recipe == "complex":
        ref, alts, reads, lens = _make_complex_sv(rng, sh, reads_per_sv, err)
        n = np.maximum(lens - sh.k + 1, 0)
        m = sum(max(len(ref) - sh.k + 1, 0) + max(len(a) - sh.k + 1, 0) for a in alts)
        return (sh.kind, int(sum(sh.blocks)), sh.genotype, ref, alts, reads, lens, np.zeros(reads_per_sv, np.int32), sh.k,
                list(sh.modes), int((n * m).sum()))
    st, ln, k = sh
    u = rng.random()
    gt = 0 if u < homref_frac else (1 if u < homref_frac + het_frac else 2)
    lc = 0.3 if (lowercase_every and i % lowercase_every == 0) else 0.0
    case = make_sv_case(rng, st, ln, gt, k=k, micro_indel=bool(rng.random() < 0.12), lowercase_frac=lc)
    if recipe == "large":
        case.mode = LARGE_MODES[st]
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
    return (st, ln, gt, case.ref_seq, [case.alt_seq], reads, lens, miss.astype(np.int32), case.k, [case.mode],
            int((n * m).sum()))


def make_workload(n_sv: int, seed: int = 20261018, types: Sequence[str] = SV_TYPES,
                  size_range: Tuple[int, int] = (50, 5000), reads_per_sv: int = 20, err: float = 0.15,
                  het_frac: float = 0.5, homref_frac: float = 0.1, max_miss: int = 0,
                  lowercase_every: int = 0, k_choices: Sequence[int] = (10,), workers: int = 0,
                  first_sv: int = 0, sv_ids: Sequence[int] = None, recipe: str = "simple") -> Workload:
    """SVs ``first_sv .. first_sv+n_sv-1`` (or exactly ``sv_ids``) of the seeded SV list, ``reads_per_sv`` reads
    each.  ``recipe``: "simple" (types cycled over DEL/TANDUP/INV/INS, sizes uniform in ``size_range``; BASELINE
    configs 2 and 5), "large" (the same SV types at 10-100 kb with a fixed scoring mode per type; config 4),
    "complex" (``types`` = keys of COMPLEX_KINDS cycled: several blocks, one task group per distinct alternative
    haplotype, the same reads scored against each; config 3).  ``workers`` > 1 generates in that many processes."""
    if recipe == "complex" and tuple(types) == SV_TYPES:
        types = COMPLEX_MIX
    types = tuple(types); k_choices = tuple(k_choices)
    ids = list(range(first_sv, first_sv + n_sv)) if sv_ids is None else [int(i) for i in sv_ids]
    n_sv = len(ids)
    jobs = [(seed, i, types, tuple(size_range), reads_per_sv, err, het_frac, homref_frac, max_miss,
             lowercase_every, k_choices, recipe) for i in ids]
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
    sv_off = np.zeros(n_sv + 1, dtype=np.int64)
    cells = 0
    n_seq = 0
    for j, (st, ln, gt, ref, alts, reads, lens, miss, k, modes, c) in enumerate(outs):
        sv_type.append(st); sv_len[j] = ln; sv_gt[j] = gt; cells += c
        na = len(alts)
        seq_parts += [ref] + list(alts) + [reads]
        seq_lens.append(np.concatenate([[len(ref)], [len(a) for a in alts], lens]).astype(np.int64))
        ids_r = np.arange(n_seq + 1 + na, n_seq + 1 + na + reads_per_sv, dtype=np.int32)
        for a in range(na):                              # every alternative haplotype against the same reads
            t_read.append(ids_r)
            t_ref.append(np.full(reads_per_sv, n_seq, np.int32)); t_alt.append(np.full(reads_per_sv, n_seq + 1 + a, np.int32))
            t_miss.append(miss)
            t_k.append(np.full(reads_per_sv, k, np.uint8)); t_mode.append(np.full(reads_per_sv, modes[a], np.uint8))
        sv_off[j + 1] = sv_off[j] + na * reads_per_sv
        n_seq += 1 + na + reads_per_sv
    seq_off = np.zeros(n_seq + 1, dtype=np.int64)
    if seq_lens:
        np.cumsum(np.concatenate(seq_lens), out=seq_off[1:])
    cat = lambda parts, dt: (np.concatenate(parts) if parts else np.zeros(0, dt))
    batch = PackedBatch(cat(seq_parts, np.uint8), seq_off, cat(t_read, np.int32), cat(t_ref, np.int32),
                        cat(t_alt, np.int32), cat(t_miss, np.int32), cat(t_k, np.uint8), cat(t_mode, np.uint8),
                        sv_off).validate()
    return Workload(batch, sv_type, sv_len, sv_gt, np.diff(sv_off).astype(np.int32), cells,
                    {"seed": seed, "types": list(types), "size_range": list(size_range), "err": err,
                     "reads_per_sv": reads_per_sv, "first_sv": first_sv, "recipe": recipe, "sv_ids": ids})
