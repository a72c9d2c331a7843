"""Host-side mirror of the reference's plugin namespace ``vapor_vali.Simple_function``.

The reference CLI does ``from vapor_vali.Simple_function import *`` (vapor_vali/vapor:322,374,470) and
calls the per-SV-type drivers one SV at a time; every driver loops over its reads and scores them one by
one on the CPU.  This module keeps the same names, argument meaning and sentinel conventions
(``[0, 0]`` = read not scorable, ``['Error', 'Error']`` = window unusable, ``'NA'`` rows), but nothing in
here scores anything: every recurrence plot, score, QS/GS/GT/GQ comes from the CUDA library behind
``include/vapor_b200.h``.  Without that library (or without a GPU) the calls raise
``VaporNativeError`` -- there is no CPU fallback.

Two ways in:

* drop-in, one call at a time -- ``dotdata``, ``calcu_vapor_single_read_score_*``, ``window_size_refine``,
  ``vapor_simple_del_Vapor(...)`` ... ``result_organize_ins``, ``write_output_main``: same signatures as the
  reference, each making a small GPU call;
* batched -- ``Session.run_events`` drives the drivers of a whole SV set as coroutines: each driver
  ``yield``s its window-QC and scoring requests, the session gathers the requests of all SVs, ships each
  kind to the GPU in one call, and resumes the drivers with the answers.  The data-dependent control
  flow of every driver (fallback to junction windows, per-allele loops) is preserved, the number of GPU
  calls is the depth of that control flow (3-5), not the number of SVs or reads.  ``vapor_b200.cli`` uses this.

Host logic restated from the reference (cited per function) is the part that *defines kernel inputs*:
window coordinates, read chopping, structure strings.  It is restated, not copied; quirks that change
inputs are kept and marked ``# quirk``.
"""
from __future__ import annotations

import math
import os
import re
import threading
from typing import Dict, Generator, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import seqio
from .engine import (Batch, Engine, GT_NAMES, MODE_ABS, MODE_ABS_AND_W10, MODE_REDEF, MODE_W10)
from ._native import VAPOR_ST_BADREAD, VAPOR_ST_SCORED

# module constants of the reference (Simple_function.pyx:19-26)
invert_base = {'A': 'T', 'T': 'A', 'C': 'G', 'G': 'C', 'N': 'N', 'a': 't', 't': 'a', 'c': 'g', 'g': 'c', 'n': 'n'}
default_flank_length = 500
default_read_length = 4000
default_max_sv_test = 10000
region_QC_Cff_default = 0.4                     # window_size_refine default (Simple_function.pyx:2030)

__all__ = [
    "invert_base", "default_flank_length", "default_read_length", "default_max_sv_test",
    "Session", "get_session", "set_session",
    "dotdata", "kmerhits", "window_size_refine", "qual_check_repetitive_region",
    "calcu_vapor_single_read_score_abs_dis_m1b", "calcu_vapor_single_read_score_within_10Perc_m1b",
    "calcu_vapor_single_read_score_directed_dis_m1b_redefine_diagnal",
    "vapor_simple_del_Vapor", "vapor_simple_inv_Vapor", "vapor_simple_tandup_Vapor", "vapor_simple_ins_Vapor",
    "vapor_simple_disdup_Vapor", "vapor_del_inv_Vapor", "vapor_dup_inv_VapoR", "vapor_long_del_inv",
    "vapor_CANNOT_CLASSIFY_VapoR",
    "result_organize_ins", "gt_estimate_log_likelihood", "log_likelihood_calcu",
    "write_output_initiate", "write_output_main", "vcf_vapor_modify", "vcf_rec_hash_modify",
    "ref_seq_readin", "chop_pacbio_read_by_pos", "cigar2alignstart_by_pos", "minimize_pacbio_read_list",
    "simple_del_chop_pacbio_read_simple_short", "simple_chop_pacbio_read_simple_short", "bam_in_decide",
    "flank_length_calculate", "reverse", "complementary", "chromos_readin", "path_modify", "path_mkdir",
    "svtype_extract", "sv_len_extract", "sv_seq_extract", "sv_insert_point_define", "chr_start_end_extract",
    "INS_length_detect", "polarity_detect", "list_unify", "unify_list", "letter_split", "block_subsplot",
    "bp_to_chr_hash", "block_around_check", "make_event_figure_1",
]


# ====================================================================================================
# small string / coordinate helpers (host only)
# ====================================================================================================
def reverse(seq):
    """Simple_function.pyx:1173."""
    return seq[::-1]


_COMP = {a: b for a, b in zip("ATGCNatgcn", "TACGNtacgn")}


_COMP_TABLE = {c: (ord(_COMP[chr(c)]) if chr(c) in _COMP else None) for c in range(0x110000 if False else 256)}


def complementary(seq):
    """Simple_function.pyx:471-478.  # quirk: characters outside ATGCN/atgcn are *dropped*, not kept."""
    if seq.isascii():
        return seq.translate(_COMP_TABLE)
    return "".join(_COMP[c] for c in seq if c in _COMP)


def flank_length_calculate(bps):
    """Simple_function.pyx:794-802: the event length below 500 bp, else 500."""
    span = int(bps[-1]) - int(bps[1])
    return span if span < 500 else 500


def path_modify(path):
    """Simple_function.pyx:1142-1145."""
    return path if path.endswith("/") else path + "/"


def path_mkdir(path):
    """Simple_function.pyx:1138-1140 (``mkdir`` without the shell)."""
    if not os.path.isdir(path):
        os.makedirs(path, exist_ok=True)


def chromos_readin(ref):
    """Contig names from ``ref.fai`` (Simple_function.pyx:356-363)."""
    if not os.path.exists(ref + ".fai"):
        seqio.build_fai(ref)
    with open(ref + ".fai") as f:
        return [line.split()[0] for line in f if line.strip()]


def list_unify(items):
    """Simple_function.pyx:1021-1025."""
    out = []
    for i in items:
        if i not in out:
            out.append(i)
    return out


unify_list = list_unify                          # Simple_function.pyx:1483-1488 is the same function


def letter_split(let):
    """'c^ba' -> ['c^', 'b', 'a'] (Simple_function.pyx:1013-1019)."""
    out: List[str] = []
    for ch in let:
        if ch == "^":
            out[-1] += ch
        else:
            out.append(ch)
    return out


def block_subsplot(bp_list, chromos):
    """['chr1','10','20','chr2','5','9'] -> [['chr1',10,20],['chr2',5,9]] (Simple_function.pyx:147-153)."""
    out: List[list] = []
    for x in bp_list:
        if x in chromos:
            out.append([x])
        else:
            out[-1].append(int(x))
    return out


def bp_to_chr_hash(bps, chromos, flank_length=500):
    """Letter -> [chrom, start, end] for consecutive breakpoints, plus the '-' and '+' flanks
    (Simple_function.pyx:98-114; the mixed int/str element types of '+' and '-' are kept)."""
    groups: List[list] = []
    for i in bps:
        if i in chromos:
            groups.append([i])
        else:
            groups[-1].append(i)
    out = {}
    rec = -1
    for g in groups:
        for j in range(len(g[2:])):
            rec += 1
            out[chr(97 + rec)] = [g[0], g[j + 1], g[j + 2]]
    last = out[sorted(out.keys())[-1]]
    out["+"] = [last[0], last[2], str(int(last[2]) + flank_length)]
    out["-"] = [out["a"][0], str(int(out["a"][1]) - flank_length), int(out["a"][1])]
    return out


def block_around_check(alt_allele, ref_allele):
    """Junctions of the alternative allele absent from the reference allele (Simple_function.pyx:91-96)."""
    n = len(letter_split(alt_allele)) + 1
    alt_path = ["-"] + letter_split(alt_allele) + ["+"]
    ref_path = ["-"] + letter_split(ref_allele) + ["+"]
    alt_j = [alt_path[j:j + 2] for j in range(n)]
    ref_j = [ref_path[j:j + 2] for j in range(n)]
    return [j for j in alt_j if j not in ref_j]


# ---- VCF field helpers (Simple_function.pyx:365-370, 833-838, 1147-1153, 1424-1456) ----------------
def chr_start_end_extract(pin):
    out = [pin[0], int(pin[1])]
    for x in pin[7].split(";"):
        if x.split("=")[0] == "END":
            out.append(int(x.split("=")[1]))
    return out


def svtype_extract(pin):
    svtype = ""
    for x in pin[7].split(";"):
        if "SVTYPE" in x:
            svtype = x.split("=")[1]
    return svtype if svtype != "" else pin[4].replace("<", "").replace(">", "")


def sv_len_extract(pin):
    out = ""
    for x in pin[7].split(";"):
        if "SVLEN" in x:
            out = x.split("=")[1]
    return out if out != "" else 0


def sv_seq_extract(pin):
    seq = ""
    for x in pin[7].split(";"):
        if x[:4] == "SEQ=":
            seq = x.split("=")[1]
    return seq


def sv_insert_point_define(pin):
    out = [0, 0]
    for x in pin[7].split(";"):
        if "insert_point=" in x:
            out = x.split("=")[1].split(":")
    return out


def INS_length_detect(pin):
    out = 0
    for x in pin[7].split(";"):
        if "SVLEN=" in x:
            out = int(x.split("=")[1])
    return out


def polarity_detect(pin):
    out = "+"
    for x in pin[7].split(";"):
        if "MEIINFO=" in x:
            out = x.split(",")[-1]
    return out


# ====================================================================================================
# host I/O that defines kernel inputs
# ====================================================================================================
def ref_seq_readin(ref, chrom, start, end, reverse_flag="FALSE"):
    """``samtools faidx ref chrom:start-end`` joined into one string (Simple_function.pyx:1203-1217),
    answered in-process by ``seqio``; ``reverse_flag != 'FALSE'`` returns the reverse complement."""
    seq = seqio.faidx(ref, chrom, int(start), int(end))
    return seq if reverse_flag == "FALSE" else reverse(complementary(seq))


_CIGAR = re.compile(r"(\d+)([MIDNSHP=X])")


def _cigar_walk(cigar, align_start, start):
    """The reference's loop, op by op (Simple_function.pyx:316-330): the definition the table-driven version below
    is tested against."""
    read_rec, align_rec = 0, align_start
    last = ""
    for m in _CIGAR.finditer(cigar):
        n, op = int(m.group(1)), m.group(2)
        if op == "S":
            read_rec += n
        elif op in "M=":
            read_rec += n
            align_rec += n
        elif op == "D":
            align_rec += n
        elif op == "I":
            read_rec += n
        last = op
        if align_rec > start - 1:
            break
    return read_rec, align_rec, last


_READ_ADV = np.zeros(256, dtype=np.int64)
_REF_ADV = np.zeros(256, dtype=np.int64)
for _c in "SM=I":
    _READ_ADV[ord(_c)] = 1
for _c in "M=D":
    _REF_ADV[ord(_c)] = 1
_IS_OP = np.zeros(256, dtype=bool)
_IS_OP[np.frombuffer(b"MIDNSHP=X", dtype=np.uint8)] = True
_POW10 = 10 ** np.arange(19, dtype=np.int64)
_cigar_tables: Dict[str, tuple] = {}


def _cigar_table(cigar):
    """(ops, cumulative read advance, cumulative reference advance) of a CIGAR string, parsed once with numpy and
    cached: a CLR read has thousands of operations and is queried by several windows."""
    t = _cigar_tables.get(cigar)
    if t is None:
        b = np.frombuffer(cigar.encode("latin-1"), dtype=np.uint8)
        is_op = _IS_OP[b]
        idx = np.nonzero(is_op)[0]
        dig_pos = np.nonzero(~is_op)[0]
        dig = b[dig_pos].astype(np.int64) - 48
        if len(idx) == 0 or (dig < 0).any() or (dig > 9).any() or (len(b) and not is_op[-1]):
            t = None if len(idx) == 0 else False
        if t is None and len(idx):
            starts = np.concatenate([[0], idx[:-1] + 1])
            ndig = idx - starts
            if (ndig <= 0).any() or (ndig > 18).any():
                t = False
            else:
                owner = np.repeat(np.arange(len(idx)), ndig)
                lens = np.add.reduceat(dig * _POW10[idx[owner] - dig_pos - 1], starts - np.arange(len(idx)))
                ops = b[idx]
                t = (ops, np.cumsum(lens * _READ_ADV[ops]), np.cumsum(lens * _REF_ADV[ops]))
        if len(_cigar_tables) > 200000:
            _cigar_tables.clear()
        _cigar_tables[cigar] = t
    return t


def cigar2alignstart_by_pos(cigar, align_start, start, end):
    """Walk the CIGAR up to reference position ``start``: returns [read offset, missed bases]
    (Simple_function.pyx:309-337).  # quirk: X, N, H, P advance nothing."""
    t = _cigar_table(cigar)
    if not t:                                   # '*', empty or malformed: the plain walk defines the answer
        read_rec, align_rec, last = _cigar_walk(cigar, align_start, start)
    else:
        ops, cum_read, cum_ref = t
        # first operation after which align_start + cum_ref > start - 1, else the last one
        i = int(np.searchsorted(cum_ref, start - 1 - align_start, side="right"))
        i = min(i, len(ops) - 1)
        read_rec, align_rec, last = int(cum_read[i]), align_start + int(cum_ref[i]), chr(ops[i])
    start_dis = int(align_rec) - start
    if last != "" and last in "M=":
        return [read_rec - start_dis, 0]
    return [read_rec, start_dis]


def chop_pacbio_read_by_pos(bam_in_new, chrom, start, end, flank_length):
    """Reads of ``samtools view bam chrom:start-end`` cut to the window: ``[[read, miss_bp, qname], ...]``
    (Simple_function.pyx:339-354).  Keeps alignments starting at or before ``start`` whose missed bases do
    not exceed half the flank and that run past the window end."""
    if seqio.native_enabled():                  # csrc/hostio.cpp: the same rules in native code
        return seqio.chop_reads([bam_in_new], chrom, start, end, flank_length, max_reads=0)
    out = []
    for rec in seqio.view(bam_in_new, chrom, start, end):
        if rec.pos < start + 1:
            align_start, miss_bp = cigar2alignstart_by_pos(rec.cigar, rec.pos, start, end)
            if not miss_bp > flank_length / 2:
                target = rec.seq[align_start:]
                if len(target) > end - start - miss_bp:
                    out.append([target[:end - start - miss_bp], miss_bp, rec.qname])
    return out


def minimize_pacbio_read_list(x, ideal_list_length=20):
    """At most 20 reads, smallest ``miss_bp`` first, file order inside one ``miss_bp`` (Simple_function.pyx:1091-1102)."""
    if len(x) <= ideal_list_length:
        return x
    by_miss: Dict[int, list] = {}
    for y in x:
        by_miss.setdefault(y[1], []).append(y)
    out: list = []
    for m in sorted(by_miss):
        if len(out) < ideal_list_length:
            out += by_miss[m]
    return out[:ideal_list_length]


def bam_in_decide(bam_in, bps):
    """One file, or every file of the directory matching an ``XXX`` / ``*`` pattern (Simple_function.pyx:69-89)."""
    if os.path.isfile(bam_in):
        return [bam_in]
    folder = "/".join(bam_in.split("/")[:-1]) + "/"
    name = bam_in.split("/")[-1]
    if "XXX" in name:
        keys = name.split("XXX")
    elif "*" in name:
        keys = name.split("*")
    else:
        print("Error: invalid name for pacbio files !")
        return []
    ext = bam_in.split(".")[-1]
    return [folder + f for f in os.listdir(folder) if f.split(".")[-1] == ext and all(k in f for k in keys)]


def simple_del_chop_pacbio_read_simple_short(bam_in, sv_info, flank_length):
    """Reads around the left breakpoint: window ``[s - f, s + f]`` (Simple_function.pyx:1378-1390)."""
    if seqio.native_enabled():                  # every file, the window cut and the 20-read cap in one native call
        return seqio.chop_reads(bam_in_decide(bam_in, sv_info), sv_info[0], int(sv_info[1]) - flank_length, int(sv_info[1]) + flank_length, flank_length)
    x = []
    for b in bam_in_decide(bam_in, sv_info):
        x += chop_pacbio_read_by_pos(b, sv_info[0], int(sv_info[1]) - flank_length, int(sv_info[1]) + flank_length, flank_length)
    return minimize_pacbio_read_list(x)


def simple_chop_pacbio_read_simple_short(bam_in, sv_info, flank_length):
    """Reads across the whole event: window ``[s - f, last + f]`` (Simple_function.pyx:1392-1401)."""
    if seqio.native_enabled():
        return seqio.chop_reads(bam_in_decide(bam_in, sv_info), sv_info[0], int(sv_info[1]) - flank_length, int(sv_info[-1]) + flank_length, flank_length)
    x = []
    for b in bam_in_decide(bam_in, sv_info):
        x += chop_pacbio_read_by_pos(b, sv_info[0], int(sv_info[1]) - flank_length, int(sv_info[-1]) + flank_length, flank_length)
    return minimize_pacbio_read_list(x)


# ====================================================================================================
# requests a driver coroutine yields, and the session that answers them on the GPU
# ====================================================================================================
XMEANS_ON_HOST = os.environ.get("VAPOR_XMEANS", "1") != "0"     # 0: size the below-diagonal dots by one bounding box


class RefineRequest:
    """``window_size_refine(seq)``: answered with ``[window_size, region_QC]`` or ``['Error', 'Error']``."""
    __slots__ = ("seq", "region_QC_Cff")

    def __init__(self, seq, region_QC_Cff=region_QC_Cff_default):
        self.seq = seq
        self.region_QC_Cff = region_QC_Cff


class ScoreRequest:
    """Score every read ``x = [read, miss_bp, qname]`` against (ref_seq, alt_seq) with k-mer size
    ``window_size`` in one of the reference's modes.  Answered with a ``ScoreAnswer``."""
    __slots__ = ("ref_seq", "alt_seq", "reads", "window_size", "mode")

    def __init__(self, ref_seq, alt_seq, reads, window_size, mode):
        self.ref_seq, self.alt_seq, self.reads, self.window_size, self.mode = ref_seq, alt_seq, reads, window_size, mode


class ScoreAnswer:
    """Per-read results of one ScoreRequest, in read order."""

    def __init__(self, stat, score, status, hits):
        self.stat, self.score, self.status, self.hits = stat, score, status, hits

    def pairs(self, which=0):
        """The ``[a, b]`` lists the reference's ``calcu_*`` function returns (``which=1``: the W10 pair of
        MODE_ABS_AND_W10).  Raises KeyError like the reference when a read holds a character invert_base rejects."""
        if (self.status == VAPOR_ST_BADREAD).any():
            raise KeyError("read holds a character outside ACGTN/acgtn (invert_base, Simple_function.pyx:1421)")
        return [_pair_to_list(self.stat[i, 2 * which], self.stat[i, 2 * which + 1]) for i in range(len(self.score))]

    def scores(self, reads=None):
        """``vapor_score_list`` as the reference's read loop builds it (``if not 0 in pair: append(1 - b/a)``,
        Simple_function.pyx:1913-1915; the simple-DEL min rule :1718-1726 for MODE_ABS_AND_W10), plus the
        read the reference would have kept as ``best_read_rec``."""
        if (self.status == VAPOR_ST_BADREAD).any():
            raise KeyError("read holds a character outside ACGTN/acgtn (invert_base, Simple_function.pyx:1421)")
        out, best = [], ""
        for i in range(len(self.score)):
            if self.status[i] == VAPOR_ST_SCORED:
                out.append(float(self.score[i]))
                if reads is not None and out[-1] == max(out):
                    best = reads[i]
        return out, best


def _pair_to_list(a, b):
    # the reference returns ints for the sentinel / count pairs and floats for means; values compare equal
    a, b = float(a), float(b)
    if a == 0 and b == 0:
        return [0, 0]
    return [a, b]


class Session:
    """Owns one ``Engine`` (one GPU) and answers driver requests in batches."""

    def __init__(self, device: int = 0, engine: Optional[Engine] = None):
        self.engine = engine if engine is not None else Engine(device)
        self.stats = {"gpu_calls": 0, "refine_requests": 0, "score_requests": 0, "reads_scored": 0, "rounds": 0}
        self.figures: List[tuple] = []            # (plt_li, scores, best_read, k, ref_seq, alt_seq, name) per event

    def close(self):
        self.engine.close()

    # ---- window_size_refine for many sequences -------------------------------------------------------
    def refine_many(self, reqs: Sequence[RefineRequest]) -> List[list]:
        """``window_size_refine`` (Simple_function.pyx:2030-2046) for every request, the self-plots of each
        k-mer size evaluated in one ``vapor_gpu_selfplot_qc`` call."""
        n = len(reqs)
        out: List[Optional[list]] = [None] * n
        seqs = [r.seq.replace("X", "") for r in reqs]
        k = [10] * n
        live = []
        for i, s in enumerate(seqs):
            if s.count("N") + s.count("n") > 100:
                out[i] = ["Error", "Error"]
            else:
                live.append(i)
        first = True
        while live:
            qc = self.engine.selfplot_qc([seqs[i] for i in live], [k[i] for i in live])
            self.stats["gpu_calls"] += 1
            nxt = []
            for row, i in zip(qc, live):
                if row[7] == VAPOR_ST_BADREAD:
                    raise KeyError("window holds a character outside ACGTN/acgtn (invert_base, Simple_function.pyx:1421)")
                H = int(row[0])
                if H == 0:
                    # len(dotdata) == 0: 'Error' at k = 10; at larger k the reference divides by zero
                    # (Simple_function.pyx:1170) -- reported as an unusable window too
                    out[i] = ["Error", "Error"]
                    continue
                lower_dots = None
                if 0.1 < float(row[2]) / float(H) < 0.5 and float(row[1]) / float(H) <= reqs[i].region_QC_Cff and XMEANS_ON_HOST:
                    # repeat-rich window: its below-diagonal dots, from the GPU, sized by the reference's X-means on the host
                    d = self.engine.dotdata(k[i], seqs[i], seqs[i])
                    self.stats["gpu_calls"] += 1
                    low = d[d[:, 0] > d[:, 1]]
                    lower_dots = (low[:, 0].tolist(), low[:, 1].tolist())
                region_qc = _qual_check_from_counts(row, len(seqs[i]), lower_dots)
                if k[i] > 30 or region_qc[0] > reqs[i].region_QC_Cff or sum(region_qc[1]) / float(len(seqs[i])) < 0.3:
                    out[i] = [k[i], region_qc]
                else:
                    k[i] += 10
                    nxt.append(i)
            live = nxt
            first = False
        return out  # type: ignore[return-value]

    # ---- scoring for many (ref, alt, reads) groups ----------------------------------------------------
    def score_many(self, reqs: Sequence[ScoreRequest]) -> List[ScoreAnswer]:
        b = Batch()
        spans = []
        for r in reqs:
            ref_id, alt_id = b.add_seq(r.ref_seq), b.add_seq(r.alt_seq)
            t0 = None
            for x in r.reads:
                t = b.add_task(b.add_seq(x[0]), ref_id, alt_id, int(x[1]), int(r.window_size), int(r.mode))
                t0 = t if t0 is None else t0
            b.end_sv(None)
            spans.append((0 if t0 is None else t0, len(r.reads)))
        pb = b.pack()
        if pb.n_task == 0:
            return [ScoreAnswer(np.zeros((0, 4)), np.zeros(0), np.zeros(0, np.uint8), np.zeros((0, 4), np.uint32)) for _ in reqs]
        res = self.engine.score(pb)
        self.stats["gpu_calls"] += 1
        self.stats["reads_scored"] += pb.n_task
        return [ScoreAnswer(res.task_stat[t0:t0 + n], res.task_score[t0:t0 + n], res.task_status[t0:t0 + n],
                            res.task_hits[t0:t0 + n]) for t0, n in spans]

    # ---- the coroutine scheduler -------------------------------------------------------------------------
    def run_events(self, coroutines: Sequence[Generator]) -> list:
        """Drive every driver coroutine to completion; returns their return values in input order."""
        n = len(coroutines)
        results: list = [None] * n
        pending: Dict[int, object] = {}
        for i, g in enumerate(coroutines):
            try:
                pending[i] = next(g)
            except StopIteration as e:
                results[i] = e.value
        while pending:
            self.stats["rounds"] += 1
            ref_ids = [i for i, r in pending.items() if isinstance(r, RefineRequest)]
            sc_ids = [i for i, r in pending.items() if isinstance(r, ScoreRequest)]
            answers: Dict[int, object] = {}
            if ref_ids:
                self.stats["refine_requests"] += len(ref_ids)
                for i, a in zip(ref_ids, self.refine_many([pending[i] for i in ref_ids])):
                    answers[i] = a
            if sc_ids:
                self.stats["score_requests"] += len(sc_ids)
                for i, a in zip(sc_ids, self.score_many([pending[i] for i in sc_ids])):
                    answers[i] = a
            nxt: Dict[int, object] = {}
            for i, a in answers.items():
                try:
                    nxt[i] = coroutines[i].send(a)
                except StopIteration as e:
                    results[i] = e.value
            pending = nxt
        return results

    def run_one(self, coroutine: Generator):
        return self.run_events([coroutine])[0]

    # ---- per-SV summaries ---------------------------------------------------------------------------------
    def summarize(self, score_lists: Sequence[Sequence[float]]):
        """QS, GS, GT, GQ for many score lists in one call (kernel 4): list of dicts, ``None`` for an empty list."""
        qs, gs, gq, gt, ns = self.engine.summarize([list(map(float, s)) for s in score_lists])
        self.stats["gpu_calls"] += 1
        out = []
        for i, s in enumerate(score_lists):
            if len(s) == 0:
                out.append(None)
                continue
            has_pos = any(float(v) > 0 for v in s)
            out.append({"QS": float(qs[i]) if has_pos else 0,          # the reference writes the int 0 (Simple_function.pyx:1228)
                        "GS": float(gs[i]), "GT": GT_NAMES[int(gt[i])], "GQ": float(gq[i]),
                        "Rec": ",".join(str(round(float(v), 2)) for v in s)})
        return out


def _qual_check_from_counts(row, seq_len, lower_dots=None):
    """``qual_check_repetitive_region`` (Simple_function.pyx:1154-1171) from the self-plot counters.

    The diagonal fraction is exact.  When 10-50 % of the dots lie below the diagonal the reference sizes them with a
    recursive X-means (:856-906, :2101-2116): ``lower_dots`` = (x list, y list) of those dots lets ``xmeans_box_sizes``
    restate that; without them the dots count as one cluster (their bounding box)."""
    H, diag, lower = int(row[0]), int(row[1]), int(row[2])
    frac = float(lower) / float(H)
    if 0.1 < frac < 0.5:
        if lower_dots is not None:
            size_cluster = xmeans_box_sizes(lower_dots[0], lower_dots[1])
        else:
            size_cluster = [float(np.sqrt((int(row[4]) - int(row[3])) * (int(row[6]) - int(row[5]))))]
    else:
        size_cluster = [0]
    return [float(diag) / float(H), size_cluster]


# ---- the reference's X-means sizing of the below-diagonal dots (host; only repeat-rich windows get here) ---------
def _bic(km, X):
    """compute_bic (Simple_function.pyx:480-516) with its clusters-with-negative-variance filter (:518-525)."""
    from scipy.spatial import distance
    centers, labels, m = km.cluster_centers_, km.labels_, km.n_clusters
    n = np.bincount(labels)
    N, d = X.shape
    var = []
    for i in range(m):
        sq = sum(distance.cdist(X[np.where(labels == i)], [centers[i]], "euclidean") ** 2)
        var.append((1.0 / (n[i] - m)) * sq if n[i] - m != 0 else float(10 ** 20) * sq)
    const_term = 0.5 * m * np.log10(N)
    keep = []
    for i in range(m):
        var[i] = [0.0 if v == -0.0 else v for v in var[i]]
        if not any(v < 0 for v in var[i]):
            keep.append(i)
    return np.sum([n[i] * np.log10(n[i]) - n[i] * np.log10(N) - ((n[i] * d) / 2) * np.log10(2 * np.pi) -
                   (n[i] / 2) * np.log10(var[i]) - ((n[i] - m) / 2) for i in keep]) - const_term


def _kmeans_split(xs, ys):
    """k_means_cluster (Simple_function.pyx:856-889): 1-4 clusters by scikit-learn KMeans, the count chosen by BIC, the
    split itself by scipy's whitened kmeans.  The library calls are made in the reference's order, so with the same
    numpy random state the split is the same."""
    from scipy.cluster.vq import kmeans, vq, whiten
    from sklearn import cluster
    if not (max(xs) - min(xs) > 10 and max(ys) - min(ys) > 10):
        return [(xs, ys)]
    pts = np.array([[xs[i], ys[i]] for i in range(len(xs))])
    ks = list(range(1, min(5, len(xs) + 1)))
    fits = [cluster.KMeans(n_clusters=i, init="k-means++").fit(pts) for i in ks]
    preds = [cluster.KMeans(n_clusters=i, init="k-means++").fit_predict(pts) for i in ks]
    bic, bic_k = [], []
    for k in ks:
        if preds[k - 1].max() < k - 1:
            continue
        b = _bic(fits[k - 1], pts)
        if abs(b) < 10 ** 8:
            bic.append(b); bic_k.append(k)
    picked = bic_k[bic.index(max(bic))]
    if picked == 1:
        return [(xs, ys)]
    white = whiten(pts)
    centroids, _ = kmeans(white, picked)
    idx, _ = vq(white, centroids)
    return [([int(v) for v in pts[idx == c, 0]], [int(v) for v in pts[idx == c, 1]]) for c in range(picked)]


def _xmeans(xs, ys):
    """X_means_cluster (Simple_function.pyx:2101-2109): split until a part no longer splits."""
    parts = [p for p in _kmeans_split(xs, ys) if not (p[0] == [] and p[1] == [])]
    if len(parts) == 1 and parts[0][0] == xs and parts[0][1] == ys:
        return [(xs, ys)]
    out = []
    for px, py in parts:
        out += _xmeans(px, py)
    return out


def xmeans_box_sizes(xs, ys):
    """sqrt(bounding-box area) of every X-means cluster of the dots (x, y) below the diagonal
    (X_means_cluster_reformat, cluster_range_decide, cluster_size_decide: Simple_function.pyx:2110-2116, 372-385).
    The clustering draws from numpy's global random state exactly like the reference; a window the libraries cannot
    cluster (the reference would raise) falls back to one bounding box."""
    xs, ys = [int(v) for v in xs], [int(v) for v in ys]
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            clusters = _xmeans(xs, ys)
    except Exception:                                   # noqa: BLE001
        clusters = [(xs, ys)]
    return [float(np.sqrt((max(cx) - min(cx)) * (max(cy) - min(cy)))) for cx, cy in clusters]


_session: Optional[Session] = None
_tls = threading.local()


def get_session() -> Session:
    """The session the drop-in (one call at a time) functions use: the calling thread's, else the process-wide one."""
    global _session
    s = getattr(_tls, "session", None)
    if s is not None:
        return s
    if _session is None:
        _session = Session(int(os.environ.get("VAPOR_DEVICE", "0")))
    return _session


def set_session(s: Optional[Session], thread_only: bool = False):
    global _session
    if thread_only:
        _tls.session = s
    else:
        _session = s


# ====================================================================================================
# drop-in scoring functions (GPU-backed)
# ====================================================================================================
def dotdata(kmerlen, seq1, seq2):
    """Recurrence plot of read ``seq1`` against structure ``seq2``: list of ``(x, y)`` tuples in the
    reference's order (Simple_function.pyx:545-549, 951-983), computed by kernels 1-2."""
    xy = get_session().engine.dotdata(int(kmerlen), seq1, seq2)
    return [(int(a), int(b)) for a, b in xy]


def kmerhits(seq1, seq2, kmerlen, nth_base=1, inversions=False):
    """Simple_function.pyx:951-983.  Only the configuration ``dotdata`` uses is on the GPU path."""
    if nth_base != 1 or not inversions:
        raise NotImplementedError("only kmerhits(seq1, seq2, k, 1, True) -- what dotdata calls -- is implemented")
    return dotdata(kmerlen, seq1, seq2)


def qual_check_repetitive_region(dotdata_qual_check):
    """Simple_function.pyx:1154-1171 on an explicit hit list (host arithmetic on a list the GPU produced)."""
    H = len(dotdata_qual_check)
    low = [(x, y) for x, y in dotdata_qual_check if x > y]
    row = [H, sum(1 for x, y in dotdata_qual_check if x == y), len(low),
           min((x for x, _ in low), default=0), max((x for x, _ in low), default=0),
           min((y for _, y in low), default=0), max((y for _, y in low), default=0)]
    return _qual_check_from_counts(row, 0, ([x for x, _ in low], [y for _, y in low]) if XMEANS_ON_HOST else None)


def window_size_refine(seq2, region_QC_Cff=region_QC_Cff_default):
    """Simple_function.pyx:2030-2046: k in {10, 20, 30, 40} from the self-plot of the window."""
    return get_session().refine_many([RefineRequest(seq2, region_QC_Cff)])[0]


def _single(mode, ref_seq, alt_seq, x, window_size):
    ans = get_session().score_many([ScoreRequest(ref_seq, alt_seq, [x], window_size, mode)])[0]
    return ans.pairs()[0]


def calcu_vapor_single_read_score_abs_dis_m1b(ref_seq, alt_seq, x, window_size):
    """Simple_function.pyx:182-203 -> ``[mean |x-y| ref, mean |x-y| alt]``, ``[1.1, 2.1]`` / ``[2.1, 1.1]`` or ``[0, 0]``."""
    return _single(MODE_ABS, ref_seq, alt_seq, x, window_size)


def calcu_vapor_single_read_score_within_10Perc_m1b(ref_seq, alt_seq, x, window_size):
    """Simple_function.pyx:277-294 -> ``[count10(alt), count10(ref)]`` or ``[0, 0]``."""
    return _single(MODE_W10, ref_seq, alt_seq, x, window_size)


def calcu_vapor_single_read_score_directed_dis_m1b_redefine_diagnal(ref_seq, alt_seq, x, window_size):
    """Simple_function.pyx:241-257 -> ``[|dir| ref, |dir| alt]`` or ``[0, 0]``."""
    return _single(MODE_REDEF, ref_seq, alt_seq, x, window_size)


# ====================================================================================================
# per-SV summary, genotype, output
# ====================================================================================================
def result_organize_ins(info_list):
    """``[key, scores] -> [key, QS, GS, Rec]`` or ``[key, 'NA', 'NA', 'NA']`` (Simple_function.pyx:1219-1231)."""
    s = get_session().summarize([info_list[1]])[0]
    if s is None:
        return [info_list[0]] + ["NA"] * 3
    return [info_list[0], s["QS"], s["GS"], s["Rec"]]


def log_likelihood_calcu(k, l, m, g, err=0.05):
    """Simple_function.pyx:2071-2077 (host form, used only for inspection; kernel 4 computes the same sums)."""
    return -k * math.log(m) + l * math.log((m - g) * err + g * (1 - err)) + (k - l) * math.log((m - g) * (1 - err) + g * err)


def gt_estimate_log_likelihood(vapor_result):
    """``[GT, GQ]`` from a row ending in ``..., GS, Rec`` (Simple_function.pyx:2054-2069).  The reference
    re-parses the rounded Rec string, so does this: the rounded scores go through kernel 4."""
    rec = [float(i) for i in vapor_result[-1].split(",")]
    s = get_session().summarize([rec])[0]
    gt = s["GT"]
    # kernel 4 applied the 0/0 -> 0/1 override with the GS of the scores it was given (the rounded ones); the
    # reference uses the GS of the row, taken from the unrounded scores (:2068), which can only be larger
    if gt == "0/0" and vapor_result[-2] > .15:
        gt = "0/1"
    return [gt, s["GQ"]]


def write_output_initiate(out_name):
    """Simple_function.pyx:2079-2082."""
    with open(out_name, "w") as fo:
        print("\t".join(["#CHR", "POS", "END", "SVTYPE", "SVID", "VaPoR_QS", "VaPoR_GS", "VaPoR_GT", "VaPoR_GQ", "VaPoR_Rec"]), file=fo)


def format_output_row(out_list, gt_gq=None):
    """One ``.vapor`` line (Simple_function.pyx:2084-2088)."""
    if "NA" not in out_list:
        gt_gq = gt_gq if gt_gq is not None else gt_estimate_log_likelihood(out_list)
        return "\t".join(str(i) for i in out_list[:-1] + list(gt_gq) + [out_list[-1]])
    return "\t".join(str(i) for i in out_list[:-1] + ["NA", "NA", "NA"])


def write_output_main(out_name, out_list):
    with open(out_name, "a") as fo:
        print(format_output_row(out_list), file=fo)


def vcf_rec_hash_modify(vcf_rec_hash):
    """key string -> list of record numbers (Simple_function.pyx:1935-1940)."""
    out: Dict[str, list] = {}
    for k1, v in vcf_rec_hash.items():
        out.setdefault(v, []).append(k1)
    return out


def vcf_vapor_modify(vcf_input, vcf_rec_hash_new):
    """Rewrite ``<vcf>.vapor`` as the annotated VCF (the second definition, Simple_function.pyx:1972-2028, is the
    live one): records gain ``;VaPor_GS=..;VaPor_GT=..;VaPor_GQ=..;VaPor_REC=..`` (that spelling) and four
    ``##INFO`` lines follow the input's last ``##INFO`` line.

    Divergence, on purpose: the reference numbers records *with* header lines in ``vcf_list_readin``
    (vapor_vali/vapor:131-134) but *without* them here (:1985), so with any header line it annotates the wrong
    record or dies with KeyError; record numbers here are file line numbers on both sides."""
    vapor_input = vcf_input + ".vapor"
    records: Dict[int, list] = {}
    meta, header = [], []
    with open(vcf_input) as fin:
        for rec, line in enumerate(fin):
            pin = line.strip().split()
            if not pin:
                continue
            if not pin[0][0] == "#":
                records[rec] = pin
            elif not pin[0] == "#CHROM":
                meta.append(pin)
            else:
                header = pin
    keep = []
    with open(vapor_input) as fin:
        for line in fin:
            pin = line.strip().split()
            if pin and pin[0] in vcf_rec_hash_new:
                for y in vcf_rec_hash_new[pin[0]]:
                    gs = round(float(pin[2]), 2) if not pin[2] == "NA" else pin[2]
                    gq = round(float(pin[4]), 2) if not pin[4] == "NA" else pin[4]
                    records[y][7] += ";VaPor_GS=" + str(gs) + ";VaPor_GT=" + str(pin[3]) + ";VaPor_GQ=" + str(gq) + ";VaPor_REC=" + str(pin[5])
                    keep.append(y)
    with open(vapor_input, "w") as fo:
        prev = ""
        for line in meta:
            joined = " ".join(line)
            cur = joined.split("=")[0]
            if prev == "##INFO" and not cur == "##INFO":
                print('##INFO=<ID=VaPoR_GS,Number=1,Type=Float,Description="VaPoR Score, representing the percentage of transverse long reads that support the prediction">', file=fo)
                print('##INFO=<ID=VaPoR_GT,Number=1,Type=String,Description="Genotype with the highest likelihood as estimated by VaPoR">', file=fo)
                print('##INFO=<ID=VaPoR_GQ,Number=1,Type=Float,Description="Genotype quality score - likelihood of the second most likely genotype on a -log10 normalized scale"', file=fo)
                print('##INFO=<ID=VaPoR_REC,Number=.,Type=Float,Description="Similarity scores assigned to each of the reads traversings the predicted SV">', file=fo)
            print(joined, file=fo)
            prev = cur
        print("\t".join(header), file=fo)
        for k1 in sorted(records):
            if k1 in keep:
                print("\t".join(str(i) for i in records[k1]), file=fo)


def make_event_figure_1(plt_li, vapor_score_list, best_read_rec, window_size, ref_seq, alt_seq, out_figure_name):
    """The reference draws a 2x2 PNG (ref/ref, alt/alt, best read/ref, best read/alt) per event
    (Simple_function.pyx:1072-1089).  The four recurrence plots come from the GPU; with ``VAPOR_FIGURES=png`` they are
    rendered (matplotlib when importable, else a built-in zlib PNG rasteriser), with ``tsv`` written as
    ``<name>.dots.tsv``, and skipped otherwise (the default: figures are not on the scoring path)."""
    how = os.environ.get("VAPOR_FIGURES", "")
    if how not in ("png", "tsv") or window_size == "Error":
        return
    if best_read_rec == "" or not vapor_score_list:
        return
    eng = get_session().engine
    read, miss = best_read_rec[0], int(best_read_rec[1])
    panels = [("ref_vs_ref", eng.dotdata(window_size, ref_seq, ref_seq)), ("alt_vs_alt", eng.dotdata(window_size, alt_seq, alt_seq)),
              ("read_vs_ref", eng.dotdata(window_size, read, ref_seq[miss:])), ("read_vs_alt", eng.dotdata(window_size, read, alt_seq[miss:]))]
    if how == "tsv":
        with open(out_figure_name + ".dots.tsv", "w") as fo:
            for name, d in panels:
                for x, y in d:
                    print(f"{name}\t{int(x)}\t{int(y)}", file=fo)
        return
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        _write_dotplot_png(out_figure_name, panels)          # no matplotlib: the built-in rasteriser
        return
    fig = plt.figure(plt_li)
    if not hasattr(fig, "add_subplot"):                          # a stand-in matplotlib (tests stub it out)
        _write_dotplot_png(out_figure_name, panels)
        return
    for i, (name, d) in enumerate(panels):
        ax = fig.add_subplot(2, 2, i + 1)
        if len(d):
            ax.plot(d[:, 0], d[:, 1], "k.", markersize=1)
        ax.set_title(name, fontsize=8)
    fig.savefig(out_figure_name)
    plt.close(fig)


def _write_dotplot_png(path, panels, side=360, pad=12):
    """2 x 2 recurrence plots as an 8-bit grey PNG, written with zlib only (x = structure position to the right,
    y = read position upwards, as in the reference's figure)."""
    import struct
    import zlib
    W = H = 2 * side + 3 * pad
    img = np.full((H, W), 255, dtype=np.uint8)
    for i, (_name, d) in enumerate(panels):
        r0, c0 = pad + (i // 2) * (side + pad), pad + (i % 2) * (side + pad)
        img[r0 - 1:r0 + side + 1, c0 - 1] = 0; img[r0 - 1:r0 + side + 1, c0 + side] = 0
        img[r0 - 1, c0 - 1:c0 + side + 1] = 0; img[r0 + side, c0 - 1:c0 + side + 1] = 0
        if len(d):
            d = np.asarray(d, dtype=np.int64)
            span = max(int(d.max()) + 1, 1)
            px = c0 + (d[:, 0] * side) // span
            py = r0 + side - 1 - (d[:, 1] * side) // span
            img[py, px] = 0
    raw = b"".join(b"\x00" + img[r].tobytes() for r in range(H))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, 0, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


# ====================================================================================================
# L2 drivers as coroutines (they yield RefineRequest / ScoreRequest and return vapor_score_list)
# ====================================================================================================
def _refine(seq):
    ans = yield RefineRequest(seq)
    return ans[0]


def _score(ref_seq, alt_seq, reads, window_size, mode, fig=None):
    """Score ``reads`` and return the driver's vapor_score_list; ``fig`` = (plt_li, out_figure_name) for the figure hook."""
    if not reads:
        return []
    ans = yield ScoreRequest(ref_seq, alt_seq, reads, window_size, mode)
    scores, best = ans.scores(reads)
    if fig is not None:
        make_event_figure_1(fig[0], scores, best, window_size, ref_seq, alt_seq, fig[1])
    return scores


def co_simple_del(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    """vapor_simple_del_Vapor (Simple_function.pyx:1701-1745)."""
    flank = flank_length_calculate(sv_info)
    out: list = []
    fig = (plt_li, out_figure_name)
    if sv_info[2] - sv_info[1] < default_max_sv_test:
        reads = simple_del_chop_pacbio_read_simple_short(bam_in, sv_info, flank)
        if len(reads) > num_reads_cff:
            ref_seq = ref_seq_readin(ref, sv_info[0], sv_info[1] - flank, sv_info[2] + flank)
            k = yield from _refine(ref_seq)
            if not k == "Error":
                alt_seq = ref_seq[:flank] + ref_seq[-flank:]
                out = yield from _score(ref_seq, alt_seq, reads, k, MODE_ABS_AND_W10, fig)   # both opinions, min rule (:1715-1726)
    else:
        reads = simple_del_chop_pacbio_read_simple_short(bam_in, sv_info, flank)
        if len(reads) > num_reads_cff:
            ref_seq = ref_seq_readin(ref, sv_info[0], sv_info[1] - flank, sv_info[1] + flank)
            k = yield from _refine(ref_seq)
            if not k == "Error":
                alt_seq = ref_seq_readin(ref, sv_info[0], sv_info[1] - flank, sv_info[1]) + ref_seq_readin(ref, sv_info[0], sv_info[2], sv_info[2] + flank)
                k = yield from _refine(alt_seq)
                if not k == "Error":
                    out = yield from _score(ref_seq, alt_seq, reads, k, MODE_W10, fig)
    return out


def co_simple_tandup(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    """vapor_simple_tandup_Vapor (Simple_function.pyx:1747-1784)."""
    flank = flank_length_calculate(sv_info)
    fig = (plt_li, out_figure_name)
    if sv_info[2] - sv_info[1] < default_max_sv_test:
        ref_seq = ref_seq_readin(ref, sv_info[0], sv_info[1] - flank, sv_info[2] + flank)
        k = yield from _refine(ref_seq)
        if not k == "Error":
            mid = ref_seq[flank:(-flank)]
            alt_seq = ref_seq[:flank] + mid + mid + ref_seq[-flank:]
            k = yield from _refine(alt_seq)
            if not k == "Error":
                reads = simple_chop_pacbio_read_simple_short(bam_in, sv_info[:2] + [sv_info[1] + 2 * (sv_info[2] - sv_info[1])], flank)
                if len(reads) > num_reads_cff:
                    return (yield from _score(ref_seq, alt_seq, reads, k, MODE_REDEF, fig))
    out: list = []
    ref_seq = ref_seq_readin(ref, sv_info[0], sv_info[2] - flank, sv_info[2] + flank)
    k = yield from _refine(ref_seq)
    if not k == "Error":
        alt_seq = ref_seq_readin(ref, sv_info[0], sv_info[2] - flank, sv_info[2]) + ref_seq_readin(ref, sv_info[0], sv_info[1], sv_info[1] + flank)
        k = yield from _refine(alt_seq)
        if not k == "Error":
            reads = simple_del_chop_pacbio_read_simple_short(bam_in, [sv_info[0], sv_info[2]], flank)
            if len(reads) > num_reads_cff:
                out = yield from _score(ref_seq, alt_seq, reads, k, MODE_W10, fig)
    return out


def co_simple_inv(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    """vapor_simple_inv_Vapor (Simple_function.pyx:1895-1933)."""
    flank = flank_length_calculate(sv_info)
    fig = (plt_li, out_figure_name)
    if sv_info[2] - sv_info[1] < default_max_sv_test:
        ref_seq = ref_seq_readin(ref, sv_info[0], sv_info[1] - flank, sv_info[2] + flank)
        k = yield from _refine(ref_seq)
        if not k == "Error":
            alt_seq = ref_seq[:flank] + reverse(complementary(ref_seq[flank:(-flank)])) + ref_seq[-flank:]
            k = yield from _refine(alt_seq)
            if not k == "Error":
                reads = simple_chop_pacbio_read_simple_short(bam_in, sv_info, flank)
                if len(reads) > num_reads_cff:
                    return (yield from _score(ref_seq, alt_seq, reads, k, MODE_ABS, fig))
    out: list = []
    ref_seq = ref_seq_readin(ref, sv_info[0], sv_info[1] - flank, sv_info[1] + flank)
    k = yield from _refine(ref_seq)
    if not k == "Error":
        alt_seq = ref_seq[:flank] + ref_seq_readin(ref, sv_info[0], sv_info[2] - flank, sv_info[2], "TRUE")
        k = yield from _refine(alt_seq)
        if not k == "Error":
            reads = simple_del_chop_pacbio_read_simple_short(bam_in, sv_info, flank)
            if len(reads) > num_reads_cff:
                out = yield from _score(ref_seq, alt_seq, reads, k, MODE_W10, fig)
    return out


def co_simple_ins(num_reads_cff, plt_li, bam_in, ref, ins_pos, ins_seq, out_figure_name, POLARITY):
    """vapor_simple_ins_Vapor (Simple_function.pyx:1856-1893)."""
    ins_seq_2 = ins_seq if POLARITY == "+" else reverse(complementary(ins_seq))
    flank = default_flank_length if len(ins_seq) > default_flank_length else len(ins_seq)
    chrom, pos = "_".join(ins_pos.split("_")[:-1]), int(ins_pos.split("_")[-1])
    out: list = []
    reads = simple_chop_pacbio_read_simple_short(bam_in, [chrom, ins_pos.split("_")[-1]] + [pos + len(ins_seq)], flank)
    if len(reads) > num_reads_cff:
        if len(ins_seq) < 5000:
            ref_seq = ref_seq_readin(ref, chrom, pos - flank, pos + flank + len(ins_seq))
            k = yield from _refine(ref_seq + ins_seq)
        else:
            ref_seq = ref_seq_readin(ref, chrom, pos - flank, pos + flank)
            k = yield from _refine(ref_seq)
        if not k == "Error":
            alt_seq = ref_seq_readin(ref, chrom, pos - flank, pos) + ins_seq_2 + ref_seq_readin(ref, chrom, pos, pos + flank)
            usable = [x for x in reads if float(x[0].count("N") + x[0].count("n")) / float(len(x[0])) < 0.1]   # :1878
            fig_alt = ref_seq[2:flank] if ins_seq_2.count("X") == len(ins_seq_2) else alt_seq            # :1892 (figure only)
            ans_scores = []
            if usable:
                ans = yield ScoreRequest(ref_seq, alt_seq, usable, k, MODE_ABS)
                ans_scores, best = ans.scores(usable)
                make_event_figure_1(plt_li, ans_scores, best, k, ref_seq, fig_alt, out_figure_name)
            out = ans_scores
    return out


def co_simple_disdup(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    """vapor_simple_disdup_Vapor (Simple_function.pyx:1786-1854).

    The reference compares the insert point, still a str when it comes from a VCF, with int breakpoints and
    raises TypeError under Python 3 (:1803); the evident intent (cf. :1789) is an int comparison, used here.
    An insert point inside the duplicated block leaves ``alt_structure`` unbound there (UnboundLocalError);
    here that event yields no scores (an 'NA' row)."""
    sv_info = list(sv_info)
    sv_info[1:3] = [int(i) for i in sv_info[1:3]]
    sv_info[4] = int(sv_info[4])
    dup_block = sv_info[:3]
    ins_point = [sv_info[3], sv_info[4]]
    flank = flank_length_calculate(dup_block)
    fig = (plt_li, out_figure_name)
    out: list = []
    bp_info = sorted([int(i) for i in sv_info[1:3] + [sv_info[4]]])
    run_flag = 0
    if sv_info[0] == sv_info[3] and max(bp_info) - min(bp_info) < default_max_sv_test:
        ref_seq = ref_seq_readin(ref, sv_info[0], min(bp_info) - flank, max(bp_info) + flank)
        k = yield from _refine(ref_seq)
        if not k == "Error":
            reads = simple_chop_pacbio_read_simple_short(bam_in, [sv_info[0]] + bp_info + [int(bp_info[-1]) + sv_info[2] - sv_info[1]], flank)
            if len(reads) > num_reads_cff:
                run_flag += 1
                if sv_info[4] > sv_info[2]:
                    alt_structure = ["a", "b", "a"]
                elif sv_info[4] < sv_info[1]:
                    alt_structure = ["b", "a", "b"]
                else:
                    return out
                a_seq = ref_seq_readin(ref, sv_info[0], bp_info[0], bp_info[1])
                b_seq = ref_seq_readin(ref, sv_info[0], bp_info[1], bp_info[2])
                alt_seq = ref_seq_readin(ref, sv_info[0], min(bp_info) - flank, min(bp_info))
                for x in alt_structure:
                    alt_seq += a_seq if x == "a" else b_seq
                alt_seq += ref_seq_readin(ref, sv_info[0], max(bp_info), max(bp_info) + flank)
                k = yield from _refine(alt_seq)
                if not k == "Error":
                    out = yield from _score(ref_seq, alt_seq, reads, k, MODE_REDEF, fig)
    if run_flag == 0:
        reads = simple_del_chop_pacbio_read_simple_short(bam_in, ins_point, flank)
        if len(reads) > num_reads_cff:
            ref_seq = ref_seq_readin(ref, ins_point[0], ins_point[1] - flank, ins_point[1] + flank)
            k = yield from _refine(ref_seq)
            if not k == "Error":
                if max(bp_info) - min(bp_info) < default_max_sv_test:
                    alt_seq = ref_seq[:flank] + ref_seq_readin(ref, dup_block[0], dup_block[1], dup_block[2]) + ref_seq[-flank:]
                    mode = MODE_ABS
                else:
                    alt_seq = ref_seq[:flank] + ref_seq_readin(ref, dup_block[0], dup_block[1], dup_block[1] + flank)
                    mode = MODE_W10
                k = yield from _refine(alt_seq)
                if not k == "Error":
                    out = yield from _score(ref_seq, alt_seq, reads, k, mode, fig)
    return out


def co_long_del_inv(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    """vapor_long_del_inv (Simple_function.pyx:1671-1690)."""
    out: list = []
    flank = 500
    ref_seq = ref_seq_readin(ref, sv_info[0][0], sv_info[0][1] - flank, sv_info[1][1] + flank)
    k = yield from _refine(ref_seq)
    if not k == "Error":
        alt_seq = ref_seq[:flank] + reverse(complementary(ref_seq_readin(ref, sv_info[1][0], sv_info[1][2] - flank, sv_info[1][2])))
        k = yield from _refine(alt_seq)
        if not k == "Error":
            reads = simple_del_chop_pacbio_read_simple_short(bam_in, sv_info[0], flank)
            if len(reads) > num_reads_cff:
                out = yield from _score(ref_seq, alt_seq, reads, k, MODE_W10, (plt_li, out_figure_name))
    return out


def co_del_inv(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    """vapor_del_inv_Vapor (Simple_function.pyx:1557-1593).

    Three calls in the reference pass the wrong number of arguments (:1585, :1591-1592) and raise TypeError;
    the evident intent -- the same drivers with ``num_reads_cff`` and ``plt_li`` -- is what runs here."""
    sv_block = [sv_info[0][0], sv_info[0][1], sv_info[-1][2]]
    flank = flank_length_calculate(sv_block)
    out: list = []
    stem, ext = ".".join(out_figure_name.split(".")[:-1]), out_figure_name.split(".")[-1]
    if sv_info[1][1] - sv_info[0][2] < 100:
        simple_pair = len(sv_info) == 2 and [i[-1] for i in sv_info] == ["del", "inv"]
        if sv_block[2] - sv_block[1] < default_max_sv_test:
            ref_seq = ref_seq_readin(ref, sv_block[0], sv_block[1] - flank, sv_block[2] + flank)
            k = yield from _refine(ref_seq)
            if not k == "Error":
                alt_seq = ref_seq[:flank]
                for x in sv_info:
                    if x[-1] == "inv":
                        alt_seq += reverse(complementary(ref_seq_readin(ref, x[0], x[1], x[2])))
                alt_seq += ref_seq[-flank:]
                k = yield from _refine(alt_seq)
                if not k == "Error":
                    reads = simple_chop_pacbio_read_simple_short(bam_in, sv_block[:2] + [sv_block[1] + len(alt_seq) - 2 * flank], flank)
                    if len(reads) > num_reads_cff:
                        out = yield from _score(ref_seq, alt_seq, reads, k, MODE_ABS, (plt_li, out_figure_name))
                    elif simple_pair:
                        out = yield from co_long_del_inv(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name)
        elif simple_pair:
            out = yield from co_long_del_inv(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name)
    else:
        for sub in sv_info:
            name = stem + "_".join(str(i) for i in sub) + "." + ext
            if "del" in sub:
                out += yield from co_simple_del(num_reads_cff, plt_li, bam_in, ref, sub[:-1], name)
            elif "inv" in sub:
                out += yield from co_simple_inv(num_reads_cff, plt_li, bam_in, ref, sub[:-1], name)
    return out


def co_dup_inv(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    """vapor_dup_inv_VapoR (Simple_function.pyx:1595-1669).  Its extra ``math.isnan`` guard (:1632) never fires
    on this path: every statistic the kernels return is a finite ratio of integers."""
    sv_info = list(sv_info)
    sv_info[1:3] = [int(i) for i in sv_info[1:3]]
    sv_info[4] = int(sv_info[4])
    dup_block = sv_info[:3]
    ins_point = [sv_info[3], int(sv_info[4])]
    flank = flank_length_calculate(dup_block)
    fig = (plt_li, out_figure_name)
    out: list = []
    if sv_info[0] == sv_info[3]:
        bp_info = sorted(sv_info[1:3] + [sv_info[4]])
        run_flag = 0
        if max(bp_info) - min(bp_info) < default_max_sv_test:
            ref_seq = ref_seq_readin(ref, sv_info[0], min(bp_info) - flank, max(bp_info) + flank)
            k = yield from _refine(ref_seq)
            if not k == "Error":
                run_flag += 1
                if sv_info[4] > sv_info[2]:
                    alt_structure = ["a", "b", "a^"]
                elif sv_info[4] < sv_info[1]:
                    alt_structure = ["b^", "a", "b"]
                else:
                    alt_structure = ["a", "a^"]
                reads = simple_chop_pacbio_read_simple_short(bam_in, [sv_info[0]] + bp_info + [bp_info[-1] + sv_info[2] - sv_info[1]], flank)
                if len(reads) > num_reads_cff:
                    alt_seq = ref_seq_readin(ref, sv_info[0], min(bp_info) - flank, min(bp_info))
                    a_seq = ref_seq_readin(ref, sv_info[0], bp_info[0], bp_info[1])
                    b_seq = ref_seq_readin(ref, sv_info[0], bp_info[1], bp_info[2])
                    for x in alt_structure:
                        alt_seq += {"a": a_seq, "b": b_seq}[x[0]] if "^" not in x else reverse(complementary({"a": a_seq, "b": b_seq}[x[0]]))
                    alt_seq += ref_seq_readin(ref, sv_info[0], max(bp_info), max(bp_info) + flank)
                    k = yield from _refine(alt_seq)
                    if not k == "Error":
                        out = yield from _score(ref_seq, alt_seq, reads, k, MODE_REDEF, fig)
        if run_flag == 0:
            ref_seq = ref_seq_readin(ref, ins_point[0], ins_point[1] - flank, ins_point[1] + flank)
            k = yield from _refine(ref_seq)
            if not k == "Error":
                reads = simple_del_chop_pacbio_read_simple_short(bam_in, ins_point, flank)
                if len(reads) > num_reads_cff:
                    if max(bp_info) - min(bp_info) < default_max_sv_test:
                        alt_seq = ref_seq[:flank] + reverse(complementary(ref_seq_readin(ref, dup_block[0], dup_block[1], dup_block[2]))) + ref_seq[-flank:]
                        mode = MODE_ABS
                    else:
                        alt_seq = ref_seq[:flank] + reverse(complementary(ref_seq_readin(ref, dup_block[0], dup_block[2] - flank, dup_block[2])))
                        mode = MODE_W10
                    k = yield from _refine(alt_seq)
                    if not k == "Error":
                        out = yield from _score(ref_seq, alt_seq, reads, k, mode, fig)
    return out


def co_cannot_classify(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    """vapor_CANNOT_CLASSIFY_VapoR (Simple_function.pyx:1490-1555): every distinct alternative haplotype is
    scored against the reference structure and all reads' scores are pooled."""
    ref_sv = sv_info[0].split("_")
    alt_sv = list_unify([i for i in sv_info[1].split("_") if i not in ref_sv])
    chromos = chromos_readin(ref)
    bp_info = block_subsplot(sv_info[2:], chromos)
    flank = max(flank_length_calculate(i) for i in bp_info)
    out: list = []
    run_flag = 0
    stem, ext = out_figure_name.split(".")[:-1], out_figure_name.split(".")[-1]
    if len(bp_info) == 1 and bp_info[0][-1] - bp_info[0][1] < default_max_sv_test:
        ref_seq = ref_seq_readin(ref, bp_info[0][0], bp_info[0][1] - flank, bp_info[0][-1] + flank)
        k = yield from _refine(ref_seq)
        if not k == "Error":
            reads = simple_chop_pacbio_read_simple_short(bam_in, bp_info[0], flank)
            let = bp_to_chr_hash(bp_info[0], chromos, flank)
            if len(reads) > num_reads_cff:
                run_flag += 1
                let_seq = {i: ref_seq_readin(ref, let[i][0], int(let[i][1]), int(let[i][-1])) for i in let}
                for alt_allele in alt_sv:
                    alt_seq = ref_seq[:flank]
                    for i in letter_split(alt_allele):
                        alt_seq += let_seq[i] if "^" not in i else reverse(complementary(let_seq[i[0]]))
                    alt_seq += ref_seq[-flank:]
                    k = yield from _refine(alt_seq)
                    if not k == "Error":
                        repeated = max([alt_allele.count(i) for i in alt_allele] + [0]) > 1
                        fig = (plt_li, ".".join(stem + [ref_sv[0] + ".vs." + alt_allele, ext]))
                        out += yield from _score(ref_seq, alt_seq, reads, k, MODE_REDEF if repeated else MODE_ABS, fig)
    if run_flag == 0:
        for alt_allele in alt_sv:
            let = bp_to_chr_hash(bp_info[0], chromos, flank)
            for jun in block_around_check(alt_allele, ref_sv[0]):
                a, b = let[jun[0][0]], let[jun[1][0]]
                if "^" not in jun[0]:
                    ref_seq_a = ref_seq_readin(ref, a[0], a[2] - flank, a[2] + flank)
                else:
                    ref_seq_a = reverse(complementary(ref_seq_readin(ref, a[0], a[1] - flank, a[1] + flank)))
                if "^" not in jun[1]:
                    ref_seq_b = ref_seq_readin(ref, b[0], b[1] - flank, b[1] + flank)
                else:
                    ref_seq_b = reverse(complementary(ref_seq_readin(ref, b[0], b[2] - flank, b[2] + flank)))
                k = yield from _refine(ref_seq_a + ref_seq_b)
                if not k == "Error":
                    alt_seq = ref_seq_a[-flank:] + ref_seq_b[:flank]
                    k = yield from _refine(alt_seq)
                    if not k == "Error":
                        anchor = [a[0], a[2]] if "^" not in jun[0] else [a[0], a[1]]
                        reads = simple_del_chop_pacbio_read_simple_short(bam_in, anchor, flank)
                        if len(reads) > 0:
                            out += yield from _score(ref_seq_a, alt_seq, reads, k, MODE_W10, None)
    return out


# ---- bulk prefetch of the drivers' primary region / read queries ---------------------------------------------
def primary_queries(name, args):
    """What a driver coroutine will ask first, known from its arguments alone: ``(ref, bam_in, [(chrom, start, end)
    faidx regions], [(chrom, start, end, flank) read windows])``.  Fallback paths (junction windows when too few reads
    span the event) are left to the per-call route."""
    regions, windows = [], []
    if name in ("co_simple_del", "co_simple_inv", "co_simple_tandup"):
        _cff, _plt, bam_in, ref, sv_info, _fig = args
        c, s, e = sv_info[0], int(sv_info[1]), int(sv_info[2])
        f = flank_length_calculate(sv_info)
        if e - s < default_max_sv_test:
            regions.append((c, s - f, e + f))
            if name == "co_simple_del":
                windows.append((c, s - f, s + f, f))
            elif name == "co_simple_inv":
                windows.append((c, s - f, e + f, f))
            else:
                windows.append((c, s - f, s + 2 * (e - s) + f, f))
        elif name == "co_simple_del":
            regions += [(c, s - f, s + f), (c, s - f, s), (c, e, e + f)]
            windows.append((c, s - f, s + f, f))
        return ref, bam_in, regions, windows
    if name == "co_simple_ins":
        _cff, _plt, bam_in, ref, ins_pos, ins_seq, _fig, _pol = args
        f = default_flank_length if len(ins_seq) > default_flank_length else len(ins_seq)
        c, pos = "_".join(ins_pos.split("_")[:-1]), int(ins_pos.split("_")[-1])
        windows.append((c, pos - f, pos + len(ins_seq) + f, f))
        regions += [(c, pos - f, pos + f + len(ins_seq)) if len(ins_seq) < 5000 else (c, pos - f, pos + f), (c, pos - f, pos), (c, pos, pos + f)]
        return ref, bam_in, regions, windows
    return None


def prefetch_events(specs, threads=0):
    """Answer the primary queries of many events in two native calls per (reference, read files) pair
    (seqio.prefetch: several host threads) before their coroutines run; the coroutines then find them in the cache."""
    if not seqio.native_enabled():
        return 0
    groups: Dict[tuple, list] = {}
    for name, args in specs:
        q = primary_queries(name, args)
        if q is None:
            continue
        ref, bam_in, regions, windows = q
        g = groups.setdefault((ref, bam_in), [[], []])
        g[0] += regions
        g[1] += windows
    n = 0
    for (ref, bam_in), (regions, windows) in groups.items():
        files = bam_in_decide(bam_in, None)
        seqio.prefetch(ref, regions, files, windows, threads=threads)
        n += len(regions) + len(windows)
    return n


# ---- drop-in wrappers: the reference signatures, one event at a time ----------------------------------
def vapor_simple_del_Vapor(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    return get_session().run_one(co_simple_del(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name))


def vapor_simple_tandup_Vapor(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    return get_session().run_one(co_simple_tandup(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name))


def vapor_simple_inv_Vapor(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    return get_session().run_one(co_simple_inv(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name))


def vapor_simple_ins_Vapor(num_reads_cff, plt_li, bam_in, ref, ins_pos, ins_seq, out_figure_name, POLARITY):
    return get_session().run_one(co_simple_ins(num_reads_cff, plt_li, bam_in, ref, ins_pos, ins_seq, out_figure_name, POLARITY))


def vapor_simple_disdup_Vapor(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    return get_session().run_one(co_simple_disdup(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name))


def vapor_del_inv_Vapor(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    return get_session().run_one(co_del_inv(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name))


def vapor_dup_inv_VapoR(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    return get_session().run_one(co_dup_inv(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name))


def vapor_long_del_inv(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    return get_session().run_one(co_long_del_inv(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name))


def vapor_CANNOT_CLASSIFY_VapoR(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name):
    return get_session().run_one(co_cannot_classify(num_reads_cff, plt_li, bam_in, ref, sv_info, out_figure_name))
