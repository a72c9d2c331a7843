"""CPU oracle for VaPoR's per-read scoring path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference algorithm in
``/root/reference/vapor_vali/Simple_function.pyx`` (cited below as ``SF:line``).
It exists so that the CUDA path can be checked bit-for-bit.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it; nothing under ``vapor_b200/`` does, and the
product path fails loudly when the CUDA library is missing instead of falling
back to this code.

Parity pinning: the reference ships no usable golden vectors for this path
(SURVEY.md section 4), so this restatement is pinned by differential testing
against the *unmodified* reference module, imported from ``/root/reference`` in
the build container (``tests/test_oracle_vs_reference.py``), and by the golden
fixtures under ``tests/golden/`` which were generated from the reference itself
by ``tests/golden/make_golden.py``.

Coordinates: a dot/hit is ``(x, y) = (structure k-mer start, read k-mer
start)`` exactly as ``kmerhits`` appends ``(i, hit)`` (SF:979).
"""
from __future__ import annotations

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

# --------------------------------------------------------------------------
# alphabet (SF:20 invert_base, SF:908-949 key_modify)
# --------------------------------------------------------------------------
CODE_INVALID = 15
_CODE = np.full(256, CODE_INVALID, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = 8 + _i
_CODE[ord("N")] = 4
_CODE[ord("n")] = 12
for _c in "RYSWKMBDHV":            # key_modify: IUPAC -> N / n (SF:909-948)
    _CODE[ord(_c)] = 4
    _CODE[ord(_c.lower())] = 12
_COMP = np.arange(16, dtype=np.uint8)   # invert_base in code space
_COMP[[0, 1, 2, 3]] = [3, 2, 1, 0]
_COMP[[8, 9, 10, 11]] = [11, 10, 9, 8]


def encode(seq) -> np.ndarray:
    """ASCII -> alphabet code after ``key_modify`` (SF:908-949)."""
    if isinstance(seq, str):
        seq = seq.encode("latin-1")
    return _CODE[np.frombuffer(bytes(seq), dtype=np.uint8)]


def _kmer_ids(arrs, k):
    """Give every distinct k-mer (row) over several window matrices one integer id."""
    allw = np.ascontiguousarray(np.concatenate(arrs, axis=0))
    v = allw.view(np.dtype((np.void, k))).ravel()
    _, inv = np.unique(v, return_inverse=True)
    out, p = [], 0
    for a in arrs:
        out.append(inv[p:p + len(a)])
        p += len(a)
    return out


def dotdata(k: int, seq1, seq2) -> np.ndarray:
    """``dotdata`` -> ``kmerhits(seq1, seq2, k, 1, True)`` (SF:545-549, SF:951-983).

    seq1 = read (hashed forward + reverse complement, SF:957-960, SF:1419-1421),
    seq2 = structure (forward only, SF:964-967).  Returns an ``(H, 2)`` int64
    array of ``(x, y)`` rows in reference order: x ascending, then y ascending,
    a palindromic read k-mer contributing the same row twice.
    Raises ``KeyError`` like ``invert_base[c]`` does (SF:1421) when the read holds
    a character outside ACGTN/acgtn after ``key_modify``.
    """
    r = encode(seq1)
    s = encode(seq2)
    n = len(r) - k + 1
    m = len(s) - k + 1
    if n >= 1 and (r == CODE_INVALID).any():
        bad = bytes(seq1.encode("latin-1") if isinstance(seq1, str) else seq1)[int(np.argmax(r == CODE_INVALID))]
        raise KeyError(chr(bad))
    if n < 1 or m < 1:
        return np.zeros((0, 2), dtype=np.int64)
    win_r = sliding_window_view(r, k)
    rc = _COMP[r][::-1]
    win_rc = sliding_window_view(rc, k)[::-1]          # row j = revcomp of read k-mer j
    win_s = sliding_window_view(s, k)
    id_r, id_rc, id_s = _kmer_ids([win_r, win_rc, win_s], k)
    keys = np.concatenate([id_r, id_rc])
    pos = np.concatenate([np.arange(n), np.arange(n)])
    order = np.lexsort((pos, keys))
    skeys, spos = keys[order], pos[order]
    valid_s = ~(win_s == CODE_INVALID).any(axis=1)      # e.g. 'X' never matches a read
    lo = np.searchsorted(skeys, id_s, "left")
    hi = np.searchsorted(skeys, id_s, "right")
    cnt = np.where(valid_s, hi - lo, 0)
    total = int(cnt.sum())
    x = np.repeat(np.arange(m), cnt)
    start = np.repeat(lo - (np.cumsum(cnt) - cnt), cnt)
    y = spos[start + np.arange(total)]
    return np.stack([x, y], axis=1).astype(np.int64)


# --------------------------------------------------------------------------
# chain clustering (SF:551-580)
# --------------------------------------------------------------------------
def _chain_group_sizes(vals: np.ndarray, dis_cff: int = 10):
    """Sort, chain while successive difference < dis_cff (SF:553-559 / 568-574).
    Returns (group id per element, size of each group), groups numbered in
    ascending value order."""
    order = np.argsort(vals, kind="stable")
    sv = vals[order]
    brk = np.concatenate([[True], np.diff(sv) >= dis_cff])
    gid_sorted = np.cumsum(brk) - 1
    sizes = np.bincount(gid_sorted)
    gid = np.empty(len(vals), dtype=np.int64)
    gid[order] = gid_sorted
    return gid, sizes


def dis_cluster_2_keepmask(vals: np.ndarray, dis_cff: int = 10) -> np.ndarray:
    """``dis_cluster_2`` (SF:566-580): True where the value's chain group has len > 10."""
    gid, sizes = _chain_group_sizes(vals, dis_cff)
    return sizes[gid] > 10


def dis_cluster_kept_order(vals: np.ndarray, dis_cff: int = 10) -> np.ndarray:
    """``dis_cluster`` (SF:551-564): indices kept, in the reference's order (kept
    groups ascending, original index order inside a group).  Groups with len > 50
    are kept; if there is none, every group tied for the maximum size."""
    gid, sizes = _chain_group_sizes(vals, dis_cff)
    keep_g = sizes > 50
    if not keep_g.any():
        keep_g = sizes == sizes.max()
    idx = np.nonzero(keep_g[gid])[0]
    return idx[np.argsort(gid[idx], kind="stable")]


def clean_dotdata_diagnal_and_anti_diagnal(dots: np.ndarray) -> np.ndarray:
    """SF:432-448.  Keep a dot unless it is 'removed' on both y-x and y+x."""
    if len(dots) == 0:
        return dots
    x, y = dots[:, 0], dots[:, 1]
    keep = dis_cluster_2_keepmask(y - x) | dis_cluster_2_keepmask(y + x)
    return dots[keep]


def clean_dotdata_diagnal_m1b(dots: np.ndarray) -> np.ndarray:
    """SF:404-416 (kept dots only; kept_region is a copy of them)."""
    if len(dots) == 0:
        return dots
    return dots[dis_cluster_kept_order(dots[:, 1] - dots[:, 0])]


def clean_dotdata_anti_diagnal_m1b(dots: np.ndarray) -> np.ndarray:
    """SF:418-430."""
    if len(dots) == 0:
        return dots
    return dots[dis_cluster_kept_order(dots[:, 1] + dots[:, 0])]


# --------------------------------------------------------------------------
# per-plot statistics (SF:705-733, 582-591, 1104-1118, 788-792, 1483-1488)
# --------------------------------------------------------------------------
def eu_dis_abs_calcu(dots: np.ndarray):
    """SF:705-708."""
    return np.mean(np.abs(dots[:, 0] - dots[:, 1]))


def eu_dis_dots_within_10perc(dots: np.ndarray) -> int:
    """SF:730-733."""
    x = dots[:, 0].astype(np.float64)
    y = dots[:, 1].astype(np.float64)
    sel = x > 0
    r = np.abs((x[sel] - y[sel]) / x[sel])
    return int((r < 0.16).sum())


def eu_dis_dir_calcu(xs: np.ndarray, ys: np.ndarray):
    """SF:710-722 (x may be fractional after re-centring)."""
    xs = np.asarray(xs, dtype=np.float64)
    ys = np.asarray(ys, dtype=np.float64)
    num = xs - ys
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(xs == 0, np.abs(num / (xs + 1.0)), np.abs(num / xs))
    sel = ratio > 0.1
    if not sel.any():
        return 0.0001
    return np.mean(num[sel])


def _number_cluster_bins(vals: np.ndarray, edges) -> np.ndarray:
    """``number_cluster`` (SF:1104-1118): bin index 0..10 of every value for the 11
    edges; a value goes to bin recb-1 for the first recb>=1 with v < edges[recb],
    else to the last bin."""
    b = np.full(len(vals), 10, dtype=np.int64)
    for recb in range(10, 0, -1):
        b[vals < edges[recb]] = recb - 1
    return b


def dis_to_diagnal_most_abundant_defined(dots: np.ndarray):
    """SF:582-591."""
    d = np.sort(dots[:, 1] - dots[:, 0])
    mn, mx = int(d[0]), int(d[-1])
    edges = [mn + i * float(mx - mn) / 10.0 for i in range(11)]
    b = _number_cluster_bins(d, edges)
    cnt = np.bincount(b, minlength=11)
    modal = np.nonzero(cnt == cnt.max())[0]
    kept2 = []
    for bi in modal:                                   # find_longest_list + unify_list (SF:788-792)
        km = d[b == bi]
        kmn, kmx = int(km[0]), int(km[-1])
        e2 = [kmn + i * float(kmx - kmn) / 10.0 for i in range(11)]
        b2 = _number_cluster_bins(km, e2)
        c2 = np.bincount(b2, minlength=11)
        for bj in np.nonzero(c2 == c2.max())[0]:
            kept2.append(km[b2 == bj])
    if len(kept2) == 1:
        return np.median(kept2[0])
    return 0


# --------------------------------------------------------------------------
# per-read score drivers (SF:182-203, 241-257, 277-294)
# --------------------------------------------------------------------------
def _up(s):
    return s.upper()


def calcu_vapor_single_read_score_abs_dis_m1b(ref_seq, alt_seq, x, window_size):
    """Mode ABS, SF:182-203."""
    ref_seq = _up(ref_seq)
    alt_seq = _up(alt_seq)
    rd = dotdata(window_size, x[0], ref_seq[x[1]:])
    ad = dotdata(window_size, x[0], alt_seq[x[1]:])
    if len(rd) > 2 and len(ad) > 2:
        if float(len(rd)) / min([float(len(ref_seq)), float(len(alt_seq))]) > 0.1:
            ref_span = float(rd[-1][0] - rd[0][0]) / float(len(ref_seq)) > 0.6
            alt_span = float(ad[-1][0] - ad[0][0]) / float(len(alt_seq)) > 0.6
            if ref_span and alt_span:
                rc = clean_dotdata_diagnal_and_anti_diagnal(rd)
                ac = clean_dotdata_diagnal_and_anti_diagnal(ad)
                if len(rc) > 0 and len(ac) > 0:
                    return [eu_dis_abs_calcu(rc), eu_dis_abs_calcu(ac)]
                return [0, 0]
            if ref_span:
                return [1.1, 2.1]
            if alt_span:
                return [2.1, 1.1]
            return [0, 0]
        return [0, 0]
    return [0, 0]


def _w10_clean(dots: np.ndarray) -> np.ndarray:
    """SF:281-288: diagonal clusters, then anti-diagonal clusters of the leftovers."""
    if len(dots) == 0:
        return dots
    d = dots[:, 1] - dots[:, 0]
    kept_idx = dis_cluster_kept_order(d)
    kept = dots[kept_idx]
    left_mask = np.ones(len(dots), dtype=bool)
    left_mask[kept_idx] = False          # membership by value == by d-group (duplicates share d)
    left = dots[left_mask]
    anti = clean_dotdata_anti_diagnal_m1b(left)
    return np.concatenate([kept, anti], axis=0)


def calcu_vapor_single_read_score_within_10Perc_m1b(ref_seq, alt_seq, x, window_size):
    """Mode W10, SF:277-294.  NB: returns [count(alt), count(ref)]."""
    rd = dotdata(window_size, x[0], ref_seq[x[1]:])
    ad = dotdata(window_size, x[0], alt_seq[x[1]:])
    if max([float(len(rd)) / float(len(ref_seq)), float(len(ad)) / float(len(alt_seq))]) > 0.1:
        rc = _w10_clean(rd)
        ac = _w10_clean(ad)
        if len(rc) > 0 and len(ac) > 0:
            return [eu_dis_dots_within_10perc(ac), eu_dis_dots_within_10perc(rc)]
        return [0, 0]
    return [0, 0]


def calcu_vapor_single_read_score_directed_dis_m1b_redefine_diagnal(ref_seq, alt_seq, x, window_size):
    """Mode REDEF, SF:241-257."""
    rd = dotdata(window_size, x[0], ref_seq[x[1]:])
    ad = dotdata(window_size, x[0], alt_seq[x[1]:])
    ok = (float(len(rd)) / float(len(ref_seq)) > 0.1 and float(len(ad)) / float(len(alt_seq)) > 0.1
          and float(rd[-1][0] - rd[0][0]) / float(len(ref_seq)) > 0.7
          and float(ad[-1][0] - ad[0][0]) / float(len(alt_seq)) > 0.7)
    if not ok:
        return [0, 0]
    rc = clean_dotdata_diagnal_and_anti_diagnal(rd)
    ac = clean_dotdata_diagnal_and_anti_diagnal(ad)
    if len(rc) > 0 and len(ac) > 0:
        ri = dis_to_diagnal_most_abundant_defined(rc)
        ai = dis_to_diagnal_most_abundant_defined(ac)
        return [abs(eu_dis_dir_calcu(rc[:, 0] + ri, rc[:, 1])),
                abs(eu_dis_dir_calcu(ac[:, 0] + ai, ac[:, 1]))]
    return [0, 0]


MODE_ABS, MODE_W10, MODE_REDEF, MODE_ABS_AND_W10 = 0, 1, 2, 3
_MODE_FN = {
    MODE_ABS: calcu_vapor_single_read_score_abs_dis_m1b,
    MODE_W10: calcu_vapor_single_read_score_within_10Perc_m1b,
    MODE_REDEF: calcu_vapor_single_read_score_directed_dis_m1b_redefine_diagnal,
}


def _pair_score(pair):
    """``if not 0 in pair: 1-float(pair[1])/float(pair[0])`` (e.g. SF:1913-1914)."""
    if 0 in pair:
        return None
    return 1 - float(pair[1]) / float(pair[0])


def score_read(mode, ref_seq, alt_seq, x, window_size, nan_guard=False):
    """Per-read score as the L2 drivers form it (a16).  ``MODE_ABS_AND_W10`` is the
    simple-DEL rule (SF:1715-1726); ``nan_guard`` is DUP_INV's extra test (SF:1632).
    Returns a float or None when the read contributes nothing."""
    if mode == MODE_ABS_AND_W10:
        s1 = _pair_score(calcu_vapor_single_read_score_abs_dis_m1b(ref_seq, alt_seq, x, window_size))
        s2 = _pair_score(calcu_vapor_single_read_score_within_10Perc_m1b(ref_seq, alt_seq, x, window_size))
        if s1 is not None and s2 is not None:
            return min([s1, s2])
        return s1 if s1 is not None else s2
    pair = _MODE_FN[mode](ref_seq, alt_seq, x, window_size)
    if nan_guard and (np.isnan(pair[0]) or np.isnan(pair[1])):
        return None
    return _pair_score(pair)


# --------------------------------------------------------------------------
# per-SV summary and genotype (SF:1219-1231, 2054-2077)
# --------------------------------------------------------------------------
def result_organize_ins(info_list):
    """SF:1219-1231."""
    if len(info_list[1]) > 0:
        pos_values = [i for i in info_list[1] if float(i) > 0]
        neg_values = [i for i in info_list[1] if not float(i) > 0]
        geno_value = float(len(pos_values)) / float(len(pos_values) + len(neg_values))
        qual_value = np.mean(pos_values) if pos_values else 0
        return [info_list[0]] + [qual_value, geno_value,
                                 ",".join([str(round(float(i), 2)) for i in info_list[1]])]
    return [info_list[0]] + ["NA" for _ in range(3)]


def log_likelihood_calcu(k, l, m, g, err=0.05):
    """SF:2071-2077."""
    out = -k * np.log(m)
    for _ in range(l):
        out += np.log((m - g) * err + g * (1 - err))
    for _ in range(k - l):
        out += np.log((m - g) * (1 - err) + g * err)
    return out


def gt_estimate_log_likelihood(vapor_result):
    """SF:2054-2069."""
    read_score_list = [float(i) for i in vapor_result[-1].split(",")]
    k = len(read_score_list)
    l = len([i for i in read_score_list if not i > 0])
    m = 2
    gt_score = [log_likelihood_calcu(k, l, m, 2), log_likelihood_calcu(k, l, m, 1),
                log_likelihood_calcu(k, l, m, 0)]
    gt_list = ["0/0", "0/1", "1/1"]
    ori = [np.exp(i - max(gt_score)) for i in gt_score]
    norm = [i / sum(ori) for i in ori]
    gt_qual = -np.log(np.median(norm)) / np.log(10)
    gt_out = gt_list[gt_score.index(max(gt_score))]
    if gt_out == "0/0" and vapor_result[-2] > .15:
        gt_out = "0/1"
    return [gt_out, gt_qual]


def summarize_sv(scores):
    """QS, GS, GT (0,1,2 = 0/0,0/1,1/1), GQ, Rec for one SV; None when no read scored."""
    row = result_organize_ins(["k", list(scores)])
    if "NA" in row:
        return None
    gt, gq = gt_estimate_log_likelihood(row)
    return {"QS": float(row[1]), "GS": float(row[2]), "GT": ["0/0", "0/1", "1/1"].index(gt),
            "GQ": float(gq), "Rec": row[3]}
