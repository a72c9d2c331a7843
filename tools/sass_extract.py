#!/usr/bin/env python
"""SASS evidence for profiles/: the TMA prologue and hot loops of the kernel-2 variants, cut out of
`cuobjdump -sass vapor_b200/csrc/libvapor_b200.so` (no GPU needed).

    python tools/sass_extract.py            -> profiles/r02_k2_tile_hot_loop.sass, profiles/r02g_k2_join_probe_loop.sass"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vapor_b200", "csrc", "libvapor_b200.so")


def functions():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;?\s*/\*", line)
        if cur and m:
            out[cur].append((int(m.group(1), 16), m.group(2).strip().rstrip(";").strip()))
    return out


def mnemonic_counts(ins):
    c = {}
    for _, t in ins:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        op = op.split(".")[0]
        c[op] = c.get(op, 0) + 1
    return dict(sorted(c.items(), key=lambda kv: -kv[1]))


def write(path, header, blocks):
    with open(path, "w") as f:
        f.write(header)
        for title, ins in blocks:
            f.write(f"\n// ---- {title} ----\n")
            for addr, t in ins:
                f.write(f"/*{addr:05x}*/  {t}\n")
    print(path)


def main():
    fns = functions()
    tile = next(v for k, v in fns.items() if "k2_tile_matchILi13ELi2E" in k)
    join = next(v for k, v in fns.items() if "k2_join_matchILi8E" in k)
    # tile kernel: TMA issue (UBLKCP) + mbarrier wait (SYNCS), then the first vote block of the hot loop = from the first
    # LDS.128 after the wait to the first VOTE
    i_tma = next(i for i, (_, t) in enumerate(tile) if "UBLKCP" in t)
    i_wait = next(i for i, (_, t) in enumerate(tile) if i > i_tma and "SYNCS" in t and "TRYWAIT" in t)
    i_lds = next(i for i, (_, t) in enumerate(tile) if i > i_wait and t.split()[0].startswith("LDS.128"))
    i_vote = next(i for i, (_, t) in enumerate(tile) if i > i_lds and t.split()[0].startswith("VOTE"))
    loop = tile[i_lds:i_vote + 1]
    cnt = mnemonic_counts(loop)
    hdr = ("// k2_tile_match<13,2> (all-pairs tile kernel, k2_mode 0), sm_100a, from cuobjdump -sass of the in-tree library.\n"
           f"// kernel: {len(tile)} instructions.  One vote block of the hot loop = 32 streamed words x 29 rows per lane:\n"
           f"// {len(loop)} instructions: {cnt}\n"
           "// ISETP.EQ.OR = one row compare per lane (alu pipe); IMAD = one Horner step of a row polynomial (fma pipe);\n"
           "// LDS.128 = four streamed words; UBLKCP = the TMA bulk copy that staged them; SYNCS = its mbarrier.\n")
    write(os.path.join(ROOT, "profiles", "r02_k2_tile_hot_loop.sass"), hdr,
          [("TMA staging of the streamed chunk", tile[max(0, i_tma - 6):i_tma + 2]), ("mbarrier wait", tile[i_wait - 1:i_wait + 3]),
           ("one vote block of the hot loop (LDS.128 ... VOTE)", loop)])
    # join kernel: UBLKCP + wait, then step 1 (membership test of one word: LDS of the bitmap word ... VOTE + STS of the survivor)
    # and step 2 (one round: two LDS.U16 bucket offsets, the entry scan, the two VOTEs that park the matched cells)
    j_tma = next(i for i, (_, t) in enumerate(join) if "UBLKCP" in t)
    j_wait = next(i for i, (_, t) in enumerate(join) if i > j_tma and "SYNCS" in t and "TRYWAIT" in t)
    j_f = next(i for i, (_, t) in enumerate(join) if i > j_wait and re.match(r"(@!?U?P\d+ )?LDS ", t))
    j_fv = next(i for i, (_, t) in enumerate(join) if i > j_f and "VOTE" in t)
    j_fs = next(i for i, (_, t) in enumerate(join) if i > j_fv and "STS" in t)
    filt = join[j_f - 4:j_fs + 1]
    j_u16 = next(i for i, (_, t) in enumerate(join) if i > j_fs and "LDS.U16" in t)
    votes = [i for i, (_, t) in enumerate(join) if i > j_u16 and "VOTE" in t]
    rnd = join[j_u16 - 4:votes[1] + 6]
    hdr = ("// k2_join_match<8> (radix-partitioned join with membership pre-filter, k2_mode 1, the default), sm_100a, from cuobjdump -sass of\n"
           f"// the in-tree library.  kernel: {len(join)} instructions.\n"
           f"// step 1, one read k-mer word per lane against the membership bitmap: {mnemonic_counts(filt)}\n"
           f"// step 2, one round of 32 survivors against their buckets: {mnemonic_counts(rnd)}\n"
           "// UBLKCP = the single TMA bulk copy that stages the table blob of the CTA; LDS = bitmap word; VOTE + POPC + STS.64 = survivor\n"
           "// compaction; LDS.U16 x2 = bucket offsets; LDS + ISETP = entry compare; VOTE x2 = queue slots of the matched cells (no atomics).\n")
    write(os.path.join(ROOT, "profiles", "r02g_k2_join_probe_loop.sass"), hdr,
          [("TMA staging of the table blob", join[max(0, j_tma - 6):j_tma + 2]), ("mbarrier wait", join[j_wait - 1:j_wait + 3]),
           ("step 1: membership test of the first of four words per lane, survivors compacted into the list", filt),
           ("step 2: one round (bucket lookup, entry scan, matched cells parked by ballot)", rnd)])


if __name__ == "__main__":
    main()
