#!/usr/bin/env python
"""End-to-end wall time of the command line on a generated data set (FASTA + SAM + BED on disk):

    python tools/cli_bench.py [--n-sv 300] [--gpus 1] [--workdir /tmp/vapor_cli_bench]

Generates the files with vapor_b200.synth_genome (seeded), runs `vapor bed` through vapor_b200.cli in-process and
prints one JSON line: wall seconds split into file loading, host driver logic + GPU calls, and writing, plus the
session's call statistics.  This is the whole drop-in path including host I/O, not the kernel benchmark (bench.py)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-sv", type=int, default=300)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--workdir", default="/tmp/vapor_cli_bench")
    args = ap.parse_args()
    from vapor_b200 import Simple_function as SF, cli, seqio, synth_genome
    t0 = time.perf_counter()
    ds = synth_genome.make_dataset(args.workdir, seed=20261018 + 7, n_simple=args.n_sv, n_complex=0, size_range=(50, 2000),
                                   coverage=args.coverage, read_len_mean=8000.0)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    if seqio.native_enabled():                                  # parse + index once (cached for the run)
        seqio.native_fasta(ds.ref_fa); seqio.native_aln(ds.sam)
    else:
        seqio.fasta(ds.ref_fa); seqio.alignments(ds.sam)
    t_load = time.perf_counter() - t0
    sessions = [SF.Session(0)] if args.gpus == 1 else args.gpus          # N > 1: one worker process per GPU
    if args.gpus == 1:
        SF.set_session(sessions[0])

    class A:
        pass
    a = A()
    a.sv_input, a.output_path, a.output_file = ds.bed, os.path.join(args.workdir, "figs"), os.path.join(args.workdir, "out.vapor")
    a.reference, a.pacbio_input, a.PB_supp, a.gpus = ds.ref_fa, ds.sam, None, args.gpus
    import contextlib, io
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        rows = cli.run_bed(a, sessions)
    t_run = time.perf_counter() - t0
    stats = {"reads_scored": None}
    if args.gpus == 1:
        stats = dict(sessions[0].stats)
        sessions[0].close()
    called = sum(1 for r in rows if len(r) > 4)
    print(json.dumps({"n_sv": args.n_sv, "gpus": args.gpus, "host_io": "native" if seqio.native_enabled() else "python", "rows": len(rows), "rows_with_scores": called,
                      "generate_s": round(t_gen, 2), "load_index_s": round(t_load, 2), "vapor_bed_s": round(t_run, 2),
                      "sv_per_s": round(len(rows) / t_run, 1), "reads_scored": stats["reads_scored"],
                      "reads_per_s": round(stats["reads_scored"] / t_run, 1) if stats["reads_scored"] else None, "session": stats}))


if __name__ == "__main__":
    main()
