"""Generate the end-to-end CLI golden tables by running the UNMODIFIED reference command line
(/root/reference/vapor_vali/vapor) on a small synthetic data set.  Build container only; commit the outputs.

    python tests/golden/make_cli_golden.py

What it does
  1. vapor_b200.synth_genome.make_dataset -> tests/golden/cli_case/{ref.fa, ref.fa.fai, reads.sam.gz, svs.bed,
     svs.vcf, svs_nohdr.vcf, truth.json}                                  (seeded, deterministic)
  2. a scratch directory OUTSIDE the repo gets: a package ``vapor_vali`` whose modules are symlinks to the
     reference's own .pyx files (nothing is copied), the no-op matplotlib stand-in from oracle/mpl_stub, and a
     ``sitecustomize`` restoring the numpy aliases SciPy dropped (``scipy.std``), and a
     ``samtools`` shim answering ``faidx`` / ``view`` in samtools' text formats (this image has no samtools)
  3. the reference CLI runs ``bed`` and ``vcf`` there; its tables land in tests/golden/cli_case/*.golden
  4. the reference's DISDUP driver cannot be reached from its CLI under Python 3 (str-vs-int TypeError,
     Simple_function.pyx:1803), so it is called directly with the insert point as an int and the score list is
     stored in disdup.golden.json
"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
CASE = os.path.join(HERE, "cli_case")
REF_PKG = "/root/reference/vapor_vali"

SHIM = r'''#!/usr/bin/env python
import sys
sys.path.insert(0, %(root)r)
from vapor_b200 import seqio
cmd = sys.argv[1]
if cmd == "faidx":
    path, region = sys.argv[2], sys.argv[3]
    chrom, rng = region.rsplit(":", 1)
    a, b = rng.split("-") if not rng.startswith("-") else ("-" + rng[1:].split("-")[0], rng[1:].split("-")[1])
    seq = seqio.fasta(path).fetch(chrom, int(a), int(b))
    print(">" + region)
    for i in range(0, len(seq), 60):
        print(seq[i:i + 60])
elif cmd == "view":
    path, region = sys.argv[2], sys.argv[3]
    chrom, rng = region.rsplit(":", 1)
    a, b = rng.split("-")
    for r in seqio.alignments(path).fetch(chrom, int(a), int(b)):
        print("\t".join([r.qname, "0", chrom, str(r.pos), "60", r.cigar, "*", "0", "0", r.seq, "*"]))
else:
    sys.exit("samtools shim: unsupported command " + cmd)
'''


def main():
    from vapor_b200 import synth_genome
    if os.path.isdir(CASE):
        shutil.rmtree(CASE)
    ds = synth_genome.make_dataset(CASE, seed=20261018, n_simple=8, n_complex=5, size_range=(60, 900), coverage=16.0,
                                   read_len_mean=6000.0, complex_types=("DEL_INV", "DUP_INV", "OTHER", "DISDUP", "OTHER2"))
    with open(ds.sam, "rb") as f, gzip.GzipFile(ds.sam + ".gz", "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)
    os.remove(ds.sam)
    sam = ds.sam + ".gz"
    # the reference CLI cannot process DISDUP records (TypeError) and mis-numbers records when the VCF has
    # header lines (vcf_vapor_modify): give it a header-less VCF without the DISDUP record
    nohdr = os.path.join(CASE, "svs_nohdr.vcf")
    with open(ds.vcf) as f, open(nohdr, "w") as g:
        for line in f:
            if not line.startswith("#") and "<DISDUP>" not in line:
                g.write(line)

    tmp = tempfile.mkdtemp(prefix="vapor_ref_cli_")
    try:
        pkg = os.path.join(tmp, "pkgs", "vapor_vali")
        os.makedirs(pkg)
        open(os.path.join(pkg, "__init__.py"), "w").close()
        os.symlink(os.path.join(REF_PKG, "Simple_function.pyx"), os.path.join(pkg, "Simple_function.py"))
        os.symlink(os.path.join(REF_PKG, "prep.pyx"), os.path.join(pkg, "prep.py"))
        # environment shim, not a change to the reference: SciPy removed the numpy aliases the reference still calls
        # (scipy.std, Simple_function.pyx:878); without them every TANDUP event dies in window_size_refine.  KMeans
        # there is unseeded -- seed numpy so the run is repeatable (its result only matters when < 40 % of a window's
        # self-plot lies on the diagonal, which none of these events reaches).
        with open(os.path.join(tmp, "pkgs", "sitecustomize.py"), "w") as f:
            f.write("import numpy, scipy\nfor _n in ('std', 'mean', 'array', 'sqrt'):\n    if not hasattr(scipy, _n): setattr(scipy, _n, getattr(numpy, _n))\nnumpy.random.seed(0)\n")
        bindir = os.path.join(tmp, "bin")
        os.makedirs(bindir)
        shim = os.path.join(bindir, "samtools")
        with open(shim, "w") as f:
            f.write(SHIM % {"root": ROOT})
        os.chmod(shim, 0o755)
        env = dict(os.environ)
        env["PATH"] = bindir + ":" + env["PATH"]
        env["PYTHONPATH"] = os.path.join(tmp, "pkgs") + ":" + os.path.join(ROOT, "oracle", "mpl_stub")
        figs = os.path.join(tmp, "figs")
        cli = os.path.join(REF_PKG, "vapor")
        out_bed = os.path.join(tmp, "bed.vapor")
        subprocess.run([sys.executable, cli, "bed", "--sv-input", ds.bed, "--output-path", figs, "--output-file", out_bed,
                        "--reference", ds.ref_fa, "--pacbio-input", sam], check=True, env=env, cwd=tmp)
        shutil.copy(out_bed, os.path.join(CASE, "svs.bed.vapor.golden"))
        work_vcf = os.path.join(tmp, "svs_nohdr.vcf")
        shutil.copy(nohdr, work_vcf)
        subprocess.run([sys.executable, cli, "vcf", "--sv-input", work_vcf, "--output-path", figs, "--output-file", "unused",
                        "--reference", ds.ref_fa, "--pacbio-input", sam], check=True, env=env, cwd=tmp)
        shutil.copy(work_vcf + ".vapor", os.path.join(CASE, "svs_nohdr.vcf.vapor.golden"))
        # DISDUP at driver level, insert point as int (the evident intent)
        code = (
            "import json, sys\n"
            "from vapor_vali.Simple_function import *\n"
            "out = {}\n"
            "for sv in json.load(open(sys.argv[1])):\n"
            "    if sv['type'] != 'DISDUP': continue\n"
            "    info = ['chr1', sv['start'], sv['end'], 'chr1', int(sv['insert_point'])]\n"
            "    out[sv['svid']] = {'sv_info': info, 'scores': vapor_simple_disdup_Vapor(3, 1, sys.argv[3], sys.argv[2], list(info), sys.argv[4])}\n"
            "json.dump(out, open(sys.argv[5], 'w'))\n")
        dd = os.path.join(tmp, "disdup.json")
        os.makedirs(figs, exist_ok=True)
        subprocess.run([sys.executable, "-c", code, os.path.join(CASE, "truth.json"), ds.ref_fa, sam, os.path.join(figs, "dd.png"), dd],
                       check=True, env=env, cwd=tmp)
        shutil.copy(dd, os.path.join(CASE, "disdup.golden.json"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    for f in sorted(os.listdir(CASE)):
        print(f, os.path.getsize(os.path.join(CASE, f)))
    make_large_case()
    make_complex_case()


def make_complex_case():
    """Third case, vcf only: 56 complex events -- DEL_INV, DUP_INV, DEL_DUP_INV (two alternative haplotypes, README.md:81),
    swapped blocks, two- and three-allele `Other=` records, and events of 10-11 kb that send vapor_CANNOT_CLASSIFY_VapoR
    down its junction-window fallback (Simple_function.pyx:1537-1555).  DISDUP records are left out: the reference CLI
    cannot process them under Python 3 (TypeError, :1803; disdup.golden.json of the first case pins that driver)."""
    from vapor_b200 import synth_genome
    case = os.path.join(HERE, "cli_case_complex")
    if os.path.isdir(case):
        shutil.rmtree(case)
    kinds = ("DEL_INV", "DUP_INV", "DEL_DUP_INV", "OTHER", "OTHER2", "OTHER3", "DUP_INV", "DEL_DUP_INV", "OTHER3", "DEL_INV",
             "OTHER2", "OTHER", "OTHER3", "OTHER_LONG")
    ds = synth_genome.make_dataset(case, seed=20261018 + 3, n_simple=0, n_complex=56, size_range=(150, 700), coverage=20.0,
                                   read_len_mean=5000.0, complex_types=kinds, kind_size={"OTHER_LONG": (7200, 7600)})
    with open(ds.sam, "rb") as f, gzip.GzipFile(ds.sam + ".gz", "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)
    os.remove(ds.sam)
    os.remove(ds.bed)
    sam = ds.sam + ".gz"
    nohdr = os.path.join(case, "svs_nohdr.vcf")
    with open(ds.vcf) as f, open(nohdr, "w") as g:
        for line in f:
            if not line.startswith("#"):
                g.write(line)
    os.remove(ds.vcf)
    tmp = tempfile.mkdtemp(prefix="vapor_ref_cli_")
    try:
        env = _reference_env(tmp)
        work_vcf = os.path.join(tmp, "svs_nohdr.vcf")
        shutil.copy(nohdr, work_vcf)
        subprocess.run([sys.executable, os.path.join(REF_PKG, "vapor"), "vcf", "--sv-input", work_vcf, "--output-path", os.path.join(tmp, "figs"),
                        "--output-file", "unused", "--reference", ds.ref_fa, "--pacbio-input", sam], check=True, env=env, cwd=tmp)
        shutil.copy(work_vcf + ".vapor", os.path.join(case, "svs_nohdr.vcf.vapor.golden"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    for f in sorted(os.listdir(case)):
        print("complex/" + f, os.path.getsize(os.path.join(case, f)))


def make_large_case():
    """Second case, bed only: events of 10-12 kb (the drivers switch to 1 kb junction windows at >= 10 kb,
    Simple_function.pyx:1728-1744, 1769-1783, 1918-1932) and a 5.5 kb insertion (short reference window, :1870-1871)."""
    from vapor_b200 import synth_genome
    case = os.path.join(HERE, "cli_case_large")
    if os.path.isdir(case):
        shutil.rmtree(case)
    ds = synth_genome.make_dataset(case, seed=20261018 + 4, n_simple=4, n_complex=0, size_range=(10050, 11500), coverage=10.0,
                                   read_len_mean=5000.0, ins_len_override=5500)
    with open(ds.sam, "rb") as f, gzip.GzipFile(ds.sam + ".gz", "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)
    os.remove(ds.sam)
    os.remove(ds.vcf)
    sam = ds.sam + ".gz"
    tmp = tempfile.mkdtemp(prefix="vapor_ref_cli_")
    try:
        env = _reference_env(tmp)
        out_bed = os.path.join(tmp, "bed.vapor")
        subprocess.run([sys.executable, os.path.join(REF_PKG, "vapor"), "bed", "--sv-input", ds.bed, "--output-path", os.path.join(tmp, "figs"),
                        "--output-file", out_bed, "--reference", ds.ref_fa, "--pacbio-input", sam], check=True, env=env, cwd=tmp)
        shutil.copy(out_bed, os.path.join(case, "svs.bed.vapor.golden"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    for f in sorted(os.listdir(case)):
        print("large/" + f, os.path.getsize(os.path.join(case, f)))


def _reference_env(tmp):
    """Scratch package + shims for running the unmodified reference (see main)."""
    pkg = os.path.join(tmp, "pkgs", "vapor_vali")
    os.makedirs(pkg)
    open(os.path.join(pkg, "__init__.py"), "w").close()
    os.symlink(os.path.join(REF_PKG, "Simple_function.pyx"), os.path.join(pkg, "Simple_function.py"))
    os.symlink(os.path.join(REF_PKG, "prep.pyx"), os.path.join(pkg, "prep.py"))
    with open(os.path.join(tmp, "pkgs", "sitecustomize.py"), "w") as f:
        f.write("import numpy, scipy\nfor _n in ('std', 'mean', 'array', 'sqrt'):\n    if not hasattr(scipy, _n): setattr(scipy, _n, getattr(numpy, _n))\nnumpy.random.seed(0)\n")
    bindir = os.path.join(tmp, "bin")
    os.makedirs(bindir)
    shim = os.path.join(bindir, "samtools")
    with open(shim, "w") as f:
        f.write(SHIM % {"root": ROOT})
    os.chmod(shim, 0o755)
    env = dict(os.environ)
    env["PATH"] = bindir + ":" + env["PATH"]
    env["PYTHONPATH"] = os.path.join(tmp, "pkgs") + ":" + os.path.join(ROOT, "oracle", "mpl_stub")
    return env


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "complex":
        make_complex_case()
    else:
        main()
