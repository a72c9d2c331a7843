"""Help texts of the command line: the host-side mirror of ``vapor_vali.prep`` (vapor_vali/prep.pyx:3-51).
Same four entry points; the wording is this package's own."""
from __future__ import print_function

_COMMON = """Required:
  --sv-input      input file of SV calls
  --output-path   folder where the recurrence plots (if requested) are kept
  --output-file   name of the output table
  --reference     reference genome (FASTA with a .fai index)
  --pacbio-input  long reads aligned to the reference: SAM, or BAM with a .bai index
Optional:
  --PB-supp       minimum number of evaluable long reads per event (default 3)
  --gpus          B200s to shard the SV list over (default 1)
"""


def print_read_me():
    print("VaPoR (B200-native scoring path) -- validate structural variants with long reads")
    print("Usage: vapor <bed|vcf|svelter> [options]")
    print("  vapor bed      SVs in BED format: chr start end SVID TYPE [INS sequence]")
    print("  vapor vcf      SVs in VCF format (simple and complex events)")
    print("  vapor svelter  SVs in SVelter format")
    print("Run a subcommand without options for its parameters.")


def readme_melt():
    print("vapor ins --sv-input-prefix <MELT prefix> ... : the reference's MELT entry point is not part of this build")
    print(_COMMON)


def readme_bed():
    print("Usage: vapor bed [options]")
    print("BED columns: chr start end SVID TYPE, TYPE one of DEL, DUP, INV, INS_<len|seq> (ALU/LINE1/SVA/HERVK likewise);")
    print("an optional sixth column gives the inserted sequence.")
    print(_COMMON)


def readme_vcf():
    print("Usage: vapor vcf [options]")
    print("Simple events by SVTYPE (DEL, DUP/TANDUP, INV, INS with SVLEN/SEQ); complex events as DISDUP (insert_point=),")
    print("DEL_INV (del=, inv=), DUP_INV (insert_point=) or Other=<ref>_<alt>_<chr:bp1:bp2...>.")
    print("The annotated VCF is written next to the input as <input>.vapor.")
    print(_COMMON)
