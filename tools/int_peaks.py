#!/usr/bin/env python
"""Integer issue-rate microbenchmarks of the library (vapor_gpu_int_peak) -> one JSON line for profiles/."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vapor_b200.engine import Engine
names = {0: "ISETP compare-accumulate (alu pipe)", 1: "LOP3 (alu pipe)", 2: "IADD3 (alu pipe)", 3: "LOP3 + IMAD independent streams (alu + fma pipes)",
         4: "the tile kernel's inner loop in isolation (circular: not used as a ceiling)", 5: "IMAD (fma pipe)",
         6: "ISETP + IMAD independent streams, streamed operand first (alu + fma pipes)"}
with Engine(0) as e:
    out = {names[w]: round(e.int_peak(w) / 1e9, 1) for w in names}
out["unit"] = "G lane-ops/s"
out["nominal_issue_limit"] = round(148 * 128 * 1.965, 1)
print(json.dumps(out))
