import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vapor_b200.engine import Engine
CASES = json.load(open("tests/golden/scoring_cases.json"))["cases"]
only = sys.argv[1:] 
eng = Engine(0)
eng.set_option("k2_mode", 1)
for c in CASES:
    if only and c["name"] not in only: continue
    if c.get("error"): continue
    for key, struct in (("dot_ref", c["ref"][c["miss"]:]), ("dot_alt", c["alt"][c["miss"]:]), ("dot_ref_upper", c["ref"].upper()[c["miss"]:])):
        print(c["name"], key, flush=True)
        d = eng.dotdata(c["k"], c["read"], struct)
        print("   n", len(d), "exp", c[key]["n"], flush=True)
