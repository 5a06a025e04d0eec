"""Debug helper: run one golden run of trace_enhanced.json on the GPU and print where it departs from the reference."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import enhanced_common as EC
from hyptokenizer_b200.tokenizer.enhanced_fast_hyperbolic_merge import EnhancedFastHyperbolicTokenizer as T

run = int(sys.argv[1])
gd = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trace_enhanced.json")))
r = gd["runs"][run]
path = os.path.join(tempfile.mkdtemp(), "corpus.txt")
open(path, "w", encoding="utf-8").write("\n".join(gd["lines"]) + "\n")
tok, merges, heads, curv = EC.run_golden(T, gd, r, path)
for k, (m, w) in enumerate(zip(merges, r["merges_ij"])):
    if m != w:
        print("first differing merge", k, m, w)
        break
for k, (g, w) in enumerate(zip(heads, r["heads"])):
    print(k, g, "\n ", w)
    if k > 12:
        break
