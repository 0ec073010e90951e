"""Rewrites the phase table of DESIGN.md from profiles/phases_r01.json (tools/phase_timing.py output)."""
import json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = json.load(open(os.path.join(ROOT, "profiles", "phases_r01.json")))
path = os.path.join(ROOT, "DESIGN.md")
text = open(path).read()
for r in rows:
    name = {"modp": "modp", "secp256k1": "secp256k1", "ristretto255": "ristretto255"}[r["group"]]
    head = f"| {name} n={r['n']} t={r['t']} |"
    cells = " | ".join(f"{v['wall_ms']:.0f} ({v['kernel_ms']:.0f})" for v in r["phases"].values())
    line = f"{head} {cells} |"
    text, k = re.subn(re.escape(head) + r".*", line, text)
    assert k == 1, head
    print(line)
open(path, "w").write(text)
