"""Line-level similarity (difflib ratio over stripped non-comment lines) of every Python file of the package and the oracle
against every Python file of the reference checkout.  Run in the build container: python scripts/copy_selfcheck.py [ref_dir]"""
import difflib, glob, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref_dir = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
norm = lambda s: [l.strip() for l in s.splitlines() if l.strip() and not l.strip().startswith("#")]
ref = {p: norm(open(p, errors="ignore").read()) for p in glob.glob(os.path.join(ref_dir, "**", "*.py"), recursive=True)}
worst = 0.0
for m in sorted(glob.glob(os.path.join(ROOT, "multi_fidelity_gpflow_b200", "*.py")) + glob.glob(os.path.join(ROOT, "oracle", "*.py"))):
    a = norm(open(m).read())
    r, p = max(((difflib.SequenceMatcher(None, a, b, autojunk=False).ratio(), p) for p, b in ref.items() if b), default=(0.0, None))
    worst = max(worst, r)
    print(f"{os.path.relpath(m, ROOT):48s} {r:.2f}  {p}")
print(f"largest ratio {worst:.2f} (the round's copy detector flags > 0.60)")
