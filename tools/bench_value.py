"""One-line view of a bench.py run (for tools/ab_run.sh):  python tools/bench_value.py --model humanoid --no-linearize ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-e2e"] + sys.argv[1:], capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    print("value %.4g ms_per_step %.4f kernel_ms %.4f contacts %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["contact_stats"]))
except Exception:
    print("FAILED", out.stderr[-400:])
