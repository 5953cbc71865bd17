"""Per-kernel times of one bench step: python scratch/kt.py [sims] -> us per simulation"""
import json, subprocess, sys
sims = sys.argv[1] if len(sys.argv) > 1 else "2048"
out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "2", "--no-cpu-baseline", "--no-e2e", "--sims-per-step", sims],
                     capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
d = json.loads(line)
n = int(sims) * d["steps"]
print("value %.1f sims/s  ms/step %.2f" % (d["value"], d["ms_per_step"]))
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms"]):
    print("  %-14s %7.2f us/sim  share %.3f  frac %.3f" % (k, 1e3 * v["ms"] / n, v["share"], v["frac"]))
