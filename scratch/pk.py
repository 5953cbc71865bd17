"""python scratch/pk.py bench.json [kernel ...]: headline and per-simulation kernel times (us) of a bench.py line."""
import json, sys
d = json.loads([x for x in open(sys.argv[1]) if x.startswith("{")][-1])
spp = d["config"]["sims_per_step_per_gpu"]
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e", d["e2e"] and round(d["e2e"]["value"], 1))
names = sys.argv[2:] or sorted(d["kernels"], key=lambda k: -d["kernels"][k]["ms"])
for k in names:
    v = d["kernels"][k]
    print(f"{k:16s} {v['ms'] / spp * 1e3:6.2f} us  share {v['share']:.3f} frac {v['frac']:.3f}")
