#!/usr/bin/env python
"""Hot source lines of one kernel from an ncu report (run here, in the build container):
    python tools/ncu_source_hot.py gpurun_out/x.ncu-rep [top]
Aggregates the per-SASS-instruction stall samples of `ncu --page source --csv` by CUDA source line."""
import collections
import csv
import io
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = sys.argv[3] if len(sys.argv) > 3 else ""      # substring of the kernel's function name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.defaultdict(lambda: collections.Counter())
text = {}
total = 0
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        fname = rows[i + 1][1] if len(rows[i + 1]) > 1 else ""
        if only and only not in fname:
            i += 3
            while i < len(rows) and not (rows[i] and rows[i][0] == "File Path"):
                i += 1
            continue
        fp, hdr = rows[i][1].split("/")[-1], rows[i + 2]
        col = {h: k for k, h in enumerate(hdr)}
        j = i + 3
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            r = rows[j]
            j += 1
            if len(r) < len(hdr) or not r[0].strip().isdigit():     # SASS rows (no line number) repeat the line totals
                continue
            # source text with quotes / commas breaks the CSV columns: index the numeric fields from the END of the row
            r = [r[0], ",".join(r[1:len(r) - len(hdr) + 2])] + r[len(r) - len(hdr) + 2:]
            try:
                n = int(r[col["# Samples"]] or 0)
            except (ValueError, IndexError):
                continue
            key = (fp, r[0])
            text[key] = r[1].strip()[:110]
            a = agg[key]
            a["samples"] += n
            total += n
            for h, k in col.items():
                if h.startswith("stall_") and "Not Issued" not in h:
                    try:
                        a[h] += int(r[k] or 0)
                    except ValueError:
                        pass
            for h in ("L1 Wavefronts Shared Excessive", "L2 Theoretical Sectors Local", "Instructions Executed"):
                if h in col:
                    try:
                        a[h] += int(float(r[col[h]] or 0))
                    except ValueError:
                        pass
        i = j
    else:
        i += 1
print(f"total samples {total}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    stalls = ", ".join(f"{h[6:]} {v}" for h, v in a.most_common() if h.startswith("stall_") and v * 20 > a["samples"])
    print(f"{a['samples'] / total * 100:5.1f}%  {key[0]}:{key[1]:>4}  inst {a['Instructions Executed']:>10}  bankx {a['L1 Wavefronts Shared Excessive']:>9}  "
          f"local {a['L2 Theoretical Sectors Local']:>9}  | {text[key]}  [{stalls}]")
