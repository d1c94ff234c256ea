"""profiles/traffic.json from an .ncu-rep (`ncu --set full`): per stage, DRAM bytes (read + write) per aspect-passing window.

    python tools/make_traffic.py gpurun_out/<rep>.ncu-rep <aspect_passing_windows_of_that_run> <label> [output.json]
"""
import csv
import json
import os
import subprocess
import sys

STAGE = [("k2_crop_resize", "k2_crop_resize"), ("k5_hist", "k5_hist"), ("k5_gram", "k5_pairs"), ("k5_pairs", "k5_pairs"), ("k5_fold", "k5_fold"),
         ("k3_masks", "k3_masks"), ("k4_score", "k4_score"), ("k1_", "k1_expand_filter")]
rep, nwin, label = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    st = next((s for pat, s in STAGE if pat in name), None)
    if st is None:
        continue
    b = sum(float(r[col[k]]) * mult[units[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    e = res.setdefault(st, {}).setdefault(name.split("(")[0], [0.0, 0])     # a stage = one launch of each of its kernels per step
    e[0] += b
    e[1] += 1
final = {}
for st, ks in res.items():
    per_step = sum(b / n for b, n in ks.values())
    final[st] = {"dram_bytes_per_window": per_step / nwin, "kernel": " + ".join(ks), "source": "%s (%d aspect-passing windows)" % (label, nwin)}
p = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
json.dump(final, open(p, "w"), indent=1)
print(json.dumps(final, indent=1))
