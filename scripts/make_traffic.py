"""profiles/traffic.json from the ncu `--set full` raw pages under profiles/: per-launch DRAM traffic (dram__bytes_read.sum +
dram__bytes_write.sum) of the step kernel, keyed by the complete workload so that bench.py's roofline.traffic can never quote another
configuration's figure:  "<config>_<lattice>_<collision>_<policy>_<nx>x<ny>x<nz>" (local extents)."""

import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = {  # key -> raw page
    "cavity_D3Q19_BGK_FP32FP32_512x512x512": "r2_ncu_tile1_d3q19_f32_raw.csv",
    "cavity_D3Q19_BGK_FP32FP16_512x512x512": "r2_ncu_tile_final_raw.csv",
    "cavity_D3Q27_KBC_FP32FP32_512x512x512": "r2_ncu_full_step_kernel_d3q27_kbc_lean_512_raw.csv",
    "cavity_D3Q27_KBC_FP32FP32_256x256x256": "r2_ncu_full_step_kernel_d3q27_kbc_lean_raw.csv",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def traffic(path):
    rows = list(csv.reader(open(path)))
    head, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        tot = sum(float(d[k]) * UNIT[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        out.append((d.get("Kernel Name", "?"), tot, float(d["gpu__time_duration.sum"]), u["gpu__time_duration.sum"]))
    return out


table = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel, ncu --set full --clock-control none, raw pages under profiles/ (made by scripts/make_traffic.py)"}
for key, name in SOURCES.items():
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        print("missing", name, file=sys.stderr)
        continue
    kernel, tot, t, tu = traffic(path)[-1]
    table[key] = {"bytes": int(round(tot)), "source": f"profiles/{name}", "kernel": kernel.split("(")[0][:80], "launch": f"{t} {tu}"}
json.dump(table, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(table, indent=1))
