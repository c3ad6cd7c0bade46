"""profiles/r01_dram_traffic.json from ncu csv logs of ONE eager forward per workload, collected with
   ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv
usage: python tools/traffic_from_ncu.py quartznet15x5=gpurun_out/traffic_qn.csv citrinet1024=... features=..."""
import csv, json, os, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASS = {"pw_gemm": "pw_gemm_pair_kernel", "dw_tma": "dw_tma_kernel", "dw_mma": "dw_tma_kernel", "dw_fast": "dw_tma_kernel",
         "logmel": "logmel_kernel", "pw_wgrad": "pw_wgrad_kernel", "wgrad_reduce": "wgrad_reduce_kernel",
         "dw_wgrad_mma": "dw_wgrad_mma_kernel", "bn_apply_fused": "bn_apply_fused_kernel",
         "bn_bwd_apply_fused": "bn_bwd_apply_fused_kernel", "bn_bwd_reduce": "bn_bwd_reduce_kernel",
         "ctc_alpha_beta": "ctc_alpha_beta_kernel"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
OUT = os.path.join(ROOT, "profiles", "r01_dram_traffic.json")
out = json.load(open(OUT)) if os.path.exists(OUT) else {}     # workloads not named on the command line are kept
for arg in sys.argv[1:]:
    wl, path = arg.split("=")
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ni, mi, ui, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value"), h.index("ID")
    per = defaultdict(lambda: defaultdict(float))
    names = {}
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        if r[mi].startswith("dram__bytes"):
            v *= UNIT.get(r[ui], 1)
            per[r[ii]]["dram"] += v
        elif r[mi].startswith("gpu__time"):
            per[r[ii]]["ns"] += v * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[ui], 1)
        names[r[ii]] = r[ni]
    agg = defaultdict(lambda: dict(launches=0, dram=0.0, ns=0.0))
    for k, d in per.items():
        cls = next((c for key, c in CLASS.items() if key in names[k]), None)
        if cls is None:
            continue
        a = agg[cls]
        a["launches"] += 1; a["dram"] += d["dram"]; a["ns"] += d["ns"]
    out[wl] = {c: {"launches": a["launches"], "dram_bytes_per_launch": a["dram"] / a["launches"],
                   "dram_bytes_per_forward": a["dram"], "gpu_time_us_per_forward_serialized": a["ns"] / 1e3}
               for c, a in agg.items()}
json.dump(out, open(OUT, "w"), indent=1)
print(json.dumps(out, indent=1))
