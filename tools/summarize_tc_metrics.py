"""Summarise tools/refresh_r02.sh's tensor-core metrics pass (ncu --csv, one row per launch and metric) per kernel:
launches, total ms, DRAM GB, time-weighted tensor-pipe activity. Writes the JSON bench.py reads `roofline.traffic` from.
usage: python tools/summarize_tc_metrics.py <metrics.csv> <out.json>"""
import collections, csv, json, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols = rows[hdr]
ci = {c: cols.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Value")}
launch = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= ci["Metric Value"]:
        continue
    d = launch.setdefault(r[ci["ID"]], {"name": re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("void ", "").replace("tg::", "")})
    d[r[ci["Metric Name"]]] = float(r[ci["Metric Value"]].replace(",", ""))
FAMILY = ("conv_igemm_kernel", "conv_halo", "wgrad_igemm", "wgrad_wide", "wgrad_halo")
per, tot = collections.OrderedDict(), {"launches": 0, "ns": 0.0, "bytes": 0.0, "tp": 0.0}
for d in launch.values():
    ns = d.get("gpu__time_duration.sum", 0.0)
    by = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    k = per.setdefault(d["name"], {"launches": 0, "ns": 0.0, "bytes": 0.0, "tp": 0.0})
    for t in (k,) + ((tot,) if d["name"].startswith(FAMILY) else ()):
        t["launches"] += 1; t["ns"] += ns; t["bytes"] += by; t["tp"] += tp * ns
fmt = lambda t: {"launches": t["launches"], "ms": round(t["ns"] / 1e6, 3), "dram_gbytes": round(t["bytes"] / 1e9, 2),
                 "tensor_pipe_active_pct_time_weighted": round(t["tp"] / max(t["ns"], 1.0), 1)}
out = {"source": "tools/refresh_r02.sh (ncu --metrics gpu__time_duration,dram__bytes_read/write,sm__pipe_tensor_cycles_active "
                 "--profile-from-start off; exactly ONE B=64 adversarial step of the final round-2 build, bracketed by "
                 "cudaProfilerStart/Stop in tools/profile_step.py)",
       "per_kernel": {k: fmt(v) for k, v in per.items()}}
out.update(fmt(tot))
out["note"] = ("top-level launches/ms/dram_gbytes/tensor_pipe: the implicit-GEMM family only (conv_igemm, conv_halo*, wgrad_igemm, "
               "wgrad_wide, wgrad_halo), the family bench.py's roofline block reports")
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(fmt(tot)))
