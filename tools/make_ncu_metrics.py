#!/usr/bin/env python
"""Developer tool: profiles/r02_ncu_metrics.json from the raw pages of the committed ncu captures.
    python tools/make_ncu_metrics.py c2=gpurun_out/ev2/raw_c2.csv c3=gpurun_out/ev2/raw_c3.csv ...
Per workload and kernel: measured DRAM bytes and executed thread-instructions of ONE launch, keyed by the hash of
the kernel sources (bench.kernel_source_hash) so that bench.py only reports them for the build they were taken on."""
import csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3}
out = {"source_hash": bench.kernel_source_hash(), "workloads": {}}
for arg in sys.argv[1:]:
    wl, path = arg.split("=")
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    def val(r, name):
        i = idx[name]
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
    per = {}
    for r in rows[2:]:
        name = re.sub(r"^void\s+", "", r[idx["Kernel Name"]])
        name = re.sub(r"^rtd::", "", name).split("<")[0].split("(")[0]
        per[name] = {"dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                     "thread_inst_executed": val(r, "smsp__inst_executed.sum") * val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                     "warp_inst_executed": val(r, "smsp__inst_executed.sum"),
                     "duration_us_under_ncu": val(r, "gpu__time_duration.sum"),
                     "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "lanes_per_inst": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                     "l1_hit_pct": val(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct")}
    out["workloads"][wl] = per
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_ncu_metrics.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
