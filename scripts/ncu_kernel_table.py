#!/usr/bin/env python
"""One row per launch of an .ncu-rep (ncu --set full): duration, DRAM / L2 bytes and GB/s, issue %, tensor %.
    python scripts/ncu_kernel_table.py <report.ncu-rep> [out.json]"""
import csv, json, subprocess, sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6, "usecond": 1.0}


def val(d, k):
    if k not in hdr:
        return None
    i = hdr.index(k)
    try:
        return float(d[i].replace(",", "")) * scale.get(units[i], 1.0)
    except ValueError:
        return None


table = []
for d in rows[2:]:
    if len(d) != len(hdr):
        continue
    us = val(d, "gpu__time_duration.sum")
    dram = (val(d, "dram__bytes_read.sum") or 0) + (val(d, "dram__bytes_write.sum") or 0)
    l2 = val(d, "lts__t_bytes.sum")
    if l2 is None and val(d, "lts__t_sectors.sum") is not None:
        l2 = 32.0 * val(d, "lts__t_sectors.sum")
    e = {"kernel": d[hdr.index("Kernel Name")].split("(")[0].replace("soswsod::", "").replace("void ", ""),
         "grid": d[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else None, "us": us,
         "dram_MB": dram / 1e6, "dram_GBps": dram / us / 1e3 if us else None,
         "l2_MB": l2 / 1e6 if l2 else None, "l2_GBps": l2 / us / 1e3 if (l2 and us) else None,
         "issue_active_pct": val(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "tensor_pipe_pct": val(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
         "dram_pct_of_peak": val(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
         "l2_pct_of_peak": val(d, "LTS.TriageCompute.lts__throughput.avg.pct_of_peak_sustained_elapsed"),
         "l2_hit_rate_pct": val(d, "lts__t_sector_hit_rate.pct")}
    table.append(e)
    print(f"{us:9.1f} us  dram {e['dram_MB']:8.1f} MB {e['dram_GBps'] or 0:7.0f} GB/s  L2 {e['l2_MB'] or 0:8.1f} MB {e['l2_GBps'] or 0:7.0f} GB/s  "
          f"L2thr {e['l2_pct_of_peak'] or 0:5.1f}%  issue {e['issue_active_pct'] or 0:5.1f}%  tensor {e['tensor_pipe_pct'] or 0:5.1f}%  grid {e['grid']}  {e['kernel'][:60]}")
if len(sys.argv) > 2:
    json.dump({"source": rep, "launches": table}, open(sys.argv[2], "w"), indent=1)
