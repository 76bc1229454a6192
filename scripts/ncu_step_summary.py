#!/usr/bin/env python
"""Turns the `ncu --set full` capture of one bench step (scripts/gpu_round.sh: prof_top.ncu-rep) into the two committed
evidence files: profiles/<tag>_ncu_full_summary.json (key raw metrics per launch) and profiles/ncu_traffic.json (DRAM
bytes per GEMM launch, read by bench.py for roofline.traffic).  Runs on the CPU box (ncu -i).

    python scripts/ncu_step_summary.py gpurun_out/v12/prof_top.ncu-rep r01_v12
"""
import csv, json, os, subprocess, sys

rep, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size']
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
launches = []
for d in rows[2:]:
    if len(d) != len(hdr):
        continue
    name = d[hdr.index("Kernel Name")]
    short = name.split("(")[0].replace("soswsod::", "")
    e = {"kernel": short}
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            e[f"{k} [{units[i]}]"] = d[i]
    def val(k):
        i = hdr.index(k)
        return float(d[i].replace(",", "")) * scale.get(units[i], 1.0)
    e["_us"] = val('gpu__time_duration.sum')
    e["_dram_bytes"] = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
    launches.append(e)
# the capture window (-s 42 -c 14 over the kernels below) is exactly one step after three warm-up steps: 2 ROI forwards,
# 9 GEMMs, 2 ROI backwards, the fused SGD launch
summary = {"source": f"ncu --set full --clock-control none --import-source on -k regex:gemm_bf16|roi_pool_fwd_fast|roi_pool_bwd_q|sgd_multi "
                     f"-s 42 -c 14, bench.py --steps 2 --warmup 3 --blocks 1 ({os.path.basename(rep)})", "launches": launches}
json.dump(summary, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.json"), "w"), indent=1)
gem = [l for l in launches if "gemm_bf16_kernel" in l["kernel"]]
# one step has 9 GEMM launches (fc6, fc7, head forward; head / fc7 / fc6 wgrad + dgrad).  The capture window starts
# inside a step's backward, so 9 consecutive launches = that step's 6 backward GEMMs + the next step's 3 forward GEMMs:
# one launch of every shape
gem = gem[:9]
traffic = {"source": f"profiles/{tag}_ncu_full_summary.json (ncu --set full, one step)",
           "gemm_bf16_kernel": {"launches_per_step": len(gem), "dram_bytes_per_step": sum(g["_dram_bytes"] for g in gem),
                                "dram_bytes_per_launch_avg": sum(g["_dram_bytes"] for g in gem) / max(len(gem), 1),
                                "fc6_fwd_dram_bytes": max(gem, key=lambda g: g["_us"])["_dram_bytes"] if gem else None,
                                "per_launch": [{"kernel": g["kernel"], "dram_bytes": g["_dram_bytes"], "us": g["_us"]} for g in gem]}}
for kind in ("roi_pool_fwd", "roi_pool_bwd", "sgd_multi"):
    ls = [l for l in launches if kind in l["kernel"]][-2:]
    def pct(l, key):
        v = next((val for kk, val in l.items() if kk.startswith(key)), None)
        return None if v is None else float(v.replace(",", ""))
    traffic[kind] = [{"kernel": l["kernel"], "dram_bytes": l["_dram_bytes"], "us": l["_us"],
                      "smem_wavefront_pct_of_peak": pct(l, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak"),
                      "issue_active_pct": pct(l, "smsp__issue_active"), "l2_sectors_pct_of_peak": pct(l, "lts__t_sectors.avg.pct"),
                      "dram_pct_of_peak": pct(l, "gpu__dram_throughput")} for l in ls]
json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
for l in launches:
    print(f"{l['_us']:9.1f} us  dram {l['_dram_bytes']/1e6:9.1f} MB  {l['kernel'][:80]}")
