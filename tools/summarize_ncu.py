#!/usr/bin/env python3
"""Turn .ncu-rep captures (read with `ncu -i ... --page raw --csv`, no GPU needed) into the small tracked summaries
under profiles/:  python tools/summarize_ncu.py gpurun_out/prof_fwd_wan42_r1.ncu-rep [...]  ->  profiles/<name>.json"""
import csv, io, json, os, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "sm__cycles_elapsed.max": "sm_cycles",
    "sm__cycles_elapsed.max.per_second": "sm_clock",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed": "xu_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed": "fma_pipe_active_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed": "alu_pipe_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__bytes_read.sum.per_second": "dram_read_rate",
    "dram__bytes_write.sum.per_second": "dram_write_rate",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "gpu_dram_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_tensor_operand_wavefronts_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__sass_inst_executed_op_tmem_ldt.sum": "tcgen05_ld_instructions",
    "smsp__sass_inst_executed_op_tmem_stt.sum": "tcgen05_st_instructions",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12,
              "Gbyte/s": 1e9, "Mbyte/s": 1e6, "Tbyte/s": 1e12, "Kbyte/s": 1e3, "byte/s": 1,
              "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1, "Ghz": 1e9, "Mhz": 1e6}

def summarize(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        rec = {"kernel": d.get("Kernel Name", "")[:80], "source_report": os.path.basename(path)}
        for k, name in KEYS.items():
            if k not in d or d[k] in ("", None):
                continue
            try:
                v = float(d[k].replace(",", ""))
            except ValueError:
                continue
            unit = u.get(k, "")
            if unit in UNIT_SCALE and name not in ("tensor_pipe_active_pct",):
                v *= UNIT_SCALE[unit]
                unit = {"ms": "s", "us": "s", "ns": "s"}.get(unit, unit.split("byte")[0] and ("B/s" if "/s" in unit else "B") if "byte" in unit else "Hz" if "hz" in unit else unit)
            rec[name] = v
            rec[name + "_unit"] = unit if unit else ("%" if name.endswith("_pct") else "")
        stalls = {k.split("issue_stalled_")[1].split("_per_issue_active")[0]: float(d[k]) for k in hdr
                  if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and d.get(k)}
        rec["warp_stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        if "dram_read" in rec and "dram_write" in rec:
            rec["dram_traffic_bytes"] = rec["dram_read"] + rec["dram_write"]
            if "duration" in rec:
                rec["dram_GBps"] = rec["dram_traffic_bytes"] / rec["duration"] / 1e9
        out.append(rec)
    return out

if __name__ == "__main__":
    os.makedirs("profiles", exist_ok=True)
    for p in sys.argv[1:]:
        recs = summarize(p)
        dst = os.path.join("profiles", os.path.basename(p).replace(".ncu-rep", ".json"))
        json.dump(recs, open(dst, "w"), indent=1)
        if "fwd_wan42" in dst and recs and "dram_traffic_bytes" in recs[0]:   # what bench.py reports as roofline.traffic
            json.dump({"la_fwd_kernel": {"dram_bytes_per_launch": recs[0]["dram_traffic_bytes"],
                                         "source": os.path.basename(dst),
                                         "command": "python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu (one launch, ncu --set full)"}},
                      open(os.path.join("profiles", "traffic.json"), "w"), indent=1)
        for r in recs:
            print(dst, r["kernel"][:40], f"{r.get('duration', 0)*1e3:.3f} ms", f"tensor {r.get('tensor_pipe_active_pct', 0):.1f}%",
                  f"dram {r.get('dram_GBps', 0):.0f} GB/s", f"traffic {r.get('dram_traffic_bytes', 0)/1e9:.3f} GB")
