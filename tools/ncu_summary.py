"""Summarises .ncu-rep captures (ncu --set full) into one JSON: python tools/ncu_summary.py out.json rep1.ncu-rep [rep2 ...]"""
import csv, io, json, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
out = {"kernels": []}
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = {"capture": rep.split("/")[-1], "Kernel Name": r[hdr.index("Kernel Name")], "Grid Size": r[hdr.index("Grid Size")], "Block Size": r[hdr.index("Block Size")]}
        for k in KEYS:
            if k in hdr:
                rec[k] = r[hdr.index(k)] + " " + units[hdr.index(k)]
        stalls = {h.split("smsp__average_warp_latency_issue_stalled_")[-1].split("_pipe")[0].replace(".ratio", "").replace("smsp__average_warps_issue_stalled_", ""): float(r[i])
                  for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("per_issue_active.ratio") and r[i] not in ("", "n/a")}
        rec["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        out["kernels"].append(rec)
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1)[:6000])
