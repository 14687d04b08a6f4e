"""Summary of an `ncu --set full` report for profiles/ (one entry per captured launch: duration, pipes, issue slots,
instruction count, DRAM bytes, registers, resident warps and the top warp-stall reasons per issued instruction):
    python tools/ncu_full_summary.py report.ncu-rep out.json "how the capture was taken" samples_per_launch [label ...]"""
import csv, io, json, subprocess, sys
rep, out, how, spl = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
labels = sys.argv[5:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "launch__registers_per_thread", "smsp__warps_active.avg.per_cycle_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
kern = []
for i, r in enumerate(rows[2:]):
  d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
  e = {"launch": labels[i] if i < len(labels) else f"launch {i}", "kernel": d["Kernel Name"][:80]}
  for k in KEYS:
    if k in d:
      e[k] = f"{d[k]} {u.get(k, '')}".strip()
  st = [(k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), float(v)) for k, v in d.items()
        if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and "not_issued" not in k and v not in ("", "n/a")]
  e["warp_stalls_per_issue_top"] = {k: round(v, 2) for k, v in sorted(st, key=lambda kv: -kv[1])[:6]}
  kern.append(e)
json.dump({"how": how, "samples_per_launch": spl, "kernels": kern}, open(out, "w"), indent=1)
print(f"{len(kern)} launches -> {out}")
