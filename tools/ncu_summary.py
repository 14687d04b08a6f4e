"""Summarise an ncu report for profiles/: python tools/ncu_summary.py report.ncu-rep out.json
(one entry per profiled launch: duration, DRAM bytes, pipe activity, issue slots, registers)."""
import csv, io, json, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
  d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
  e = {"kernel": d["Kernel Name"]}
  for k in KEYS:
    if k in d:
      e[k] = f"{d[k]} {u.get(k, '')}".strip()
  out.append(e)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(f"{len(out)} launches -> {sys.argv[2]}")
