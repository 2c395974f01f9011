"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of metrics profiles/ cites."""
import csv
import subprocess
import sys

rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_kernels.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "dram__bytes.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct"]
stalls = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
for r in data:
    print("=" * 100)
    print(r[idx["Kernel Name"]])
    for w in WANT:
        if w in idx:
            print(f"  {w:92s} {r[idx[w]]:>14s} {units[idx[w]]}")
    top = sorted(((float(r[idx[s]] or 0), s) for s in stalls), reverse=True)[:4]
    for v, s in top:
        print(f"  stall {s.split('issue_stalled_')[1].split('_per_issue')[0]:86s} {v:14.2f} warps/issue")
