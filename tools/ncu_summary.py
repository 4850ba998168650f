"""Summarise an .ncu-rep: per kernel, key throughput / stall metrics (raw page) and top stall source lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:60])
    for k in keys:
        if k in d:
            print("  %-75s %s" % (k, d[k]))
    st = sorted(((float(d[s]), s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for s in stall if d.get(s)), reverse=True)
    print("  stalls/issue:", ", ".join("%s %.2f" % (n, v) for v, n in st[:8]))
