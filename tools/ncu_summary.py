"""Print the headline counters of the first kernel in an ncu report.
usage: python tools/ncu_summary.py <file.ncu-rep> [world_steps]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ws = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
for k in keys:
    if k in d: print(f"{k:75s} {d[k][0]} {d[k][1]}")
print("-- warp stall reasons (pct of warp-active cycles per issue slot)")
st = sorted(((float(v[0].replace(',', '')), h) for h, v in d.items() if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")), reverse=True)
for v, h in st[:10]: print(f"   {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):28s} {v:.3f}")
if ws:
    n = float(d["smsp__inst_executed.sum"][0].replace(',', ''))
    t = float(d["gpu__time_duration.sum"][0].replace(',', ''))
    u = d["gpu__time_duration.sum"][1]
    print(f"warp-instr per world-step {n/ws:,.0f}; time {t} {u}")
    try:
        fl = sum(float(d[k][0].replace(',', '')) * m for k, m in (("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", 2), ("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", 1), ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", 1)))
        print(f"executed FP32 flop per world-step (ffma*2+fmul+fadd) {fl/ws:,.0f}")
    except KeyError: pass
