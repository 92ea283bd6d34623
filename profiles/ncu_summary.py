"""Condense an .ncu-rep (one kernel) into the text summary committed under profiles/:
python profiles/ncu_summary.py gpurun_out/x.ncu-rep > profiles/r2_ncu_x.txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
KEEP = ("gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__icc_request_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
print(f"# {rep}")
for h, v in zip(hdr, vals):
    if h in ("Kernel Name", "Block Size", "Grid Size") or h in KEEP:
        print(f"{h} = {v}")
print("# stall reasons (warps per issue-active cycle)")
for h, v in zip(hdr, vals):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and float(v) >= 0.05:
        print(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} = {float(v):.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    hdr, data = rows[1], rows[2:]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    tot = sum(int(r[isamp] or 0) for r in data)
    print(f"# source view: {len(data)} SASS instructions, {tot} samples; the 25 most sampled")
    for k in sorted(sorted(range(len(data)), key=lambda k: -int(data[k][isamp] or 0))[:25]):
        print(f"{k:6d} samples {int(data[k][isamp] or 0):6d} ({100 * int(data[k][isamp] or 0) / max(tot, 1):4.1f} %) executed {data[k][iex]:>9s}  {data[k][isrc][:90]}")
    mix = collections.Counter()
    for r in data:
        t = r[isrc].split()
        if t:
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            mix[op] += int(r[iex] or 0)
    print("# executed warp-instructions by opcode (top 16):", ", ".join(f"{o} {c}" for o, c in mix.most_common(16)))
