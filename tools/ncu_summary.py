"""Key metrics + stall / opcode picture of one kernel of an .ncu-rep: python tools/ncu_summary.py REPORT [kernel-id]"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=" * 100)
    for k in want:
        for kk in hdr:
            if kk == k or kk.startswith(k):
                print(f"{kk:86s} {d[kk]}")
                break
    for k in ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
              "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
              "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
              "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "smsp__thread_inst_executed_per_inst_executed.ratio"):
        if k in d:
            print(f"{k:86s} {d[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, sass, kernels = None, [], 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels += 1
        if kernels > 1:
            break
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].startswith("0x"):
        sass.append(dict(zip(hdr, r)))
if sass:
    tot = sum(int(s["# Samples"]) for s in sass)
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(s[k]) for s in sass) for k in stalls}
    print("-" * 100)
    print("warp-state samples (first kernel of the report):", tot)
    for k, v in sorted(agg.items(), key=lambda x: -x[1]):
        if v:
            print(f"  {k:28s} {v:8d} {100 * v / tot:5.1f}%")
    op, smp = collections.Counter(), collections.Counter()
    for s in sass:
        o = s["Source"].split()
        name = (o[1] if o[0].startswith("@") else o[0]).split(".")[0]
        op[name] += int(s["Instructions Executed"]); smp[name] += int(s["# Samples"])
    tt = sum(op.values())
    print("warp instructions executed:", tt)
    for k, v in op.most_common(24):
        print(f"  {k:12s} {v:12d} {100 * v / tt:5.1f}%   samples {100 * smp[k] / tot:5.1f}%")
