"""Summarise an .ncu-rep (raw page + SASS page) into text: key metrics, per-phase instruction mix.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xxx.txt]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:86s} {r[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 3:
    hdr = rows[1]; data = rows[2:]
    ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot = sum(int(r[ie]) for r in data if len(r) > ie and r[ie].isdigit())
    print(f"\ntotal warp instructions executed: {tot}  (SASS instructions: {len(data)})")
    reg = 0
    regs = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for r in data:
        if len(r) <= ie or not r[ie].isdigit():
            continue
        op = [t for t in r[ia].split() if not t.startswith("@")]
        o = op[0] if op else ""
        regs[reg][0] += int(r[ie]); regs[reg][1] += int(r[isamp]); regs[reg][2][o.split(".")[0]] += int(r[ie])
        if o.startswith("BAR"):
            reg += 1
    for k, v in regs.items():
        print(f"region {k} (between barriers): {v[0]} warp-inst = {100 * v[0] / max(tot,1):.1f}%  stall samples {v[1]}")
        print("    ", ", ".join(f"{a}:{b}" for a, b in v[2].most_common(12)))
