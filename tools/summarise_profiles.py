"""Turn the .ncu-rep captures in gpurun_out/ into the text summaries committed under profiles/ (run here, no GPU needed).
  python tools/summarise_profiles.py gpurun_out/r2_prof_brick.ncu-rep profiles/r2_ncu_k_fill_brick_256cubed.txt"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def source(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    for i, r in enumerate(rows):
        if "Source" in r and "# Samples" in r:
            return r, rows[i + 1:]
    return None, []


def main(rep, dst):
    h, units, vals = raw(rep)
    lines = []
    for v in vals:
        name = v[h.index("Kernel Name")]
        lines.append(f"== kernel: {name}")
        for i, k in enumerate(h):
            if k in KEYS or ("issue_stalled" in k and "per_issue_active" in k and "not_issued" not in k and float(v[i] or 0) > 0.05) or "dmma" in k.lower():
                lines.append(f"  {k:92s} {v[i]:>16s} {units[i]}")
    hd, src = source(rep)
    if hd:
        ia, isamp, iex = hd.index("Source"), hd.index("# Samples"), hd.index("Instructions Executed")
        tot = sum(int(r[isamp] or 0) for r in src) or 1
        ex = sum(int(r[iex] or 0) for r in src)
        lines.append("")
        lines.append(f"total warp instructions executed: {ex}  (SASS instructions: {len(src)}); stall samples: {tot}")
        mix = {}
        for r in src:
            op = r[ia].strip().split()
            op = [o for o in op if not o.startswith("@")]
            if op:
                mix[op[0].split(".")[0]] = mix.get(op[0].split(".")[0], 0) + int(r[iex] or 0)
        top = sorted(mix.items(), key=lambda kv: -kv[1])[:14]
        lines.append("instruction mix (warp-level, executed): " + ", ".join(f"{k}:{v}" for k, v in top))
        lines.append("top stall sites (share of samples | executed | SASS):")
        for i in sorted(sorted(range(len(src)), key=lambda i: -int(src[i][isamp] or 0))[:14]):
            lines.append(f"  {100 * int(src[i][isamp] or 0) / tot:5.1f}%  {src[i][iex]:>10s}  {src[i][ia].strip()[:100]}")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
