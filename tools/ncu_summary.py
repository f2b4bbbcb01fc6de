"""Summarise an ncu report of mcalf_fast_kernel into a text file for profiles/ (run where ncu is installed).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_summary_rNN.txt [samples_per_launch]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "launch__block_size",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    nsamp = int(sys.argv[3]) if len(sys.argv) > 3 else None
    lines = ["ncu summary of %s" % rep, ""]
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    for i, name in enumerate(hdr):
        if any(k in name for k in KEYS) and ".min" not in name and ".max" not in name and ".sum.pct" not in name:
            lines.append("%-95s %s %s" % (name, vals[i], units[i]))
    src = page(rep, "source")
    h, data = src[1], src[2:]
    iS, iN, iSm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    tot = sum(int(r[iN]) for r in data)
    ts = sum(int(r[iSm]) for r in data) or 1
    lines += ["", "SASS regions (60 instructions each): share of executed warp-instructions, share of stall samples, dominant opcodes",
              "total warp-instructions %d%s" % (tot, ("  (%.1f per sample)" % (tot / nsamp)) if nsamp else "")]
    for b0 in range(0, len(data), 60):
        blk = data[b0:b0 + 60]
        e = sum(int(r[iN]) for r in blk)
        s = sum(int(r[iSm]) for r in blk)
        if e / tot < 0.004 and s / ts < 0.004:
            continue
        ops = {}
        for r in blk:
            t = r[iS].split()
            op = t[0] if not t[0].startswith("@") else t[1]
            ops[op] = ops.get(op, 0) + 1
        top = ", ".join("%s x%d" % kv for kv in sorted(ops.items(), key=lambda x: -x[1])[:5])
        lines.append("%5d  exec %5.2f%%  samples %5.2f%%   %s" % (b0, e / tot * 100, s / ts * 100, top))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
