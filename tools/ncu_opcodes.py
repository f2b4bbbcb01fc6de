"""Opcode histogram of mcalf_fast_kernel from an ncu report (source page, SASS view), and the executed FP32
flop count DERIVED FROM THE COUNTERS -- the cross-check of bench.py's per-path flop model (run where ncu is
installed; the report comes from the GPU box).

    python tools/ncu_opcodes.py gpurun_out/prof.ncu-rep profiles/ncu_opcodes_rNN.csv profiles/ncu_opcodes_rNN.json \
           <samples per launch> [kernel ms]

ncu's `smsp__sass_thread_inst_executed_op_{ffma,fmul,fadd}` counters do not see the packed FFMA2 / FMUL2 /
FADD2 opcodes of sm_100 (two IEEE fp32 operations per lane per instruction), so the count is taken from the
per-instruction "Thread Instructions Executed" column of the source page instead:

    flop = 4 FFMA2 + 2 (FMUL2 + FADD2) + 2 FFMA + FMUL + FADD            (FMA-pipe work: the roofline's flops)
         [+ FMNMX + MUFU + FSEL/FSETP are listed but NOT counted: they run on other pipes]

FMA-pipe occupancy: FFMA2/FMUL2/FADD2 hold the pipe for two issue cycles (the FFMA-only microbenchmark and an
FFMA2-only one reach the same flop rate), so pipe cycles = 2 (packed) + 1 (scalar fp32) + IMAD.
"""
import csv
import io
import json
import subprocess
import sys

PACKED = {"FFMA2": 4, "FMUL2": 2, "FADD2": 2}
SCALAR = {"FFMA": 2, "FMUL": 1, "FADD": 1}
OTHER_FP = ("FMNMX", "MUFU", "FSEL", "FSETP", "FCHK", "F2F", "F2I", "I2F", "I2FP", "FRND")
FP64 = {"DFMA": 2, "DMUL": 1, "DADD": 1}


def source_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # the first line names the kernel, the second is the header
    start = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    return rows[start], rows[start + 1:]


def main():
    rep, dst_csv, dst_json = sys.argv[1], sys.argv[2], sys.argv[3]
    nsamp = int(sys.argv[4])
    kernel_ms = float(sys.argv[5]) if len(sys.argv) > 5 else None
    h, data = source_page(rep)
    iS, iN, iT, iP = (h.index(k) for k in ("Source", "Instructions Executed", "Thread Instructions Executed",
                                            "Predicated-On Thread Instructions Executed"))
    ops = {}
    for r in data:
        if len(r) <= iP or not r[iN].isdigit():
            continue
        t = r[iS].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        d = ops.setdefault(op, [0, 0, 0])
        d[0] += int(r[iN])
        d[1] += int(r[iT])
        d[2] += int(r[iP])
    total = sum(v[0] for v in ops.values())
    with open(dst_csv, "w") as fh:
        fh.write("opcode,warp_instructions,thread_instructions,predicated_on_thread_instructions,share_of_warp_instructions,warp_instructions_per_logL\n")
        for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0]):
            fh.write("%s,%d,%d,%d,%.5f,%.2f\n" % (k, v[0], v[1], v[2], v[0] / total, v[0] / nsamp))
    flop = sum(ops.get(k, [0, 0, 0])[2] * w for k, w in {**PACKED, **SCALAR}.items())
    flop_other = sum(ops.get(k, [0, 0, 0])[2] for k in OTHER_FP)
    flop64 = sum(ops.get(k, [0, 0, 0])[2] * w for k, w in FP64.items())
    fp_warp = sum(ops.get(k, [0, 0, 0])[0] for k in list(PACKED) + list(SCALAR))
    pipe_cycles = sum(2 * ops.get(k, [0, 0, 0])[0] for k in PACKED) + sum(ops.get(k, [0, 0, 0])[0] for k in SCALAR) \
        + ops.get("IMAD", [0, 0, 0])[0]
    out = {
        "source": rep, "samples_per_launch": nsamp, "warp_instructions": total,
        "warp_instructions_per_logL": total / nsamp,
        "fp32_flop_per_logL": flop / nsamp,
        "fp32_flop_convention": "4*FFMA2 + 2*(FMUL2+FADD2) + 2*FFMA + FMUL + FADD, predicated-on thread instructions",
        "other_fp_thread_instructions_per_logL": flop_other / nsamp,
        "fp64_flop_per_logL": flop64 / nsamp,
        "fp32_arith_share_of_warp_instructions": fp_warp / total,
        "fma_pipe_issue_cycles_per_logL": pipe_cycles / nsamp,
        "fma_pipe_share_of_issue_slots": pipe_cycles / total,
        "top_opcodes": {k: round(v[0] / nsamp, 1) for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:24]},
    }
    if kernel_ms:
        out["kernel_ms"] = kernel_ms
        out["fp32_tflops"] = flop / (kernel_ms * 1e-3) / 1e12
    json.dump(out, open(dst_json, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
