"""Condenses an .ncu-rep (ncu --set full) into the few lines that are committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid (CTAs)"),
    ("launch__block_size", "block (threads)"),
    ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "SM issue-slot utilisation %"),
    ("sm__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction (of 32)"),
    ("smsp__thread_inst_executed_pred_on_per_inst_executed.ratio", "pred-on threads / warp instruction"),
    ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "FADD thread-instructions"),
    ("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "FMUL thread-instructions"),
    ("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "FFMA thread-instructions"),
    ("sm__inst_executed_pipe_xu.sum", "XU (MUFU) warp instructions"),
    ("sm__inst_executed_pipe_fma.sum", "FMA-pipe warp instructions"),
    ("sm__inst_executed_pipe_alu.sum", "ALU-pipe warp instructions"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler / cycle"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
]

SCALE = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6,
         "second": 1e3}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for launch in rows[2:]:
        get = {h: (launch[i], units[i]) for i, h in enumerate(hdr)}
        print(f"kernel {get['Kernel Name'][0]}")
        for key, label in WANT:
            if key in get:
                v, u = get[key]
                print(f"  {label:48s} {v} {u}   [{key}]")
        # --set full reports the FP32 op mix as chip-wide thread-instructions per cycle
        def rate(op):
            return float(get[f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"][0])
        fadd, fmul, ffma = rate("fadd"), rate("fmul"), rate("ffma")
        peak = float(get["sm__sass_thread_inst_executed_op_ffma_pred_on.sum.peak_sustained"][0]) * 2
        cycles = float(get["sm__cycles_elapsed.max"][0])
        t, tu = get["gpu__time_duration.sum"]
        ms = float(t) * SCALE.get(tu, 1.0)
        per_cycle = fadd + fmul + 2 * ffma
        tf = per_cycle * cycles / (ms * 1e-3) / 1e12
        print(f"  {'FADD / FMUL / FFMA thread-inst per cycle (chip)':48s} {fadd:.0f} / {fmul:.0f} / {ffma:.0f}")
        print(f"  {'hardware FP32 FLOP/cycle (fadd + fmul + 2 ffma)':48s} {per_cycle:.0f} of {peak:.0f} peak = "
              f"{per_cycle / peak * 100:.1f} % of FP32 peak  (~{tf:.1f} TFLOP/s)")
        for pipe in ("xu", "fma", "alu", "cbu", "lsu"):
            k = f"sm__inst_executed_pipe_{pipe}.avg.pct_of_peak_sustained_active"
            if k in get:
                print(f"  {'pipe ' + pipe + ' % of peak':48s} {float(get[k][0]):.1f}")
        eff = float(get["smsp__thread_inst_executed_per_inst_executed.ratio"][0]) / 32
        print(f"  {'warp execution efficiency':48s} {eff * 100:.1f} %")
        print()


if __name__ == "__main__":
    main()
