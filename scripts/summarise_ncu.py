"""Turn an .ncu-rep capture into the text summary committed under profiles/ (key raw metrics + SASS opcode mix).

    python scripts/summarise_ncu.py gpurun_out/prof.ncu-rep profiles/r1_xxx.txt --ray-steps 7928453664
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
ray_steps = float(sys.argv[sys.argv.index("--ray-steps") + 1]) if "--ray-steps" in sys.argv else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, r = rows[0], rows[1], rows[2]
g = lambda k: r[h.index(k)] if k in h else "n/a"
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none, kernel: {r[h.index('Kernel Name')]}\n# source: {rep}\n")
    for k in KEYS:
        f.write(f"{k} = {g(k)} [{u[h.index(k)] if k in h else ''}]\n")
    st = {k: float(g(k)) for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
    f.write("\n# warp stall reasons (average warps stalled per issue-active cycle)\n")
    for k, v in sorted(st.items(), key=lambda kv: -kv[1]):
        if v > 0.01:
            f.write(f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} = {v:.3f}\n")
    try:
        cyc = float(g("sm__cycles_elapsed.avg"))
        ops = {k: float(g(f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed")) for k in ("dfma", "dmul", "dadd")}
        ghz = cyc / (float(g("gpu__time_duration.sum")) * 1e-3) / 1e9
        tf = (2 * ops["dfma"] + ops["dmul"] + ops["dadd"]) * ghz * 1e9 / 1e12
        f.write(f"\n# derived\nsm_clock_GHz = {ghz:.3f}\nfp64_flops_ncu_counted_TFLOPs = {tf:.2f}   (2*DFMA + DMUL + DADD thread-ops per second)\n")
        f.write(f"fp64_thread_ops_per_cycle = {sum(ops.values()):.1f} of 9472 (148 SM x 64 lanes)\n")
        if ray_steps:
            f.write(f"fp64_thread_ops_per_ray_step = dfma {ops['dfma'] * cyc / ray_steps:.1f}, dmul {ops['dmul'] * cyc / ray_steps:.1f}, "
                    f"dadd {ops['dadd'] * cyc / ray_steps:.1f}\n")
    except Exception as e:  # noqa: BLE001
        f.write(f"# derived metrics unavailable: {e}\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    hi = [i for i, rr in enumerate(srows) if rr and rr[0] == "Address"]
    if hi:
        hh = srows[hi[0]]
        ci = {n: i for i, n in enumerate(hh)}
        end = hi[1] - 1 if len(hi) > 1 else len(srows)
        by, samp = collections.Counter(), collections.Counter()
        for rr in srows[hi[0] + 1:end]:
            if len(rr) < len(hh):
                continue
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", rr[ci["Source"]].strip())
            op = m.group(2) if m else "?"
            by[op] += int(rr[ci["Thread Instructions Executed"]] or 0)
            samp[op] += int(rr[ci["# Samples"]] or 0)
        tot, stot = sum(by.values()), max(1, sum(samp.values()))
        f.write("\n# SASS opcode mix (thread instructions executed; share of stall samples)\n")
        for op, n in by.most_common(24):
            per = f"{n / ray_steps:8.1f} per ray-step" if ray_steps else ""
            f.write(f"{op:10s} {100 * n / tot:6.2f} % {per}   samples {100 * samp[op] / stot:5.2f} %\n")
        if ray_steps:
            f.write(f"TOTAL      {tot / ray_steps:.1f} thread instructions per ray-step\n")
            # the issue model: an FP64 instruction holds the SMSP's dispatch port for two cycles, anything else
            # for one -> cycles per warp-step and SMSP = 2 F + O, to be compared with the measured active cycles
            F = sum(by[k] for k in ("DFMA", "DMUL", "DADD", "DSETP")) / ray_steps
            O = tot / ray_steps - F
            try:
                n_smsp = 4 * 148
                measured = float(g("sm__cycles_active.avg")) / (ray_steps / 32.0 / n_smsp)
                f.write(f"\n# issue model: F = {F:.1f} FP64-pipe instructions (incl. DSETP), O = {O:.1f} others per ray-step\n"
                        f"issue_model_cycles_per_warp_step = 2 F + O = {2 * F + O:.0f}\n"
                        f"measured_cycles_per_warp_step_per_SMSP = {measured:.0f}   (sm__cycles_active / (ray-steps / 32 / 592 SMSPs))\n")
            except Exception as e:  # noqa: BLE001
                f.write(f"# issue model unavailable: {e}\n")
print("wrote", out)
