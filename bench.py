#!/usr/bin/env python
"""bench.py -- headline benchmark: rays/s of the lamppost emissivity workload (a = 0.998, h = 10 rg, RK4).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of rays: the fused CUDA kernel generates the rays of
this rank's shard of the PointSource grid, integrates every geodesic (RK4, reference step schedule), computes
redshifts and bins the disc hits into the radial emissivity histogram; with N > 1 the per-GPU partial
histograms are then combined with one NCCL all-reduce.  Weak scaling: RAYS_PER_GPU rays per rank per step
(1.25e8 by default = the 1e9-ray BASELINE workload on 8 GPUs), the global grid gets finer in cos(alpha).

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SPIN = 0.998
SOURCE = [0.0, 10.0, 1e-3, 1.5707]       # BASELINE config 1 / SURVEY 8d: lamppost at h = 10 rg
COS0, COSMAX = -0.995, 0.995
BETA0, BETAMAX = -3.14159265358979, 3.14159265358979
NR = 100
FLOP_PER_RK4_STEP = 259.0                  # algorithmic FP64 operations per RK4 step
METRIC = "rays/sec (lamppost emissivity, a=0.998)"


def profiled_dram_traffic(math):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fused kernel from the committed ncu --set full capture
    (profiles/, one launch of 1.5e7 rays); None if the summary is missing."""
    path = os.path.join(ROOT, "profiles", f"r2_emis_rk4_{math}.txt")
    try:
        total, unit = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for line in open(path):
            if line.startswith(("dram__bytes_read.sum", "dram__bytes_write.sum")):
                val, u = line.split("=")[1].split()
                total += float(val) * unit[u.strip("[]")]
        return total
    except Exception:
        return None


def hbm_view(traffic_bytes):
    """The HBM side of the roofline for completeness: the fused kernel moves (almost) nothing."""
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "MEASURED_PEAKS.json"
    except Exception:
        peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    return {"peak_gbs": peak, "peak_source": src, "dram_bytes_profiled_launch": traffic_bytes,
            "note": "profiled launch: 1.5e7 rays in ~0.28 s -> ~0.05 GB/s, 1e-5 of the HBM roofline: the kernel is FP64-issue-bound"}


def grid_for(total_rays, n_beta=10000):
    """PointSource grid with `total_rays` live rays: n_beta azimuths x (total_rays / n_beta) polar rows."""
    n_rows = max(1, int(round(total_rays / n_beta)))
    dcos = (COSMAX - COS0) / n_rows
    dbeta = (BETAMAX - BETA0) / n_beta
    return dcos, dbeta, n_rows * n_beta


# ---------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi in the background during the timed region)
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                power.append(float(r[3]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}


# ---------------------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle/_ref when the reference compiled, else the oracle port)
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_run(n_rays, end_to_end=False):
    """Time run_raytrace() of the reference CPU path on a lamppost grid of ~n_rays rays with all host threads,
    exactly as src/tests/integrator_perf_test.cpp:74-79 does (chrono around run_raytrace only)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bindings as ob
    from raytrace_cpu_b200 import _abi as A

    dcos, dbeta, _ = grid_for(n_rays, n_beta=max(100, int(round(math.sqrt(n_rays)))))
    src = ob.make_pointsource(SOURCE, dcos, dbeta, cosalpha0=COS0, cosalphamax=COSMAX, beta0=BETA0, betamax=BETAMAX)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    if ob.ref_available():
        kind = "reference"
        ref = ob.Ref()
        cores = ref.set_num_threads(cores)      # torchrun exports OMP_NUM_THREADS=1: use every host thread anyway
        t_all0 = time.perf_counter()
        h = ref.pointsource(src, SPIN, A.TOL)
        h.redshift_start()
        secs = h.run(ob.make_stop(), ob.make_run(A.RK4))
        h.range_phi()
        h.redshift(-1.0)
        t_all = time.perf_counter() - t_all0
        rays = h.rays(copy=False)
        traced = int((rays["steps"] != -1).sum())
        steps = int(np.abs(rays["steps"][rays["steps"] != -1].astype(np.int64)).sum())
        del rays, h
    else:
        kind = "port"
        orc = ob.Oracle()
        tr = orc.tracer(SPIN, A.TOL)
        t_all0 = time.perf_counter()
        rays = orc.pointsource_rays(tr, src)
        orc.redshift_start(rays, tr, 0.0)
        t0 = time.perf_counter()
        orc.trace(rays, tr, ob.make_stop(), ob.make_run(A.RK4), cores)
        secs = time.perf_counter() - t0
        orc.range_phi(rays)
        orc.redshift(rays, tr, -1.0)
        t_all = time.perf_counter() - t_all0
        traced = int((rays["steps"] != -1).sum())
        steps = int(np.abs(rays["steps"][rays["steps"] != -1].astype(np.int64)).sum())
    return {"kind": kind, "cores": cores, "rays": traced, "steps": steps, "trace_seconds": secs, "e2e_seconds": t_all}


def cpu_baseline(budget_s=15.0):
    """Bounded sample of the same workload on the host cores: a pilot sizes the sample to ~budget_s."""
    pilot = cpu_reference_run(60_000)
    rate = pilot["rays"] / max(pilot["trace_seconds"], 1e-6)
    n = int(min(4_000_000, max(125_000, rate * budget_s)))
    run = cpu_reference_run(n)
    return {
        "value": run["rays"] / run["trace_seconds"], "unit": "rays/s", "cores": run["cores"], "kind": run["kind"],
        "sample": f"lamppost h=10 a=0.998 RK4 grid of {run['rays']} rays ({run['steps'] / run['rays']:.0f} steps/ray), "
                  f"run_raytrace() only timed with chrono as integrator_perf_test.cpp:74-79, OMP threads = {run['cores']}, "
                  f"-O2 -fopenmp -ffp-contract=off; {run['trace_seconds']:.2f} s "
                  f"(incl. serial init/redshift: {run['e2e_seconds']:.2f} s)",
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pilot = cpu_reference_run(60_000)
    rate = pilot["rays"] / max(pilot["trace_seconds"], 1e-6)
    n = int(min(4_000_000, max(60_000, rate * args.ref_step_seconds)))
    for _ in range(args.warmup):
        cpu_reference_run(n)
    t, rays, steps, last = 0.0, 0, 0, None
    for _ in range(args.steps):
        last = cpu_reference_run(n)
        t += last["trace_seconds"]
        rays += last["rays"]
        steps += last["steps"]
    value = rays / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"lamppost PointSource h=10 rg, a=0.998, RK4, theta_max=pi/2, r_max=1000: bounded sample of "
                               f"{last['rays']} rays per step on the host CPU (rays/s is size independent)",
                   "rays_per_step": last["rays"], "steps_per_ray": steps / rays},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": last["cores"], "kind": last["kind"],
                         "sample": f"{last['rays']} rays per step, run_raytrace() only, all {last['cores']} host threads"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kerrtrace", choices=["kerrtrace", "reference"])
    ap.add_argument("--rays-per-gpu", type=float, default=1.25e8, help="non-default values are for profiling only")
    ap.add_argument("--e2e-rays", type=float, default=3.2e7, help="rays per rank per end-to-end step (host AoS buffers)")
    ap.add_argument("--math", default="exact", choices=["exact", "fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-mode", action="store_true", help="skip the informational run in the other arithmetic mode")
    ap.add_argument("--ref-step-seconds", type=float, default=6.0)
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    import raytrace_cpu_b200 as K
    from raytrace_cpu_b200 import _abi as A

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries exactly one JSON line: while the benchmark runs, file descriptor 1 points at stderr, so that whatever a
    # library writes there (NCCL prints its version on communicator creation) cannot precede the line; restored before printing
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    ctx = K.get_context(local_rank)
    # everything (torch fills, our kernels, NCCL) is ordered on ONE explicit stream, and the CUDA events that
    # time the kernels are recorded on that same stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    math_mode = A.MATH_FAST if args.math == "fast" else A.MATH_EXACT

    # ---- workload: global grid of world * rays_per_gpu rays, chunks of rays dealt round-robin ---------------
    total = int(args.rays_per_gpu) * world
    dcos, dbeta, live = grid_for(total)
    src = K.pointsource_desc(SOURCE, 0.0, dcos, dbeta, COS0, COSMAX, BETA0, BETAMAX)
    n_slots = ctx.L.kt_pointsource_count(C.byref(src), None, None)
    bins = K.emissivity_bins(SPIN, Nr=NR, src=src)
    shard = A.Shard()
    shard.begin, shard.end, shard.rank, shard.nranks, shard.chunk = 0, n_slots, rank, world, 0
    hist = torch.zeros(5 * NR + A.NCOUNTERS, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(ev0=None, ev1=None, mode=None):
        hist.zero_()
        if ev0 is not None:
            ev0.record(stream)
        K.emissivity_dev(src, SPIN, bins, hist.data_ptr(), method=K.Integrator.RK4, shard=shard, ctx=ctx,
                         math_mode=math_mode if mode is None else mode)
        if ev1 is not None:
            ev1.record(stream)
        if world > 1:
            dist.all_reduce(hist)          # NCCL sum of the per-GPU partial histograms (5*Nr+8 doubles)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_tflops = ctx.fp64_peak_tflops()
    for _ in range(args.warmup):
        step()
        flush.zero_()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record(stream)
    for i in range(args.steps):
        step(*kev[i])
        flush.zero_()                      # L2 flush between timed iterations
    t1.record(stream)
    barrier()
    elapsed_ms = t0.elapsed_time(t1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    launches = ctx.kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([elapsed_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    elapsed_ms, kernel_ms = tmax.tolist()

    h = hist.cpu().numpy()
    res = A.EmisResult.from_packed(NR, h)
    cnt = res.counters                     # after the all-reduce: whole-job counters of the LAST step
    traced_per_step = int(cnt[A.CNT_TRACED])
    steps_per_step = int(cnt[A.CNT_STEPS])
    value = traced_per_step * args.steps / (elapsed_ms * 1e-3)

    # ---- for information: the same steps in the other arithmetic mode (not the headline) -------------
    other = None
    if not args.no_other_mode:
        omode, oname = (A.MATH_EXACT, "exact") if args.math == "fast" else (A.MATH_FAST, "fast")
        step(mode=omode)
        flush.zero_()
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record(stream)
        for i in range(args.steps):
            step(mode=omode)
            flush.zero_()
        o1.record(stream)
        barrier()
        oms = torch.tensor([o0.elapsed_time(o1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(oms, op=dist.ReduceOp.MAX)
        ocnt = A.EmisResult.from_packed(NR, hist.cpu().numpy()).counters
        other = {"math": oname, "value": int(ocnt[A.CNT_TRACED]) * args.steps / (oms.item() * 1e-3), "unit": "rays/s",
                 "ms_per_step": oms.item() / args.steps, "warmup": 1,
                 # same accounting as `roofline`: 259 algorithmic flop per RK4 step of the reference's expressions, per GPU
                 "roofline_frac": FLOP_PER_RK4_STEP * (int(ocnt[A.CNT_STEPS]) / world) / (oms.item() / args.steps * 1e-3) / 1e12 / peak_tflops,
                 "note": "same workload and timing rules; 'fast' = the reference's right-hand side evaluated through algebraically "
                         "equivalent, cheaper expressions with FMA contraction (kt_trace.cuh: rk4_step_lean); same step schedule; "
                         "per-ray deviations are in the class of the reference's own -march=native build (0.26 % of rays beyond "
                         "1e-9 on the reference's test grid, the same rays the FMA-contracted reference moves)"}

    # ---- end to end through the reference-facing C ABI call with HOST buffers (kt_trace on Ray<double>[]) ----
    e2e = None
    if not args.no_e2e:
        # rank r owns polar rows r, r+world, r+2*world, ... of a global grid of world*e2e_rays rays
        g_dcos, e_dbeta, _ = grid_for(int(args.e2e_rays) * world, n_beta=2000)
        tr = K.api.make_tracer(SPIN, A.TOL)
        # the app's own set-up (PointSource ctor + redshift_start) happens once, outside the timed region
        mine = K.PointSource(SOURCE, 0.0, SPIN, A.TOL, g_dcos * world, e_dbeta, COS0 + rank * g_dcos, COSMAX, BETA0, BETAMAX, ctx=ctx)
        mine.redshift_start()
        n_e = mine.get_count()
        # `mine` keeps the pristine initial rays in HBM; before every step (outside the timed region) they are
        # written back into the application-side host array `work` (pinned, like the shim's Raytracer::rays)
        work = torch.empty(n_e * 144, dtype=torch.uint8).pin_memory()
        stop = K.api.thetalim_stop(math.pi / 2)
        run = K.api.make_run(K.Integrator.RK4, 1000.0, -1, math_mode)
        times = []
        for i in range(args.warmup + args.steps):
            K.lib.check(ctx.L.kt_rayset_download(mine._rs, C.c_void_p(work.data_ptr()), n_e))
            barrier()
            ts = time.perf_counter()
            K.lib.check(ctx.L.kt_trace(ctx.h, C.byref(tr), C.byref(stop), C.byref(run), C.c_void_p(work.data_ptr()), n_e))
            torch.cuda.synchronize()
            te = time.perf_counter() - ts
            if i >= args.warmup:
                times.append(te)
        tt = torch.tensor([sum(times)], dtype=torch.float64, device=dev)
        nn = torch.tensor([float(n_e)], dtype=torch.float64, device=dev)
        by_rank = [1e3 * sum(times) / args.steps]
        if world > 1:
            gathered = [torch.zeros_like(tt) for _ in range(world)]
            dist.all_gather(gathered, tt)
            by_rank = [1e3 * g.item() / args.steps for g in gathered]
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(nn)
        rays_e = work.numpy().view(A.RAY_DTYPE)
        traced_e = int((rays_e["steps"] != -1).sum())
        frac_live = traced_e / max(n_e, 1)
        zero_copy = os.environ.get("KT_ZEROCOPY", "1") != "0"
        if zero_copy:
            # the kernel reads every record whole (a warp copies the 32 consecutive 144-byte records of a refill as one coalesced
            # span) and writes back t..pphi + the six ints of every traced ray (88 bytes)
            h2d = n_e * 144
            d2h = traced_e * 88
        else:
            h2d = d2h = n_e * 144
        e2e = {"value": nn.item() * frac_live * args.steps / tt.item(), "unit": "rays/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "call": "kt_trace(ctx, tracer, stop, run, Ray<double>* host_rays, n)  == Raytracer<double>::run_raytrace on rays[]",
               "transfer": ("zero-copy: the persistent kernel reads the records from the pinned host array over PCIe when a warp refills "
                            "(32 consecutive records per coalesced load) and writes a ray's results back when it retires "
                            "(no staging copy)") if zero_copy
                           else "staged: cudaMemcpyAsync H2D -> kernel -> cudaMemcpyAsync D2H",
               "rays_per_step_per_gpu": int(n_e), "ms_per_step": 1e3 * tt.item() / args.steps,
               "ms_per_step_by_rank": [round(x, 1) for x in by_rank]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()

    if rank == 0:
        alg_tflops = FLOP_PER_RK4_STEP * (steps_per_step / world) / (kernel_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"lamppost PointSource h=10 rg, a=0.998, RK4 (reference step schedule, math={args.math}), theta_max=pi/2, "
                            f"r_max=1000, Nr=100 log bins: {traced_per_step} live rays per step ({int(args.rays_per_gpu)} per GPU), "
                            f"fused init+trace+redshift+binning kernel, NCCL all-reduce of the histogram",
                "rays_per_step": traced_per_step, "rays_per_gpu": int(args.rays_per_gpu), "steps_per_ray": steps_per_step / traced_per_step,
                "grid": {"dcosalpha": dcos, "dbeta": dbeta, "slots": int(n_slots)}, "shard": "chunks of <= 65536 rays dealt round-robin over ranks (kt_shard, automatic chunk size)",
                "math": args.math,
                "l2": "rays are generated on device (no input buffers); a 256 MB buffer is written between timed steps anyway",
            },
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {
                "bound": "fp64", "achieved": alg_tflops, "peak": peak_tflops, "unit": "TFLOP/s", "frac": alg_tflops / peak_tflops,
                "traffic": profiled_dram_traffic(args.math),
                "traffic_note": "DRAM bytes (read + write) of ONE launch of this kernel on 1.5e7 rays from the committed ncu --set full "
                                "capture (profiles/r2_emis_rk4_*.txt): the kernel has no input or output arrays, the few MB are "
                                "local-memory lines evicted from L2; algorithmic HBM traffic is the 4 KB histogram",
                "how": f"achieved = {FLOP_PER_RK4_STEP:.0f} algorithmic FP64 flop per RK4 step (SURVEY 8d) x {steps_per_step // world} steps "
                       f"per launch / {kernel_ms:.1f} ms (CUDA events around the fused kernel, mean over timed steps, max over ranks); "
                       "peak = DFMA micro-benchmark kt_fp64_peak() measured in this run (MEASURED_PEAKS.json has no FP64 entry); "
                       "HBM traffic of the kernel is ~0 (rays never leave the SM), see profiles/ for the ncu-counted FP64 pipe view. "
                       "The binding resource is the SMSP dispatch port: an FP64 instruction holds it for two cycles, anything else "
                       "for one, and the measured cycles per warp-step equal 2 x (FP64 instructions) + (other instructions) to 1 % "
                       "(DESIGN.md section 4), so frac is capped by the instruction mix of an IEEE division (6 FP64 + 5 other) "
                       "and root (8 + 4), not by the DFMA rate",
                "kernel_ms": kernel_ms,
                "hbm": hbm_view(profiled_dram_traffic(args.math)),
            },
            "cpu_baseline": cpu,
            "other_math_mode": other,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
