#!/usr/bin/env python3
"""Benchmark of the hot path: a 2^20-member Lorenz ensemble (rho sweep), rktp64,
fixed h = 1/250, t in [0, 50]  (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one full pass of the hot path over the ensemble: every member is
integrated from t = 0 to t = 50 (12 501 RK steps each).  `value` is RK steps/s with
inputs resident in HBM (CUDA events, max over ranks); `e2e` is the same metric
through the C ABI with host buffers (pinned H2D of y0/params and D2H of the final
states inside the timed region).  Under torchrun each rank integrates its own
2^20-member shard of the (N x 2^20)-member ensemble (weak scaling, no data-path
collective) and the ranks all-reduce one ensemble statistic over NCCL.

`--impl reference` times the reference's CPU path.  The Rust crate cannot be built
in this image (no cargo/rustc, un-vendored git dependency), so that arm runs the
C++ restatement under oracle/ (kind "port") on all host cores, on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_MEMBERS = 1 << 20
T_END = 50.0
H = 1.0 / 250.0
STEPS_PER_MEMBER = 12501  # 12500 x (1/250) accumulates to one ulp short of 50, then one closing step
FLOPS_PER_STEP = 252.0  # Lorenz + rktp64, FMA = 2, structural zeros skipped
# DRAM bytes of ONE launch of the headline kernel at 2^20 members (ncu --set full, profiles/r1_run3_*):
# 52.07 MB read + 8.42 MB written
ODE_DRAM_TRAFFIC_PER_LAUNCH = 52074752 + 8422656
WORKLOAD = "lorenz_2^20_members_rho_sweep_rktp64_h=1/250_t=[0,50]"


def lorenz_inputs(n, rank=0, world=1):
    """rho sweep over the GLOBAL ensemble; this rank's contiguous shard."""
    total = n * world
    i = np.arange(rank * n, (rank + 1) * n, dtype=np.float64)
    rho = 20.0 + 20.0 * i / (total - 1)
    prm = np.stack([np.full(n, 10.0), rho, np.full(n, 8.0 / 3.0)], axis=1)
    y0 = np.tile(np.array([1.0, 2.0, 20.0]), (n, 1))
    return np.ascontiguousarray(y0), np.ascontiguousarray(prm)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_rate(n_sample, threads):
    """RK steps/s of the CPU restatement on `threads` host threads (bounded sample)."""
    import cases
    from oracle import oracle as orc
    orc.build()
    s = cases.lorenz(n=n_sample, t1=T_END, h=H, arith="strict")
    t0 = time.perf_counter()
    r = orc.run_solver(s, n_threads=threads)
    dt = time.perf_counter() - t0
    steps = int(r.steps.sum())
    return steps / dt, dt, steps


DDE_MEMBERS = 1 << 18
DDE_DRAM_TRAFFIC_PER_LAUNCH = 50509854000 + 51415350000  # ncu --set full, profiles/r1_final_ncu_full_dde_fixed.txt
# Algorithmic bytes per RK step: one coefficient record (y, d_1..d_4) = 40 B written + one read (N=1, I=5).
# SURVEY.md §8(d) quotes 48 B records (96 B/step) because it counts the record's start time; with a fixed
# step the time grid is regenerated, so the start time carries no information and is not stored.  The
# roofline fraction is taken on the 80 B the data structure needs; the 96 B figure is reported beside it.
DDE_BYTES_PER_STEP = 80.0
DDE_BYTES_SURVEY_PER_STEP = 96.0


def dde_leg(d, torch, local_rank, K=3, W=2, rank=0, world=1):
    """Secondary workload (BASELINE configs[2]): 2^18-member constant-delay DDE
    x' = a x + b x(t-1), k sweep, rktp64, h = 0.02, t in [0, 100]; HBM-bound.
    Under torchrun every rank integrates its own 2^18-member range of the (world x 2^18)-member
    sweep (weak scaling, no data-path collective); the time is the max over ranks."""
    import cases
    n = DDE_MEMBERS
    k = 0.5 + (rank * n + np.arange(n)) / (n * world - 1.0)
    prm = cases.dde_params(k)
    problem = (d.Solver.new().max_delay(1.0).initial(d.InitFn(d.SIN)).interval(0.0, 100.0)
               .stepsize(0.02).equation(d.systems.DdeLinear, prm).arith("fast").problem())
    ens = d.Ensemble(problem, n, device=local_rank)
    d_prm = torch.from_numpy(np.ascontiguousarray(prm.T)).cuda()
    d_y0 = torch.zeros(n, dtype=torch.float64, device="cuda")
    ens.set_initial_device(d_y0.data_ptr())
    ens.set_params_device(d_prm.data_ptr())
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(W):
        ens.run(stream)
    torch.cuda.synchronize()
    ms = []
    for _ in range(K):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ens.run(stream)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    tot = ens.totals()
    t_fin, y_fin, _ = ens.final()
    err = float(np.max(np.abs(y_fin[:, 0] - np.sin(k * t_fin))))
    ens.close()
    return {"mean_ms": float(np.mean(ms)), "err": err, "steps": int(tot["steps"]), "n_failed": int(tot["n_failed"])}


def dde_report(local, world, torch, dist):
    """Combine the ranks' DDE legs (max of the time and the error, sums of the counts).  The collectives
    live here, outside the try block of the compute, so a rank whose leg failed still takes part."""
    n = DDE_MEMBERS
    ok = local is not None and "error" not in local
    vals = [local["mean_ms"], local["err"], 0.0] if ok else [0.0, 0.0, 1.0]
    sums = [float(local["steps"]), float(local["n_failed"])] if ok else [0.0, 0.0]
    if world > 1:
        mx = torch.tensor(vals, dtype=torch.float64, device="cuda")
        sm = torch.tensor(sums, dtype=torch.float64, device="cuda")
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        vals, sums = [float(x) for x in mx.tolist()], [float(x) for x in sm.tolist()]
    if vals[2] != 0.0:
        return local if (local is not None and "error" in local) else {"error": "the DDE leg failed on another rank"}
    mean_ms, err = vals[0], vals[1]
    steps, n_failed = int(sums[0]), int(sums[1])
    rate = steps / (mean_ms * 1e-3)
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm, src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        hbm, src = 6650.0, "fallback 6.65 TB/s (of fallback)"
    achieved = DDE_BYTES_PER_STEP * rate / 1e9 / world   # per GPU, against one GPU's copy bandwidth
    return {
        "workload": "dde_linear_2^18_members_per_gpu_k_sweep_tau=1_rktp64_h=0.02_t=[0,100]",
        "metric": "rk_steps_per_sec", "value": rate, "unit": "RK steps/s", "ms_per_step": mean_ms,
        "n_gpus": world, "scaling": "weak",
        "rk_steps_per_member": int(steps // (n * world)), "n_failed": n_failed,
        "max_abs_err_vs_exact_sin": err,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                     "frac": achieved / hbm, "traffic": DDE_DRAM_TRAFFIC_PER_LAUNCH,
                     "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one launch, "
                                       "profiles/r1_final_ncu_full_dde_fixed.txt (algorithmic bytes per launch: "
                                       f"{int(DDE_BYTES_PER_STEP * steps // world)})",
                     "peak_source": src,
                     "algorithmic_bytes_per_rk_step": DDE_BYTES_PER_STEP,
                     "survey_bytes_per_rk_step": DDE_BYTES_SURVEY_PER_STEP,
                     "frac_at_survey_bytes": DDE_BYTES_SURVEY_PER_STEP * rate / 1e9 / world / hbm,
                     "kernel": "dde_fixed2_kernel<SysDdeLinear,7,5,rktp64-sparsity> (two members per thread)"},
    }


def other_config_legs(d, peak_tflops):
    """The remaining BASELINE.json configs (parity-test cases, not the headline): one timed pass each
    after one warm-up, kernel time from the library's CUDA events (dfb_totals.kernel_ms)."""
    import cases
    KIND = {0: "integrate_general_kernel", 1: "ode_fixed_kernel", 2: "dde_fixed_kernel",
            3: "ode_adaptive_kernel", 4: "ode_fixed_periodic_kernel", 5: "ode_event_kernel"}

    def timed(make):
        ens = make().build()
        ens.run(); ens.synchronize()
        ms = []
        for _ in range(2):
            ens.run(); ens.synchronize()
            ms.append(ens.totals()["kernel_ms"])
        tot = ens.totals()
        ens.close()
        m = float(np.mean(ms))
        return {"kernel": KIND.get(tot["kernel_kind"], "?"), "kernel_ms": m, "rk_steps": tot["steps"],
                "rejected": tot["rejected"], "events": tot["events"], "n_failed": tot["n_failed"],
                "rk_steps_per_s": tot["steps"] / (m * 1e-3), "events_per_s": tot["events"] / (m * 1e-3)}

    out = {}
    auto = d.AutomaticStepsize(1e-2, (1e-6, 1e-1), [1e-9] * 3, [1e-9] * 3, 4, 0.9, (0.2, 5.0))
    legs = {
        # configs[1], adaptive half: 307 flop per ATTEMPTED step (252 + error estimate + controller)
        "lorenz_2^20_adaptive_rktp64_tol1e-9": (lambda: cases.lorenz(n=1 << 20, t1=T_END, arith="fast").stepsize(auto), 307.0),
        # configs[3]
        "bouncing_ball_2^19_t=[0,8]": (lambda: cases.bouncing_ball(n=1 << 19, t1=8.0, arith="fast", events=0), None),
        "egg_billiard_2^19_200_reflections": (lambda: cases.egg_billiard(n=1 << 19, reflections=200, arith="fast", events=0), None),
        # configs[4]: Lorenz + 3x3 variational block (N = 12), QR every 0.5; flop/step by the same
        # formula as the headline: 12(2*21+6) + 12*13 + 13 + 7*F_rhs, F_rhs = 8 + 45 + 1
        "lyapunov_lorenz_2^18_rho_grid_t=[0,100]": (lambda: cases.lyap_lorenz(100.0, rho=np.linspace(20.0, 40.0, 1 << 18), arith="fast"), 1123.0),
        # configs[4] on a delay-equation grid: x' = a x(t-1) + b x'(t-1), every 5th step sampled for the
        # segment-norm exponent estimator (diffurch_b200/lyapunov.py); the kernel time is what is reported
        "lyapunov_ndde_2^16_ab_grid_t=[0,100]": (lambda: d.lyapunov.linear_delay_grid(
            d.systems.NddeLinear, *[x.ravel() for x in np.meshgrid(np.linspace(-2.0, 0.5, 256), np.linspace(-0.6, 0.6, 256))])[0], None),
    }
    for name, (mk, flops) in legs.items():
        try:
            r = timed(mk)
            if flops:
                att = r["rk_steps"] + r["rejected"]
                r["algorithmic_flops_per_attempted_step"] = flops
                r["achieved_tflops"] = flops * att / (r["kernel_ms"] * 1e-3) / 1e12
                r["frac_of_fp64_peak"] = r["achieved_tflops"] / peak_tflops
            out[name] = r
        except Exception as e:  # the headline line must still print
            out[name] = {"error": repr(e)}
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_sample = int(os.environ.get("DFB_BENCH_REF_SAMPLE", 2048 * cores))  # members per step (tests shrink it)
    for _ in range(args.warmup):
        cpu_oracle_rate(max(64, n_sample // 8), cores)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt, _ = cpu_oracle_rate(n_sample, cores)
        rates.append(r)
        times.append(dt)
    value = float(np.mean(rates))
    sample = (f"{n_sample} of the 2^20 members per step (same rho sweep end points, same t range), "
              f"{cores} std::thread workers, one Solver per member")
    line = {
        "impl": "reference", "metric": "rk_steps_per_sec", "value": value, "unit": "RK steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(times) * 1e3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference CPU path; C++ restatement (Rust toolchain unavailable)"},
        "cpu_baseline": {"value": value, "unit": "RK steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "RK steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--members", type=int, default=N_MEMBERS, help=argparse.SUPPRESS)
    ap.add_argument("--arith", default="fast", choices=["fast", "strict"], help=argparse.SUPPRESS)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: diffurch_b200 has no CPU path"}))
        return 1
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import cases
    import diffurch_b200 as d

    n = args.members
    W = max(args.warmup, 3)
    K = args.steps
    y0, prm = lorenz_inputs(n, rank, world)
    problem = (d.Solver.new().equation(d.systems.Lorenz, prm).initial(y0).interval(0.0, T_END)
               .stepsize(H).arith(args.arith).problem())
    ens = d.Ensemble(problem, n, device=local_rank)

    # ---- inputs resident in HBM (SoA) for the `value` leg ----
    d_y0 = torch.from_numpy(np.ascontiguousarray(y0.T)).cuda()
    d_prm = torch.from_numpy(np.ascontiguousarray(prm.T)).cuda()
    ens.set_initial_device(d_y0.data_ptr())
    ens.set_params_device(d_prm.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        ens.run(stream)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kernel_ms = []
    stat = torch.zeros(1, dtype=torch.float64, device="cuda")
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(K):
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record()
        ens.run(stream)
        ev[k][1].record()
    # ensemble statistic across ranks: mean final z (the only collective; NCCL over NVLink)
    _, d_yfin, _ = ens.final_device_ptrs()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    tot = ens.totals()
    steps_per_pass = tot["steps"]
    assert steps_per_pass == n * STEPS_PER_MEMBER, (steps_per_pass, n * STEPS_PER_MEMBER)
    assert tot["n_failed"] == 0
    t_fin, y_fin, _ = ens.final()
    local_stat = float(np.mean(y_fin[:, 2]))
    t_tensor = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    s_tensor = torch.tensor([local_stat], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_tensor, op=dist.ReduceOp.MAX)
        dist.all_reduce(s_tensor, op=dist.ReduceOp.SUM)
    max_ms = float(t_tensor.item())
    mean_z = float(s_tensor.item()) / world
    value = float(steps_per_pass) * world * K / (max_ms * 1e-3)

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ----
    y0_pin = torch.from_numpy(y0).pin_memory().numpy()
    prm_pin = torch.from_numpy(prm).pin_memory().numpy()
    ens2 = d.Ensemble(problem, n, device=local_rank)
    out_t = torch.empty(n, dtype=torch.float64).pin_memory().numpy()       # results land in pinned memory too
    out_y = torch.empty((n, 3), dtype=torch.float64).pin_memory().numpy()
    e2e_times = []
    for it in range(2 + K):
        barrier()
        t0 = time.perf_counter()
        ens2.set_initial(y0_pin)
        ens2.set_params(prm_pin)
        ens2.run(stream)
        tf, yf, _ = ens2.final(want_aux=False, out=(out_t, out_y))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= 2:
            e2e_times.append(dt)
    e2e_t = torch.tensor([float(np.sum(e2e_times))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = float(steps_per_pass) * world * K / float(e2e_t.item())
    assert np.array_equal(yf, y_fin), "e2e result differs from the HBM-resident run"
    h2d = int(y0.nbytes + prm.nbytes)
    d2h = int(tf.nbytes + yf.nbytes)

    # ---- secondary workload (configs[2]); every rank takes part, rank 0 reports ----
    secondary = None
    if not os.environ.get("DFB_BENCH_SKIP_DDE"):
        try:
            local_dde = dde_leg(d, torch, local_rank, rank=rank, world=world)
        except Exception as e:  # the headline line must still print
            local_dde = {"error": repr(e)}
        secondary = dde_report(local_dde, world, torch, dist)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel ----
    import ctypes as C
    tf64, mhz = C.c_double(), C.c_double()
    d.lib.dfb_measure_fp64_peak(local_rank, C.byref(tf64), C.byref(mhz))
    kern_avg_ms = total_ms / K  # one launch per step; events bracket exactly that launch
    achieved_tflops = FLOPS_PER_STEP * steps_per_pass / (kern_avg_ms * 1e-3) / 1e12
    # Denominator: MEASURED_PEAKS.json has no FP64 entry and B200_PROFILING.md states no FP64
    # fallback, so the peak is measured here (DFMA-chain microbenchmark, best of 2 operand
    # patterns x 2 chain counts x 4 occupancies).  The pipe's architectural rate at the clock
    # observed during the timed region (SMs x 64 DFMA/clk x 2 x f) is reported next to it, and
    # the LARGER of the two is the denominator, so the fraction can only be understated.
    props = torch.cuda.get_device_properties(local_rank)
    sm_mhz_obs = (clocks or {}).get("sm_mhz") or 1965.0
    arch_peak = props.multi_processor_count * 64 * 2 * sm_mhz_obs * 1e6 / 1e12
    peak = max(tf64.value, arch_peak)
    roofline = {
        "bound": "fp64", "achieved": achieved_tflops, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved_tflops / peak,
        "traffic": ODE_DRAM_TRAFFIC_PER_LAUNCH if n == N_MEMBERS else None,
        "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one launch, "
                          "profiles/r1_run3_ncu_full_ode_fixed_lorenz.txt (algorithmic HBM bytes per launch: "
                          f"{100 * n}; the kernel is FP64-pipe-bound, not HBM-bound)",
        "peak_source": "max(DFMA-chain microbenchmark measured in this run, SMs x 64 DFMA/clk x 2 x SM clock "
                       "sampled during the timed region); MEASURED_PEAKS.json has no FP64 entry",
        "peak_measured_dfma_chain": tf64.value,
        "peak_architectural_at_observed_clock": arch_peak,
        "peak_sm_mhz": mhz.value,
        "algorithmic_flops_per_rk_step": FLOPS_PER_STEP,
        "kernel": "ode_fixed_kernel<SysLorenz,7,rktp64-sparsity,FAST>",
        "kernel_ms": kern_avg_ms,
        "hbm_bytes_per_member": 8 * (3 + 3) + 8 * (1 + 3) + 8 + 8 + 4,
    }

    # ---- CPU baseline (oracle "port"), bounded sample on this box's host cores ----
    cores = os.cpu_count() or 1
    probe_rate, _, _ = cpu_oracle_rate(256 * cores, cores)       # sizes the sample for ~15 s of CPU work
    n_sample = int(max(1024 * cores, min(N_MEMBERS, float(os.environ.get("DFB_BENCH_CPU_SECONDS", "15")) * probe_rate / STEPS_PER_MEMBER)))
    n_sample -= n_sample % cores
    cpu_rate, cpu_dt, _ = cpu_oracle_rate(n_sample, cores)
    cpu1_rate, _, _ = cpu_oracle_rate(64, 1)
    cpu_baseline = {
        "value": cpu_rate, "unit": "RK steps/s", "cores": cores, "kind": "port",
        "sample": f"{n_sample} members x 12501 steps ({cpu_dt:.1f} s), C++ restatement of the reference "
                  f"CPU path (Rust toolchain unavailable), std::thread over members",
        "single_core_ns_per_step": 1e9 / cpu1_rate,
    }

    others = None
    if world == 1 and not os.environ.get("DFB_BENCH_SKIP_OTHERS"):
        try:
            others = other_config_legs(d, peak)
        except Exception as e:
            others = {"error": repr(e)}

    line = {
        "metric": "rk_steps_per_sec", "value": value, "unit": "RK steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": max_ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "members_per_gpu": n, "rk_steps_per_member": STEPS_PER_MEMBER,
                   "arith": args.arith, "tableau": "rktp64 (S=7)", "sharding": f"trajectory range x{world}",
                   "l2_flush": "256 MiB memset before every timed step (outside the event pair)",
                   "ensemble_stat_mean_final_z": mean_z},
        "e2e": {"value": e2e_value, "unit": "RK steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": float(np.mean(e2e_times) * 1e3)},
        "gpu_launches": K,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "wall_s_timed_region": t_wall,
        "secondary_dde": secondary,
        "other_configs": others,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
