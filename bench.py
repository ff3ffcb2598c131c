#!/usr/bin/env python3
"""Benchmark of the hot path on the two BASELINE ensembles (BASELINE.json configs[1], configs[2];
SURVEY.md §8d configs 2 and 3):

  * 2^20-member Lorenz ensemble (rho sweep), rktp64, fixed h = 1/250, t in [0, 50]  — FP64-pipe bound;
  * 2^18-member constant-delay DDE ensemble x' = a x + b x(t - 1) (k sweep), rktp64, h = 0.02,
    t in [0, 100]                                                                    — HBM bound.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling strong|weak]

One "step" = one full pass of the hot path over the ensemble (every member integrated over the whole
interval: 12 501 RK steps per Lorenz member, 5 001 per DDE member).  `value` is RK steps/s of the Lorenz
ensemble with inputs resident in HBM (CUDA events on the launching stream, max over ranks); `e2e` is the
same metric through the C ABI with host buffers (pinned H2D of y0/params and D2H of the final states
inside the timed region).  The DDE ensemble is reported with the same three blocks under
`roofline.dde`, `cpu_baseline.dde`, `e2e.dde` (and `secondary_dde` for its value).

Multi-GPU (torchrun, one rank per GPU).  Default `--scaling strong`: the BASELINE ensembles (2^20 / 2^18
members TOTAL) are split by contiguous member range, every rank integrates total/N members, no data-path
collective; the ranks all-reduce one ensemble statistic over NCCL.  `--scaling weak` gives every rank a
full-size shard; under the default a short weak-scaling pass is also run and reported as `weak_scaling`.

`--impl reference` times the reference's CPU path.  The Rust crate cannot be built in this image (no
cargo/rustc, un-vendored git dependency), so that arm runs the C++ restatement under oracle/ (kind
"port", built -O3 -ffp-contract=off) on all host cores, on a bounded sample of the same workload.
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

DDE_MEMBERS = 1 << 18
DDE_STEPS_PER_MEMBER = 5001
DDE_WORKLOAD = "dde_linear_2^18_members_k_sweep_tau=1_rktp64_h=0.02_t=[0,100]"
DDE_DRAM_TRAFFIC_PER_LAUNCH = 48981858000 + 50912352000  # ncu --set full, profiles/r2_final_ncu_full_dde_fixed.txt
# Algorithmic bytes per RK step: one coefficient record (y, d_1..d_4) = 40 B written + one read (N=1, I=5).
# SURVEY.md §8(d) quotes 48 B records (96 B/step) because it counts the record's start time; with a fixed
# step the time grid is regenerated, so the start time carries no information and is not stored.  The
# roofline fraction is taken on the 80 B the data structure needs; the 96 B figure is reported beside it.
DDE_BYTES_PER_STEP = 80.0
DDE_BYTES_SURVEY_PER_STEP = 96.0


def shard_bounds(total, rank, world):
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def lorenz_inputs(total, lo, hi):
    """rho sweep over the GLOBAL ensemble of `total` members; rows [lo, hi)."""
    i = np.arange(lo, hi, dtype=np.float64)
    n = hi - lo
    rho = 20.0 + 20.0 * i / (total - 1)
    prm = np.stack([np.full(n, 10.0), rho, np.full(n, 8.0 / 3.0)], axis=1)
    y0 = np.tile(np.array([1.0, 2.0, 20.0]), (n, 1))
    return np.ascontiguousarray(y0), np.ascontiguousarray(prm)


def dde_k(total, lo, hi):
    return 0.5 + np.arange(lo, hi, dtype=np.float64) / (total - 1.0)


def make_config(scaling, world, arith="fast"):
    """Identical for the native and the reference arm (the driver compares the dicts)."""
    return {"workload": WORKLOAD, "members_total": N_MEMBERS if scaling == "strong" else N_MEMBERS * world,
            "rk_steps_per_member": STEPS_PER_MEMBER, "tableau": "rktp64 (S=7)", "arith": arith,
            "scaling": scaling, "secondary_workload": DDE_WORKLOAD}


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


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle "port"): the only places bench.py executes oracle/
# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(n_sample, threads):
    """RK steps/s of the CPU restatement on `threads` host threads (bounded sample of the Lorenz sweep)."""
    import cases
    from oracle import oracle as orc
    orc.build()
    s = cases.lorenz(n=n_sample, t1=T_END, h=H, arith="strict")
    t0 = time.perf_counter()
    r = orc.run_solver(s, n_threads=threads)
    dt = time.perf_counter() - t0
    steps = int(r.steps.sum())
    return steps / dt, dt, steps


def cpu_oracle_rate_dde(n_sample, threads):
    """The same for the DDE ensemble: `n_sample` members spread over the k sweep, t in [0, 100]."""
    import cases
    from oracle import oracle as orc
    orc.build()
    k = 0.5 + np.arange(n_sample) / max(n_sample - 1.0, 1.0)
    s = cases.eq_dde_sin(k=k, t1=100.0, trace=False, arith="strict")
    t0 = time.perf_counter()
    r = orc.run_solver(s, n_threads=threads)
    dt = time.perf_counter() - t0
    steps = int(r.steps.sum())
    return steps / dt, dt, steps


def sized_cpu_baseline(rate_fn, steps_per_member, cores, seconds, label):
    probe_rate, _, _ = rate_fn(128 * cores, cores)  # sizes the sample for ~`seconds` of CPU work
    n_sample = int(max(32 * cores, seconds * probe_rate / steps_per_member))
    n_sample -= n_sample % cores
    rate, dt, _ = rate_fn(n_sample, cores)
    rate1, _, _ = rate_fn(8, 1)
    return {"value": rate, "unit": "RK steps/s", "cores": cores, "kind": "port",
            "sample": f"{n_sample} members x {steps_per_member} steps of {label} ({dt:.1f} s), C++ restatement of the "
                      f"reference CPU path built -O3 -ffp-contract=off (Rust toolchain unavailable), std::thread "
                      f"over members, one Solver per member",
            "single_core_ns_per_step": 1e9 / rate1}


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
              f"{cores} std::thread workers, one Solver per member; oracle built -O3 -ffp-contract=off")
    cpu = {"value": value, "unit": "RK steps/s", "cores": cores, "kind": "port", "sample": sample}
    if not os.environ.get("DFB_BENCH_SKIP_DDE"):
        n_dde = max(cores, int(os.environ.get("DFB_BENCH_REF_SAMPLE", 4096 * cores)) // 4)
        dr, ddt, _ = cpu_oracle_rate_dde(n_dde, cores)
        cpu["dde"] = {"value": dr, "unit": "RK steps/s", "cores": cores, "kind": "port",
                      "sample": f"{n_dde} members of the k sweep x {DDE_STEPS_PER_MEMBER} steps ({ddt:.1f} s)"}
    line = {
        "impl": "reference", "metric": "rk_steps_per_sec", "value": value, "unit": "RK steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(times) * 1e3), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args.scaling, world),
        "note": "reference CPU path; C++ restatement (Rust toolchain unavailable); each step a bounded sample, as a rate",
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "RK steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    """torch / distributed plumbing shared by the legs."""

    def __init__(self, torch, dist, rank, world, local_rank):
        self.torch, self.dist, self.rank, self.world, self.local_rank = torch, dist, rank, world, local_rank
        self.stream = torch.cuda.current_stream().cuda_stream
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return [float(x) for x in t.tolist()]


def timed_resident(ctx, ens, K, W):
    """W warm-ups, then K passes with device-resident inputs; per-pass CUDA-event times (ms) on the
    launching stream, an L2 flush before every pass (outside the event pair)."""
    torch = ctx.torch
    for _ in range(W):
        ens.run(ctx.stream)
    ctx.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    ctx.barrier()
    t0 = time.perf_counter()
    for k in range(K):
        ctx.flush.zero_()
        ev[k][0].record()
        ens.run(ctx.stream)
        ev[k][1].record()
    ctx.barrier()
    wall = time.perf_counter() - t0
    return [a.elapsed_time(b) for a, b in ev], wall


def timed_e2e(ctx, ens, y0_pin, prm_pin, out, K, warm=2):
    """The call a user makes, host buffers in and out: set_initial + set_params (pinned H2D + device
    transpose), run, get_final (device transpose + D2H into pinned memory).  Wall clock, barrier +
    synchronize on both sides of every pass.  Then two more passes with a synchronize between the phases
    to attribute the time (those are not part of the e2e number)."""
    torch = ctx.torch
    times = []
    res = None
    for it in range(warm + K):
        ctx.barrier()
        t0 = time.perf_counter()
        ens.set_initial(y0_pin)
        ens.set_params(prm_pin)
        ens.run(ctx.stream)
        res = ens.final(want_aux=False, out=out)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= warm:
            times.append(dt)
    phases = {"h2d_and_transpose_ms": [], "kernel_ms": [], "transpose_and_d2h_ms": []}
    for _ in range(2):
        ctx.barrier()
        t0 = time.perf_counter()
        ens.set_initial(y0_pin)
        ens.set_params(prm_pin)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ens.run(ctx.stream)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ens.final(want_aux=False, out=out)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        phases["h2d_and_transpose_ms"].append((t1 - t0) * 1e3)
        phases["kernel_ms"].append((t2 - t1) * 1e3)
        phases["transpose_and_d2h_ms"].append((t3 - t2) * 1e3)
    return times, res, {k: float(np.mean(v)) for k, v in phases.items()}


def pinned(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


def lorenz_leg(ctx, d, total, K, W, arith, want_e2e=True):
    """Lorenz ensemble of `total` members split over the ranks.  Returns this rank's measurements with the
    cross-rank reductions already applied."""
    torch = ctx.torch
    lo, hi = shard_bounds(total, ctx.rank, ctx.world)
    n = hi - lo
    y0, prm = lorenz_inputs(total, lo, hi)
    problem = (d.Solver.new().equation(d.systems.Lorenz, prm).initial(y0).interval(0.0, T_END)
               .stepsize(H).arith(arith).problem())
    ens = d.Ensemble(problem, n, device=ctx.local_rank)
    d_y0 = torch.from_numpy(np.ascontiguousarray(y0.T)).cuda()
    d_prm = torch.from_numpy(np.ascontiguousarray(prm.T)).cuda()
    ens.set_initial_device(d_y0.data_ptr())
    ens.set_params_device(d_prm.data_ptr())
    step_ms, wall = timed_resident(ctx, ens, K, W)
    tot = ens.totals()
    assert tot["steps"] == n * STEPS_PER_MEMBER, (tot["steps"], n * STEPS_PER_MEMBER)
    assert tot["n_failed"] == 0
    _, y_fin, _ = ens.final()
    total_ms = float(sum(step_ms))
    max_ms, = ctx.reduce([total_ms], "MAX")
    sum_z, sum_steps = ctx.reduce([float(np.sum(y_fin[:, 2])), float(tot["steps"])], "SUM")
    out = {"n_local": n, "total_ms_local": total_ms, "max_ms": max_ms, "steps_all_ranks_per_pass": sum_steps,
           "value": sum_steps * K / (max_ms * 1e-3), "mean_final_z": sum_z / total, "wall_s": wall,
           "kernel_ms_local": total_ms / K}
    if want_e2e:
        ens2 = d.Ensemble(problem, n, device=ctx.local_rank)
        y0_pin, prm_pin = pinned(torch, y0), pinned(torch, prm)
        out_t = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
        out_y = torch.empty((n, 3), dtype=torch.float64).pin_memory().numpy()
        times, (tf, yf, _), phases = timed_e2e(ctx, ens2, y0_pin, prm_pin, (out_t, out_y), K)
        assert np.array_equal(yf, y_fin), "e2e result differs from the HBM-resident run"
        e2e_s, = ctx.reduce([float(np.sum(times))], "MAX")
        ph = ctx.reduce([phases["h2d_and_transpose_ms"], phases["kernel_ms"], phases["transpose_and_d2h_ms"]], "MAX")
        out["e2e"] = {"value": sum_steps * K / e2e_s, "unit": "RK steps/s",
                      "h2d_bytes_per_step": int(y0.nbytes + prm.nbytes), "d2h_bytes_per_step": int(tf.nbytes + yf.nbytes),
                      "ms_per_step": e2e_s / K * 1e3,
                      "phases_ms_max_over_ranks": {"h2d_and_transpose": ph[0], "kernel": ph[1], "transpose_and_d2h": ph[2],
                                                   "note": "separate passes with a synchronize between the phases"}}
        ens2.close()
    ens.close()
    return out


def dde_leg(ctx, d, total, K, W):
    """Constant-delay DDE ensemble of `total` members split over the ranks (BASELINE configs[2])."""
    import cases
    torch = ctx.torch
    lo, hi = shard_bounds(total, ctx.rank, ctx.world)
    n = hi - lo
    k = dde_k(total, lo, hi)
    prm = np.ascontiguousarray(cases.dde_params(k))
    y0 = np.zeros((n, 1))
    problem = (d.Solver.new().max_delay(1.0).initial(d.InitFn(d.SIN)).interval(0.0, 100.0)
               .stepsize(0.02).equation(d.systems.DdeLinear, prm).arith("fast").problem())
    ens = d.Ensemble(problem, n, device=ctx.local_rank)
    d_prm = torch.from_numpy(np.ascontiguousarray(prm.T)).cuda()
    d_y0 = torch.zeros(n, dtype=torch.float64, device="cuda")
    ens.set_initial_device(d_y0.data_ptr())
    ens.set_params_device(d_prm.data_ptr())
    step_ms, _ = timed_resident(ctx, ens, K, W)
    tot = ens.totals()
    t_fin, y_fin, _ = ens.final()
    err = float(np.max(np.abs(y_fin[:, 0] - np.sin(k * t_fin))))
    # e2e: host buffers through dfb_set_initial / dfb_set_params -> dfb_run -> dfb_get_final
    y0_pin, prm_pin = pinned(torch, y0), pinned(torch, prm)
    out_t = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    out_y = torch.empty((n, 1), dtype=torch.float64).pin_memory().numpy()
    times, (tf, yf, _), phases = timed_e2e(ctx, ens, y0_pin, prm_pin, (out_t, out_y), K)
    same = bool(np.array_equal(yf, y_fin))
    ens.close()
    return {"total_ms": float(sum(step_ms)), "e2e_s": float(np.sum(times)), "err": err, "steps": int(tot["steps"]),
            "n_failed": int(tot["n_failed"]) + (0 if same else 1), "h2d": int(y0.nbytes + prm.nbytes),
            "d2h": int(tf.nbytes + yf.nbytes), "phases": phases, "n_local": n}


def dde_report(local, ctx, total, K, scaling):
    """Combine the ranks' DDE legs (max of the times and the error, sums of the counts).  The collectives
    live here, outside the try block of the compute, so a rank whose leg failed still takes part."""
    world = ctx.world if ctx is not None else 1
    ok = local is not None and "error" not in local
    mx = [local["total_ms"], local["e2e_s"], local["err"], 0.0] if ok else [0.0, 0.0, 0.0, 1.0]
    sm = [float(local["steps"]), float(local["n_failed"])] if ok else [0.0, 0.0]
    if world > 1:
        mx = ctx.reduce(mx, "MAX")
        sm = ctx.reduce(sm, "SUM")
    if mx[3] != 0.0:
        return local if (local is not None and "error" in local) else {"error": "the DDE leg failed on another rank"}
    total_ms, e2e_s, err = mx[0], mx[1], mx[2]
    steps, n_failed = int(sm[0]), int(sm[1])
    rate = steps * K / (total_ms * 1e-3)
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm, src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        hbm, src = 6650.0, "fallback 6.65 TB/s (of fallback)"
    achieved = DDE_BYTES_PER_STEP * rate / 1e9 / world   # per GPU, against one GPU's copy bandwidth
    members_total = total if scaling == "strong" else total * world
    return {
        "workload": DDE_WORKLOAD, "members_total": members_total, "members_per_gpu": members_total // world,
        "metric": "rk_steps_per_sec", "value": rate, "unit": "RK steps/s", "ms_per_step": total_ms / K,
        "n_gpus": world, "scaling": scaling, "steps": K,
        "rk_steps_per_member": int(steps // members_total), "n_failed": n_failed,
        "max_abs_err_vs_exact_sin": err,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                     "frac": achieved / hbm,
                     "traffic": DDE_DRAM_TRAFFIC_PER_LAUNCH if (world == 1 and members_total == DDE_MEMBERS) else None,
                     "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one 2^18-member "
                                       "launch, profiles/r2_final_ncu_full_dde_fixed.txt (algorithmic bytes per launch: "
                                       f"{int(DDE_BYTES_PER_STEP * steps // world)})",
                     "peak_source": src, "per_gpu": True,
                     "algorithmic_bytes_per_rk_step": DDE_BYTES_PER_STEP,
                     "survey_bytes_per_rk_step": DDE_BYTES_SURVEY_PER_STEP,
                     "frac_at_survey_bytes": DDE_BYTES_SURVEY_PER_STEP * rate / 1e9 / world / hbm,
                     "kernel": "dde_fixed2_kernel / dde_fixed_kernel <SysDdeLinear,7,5,rktp64-sparsity> "
                               "(members per thread and block size chosen from the shard size)",
                     "kernel_ms": total_ms / K},
        "e2e": {"value": steps * K / e2e_s, "unit": "RK steps/s", "h2d_bytes_per_step": local["h2d"],
                "d2h_bytes_per_step": local["d2h"], "ms_per_step": e2e_s / K * 1e3,
                "phases_ms_rank0": local["phases"]},
    }


# SURVEY §8(d) step formula: sum_i N(2 nnz(a_i) + 1) + N(2 nnz(b) + 1) + (2(S-1) + 1) + S F_rhs, rktp64
def step_flops(N, f_rhs):
    return N * (2 * 21 + 6) + N * (2 * 6 + 1) + 13 + 7 * f_rhs


def other_config_legs(d, peak_tflops):
    """The remaining BASELINE.json configs (parity-test cases, not the headline): one warm-up, then two
    timed passes each; kernel time from the library's CUDA events (dfb_totals.kernel_ms)."""
    import cases
    KIND = {0: "integrate_general_kernel", 1: "ode_fixed_kernel", 2: "dde_fixed_kernel",
            3: "ode_adaptive_kernel", 4: "ode_fixed_periodic_kernel", 5: "ode_event_kernel",
            6: "dde_event_kernel"}

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
        r = {"kernel": KIND.get(tot["kernel_kind"], "?"), "kernel_ms": m, "rk_steps": tot["steps"],
             "rejected": tot["rejected"], "events": tot["events"], "n_failed": tot["n_failed"],
             "rk_steps_per_s": tot["steps"] / (m * 1e-3), "events_per_s": tot["events"] / (m * 1e-3)}
        if tot.get("locate_iters"):
            r["locate_iters"] = tot["locate_iters"]
        return r

    out = {}
    auto = d.AutomaticStepsize(1e-2, (1e-6, 1e-1), [1e-9] * 3, [1e-9] * 3, 4, 0.9, (0.2, 5.0))
    # (builder, flop per attempted step, flop per location-loop iteration [Horner interpolant N(I-1) FMA +
    #  detector + midpoint/width updates], N, detector evaluations per step)
    legs = {
        # configs[1], adaptive half: 307 flop per ATTEMPTED step (252 + error estimate + controller)
        "lorenz_2^20_adaptive_rktp64_tol1e-9": (lambda: cases.lorenz(n=1 << 20, t1=T_END, arith="fast").stepsize(auto), 307.0, 0.0),
        # configs[3]: §8d step formula with N = 2 (F_rhs = 1) / N = 4 (F_rhs = 0); a bisection iteration is
        # 4 FMA (Horner, I = 5) per component the detector reads + detector (1 / 6 flop) + 3 (midpoint, width):
        # the ball's detectors read one component each (the kernel folds the other one away since the locators
        # are compile-time constants), the egg's boundary function reads x and y of its N = 4
        "bouncing_ball_2^19_t=[0,8]": (lambda: cases.bouncing_ball(n=1 << 19, t1=8.0, arith="fast", events=0), step_flops(2, 1), 1 * 8 + 1 + 3.0),
        "egg_billiard_2^19_200_reflections": (lambda: cases.egg_billiard(n=1 << 19, reflections=200, arith="fast", events=0), step_flops(4, 0), 2 * 8 + 6 + 3.0),
        # configs[4]: Lorenz + 3x3 variational block (N = 12), QR every 0.5; flop/step by the same
        # formula as the headline: 12(2*21+6) + 12*13 + 13 + 7*F_rhs.  F_rhs = 8 + 33 + 1 = 42: the Jacobian
        # product WITHOUT its structural zero / unit entries (11 flop per column) — what the FAST kernel
        # executes since round 2; the reference's dense 3x3 gemm (15 flop per column, F_rhs = 54) would read
        # 1123 flop/step and flatter the fraction by 8 %
        "lyapunov_lorenz_2^18_rho_grid_t=[0,100]": (lambda: cases.lyap_lorenz(100.0, rho=np.linspace(20.0, 40.0, 1 << 18), arith="fast"), 1039.0, 0.0),
        # configs[4] on a delay-equation grid: x' = a x(t-1) + b x'(t-1), every 5th step sampled for the
        # segment-norm exponent estimator (diffurch_b200/lyapunov.py); the kernel time is what is reported
        "lyapunov_ndde_2^16_ab_grid_t=[0,100]": (lambda: d.lyapunov.linear_delay_grid(
            d.systems.NddeLinear, *[x.ravel() for x in np.meshgrid(np.linspace(-2.0, 0.5, 256), np.linspace(-0.6, 0.6, 256))])[0], None, 0.0),
        # DDE + located events (examples/dde-relay-clamp.rs), sigma sweep
        "dde_relay_clamp_2^16_alpha_sweep_t=[0,30]": (lambda: cases.dde_relay_clamp_ensemble(1 << 16, arith="fast"), None, 0.0),
        "dde_relay_clamp_2^18_alpha_sweep_t=[0,30]": (lambda: cases.dde_relay_clamp_ensemble(1 << 18, arith="fast"), None, 0.0),
    }
    # dde_event_kernel: one coefficient record (t, y[N], d_1..d_4[N]) = 8 (1 + N I) B written and one read per
    # committed step (N = 2, I = 5: 88 B each)
    hbm_bytes_per_step = {"dde_relay_clamp_2^16_alpha_sweep_t=[0,30]": 176.0, "dde_relay_clamp_2^18_alpha_sweep_t=[0,30]": 176.0}
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm_peak = 6650.0
    for name, (mk, flops, iter_flops) in legs.items():
        try:
            r = timed(mk)
            if flops:
                att = r["rk_steps"] + r["rejected"]
                r["algorithmic_flops_per_attempted_step"] = flops
                total_flops = flops * att
                if iter_flops and r.get("locate_iters"):
                    r["mean_bisection_iters_per_event"] = r["locate_iters"] / max(r["events"], 1)
                    r["algorithmic_flops_per_bisection_iter"] = iter_flops
                    total_flops += iter_flops * r["locate_iters"]
                r["achieved_tflops"] = total_flops / (r["kernel_ms"] * 1e-3) / 1e12
                r["frac_of_fp64_peak"] = r["achieved_tflops"] / peak_tflops
            if name in hbm_bytes_per_step:
                r["algorithmic_hbm_bytes_per_step"] = hbm_bytes_per_step[name]
                r["achieved_gbs"] = hbm_bytes_per_step[name] * r["rk_steps"] / (r["kernel_ms"] * 1e-3) / 1e9
                r["frac_of_hbm_peak"] = r["achieved_gbs"] / hbm_peak
                if r.get("locate_iters"):
                    r["mean_bisection_iters_per_event"] = r["locate_iters"] / max(r["events"], 1)
            out[name] = r
        except Exception as e:  # the headline line must still print
            out[name] = {"error": repr(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--members", type=int, default=N_MEMBERS, help=argparse.SUPPRESS)      # total (strong) / per GPU (weak)
    ap.add_argument("--dde-members", type=int, default=DDE_MEMBERS, help=argparse.SUPPRESS)
    ap.add_argument("--arith", default="fast", choices=["fast", "strict"], help=argparse.SUPPRESS)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return run_reference(args, rank, world)

    # stdout carries exactly ONE line (the JSON): whatever libraries print there (NCCL's version banner
    # under NCCL_DEBUG=VERSION, for one) goes to stderr instead
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        json_out.write(json.dumps({"error": "no CUDA device: diffurch_b200 has no CPU path"}) + "\n")
        json_out.flush()
        return 1
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import diffurch_b200 as d

    ctx = Ctx(torch, dist, rank, world, local_rank)
    W = max(args.warmup, 3)
    K = args.steps
    scaling = args.scaling
    total = args.members if scaling == "strong" else args.members * world
    dde_total = args.dde_members if scaling == "strong" else args.dde_members * world

    # ---- headline: Lorenz ensemble ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L = lorenz_leg(ctx, d, total, K, W, args.arith)
    clocks = sampler.stop() if rank == 0 else None

    # ---- secondary workload (configs[2]); every rank takes part, rank 0 reports ----
    secondary = None
    if not os.environ.get("DFB_BENCH_SKIP_DDE"):
        try:
            local_dde = dde_leg(ctx, d, dde_total, max(3, min(K, 5)), 2)
        except Exception as e:  # the headline line must still print
            local_dde = {"error": repr(e)}
        secondary = dde_report(local_dde, ctx, args.dde_members, max(3, min(K, 5)), scaling)

    # ---- the other scaling mode, short (multi-GPU runs only) ----
    other_scaling = None
    if world > 1 and not os.environ.get("DFB_BENCH_SKIP_WEAK"):
        alt = "weak" if scaling == "strong" else "strong"
        alt_total = args.members * world if alt == "weak" else args.members
        A = lorenz_leg(ctx, d, alt_total, 3, 3, args.arith, want_e2e=False)
        other_scaling = {"scaling": alt, "members_total": alt_total, "members_per_gpu": alt_total // world,
                         "value": A["value"], "unit": "RK steps/s", "ms_per_step": A["max_ms"] / 3, "steps": 3}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel of the headline ----
    import ctypes as C
    tf64, mhz = C.c_double(), C.c_double()
    d.lib.dfb_measure_fp64_peak(local_rank, C.byref(tf64), C.byref(mhz))
    kern_avg_ms = L["kernel_ms_local"]  # one launch per step; events bracket exactly that launch (this rank)
    achieved_tflops = FLOPS_PER_STEP * L["n_local"] * STEPS_PER_MEMBER / (kern_avg_ms * 1e-3) / 1e12
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
        "frac": achieved_tflops / peak, "per_gpu": True,
        "traffic": ODE_DRAM_TRAFFIC_PER_LAUNCH if L["n_local"] == N_MEMBERS else None,
        "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one 2^20-member launch, "
                          "profiles/r1_run3_ncu_full_ode_fixed_lorenz.txt (algorithmic HBM bytes per launch: "
                          f"{100 * L['n_local']}; the kernel is FP64-pipe-bound, not HBM-bound)",
        "peak_source": "max(DFMA-chain microbenchmark measured in this run, SMs x 64 DFMA/clk x 2 x SM clock "
                       "sampled during the timed region); MEASURED_PEAKS.json has no FP64 entry",
        "peak_measured_dfma_chain": tf64.value,
        "peak_architectural_at_observed_clock": arch_peak,
        "peak_sm_mhz": mhz.value,
        "algorithmic_flops_per_rk_step": FLOPS_PER_STEP,
        "kernel": "ode_fixed_kernel<SysLorenz,7,rktp64-sparsity,FAST> (members per thread and block size chosen "
                  "from the shard size, csrc/host/launch_geometry.hpp)",
        "kernel_ms": kern_avg_ms,
        "hbm_bytes_per_member": 8 * (3 + 3) + 8 * (1 + 3) + 8 + 8 + 4,
    }

    # ---- CPU baselines (oracle "port"), bounded samples on this box's host cores ----
    cores = os.cpu_count() or 1
    cpu_s = float(os.environ.get("DFB_BENCH_CPU_SECONDS", "12"))
    cpu_baseline = sized_cpu_baseline(cpu_oracle_rate, STEPS_PER_MEMBER, cores, cpu_s, "the Lorenz rho sweep")
    if secondary and "error" not in secondary:
        try:
            cpu_baseline["dde"] = sized_cpu_baseline(cpu_oracle_rate_dde, DDE_STEPS_PER_MEMBER, cores, cpu_s * 0.75,
                                                     "the DDE k sweep")
        except Exception as e:
            cpu_baseline["dde"] = {"error": repr(e)}

    others = None
    if world == 1 and not os.environ.get("DFB_BENCH_SKIP_OTHERS"):
        try:
            others = other_config_legs(d, peak)
        except Exception as e:
            others = {"error": repr(e)}

    e2e = L["e2e"]
    if secondary and "error" not in secondary:
        roofline["dde"] = secondary.pop("roofline")
        e2e["dde"] = secondary.pop("e2e")
    config = make_config(scaling, world, args.arith)
    if args.members != N_MEMBERS:
        config["members_total"] = total
        config["workload"] = f"lorenz_{total}_members_rho_sweep_rktp64_h=1/250_t=[0,50] (reduced: --members)"
    line = {
        "metric": "rk_steps_per_sec", "value": L["value"], "unit": "RK steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": L["max_ms"] / K,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config,
        "sharding": {"members_per_gpu": L["n_local"], "by": f"contiguous member range x{world}",
                     "data_path_collective": None,
                     "l2_flush": "256 MiB memset before every timed step (outside the event pair)",
                     "ensemble_stat_mean_final_z_allreduced": L["mean_final_z"]},
        "e2e": e2e,
        "gpu_launches": K,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "wall_s_timed_region": L["wall_s"],
        "secondary_dde": secondary,
        "other_scaling": other_scaling,
        "other_configs": others,
    }
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
