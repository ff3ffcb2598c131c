#!/usr/bin/env python3
"""Strong-scaling shards on ONE GPU.  The ensemble shards by member range with no data-path collective,
so what one of N GPUs does in a strong-scaling run of the BASELINE ensembles is exactly a (total / N)-member
launch: this tool times those launches (kernel time from the library's CUDA events) for N = 1, 2, 4, 8,
for the geometry the launcher picks and for every forced (members per thread, block) alternative, and
prints the predicted N-GPU speed-up.  Used to tune csrc/host/launch_geometry.hpp; output goes to profiles/.

    python tools/bench_shards.py [--sweep]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
import cases  # noqa: E402
import diffurch_b200 as d  # noqa: E402


def timed(problem_fn, reps=3):
    ens = problem_fn()
    ens.run()
    ens.synchronize()
    ms = []
    for _ in range(reps):
        ens.run()
        ens.synchronize()
        ms.append(ens.totals()["kernel_ms"])
    tot = ens.totals()
    ens.close()
    return float(np.min(ms)), float(np.mean(ms)), tot


def lorenz_shard(total, world):
    lo, hi = bench.shard_bounds(total, 0, world)
    y0, prm = bench.lorenz_inputs(total, lo, hi)
    return (d.Solver.new().equation(d.systems.Lorenz, prm).initial(y0).interval(0.0, bench.T_END)
            .stepsize(bench.H).arith("fast").build())


def dde_shard(total, world):
    lo, hi = bench.shard_bounds(total, 0, world)
    k = bench.dde_k(total, lo, hi)
    return (d.Solver.new().max_delay(1.0).initial(d.InitFn(d.SIN)).interval(0.0, 100.0).stepsize(0.02)
            .equation(d.systems.DdeLinear, cases.dde_params(k)).arith("fast").build())


def with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true", help="also force every (tpt, block) alternative")
    args = ap.parse_args()
    out = {"lorenz": {}, "dde": {}}
    base = {}
    for name, total, mk, tpt_var, blk_var in (("lorenz", bench.N_MEMBERS, lorenz_shard, "DFB_ODE_TPT", "DFB_ODE_BLOCK"),
                                              ("dde", bench.DDE_MEMBERS, dde_shard, "DFB_DDE_TPT", "DFB_DDE_BLOCK")):
        for world in (1, 2, 4, 8):
            mn, mean, tot = timed(lambda: mk(total, world))
            row = {"members": total // world, "auto_ms_min": mn, "auto_ms_mean": mean,
                   "steps_per_s": tot["steps"] / (mn * 1e-3)}
            if world == 1:
                base[name] = mn
            row["predicted_speedup_vs_1gpu"] = base[name] / mn
            row["predicted_efficiency"] = base[name] / mn / world
            if args.sweep:
                alt = {}
                for tpt in (1, 2):
                    for blk in (32, 64, 128):
                        m2, _, _ = with_env({tpt_var: tpt, blk_var: blk}, lambda: timed(lambda: mk(total, world), reps=2))
                        alt[f"tpt{tpt}_block{blk}"] = m2
                row["forced_ms_min"] = alt
            out[name][f"n_gpus={world}"] = row
            print(name, world, json.dumps(row), file=sys.stderr, flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
