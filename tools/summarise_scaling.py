#!/usr/bin/env python3
"""Strong-scaling summary from bench.py lines measured at N = 1, 2, 4, 8 (one JSON line per file):

    python tools/summarise_scaling.py n1.json n2.json n4.json n8.json > profiles/..._strong_scaling_1_2_4_8.json

Speed-ups are relative to the N = 1 file given; nothing is measured here."""
import json
import sys


def line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def main(paths):
    rows = [line(p) for p in paths]
    rows.sort(key=lambda d: d["n_gpus"])
    base = rows[0]
    assert base["n_gpus"] == 1, "first file must be the one-GPU line"
    out = []
    for d in rows:
        dde = d.get("secondary_dde") or {}
        bdde = base.get("secondary_dde") or {}
        e2e, be2e = d["e2e"], base["e2e"]
        out.append({
            "n_gpus": d["n_gpus"],
            "scaling": d["scaling"],
            "members_per_gpu": d["config"].get("members_per_gpu"),
            "lorenz_steps_per_s": d["value"],
            "lorenz_speedup": d["value"] / base["value"],
            "lorenz_ms_per_pass": d["ms_per_step"],
            "lorenz_e2e_steps_per_s": e2e["value"],
            "lorenz_e2e_speedup": e2e["value"] / be2e["value"],
            "e2e_phases_ms": e2e.get("phases_ms_max_over_ranks"),
            "roofline_frac_per_gpu": d["roofline"]["frac"],
            "dde_steps_per_s": dde.get("value"),
            "dde_speedup": (dde["value"] / bdde["value"]) if dde.get("value") and bdde.get("value") else None,
            "dde_ms_per_pass": dde.get("ms_per_step"),
            "dde_hbm_frac_per_gpu": d["roofline"].get("dde", {}).get("frac"),
            "dde_e2e_steps_per_s": (e2e.get("dde") or {}).get("value"),
            "dde_e2e_speedup": ((e2e["dde"]["value"] / be2e["dde"]["value"])
                                if (e2e.get("dde") and be2e.get("dde")) else None),
            "other_scaling": d.get("other_scaling"),
            "clocks": d.get("clocks"),
        })
    json.dump(out, sys.stdout, indent=1)
    sys.stdout.write("\n")


if __name__ == "__main__":
    main(sys.argv[1:])
