"""User-system plugins used by the GPU tests.  They are compiled by `__graft_entry__.build()`
(nvcc cross-compiles on the CPU-only build box); the built .so files travel to the GPU box."""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
PLUGINS = {"van_der_pol": "SysVanDerPol", "user_lorenz": "SysUserLorenz", "user_dde": "SysUserDde",
           "switch_parameter": "SysSwitchParameter"}


def header(name):
    return os.path.join(HERE, name + ".cuh")


def build_all(verbose=False):
    from diffurch_b200 import user
    out = {}
    for name, struct in PLUGINS.items():
        out[name] = user.build_user_system(header(name), struct, verbose=verbose)
        if verbose:
            print("[plugins]", name, "->", out[name])
    return out
