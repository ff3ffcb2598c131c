"""bindings/rust/diffurch-b200-sys cannot be compiled here (no cargo/rustc): keep its `#[repr(C)]`
structs and `extern "C"` list in step with the ctypes mirror (diffurch_b200/_abi.py), which IS
exercised end to end by the GPU tests, and with the symbols the library exports."""
import os
import re

import diffurch_b200 as d
from diffurch_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "bindings", "rust", "diffurch-b200-sys", "src", "lib.rs")).read()


def rust_fields(name):
    body = re.search(r"pub struct %s \{(.*?)\n\}" % name, SRC, re.S).group(1)
    return [m.group(1) for m in re.finditer(r"pub (\w+):", body)]


def test_struct_fields_match_the_ctypes_mirror():
    pairs = {"dfb_tableau": abi.Tableau, "dfb_locator": abi.Locator, "dfb_auto_stepsize": abi.AutoStepsize,
             "dfb_problem": abi.Problem, "dfb_system_info": abi.SystemInfo, "dfb_stats": abi.Stats,
             "dfb_totals": abi.Totals}
    for rust_name, ctype in pairs.items():
        assert rust_fields(rust_name) == [f[0] for f in ctype._fields_], rust_name


def test_extern_block_lists_every_exported_symbol():
    declared = set(re.findall(r"pub fn (dfb_\w+)\(", SRC))
    assert declared == set(abi.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(d.lib, sym)
    assert int(re.search(r"DFB_ABI_VERSION: i32 = (\d+)", SRC).group(1)) == abi.ABI_VERSION
