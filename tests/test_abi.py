"""CPU-only: the C-ABI shared library loads and exports every symbol include/mfb200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "mfb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mfb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import ctypes

    import monofonic_b200 as M

    path = os.path.join(ROOT, "monofonic_b200", "libmfb200.so")
    if not os.path.exists(path):
        import __graft_entry__ as ge

        ge.build()
    lib = M.load()
    syms = header_symbols()
    assert len(syms) > 50
    for s in syms:
        assert hasattr(lib, s), s
    # the ctypes table covers the whole header too
    from monofonic_b200 import _lib

    assert set(_lib.EXPORTED) == set(syms)
    assert lib.mfb_build_info() == b"cuda sm_100a"


def test_no_cpu_fallback():
    """without a GPU the product refuses to run (mfb_init fails loudly)"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import monofonic_b200 as M

    lib = M.load()
    rc = lib.mfb_init(0)
    assert rc != 0
    assert b"CUDA" in lib.mfb_last_error() or b"cuda" in lib.mfb_last_error() or b"device" in lib.mfb_last_error()


def test_product_never_imports_oracle_or_emulator():
    pkg = os.path.join(ROOT, "monofonic_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) and f != "mfb_platform.h":
                txt = open(os.path.join(dirpath, f)).read()
                assert "lpt_oracle" not in txt and "libmfb_emu" not in txt, f


def test_bench_touches_the_oracle_only_in_its_cpu_legs():
    """bench.py may execute oracle/ only as the cpu_baseline / --impl reference checker, never in the measured arm"""
    import ast

    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    users = set()
    for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        seg = ast.get_source_segment(src, fn)
        if "lpt_oracle" in seg or '"oracle"' in seg:
            users.add(fn.name)
    # cpu_reference_run: the oracle port; cpu_compiled_reference_run: oracle/_ref/ref_bench (the reference's own sources)
    assert users == {"cpu_reference_run", "cpu_compiled_reference_run"}, users
    # module level: no import of, and no sys.path entry for, oracle/ (the module docstring may talk about it)
    for node in tree.body:
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            names = [a.name for a in node.names] + [getattr(node, "module", None) or ""]
            assert not any("oracle" in x for x in names), names
        elif not isinstance(node, (ast.FunctionDef, ast.Expr)):
            assert "oracle" not in (ast.get_source_segment(src, node) or ""), ast.get_source_segment(src, node)
