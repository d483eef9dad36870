"""CPU: the C-ABI shared library loads, exports every symbol include/nngp_b200.h declares, and refuses to
compute without a CUDA device (no CPU fallback).  No compute calls are made here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "nngp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nngp_[a-z0-9_]+)\s*\(", txt)) - {"nngp_ctx"})


def test_every_declared_symbol_is_exported():
    from spnngp_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/nngp_b200.h but not exported"
    assert sorted(_lib.ABI_SYMBOLS) == syms
    assert lib.nngp_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    from spnngp_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the no-device path cannot be exercised")
    lib = _lib.load()
    h = C.c_void_p()
    st = lib.nngp_create(-1, None, C.byref(h))
    assert st == -2 and not h.value  # NNGP_ENODEV
    assert b"no CPU fallback" in lib.nngp_last_error(None)
    from spnngp_b200 import Engine, NNGPError
    with pytest.raises(NNGPError):
        Engine()


def test_r_glue_exports_the_reference_call_symbols():
    path = os.path.join(ROOT, "spnngp_b200", "libspNNGP_b200.so")
    assert os.path.exists(path), "run __graft_entry__.build()"
    glue = C.CDLL(path)
    for s in ("rNNGP", "cNNGP", "rNNGPPredict", "mkNNIndx", "mkNNIndxCB",  # accelerated (src/spNNGP.h, src/nn.h)
              "sNNGP", "sNNGPLogit", "cSLGP", "sNNGPPredict", "PGLogit", "rNNGPReplicated", "mkUIndx",  # stubs
              "cNNGPKFold",  # added: the k-fold loop of R/spConjNNGP.R in one call
              "R_init_spNNGP"):
        assert hasattr(glue, s)


def test_product_code_never_touches_the_oracle():
    """Nothing under spnngp_b200/ may import, link or execute oracle/ (task rule)."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "spnngp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"oracle[/.]|liboracle|pyoracle|_ref/", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_tools_do_not_use_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/ (task rule): the developer probes
    under tools/ get their neighbour lists from the engine's own search."""
    bad = []
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith((".py", ".sh")):
            txt = open(os.path.join(ROOT, "tools", f), errors="ignore").read()
            if re.search(r"from oracle|import oracle|pyoracle|oracle/", txt):
                bad.append(f)
    assert not bad, bad


def test_bench_reference_arm_prints_one_json_line(tmp_path):
    """bench.py --impl reference (the reference's own CPU path through oracle/_ref, no GPU): stdout carries exactly one
    JSON line with the contract's keys even when a library would print a banner there (NCCL_DEBUG=VERSION), and without a
    CUDA device the GPU arm refuses to run instead of falling back."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "libspnngp_ref.so")):
        pytest.skip("oracle/_ref not built")
    env = dict(os.environ, NCCL_DEBUG="VERSION")
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-budget-s", "5"], capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert p.returncode == 0, p.stderr[-500:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rNNGP_mcmc_iter_per_s" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["workload"].endswith("(BASELINE.json configs[2])")
    import torch
    if not torch.cuda.is_available():
        q = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                           timeout=600, cwd=root)
        assert q.returncode != 0 and "no CPU fallback" in (q.stderr + q.stdout)


def test_every_python_tool_parses():
    """tools/*.py, bench.py and the package are never imported by the CPU suite as a whole (they need a GPU): at least
    every file must be valid Python."""
    import ast
    import glob
    files = glob.glob(os.path.join(ROOT, "tools", "*.py")) + glob.glob(os.path.join(ROOT, "spnngp_b200", "*.py"))
    files += [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    assert len(files) > 20
    for f in files:
        ast.parse(open(f).read(), filename=f)
