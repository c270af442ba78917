"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/sla_b200.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sla_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sla_[a-z0-9_A-Z]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from sparse_linear_algebra_b200 import _lib

    return _lib


def test_header_symbols_exported(lib):
    L = C.CDLL(lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/sla_b200.h but not exported by libsla_b200.so"


def test_bindings_cover_header(lib):
    assert sorted(lib.SIGNATURES) == _declared_symbols()


def test_no_oracle_in_product():
    """The product path must never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "sparse_linear_algebra_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"


def test_fails_loudly_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import sparse_linear_algebra_b200 as sla

    with pytest.raises(sla.SlaError) as e:
        sla.Context(0)
    assert e.value.status == lib.SLA_ERR_CUDA
    assert "no CPU path" in str(e.value)


def test_solve_opts_defaults(lib):
    L = lib.load()
    o = lib.SolveOpts()
    L.sla_solve_opts_default(C.byref(o))
    # nits = 200, tolAbs = 1e-6, tolRel = 1e-4   (Sparse.hs:1034-1036)
    assert (o.max_iters, o.tol_abs, o.tol_rel, o.true_residual, o.check_every) == (200, 1e-6, 1e-4, 1, 1)


def _build_c_smoke(tmp_path):
    import subprocess

    exe = os.path.join(str(tmp_path), "abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                           "-L", os.path.join(ROOT, "sparse_linear_algebra_b200"), "-lsla_b200", "-lm",
                           "-Wl,-rpath," + os.path.join(ROOT, "sparse_linear_algebra_b200"), "-o", exe])
    return exe


def test_c_program_links_against_the_abi(lib, tmp_path):
    """A plain C99 program compiles against include/sla_b200.h and links to the .so; without a GPU it exits 77."""
    import subprocess

    exe = _build_c_smoke(tmp_path)
    rc = subprocess.run([exe], capture_output=True, text=True).returncode
    assert rc in (0, 77)


@pytest.mark.gpu
def test_c_program_runs_on_gpu(lib, tmp_path):
    import subprocess

    exe = _build_c_smoke(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "abi_smoke ok" in p.stdout


def _build_cpp_mirror(tmp_path):
    import subprocess

    exe = os.path.join(str(tmp_path), "mirror_smoke")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "mirror_smoke.cpp"),
                           "-L", os.path.join(ROOT, "sparse_linear_algebra_b200"), "-lsla_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "sparse_linear_algebra_b200"), "-o", exe])
    return exe


def test_cpp_mirror_compiles(lib, tmp_path):
    """The header-only C++ host mirror (include/sla_b200.hpp) compiles and links; without a GPU it exits 77."""
    import subprocess

    rc = subprocess.run([_build_cpp_mirror(tmp_path)], capture_output=True, text=True).returncode
    assert rc in (0, 77)


@pytest.mark.gpu
def test_cpp_mirror_runs_reference_specs_on_gpu(lib, tmp_path):
    import subprocess

    p = subprocess.run([_build_cpp_mirror(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "mirror_smoke ok" in p.stdout


def test_python_wrappers_check_buffer_lengths():
    """the wrappers pass raw pointers, so inconsistent lengths must be refused before the library reads past a buffer
    (checked before any context is needed: runs without a GPU)."""
    import pytest

    import sparse_linear_algebra_b200 as sla

    with pytest.raises(ValueError):
        sla.SpMatrix.fromCOO((3, 3), [0, 1], [0, 1, 2], [1.0, 2.0])
    with pytest.raises(ValueError):
        sla.SpMatrix.fromCSR(3, 3, [0, 1, 2], [0, 1], [1.0, 2.0])            # row_ptr too short
    with pytest.raises(ValueError):
        sla.SpMatrix.fromCSR(2, 3, [0, 1, 2], [0, 1], [1.0])                 # col / val differ
    with pytest.raises(ValueError):
        sla.SpMatrix.mkDiagonal(4, [1.0, 2.0])
