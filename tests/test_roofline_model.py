"""The byte counts bench.py and DESIGN.md quote are the ones SURVEY.md §8(d) derives (scripts/roofline_model.py)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_bytes_match_the_survey():
    rm = _load(os.path.join(ROOT, "scripts", "roofline_model.py"), "roofline_model")
    import bench

    n2, z2 = 10_000_000, 320_000_000
    g = 4096
    n3, z3 = g * g, 5 * g * g - 4 * g
    assert rm.b_spmv(n2, z2) == bench.spmv_bytes(n2, z2) == 4_040_000_004            # 4.040 GB
    assert z3 == 83_869_696 and rm.b_spmv(n3, z3) == 12 * z3 + 20 * n3 + 4
    assert rm.b_bicgstab(n3, z3) == 24 * z3 + 168 * n3                               # 4.831 GB per bicgstabStep
    assert abs(rm.b_bicgstab(n3, z3) / 1e9 - 4.831) < 1e-3
    assert abs(rm.b_arnoldi_cycle(4_000_000, 256_000_000, 30) / 1e9 - 128.2) < 0.1   # 30 B_spmv + 8400 n
    assert rm.b_arnoldi_cycle(4_000_000, 256_000_000, 30) == 30 * rm.b_spmv(4_000_000, 256_000_000) + 8400 * 4_000_000
    assert abs(rm.b_spmm(n2, z2, 128) / 1e9 - 7.08) < 0.01


def test_request_port_floor_of_cfg2():
    """>= 350 M requests per cfg-2 (#>) at one request per SM-cycle -> >= 1.20 ms -> <= ~51 % of the HBM peak."""
    rm = _load(os.path.join(ROOT, "scripts", "roofline_model.py"), "roofline_model")
    req = rm.spmv_requests(10_000_000, 320_000_000, panels=2)
    assert 350e6 <= req <= 356e6
    t_ms = req / (148 * 1.965e9) * 1e3
    assert 1.20 <= t_ms <= 1.23
    assert 0.50 <= rm.b_spmv(10_000_000, 320_000_000) / (t_ms * 1e-3) / 1e9 / 6534.8 <= 0.52


def test_gather_microbenchmark_compiles(tmp_path):
    """scripts/microbench/gather_paths.cu (the request-port probe of DESIGN.md §3.1) cross-compiles for sm_100a and its
    SASS uses the three paths it claims to measure."""
    import shutil
    import subprocess

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        import pytest

        pytest.skip("nvcc not available")
    exe = os.path.join(str(tmp_path), "gather_paths")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-o", exe,
                           os.path.join(ROOT, "scripts", "microbench", "gather_paths.cu")])
    cuobjdump = os.path.join(os.path.dirname(nvcc), "cuobjdump")
    if os.path.exists(cuobjdump):
        sass = subprocess.run([cuobjdump, "-sass", exe], capture_output=True, text=True).stdout
        for mnemonic in ("LDG.E.NA", "LDGSTS", "UBLKCP"):
            assert mnemonic in sass, mnemonic
