"""One host process driving several GPUs (csrc/multi.cu, sla_init_multi): the global objects must give the single-GPU results —
(#>) bit for bit (NCCL exchange keeps the ascending fold), dots / solvers within the reduction-order tolerance.  Needs >= 2 GPUs."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def multi():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from sparse_linear_algebra_b200 import multi as mm

    ctx = mm.MultiContext(2)
    yield mm, ctx
    ctx.close()


@pytest.mark.parametrize("kind,n,k,band", [("uniform", 20000, 16, 0), ("laplace", 96 * 96, 5, 96), ("ragged", 1001, 8, 0)])
def test_multi_matches_single_gpu(multi, ora, kind, n, k, band):
    import sparse_linear_algebra_b200 as sla

    mm, m = multi
    gk = sla.GEN_LAPLACE2D if kind == "laplace" else sla.GEN_UNIFORM
    ok = ora.GEN_LAPLACE2D if kind == "laplace" else ora.GEN_UNIFORM
    seed = 0x5EED0041
    A = m.generate(gk, n, k, seed, band)
    x = m.generate_vector(n, seed + 1)
    Ao = ora.SpMatrix.synth(ok, n, k, seed, band)
    xo = ora.SpVector.synth(seed + 1, n)
    assert A.dim == (n, n) and A.nnz == Ao.nnz
    y = (A @ x).toDenseListSV()
    yo = Ao.matVec(xo).toDenseListSV()
    assert y.tobytes() == yo.tobytes()                                       # (#>) bit-exact across the two GPUs
    d, do = x.dot(A @ x), xo.dot(Ao.matVec(xo))
    assert abs(d - do) <= 1e-12 * abs(do) + 1e-300
    assert abs(x.norm2() - xo.norm2()) <= 1e-13 * xo.norm2()
    # the same matrix from a global host CSR
    rp, ci, va = Ao.toCSR()
    A2 = m.fromCSR(n, n, rp, ci, va)
    assert (A2 @ x).toDenseListSV().tobytes() == yo.tobytes()
    # BiCGSTAB trajectory, pure steps keep their argument
    b, bo = A @ x, Ao.matVec(xo)
    st0 = mm.bicgsInit(A, b, m.zeros(n))
    rhat = m.vector(st0.field(1, n))
    sto = ora.bicgsInit(Ao, bo, ora.SpVector.mkSpVR(n, np.zeros(n)))
    rhato = bo - Ao.matVec(ora.SpVector.mkSpVR(n, np.zeros(n)))
    x_before = st0.field(0, n)
    st = st0
    for it in range(4):
        st = mm.bicgstabStep(A, rhat, st, pure=True)
        sto = ora.bicgstabStep(Ao, rhato, sto)
        xr = sto.x.toDenseListSV()
        assert np.abs(st.field(0, n) - xr).max() <= 1e-10 * np.abs(xr).max()
    assert st0.field(0, n).tobytes() == x_before.tobytes()
    if kind != "laplace":
        xs, its, res = mm.linSolve0(sla.BICGSTAB_, A, b, m.vector(np.full(n, 0.1)), info=True)
        xso, ito, _ = ora.linSolve0(ora.BICGSTAB_, Ao, bo, ora.SpVector.mkSpVR(n, [0.1] * n), info=True)
        assert its == ito and np.abs(xs.toDenseListSV() - xso.toDenseListSV()).max() <= 1e-10
        xg, itg, resg = mm.gmres(A, b, m.zeros(n), restart=20, tol_abs=1e-10, tol_rel=1e-12, info=True)
        assert resg <= 1e-8
        Q, H, brk = mm.arnoldi(A, x, 6)
        Qo, Ho = ora.arnoldi(Ao, xo, 6)
        assert H.shape == Ho.shape and np.abs(H - Ho).max() <= 1e-9 * np.abs(Ho).max() and np.abs(Q - Qo).max() <= 1e-9


def test_multi_errors(multi):
    import sparse_linear_algebra_b200 as sla

    mm, m = multi
    A = m.generate(sla.GEN_UNIFORM, 1000, 4, 3)
    with pytest.raises(sla.MatVecSizeMismatchException):
        A @ m.zeros(999)
    assert m.p2p in (True, False) and m.launches > 0
    with pytest.raises(sla.SlaError):
        mm.MultiContext(64)


def test_c_program_drives_two_gpus(tmp_path):
    """tests/c/multi_smoke.c: a plain C program, one process, two GPUs, only include/sla_b200.h."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    exe = tmp_path / "multi_smoke"
    lib_dir = os.path.join(ROOT, "sparse_linear_algebra_b200")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "multi_smoke.c"),
                           "-L", lib_dir, "-lsla_b200", "-Wl,-rpath," + lib_dir, "-lm", "-o", str(exe)])
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "MULTI_SMOKE OK" in p.stdout
