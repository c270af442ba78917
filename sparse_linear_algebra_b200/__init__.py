"""sparse_linear_algebra_b200 — B200 (sm_100a) backend for the hot path of ocramz/sparse-linear-algebra.

Only what the path needs lives here: csrc/ (hand-written CUDA kernels + the C ABI of include/sla_b200.h,
built in-tree into libsla_b200.so) and sparse.py, the host-side mirror of the reference's operator surface.
Importing the package loads the shared library and fails loudly if it is missing; there is no CPU fallback.
"""
from . import _lib
from ._lib import LIB_PATH, SolveOpts

_lib.load()   # fail loudly when the CUDA extension is missing

from .sparse import (  # noqa: E402
    BCG_, BF16, BICGSTAB_, CGNE_, CGS_, F64, GMRES_, GEN_BANDED, GEN_BLOCK16, GEN_LAPLACE2D, GEN_UNIFORM, Context, DenseBlock, DenseMatrix, IterE,
    KrylovState, MatVecSizeMismatchException, NeedsPivoting, OutOfBoundsIndexError, SlaError, SpMatrix, SpVector, arnoldi,
    backslash, bicgsInit, bicgstabStep, cgneInit, cgneStep, cgsInit, cgsStep, default_context, diagPartitions, gmres,
    ilu0Pre, jacobiPre, linSolve0, linSolve0Host, mSsorPre, set_default_context, triLowerSolve, triUpperSolve,
)
