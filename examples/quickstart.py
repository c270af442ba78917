"""Quick start: the README session of ocramz/sparse-linear-algebra (README.md:97-241) on the B200 backend.
Run on a machine with a CUDA device after `python -c "import __graft_entry__ as g; g.build()"`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparse_linear_algebra_b200 as sla

amat = sla.SpMatrix.fromListSM((3, 3), [(0, 0, 2), (1, 0, 4), (1, 1, 3), (1, 2, 2), (2, 2, 5)])
b = sla.SpVector.fromListDenseSV(3, [3, 2, 5])

x = sla.linSolve0(sla.BICGSTAB_, amat, b, sla.SpVector.fromListSV(3, []))      # λ> x <- linSolve0 BICGSTAB_ amat b x0
print("x          =", x.toDenseListSV())                                          # 1.50, -2.00, 1.00
print("amat #> x  =", (amat @ x).toDenseListSV())                                 # 3.00, 2.00, 5.00

x0 = sla.SpVector.fromListSV(3, [])
st = sla.bicgsInit(amat, b, x0)                                                   # λ> let initState = bicgsInit amat b x0
r0hat = st.r.copy()                                                               # λ> let r0hat = b ^-^ (amat #> x0)
for _ in range(3):
    sla.bicgstabStep(amat, r0hat, st)                                             # λ> iterate (bicgstabStep amat r0hat) ...
print("3 steps    =", st.x.toDenseListSV(), " ||r|| =", st.r.norm2())

print("amat <\\> b =", sla.backslash(amat, b).toDenseListSV())                    # GMRES(30)
print("x <.> x    =", x.dot(x), "  transpose rows:", amat.transpose().toCSR()[0].tolist())
