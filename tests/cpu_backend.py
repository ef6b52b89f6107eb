"""TEST INFRASTRUCTURE: a CPU stand-in for the feLinearSystem backend (oracle assembly + scipy sparse LU) with the
method names of feng_b200.linear_system.LinearSystemB200, so the HOST Newton loop can be exercised without a GPU and
compared with the reference's solveNewtonRaphson (which the harness runs with an Eigen SparseLU stub backend)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import fe_oracle as O


class OracleLinearSystem:
    def __init__(self, opb, ia, ja, constraint_rows=()):
        self.pb, self.ia, self.ja = opb, ia, ja
        self.rows = np.asarray(constraint_rows, np.int64)
        self.vals = np.zeros(ja.shape[0])
        self.rhs = np.zeros(opb.n_inc)
        self.du = np.zeros(opb.n_inc)
        self._recompute = True
        self._sol = None

    def getRecomputeStatus(self): return self._recompute
    def setRecomputeStatus(self, f): self._recompute = bool(f)
    def getSystemSize(self): return self.pb.n_inc
    def getRHSMaxNorm(self): return float(np.abs(self.rhs).max())

    def setToZero(self):
        if self._recompute:
            self.vals[:] = 0.0
        self.rhs[:] = 0.0

    def assembleResiduals(self, sol, **kw):
        self._sol = sol
        _, r = O.assemble(self.pb, self.ia, self.ja, sol, matrix=False)
        self.rhs += r

    def assembleMatrices(self, sol, **kw):
        v, _ = O.assemble(self.pb, self.ia, self.ja, sol, residual=False)
        self.vals += v

    def constrainEssentialComponents(self, sol=None):
        self.vals, self.rhs = O.constrain(self.ia, self.ja, self.vals, self.rhs, self.rows)

    def applyPeriodicity(self): pass

    def solve(self):
        n = self.pb.n_inc
        A = sp.csr_matrix((self.vals, self.ja, self.ia), shape=(n, n))
        self.du = spla.spsolve(A.tocsc(), self.rhs)
        return True, float(np.abs(self.du).max()), float(np.abs(self.rhs).max()), \
            float(np.abs(A @ self.du - self.rhs).max()), 1

    def correctSolution(self, sol, correctSolutionDot=False):
        sol[:self.pb.n_inc] += self.du
