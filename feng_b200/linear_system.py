"""Python mirror of the reference's feLinearSystem interface on top of the C ABI, plus the Newton loop that drives it.

`LinearSystemB200` keeps the method names, argument meaning and call order of class feLinearSystem
(src/feLinearSystem.h:43-169) so that tests read like the reference's own drivers; `solve_newton_raphson` restates
solveNewtonRaphson for stationary problems (src/feNonLinearSolver.cpp:38-182).  The production drop-in is the C++
adapter (adapter/feLinearSystemB200.h); both go through the same libfeng_b200.so entry points.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import capi
from .problems import HostProblem, form_layout


@dataclass
class NLSolverOptions:
    """feNLSolverOptions (src/feNonLinearSolver.h:12-40)"""
    tolResidual: float = 1e-10
    tolCorrection: float = 1e-10
    tolDivergence: float = 1e4
    maxIter: float = 20
    recomputeJacobianEveryNsteps: int = 3
    residualDecrease: float = 1e-1


class LinearSystemB200:
    def __init__(self, pb: HostProblem, device: int = 0, colors: np.ndarray | None = None,
                 constraint_rows: np.ndarray | None = None, device_pattern: bool = False, partition=None):
        self.pb = pb
        self.sys = capi.System(device)
        s = self.sys
        s.set_mesh(pb.dim, pb.mesh.xyz, pb.mesh.cells)
        s.set_quadrature(pb.w)
        self.su = s.add_space(pb.LU.shape[1], pb.ncomp, pb.adrU, pb.LU, pb.dLU)
        self.sp = -1
        if pb.adrP is not None:
            self.sp = s.add_space(pb.LP.shape[1], 1, pb.adrP, pb.LP, pb.dLP)
        if pb.chns is not None:
            # monolithic CHNS form on {U, P, Phi, Mu} (src/CHNS_Solver.cpp:236-420)
            self.sf = s.add_space(pb.LF.shape[1], 1, pb.adrF, pb.LF, pb.dLF)
            self.sm = s.add_space(pb.LF.shape[1], 1, pb.adrM, pb.LF, pb.dLF)
        if not device_pattern:
            if pb.ia is None:
                pb.build_pattern()
            s.set_pattern(pb.n_inc, pb.n_dof, pb.ia, pb.ja)
        if pb.chns is not None:
            s.add_form_chns(self.su, self.sp, self.sf, self.sm, pb.chns)
        for f in pb.forms:
            rows, cols = form_layout(f.kind)
            if rows == ("P",):                       # MIXED_DIVERGENCE is declared on {p, u}
                su, sp = self.sp, self.su
            else:
                su, sp = self.su, (self.sp if "P" in cols else -1)
            s.add_form(f.kind, su, sp, f.coeff, f.param, f.source)
        if device_pattern:
            s.build_pattern(pb.n_inc, pb.n_dof)      # feEZCompressedRowStorage rules on the device
        if colors is not None:
            s.set_colors(int(colors.max()) + 1, colors)
        if constraint_rows is not None and len(constraint_rows):
            s.set_constraints(constraint_rows)
        s.finalize()
        self.partition = partition
        if partition is not None and partition.world > 1:
            # one NCCL communicator per system, its unique id broadcast through torch.distributed
            import torch
            import torch.distributed as dist
            ident = torch.zeros(128, dtype=torch.uint8)
            if partition.rank == 0:
                ident = torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8).clone()
            dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
            ident = ident.to(dev)
            dist.broadcast(ident, 0)
            s.comm_init(bytes(ident.cpu().numpy().tobytes()), partition.rank, partition.world)
            s.set_halo(partition.owned, partition.neighbors, partition.send_ptr, partition.send_idx, partition.recv_ptr,
                       partition.recv_idx)
        # feLinearSystem defaults (src/feLinearSystem.h:60-69)
        self._recomputeMatrix = True
        self._rel_tol, self._abs_tol, self._div_tol, self._max_iter = 1e-8, 1e-14, 1e6, 10000
        self.restart = 30
        self.pc = capi.PC_AUTO
        self.last_info = None
        # device-resident state (row N3): after the first upload the device copy, which correctSolution keeps current, is
        # authoritative and assemble* no longer re-uploads the caller's vector
        self.device_resident = False
        self.uploads = 0
        self._device_current = False

    def sys_M(self) -> int:
        """rows of the fused element system (15 for 2-D P2/P1, 34 for 3-D)"""
        return self.pb.adrU.shape[1] + (0 if self.pb.adrP is None else self.pb.adrP.shape[1])

    # ---- setters of the base class -----------------------------------------------------------------
    def setAbsoluteTol(self, v): self._abs_tol = v
    def setRelativeTol(self, v): self._rel_tol = v
    def setDivergenceTol(self, v): self._div_tol = v
    def setMaxIter(self, v): self._max_iter = int(v)
    def getRecomputeStatus(self): return self._recomputeMatrix
    def setRecomputeStatus(self, flag): self._recomputeMatrix = bool(flag)

    # ---- the virtuals ----------------------------------------------------------------------------------
    def getSystemSize(self): return self.sys.n_inc
    def getRHSMaxNorm(self): return self.sys.rhs_max_norm()
    def getResidualMaxNorm(self): return self.sys.du_max_norm()

    def setToZero(self):
        self.sys.set_to_zero(3 if self._recomputeMatrix else 1)

    def setMatrixToZero(self): self.sys.set_to_zero(2)
    def setResidualToZero(self): self.sys.set_to_zero(1)

    def _upload(self, sol, sol_dot=None, c0=0.0, t=0.0):
        if self.device_resident and self._device_current and sol_dot is None:
            return
        self.sys.set_solution(sol, sol_dot, c0, t)
        self.uploads += 1
        self._device_current = True

    def assemble(self, sol, sol_dot=None, c0=0.0, t=0.0, assembleOnlyTransientMatrices=False):
        self._upload(sol, sol_dot, c0, t)
        self.sys.assemble(3 if self._recomputeMatrix else 1, assembleOnlyTransientMatrices)

    def assembleMatrices(self, sol, sol_dot=None, c0=0.0, t=0.0, assembleOnlyTransientMatrices=False, upload=True):
        if upload:
            self._upload(sol, sol_dot, c0, t)
        self.sys.assemble(2, assembleOnlyTransientMatrices)

    def assembleResiduals(self, sol, sol_dot=None, c0=0.0, t=0.0, upload=True):
        if upload:
            self._upload(sol, sol_dot, c0, t)
        self.sys.assemble(1, False)

    def constrainEssentialComponents(self, sol=None): self.sys.constrain()
    def applyPeriodicity(self): self.sys.apply_periodicity()
    def permute(self): pass

    def solve(self):
        """-> (success, normDx, normResidual, normAxb, nIter) as feLinearSystem::solve"""
        info = self.sys.solve(self._rel_tol, self._abs_tol, self._div_tol, self._max_iter, self.restart, self.pc,
                              raise_on_fail=False)
        self.last_info = info
        return bool(info.converged), info.norm_dx, info.norm_rhs, info.norm_axb, info.iterations

    def correctSolution(self, sol: np.ndarray, correctSolutionDot=False):
        self.sys.correct_solution(sol, correctSolutionDot)

    def writeMatrix(self): return self.sys.get_matrix_values()
    def writeRHS(self): return self.sys.get_rhs()
    def writeResidual(self): return self.sys.get_du()


def solve_newton_raphson(system: LinearSystemB200, sol: np.ndarray, opt: NLSolverOptions = NLSolverOptions(),
                         verbose: bool = False):
    """Stationary Newton loop, restating src/feNonLinearSolver.cpp:38-182 step by step.  Returns (status, history)."""
    stop, it = False, 0
    normCorrection = normResidual = normAxb = 0.0
    history = []
    system.setRecomputeStatus(True)
    while not stop:
        system.setToZero()                                            # :77
        system.assembleResiduals(sol)                                 # :80
        normResidual = system.getRHSMaxNorm()                         # :81
        if it > 0 and normResidual <= opt.tolResidual:                # :82-87
            break
        if system.getRecomputeStatus():
            system.assembleMatrices(sol, upload=False)                # :91
        system.constrainEssentialComponents(sol)                      # :94
        system.applyPeriodicity()                                     # :95
        ok, normCorrection, normResidual, normAxb, nIter = system.solve()   # :98
        if not ok:
            return -1, history
        if normResidual > opt.tolDivergence:
            return -2, history
        system.correctSolution(sol)                                   # :123
        it += 1
        history.append(dict(iter=it, normAxb=normAxb, linearIter=nIter, normCorrection=normCorrection,
                            normResidual=normResidual))
        if verbose:
            print(f"It. {it:2d} : ||J*du - NL|| = {normAxb:10.10e} ({nIter:4d} iter.)  ||du|| = {normCorrection:10.10e}"
                  f"  ||NL(u)|| = {normResidual:10.10e}")
        system.setRecomputeStatus(True)                               # stationary: always recompute (:19-20)
        stop = normResidual <= opt.tolResidual or normCorrection <= opt.tolCorrection or it > opt.maxIter
    if normResidual <= opt.tolResidual or normCorrection <= opt.tolCorrection:
        return 0, history
    return -3, history
