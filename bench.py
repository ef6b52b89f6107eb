#!/usr/bin/env python
"""Benchmark of the hot path: P2/P1 Navier-Stokes Jacobian + residual assembly (Melem/s) and Newton-step pieces.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload t3d|t2d] [--n SIZE]

A "step" is one pass of the hot path over the whole mesh: setToZero + fused Jacobian+residual assembly of every
element (what solveNewtonRaphson does each iteration, src/feNonLinearSolver.cpp:77-91).  `value` = elements
assembled per second with every input resident in HBM; `e2e` = the same through the C ABI with HOST buffers (the
solution vector is copied host->device inside the timed region, the rhs max-norm is read back).  One process per GPU;
for N > 1 every rank owns a slab of a [0,1]^2 x [0,N] box (t3d) or a strip of a [0,1] x [0,N] rectangle (t2d): weak
scaling, owner-computes with one ghost layer of elements, no data-path collective in assembly.  Default workload: T3D(92)
per GPU = 4.67 M tetrahedra, 19.8 M DOF, the "~20 M-DOF tetrahedral cube" of BASELINE.json.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ns_p2p1_jacobian_residual_assembly"
UNIT = "Melem/s"
MU, RHO = 1.0 / 40.0, 1.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_problem(workload, n, rank, world, strong=False):
    from feng_b200 import mesh as M, partition as PT, problems as PB
    part = None
    if workload == "t2d":
        # strip r of the [0,1] x [0,world] domain: n x n owned cells plus one ghost row of cells towards each
        # neighbour (owner-computes); rows owned by one rank, halo plan for the SpMV input
        pb, part = PT.strip_problem(n, rank, world, "ns_div", 8, 1, MU, RHO, build_pattern=False, with_source=False)
        owned = 2 * n * n
        name = f"T2D({n}) P2/P1 Navier-Stokes (convU+divU+divSigma), Kovasznay Re=40 + noise, quad deg 8 (16 pts)"
    elif workload == "chns":
        # config 5: monolithic CHNS_Abels on [U (P2), P, Phi, Mu (P1)] with finite-difference Jacobian (22 residual
        # evaluations per element, src/feBilinearForm.cpp:388-428), BDF2-like state (solDot, c0) supplied by the host
        m, _ = PT.strip_mesh(n, rank, world)
        pb = PB.chns(m, chns_model(), 8, 1, MU, RHO, build_pattern=False)
        owned = m.n_cells if world == 1 else 2 * n * n
        name = (f"T2D({n}) transient CHNS_Abels P2/P1/P1/P1, residual + finite-difference Jacobian (M = 21 columns), "
                f"quad deg 8 (16 pts)")
    else:
        # slab r of the [0,1]^2 x [0,world] box: n^3 owned cells of 6 Kuhn tetrahedra plus one ghost layer of cells
        pb, part = PT.slab_problem(n, rank, world, "ns_div", 6, 3, MU, RHO, build_pattern=False, with_source=False, strong=strong)
        owned = 6 * n * n * int(np.diff(PT.layer_bounds(n, world, strong))[rank])
        name = (f"T3D({n}) P2/P1 Navier-Stokes tetrahedra (convU+divU+divSigma), trigonometric field + noise, "
                f"quad deg 6 (24 pts)" + (f", ONE cube cut into {world} slabs" if strong and world > 1 else ""))
    sol = PB.perturb_unknowns(pb)
    return pb, sol, owned, name, part


def algorithmic_bytes_per_element(pb, nnz):
    """index gather + vertex coordinates + local solution + one write of every CSR value and
    of every rhs entry (no source table, no transient term in this workload)."""
    n_loc = pb.adrU.shape[1] + pb.adrP.shape[1]
    transient = 0
    if pb.chns is not None:                # + Phi, Mu tables and the time-derivative vector of the transient form
        n_loc += pb.adrF.shape[1] + pb.adrM.shape[1]
        transient = 1
    nE = pb.mesh.n_cells
    return 4 * n_loc + 8 * pb.dim * (pb.dim + 1) + 8 * n_loc * (1 + transient) + 8 * nnz / nE + 8 * pb.n_inc / nE


def chns_model():
    from feng_b200 import problems as PB
    return PB.ChnsModel(rhoA=1.0, rhoB=0.8, viscA=0.02, viscB=0.05, mobility=1e-3, sigma=0.1, epsilon=0.05, force=(0.0, -0.5))


def chns_state(pb):
    """BDF-like transient state of the CHNS workload: perturbed unknowns, a time-derivative vector and c0 = 1.5 / dt."""
    from feng_b200 import problems as PB
    sol = PB.perturb_unknowns(pb)
    dot = np.random.default_rng(11).uniform(-1.0, 1.0, pb.n_dof)
    return sol, dot, 150.0


# FP64 work the shipped kernels EXECUTE per element, from ncu instruction counts (DFMA = 2 flop).  t3d: counted over every launch
# of one assembly pass at the bench size by scripts/gpu_r02e.sh and read from profiles/traffic.json ("t3d_flop_per_element",
# stamped with its source); the others are the round-1 captures (profiles/README.md: 2-D row-owner kernels r01b, CHNS kernel r01o)
# until they are re-captured.
EXECUTED_FLOP_PER_ELEMENT = {"t2d": 5.0e3, "t3d": 27.0e3, "chns": 113.0e3}


def fp64_roofline(workload, peak_tflops, n_elm, kernel_ms):
    """The second roofline SURVEY.md section 8(d) asks for: the assembly is not HBM-bound, so the executed FP64 rate is
    reported against the DFMA peak measured on this GPU in the same run."""
    flop, src = EXECUTED_FLOP_PER_ELEMENT[workload], "round-1 ncu capture (profiles/README.md)"
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        if tj.get(f"{workload}_flop_per_element"):
            flop, src = float(tj[f"{workload}_flop_per_element"]), tj.get(f"{workload}_source", "profiles/traffic.json")
    achieved = flop * n_elm / (kernel_ms * 1e-3) / 1e12
    return {"measured_dfma_peak_tflops": peak_tflops, "executed_flop_per_element": flop, "achieved_tflops": achieved,
            "frac": achieved / peak_tflops if peak_tflops else None, "flop_source": src,
            "note": "flop per element = executed FP64 instructions of the profiled launches (ncu), FMA counted as 2"}


def cpu_reference_baseline(n_cpu, reps, threads=None, chns=False):
    """The reference's own CPU assembly (oracle/_ref = unmodified feNG compiled here) on a bounded sample."""
    from feng_b200 import mesh as M, problems as PB
    from oracle import ref
    if not ref.available():
        return None
    import tempfile
    m = M.square_mesh(n_cpu)
    path = os.path.join(tempfile.gettempdir(), f"bench_t2d_{n_cpu}_{os.getpid()}.msh")
    M.write_msh(m, path)
    if threads:
        ref.set_threads(threads)
    if chns:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from test_chns import _oracle_problem
        pb = PB.chns(m, chns_model(), 8, 1, MU, RHO, build_pattern=False)
        P = ref.RefProblem(path, "chns", 2, 8, 1, MU, RHO, chns=_oracle_problem(pb).prm.as_array())
        os.remove(path)
        sol, dot, c0 = chns_state(pb)
        P.set_solution(sol, dot, c0, 0.0)
        P.assemble()
        times = []
        for _ in range(reps):
            _, _, sec = P.assemble()
            times.append(float(sec[0] + sec[1]))
        nE = m.n_cells
        P.close()
        return {"times": times, "n_elm": nE, "cores": ref.max_threads(),
                "sample": f"T2D({n_cpu}) = {nE} triangles, CHNS_Abels residual + finite-difference Jacobian through "
                          f"feBilinearForm::computeMatrixFiniteDifference, {reps} passes"}
    P = ref.RefProblem(path, "ns_div", 2, 8, field=1, mu=MU, rho=RHO, p_essential=False)
    os.remove(path)
    pb = PB.taylor_hood(m, "ns_div", 8, 1, MU, RHO, build_pattern=False, with_source=True)
    P.set_solution(PB.perturb_unknowns(pb))
    P.assemble()                                     # warm-up
    times = []
    for _ in range(reps):
        _, _, sec = P.assemble()
        times.append(float(sec[0] + sec[1]))
    nE = m.n_cells
    P.close()
    return {"times": times, "n_elm": nE, "cores": ref.max_threads(),
            "sample": f"T2D({n_cpu}) = {nE} triangles, same forms + the reference's zero vector-source form, "
                      f"Jacobian+residual, {reps} passes"}


def cpu_port_baseline(n_cpu, reps, threads=None):
    """3-D workload: the reference has no vector-valued space on tetrahedra (src/feSpace.cpp:762-767 instantiates the 2-D
    ones only), so its CPU arm is oracle/port_cpp.cpp: a C++/OpenMP restatement of the reference's own assembly organisation
    (one traversal per weak form, quadrature-point-major loops, colour loop + sorted scatter) generalised to dim = 3, pinned
    on the numpy oracle and, in 2-D, on the compiled reference (tests/test_port_cpp.py).  All host threads, bounded sample."""
    from feng_b200 import coloring, mesh as M, problems as PB
    from oracle import port
    if not port.available():
        return None
    if threads:
        port.set_threads(threads)
    m = M.cube_mesh(n_cpu)
    pb = PB.taylor_hood(m, "ns_div", 6, 3, MU, RHO, build_pattern=False, with_source=False)
    P = port.PortProblem(pb, coloring.color_elements(m.cells, m.n_vertices))
    sol = PB.perturb_unknowns(pb)
    P.assemble(sol)                                  # warm-up
    times = [P.assemble(sol)[2] for _ in range(reps)]
    return {"times": times, "n_elm": m.n_cells, "cores": port.max_threads(), "kind": "port", "port": "oracle/port_cpp.cpp (C++/OpenMP)",
            "sample": f"T3D({n_cpu}) = {m.n_cells} tetrahedra, same forms (convU+divU+divSigma), Jacobian+residual, colour loop + "
                      f"sorted scatter, C++/OpenMP restatement of the reference's assembly path (the reference has no 3-D vector "
                      f"spaces), {reps} passes"}


def cpu_reference_3d_scalar(n_cpu, reps):
    """Reference-code anchor in 3-D: the unmodified feSysElm_Diffusion<3> + Source on P2 tetrahedra (the only 3-D weak forms the
    reference has, src/feSysElm.cpp:586-589), same colour loop and scatter, on a T3D(n) cube written as a .msh file."""
    from feng_b200 import mesh as M
    from oracle import ref
    if not ref.available():
        return None
    import tempfile
    m = M.cube_mesh(n_cpu)
    path = os.path.join(tempfile.gettempdir(), f"bench_t3d_{n_cpu}_{os.getpid()}.msh")
    M.write_msh(m, path)
    P = ref.RefProblem(path, "diffusion", 2, 4, field=0, mu=1.0)
    os.remove(path)
    P.assemble()
    times = []
    for _ in range(reps):
        _, _, sec = P.assemble()
        times.append(float(sec[0] + sec[1]))
    nE = m.n_cells
    P.close()
    per = sum(times) / len(times)
    return {"value": nE / per / 1e6, "unit": UNIT, "cores": ref.max_threads(), "kind": "reference",
            "sample": f"T3D({n_cpu}) = {nE} tetrahedra, scalar P2 feSysElm_Diffusion<3> + Source (10 x 10 local system, quad deg 4), "
                      f"Jacobian+residual, {reps} passes"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    if args.workload == "t3d":
        base = cpu_port_baseline(args.cpu_n3, max(args.steps, 1) + args.warmup)
        kind, timing = "port", "oracle/port_cpp.cpp (C++/OpenMP, all host threads), host steady_clock around the assembly"
    else:
        base = cpu_reference_baseline(args.cpu_n_chns if args.workload == "chns" else args.cpu_n, max(args.steps, 1) + args.warmup,
                                      chns=args.workload == "chns")
        kind, timing = "reference", ("reference's own colour loop over computeMatrix/computeResidual + restated "
                                     "Pardiso-style scatter, OpenMP, host steady_clock")
    if base is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libfeng_ref.so not built"}))
        return
    t = base["times"][args.warmup:]
    per = sum(t) / len(t)
    val = base["n_elm"] / per / 1e6
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(t), "warmup": args.warmup,
            "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": base["sample"], "timing": timing},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": base["cores"], "kind": kind,
                             "sample": base["sample"]},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="t3d", choices=["t2d", "t3d", "chns"],
                    help="t3d: the ~20 M-DOF tetrahedral cube BASELINE.json quotes the metric on (default); t2d: refined triangles")
    ap.add_argument("--n", "--size", dest="n", type=int, default=0)
    ap.add_argument("--cpu-n", type=int, default=256, help="T2D size of the reference CPU sample")
    ap.add_argument("--cpu-n-chns", type=int, default=96, help="T2D size of the reference CPU sample of the CHNS workload")
    ap.add_argument("--cpu-n3", type=int, default=24, help="T3D size of the C++ port CPU sample")
    ap.add_argument("--solve-maxit", type=int, default=2000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="skip the Newton-step timing (assembly + GMRES)")
    ap.add_argument("--assembly", default="auto", choices=["auto", "scatter", "gather"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = one T3D(n) cube per GPU (default); strong = ONE T3D(n) cube cut into N slabs")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity_check block")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    n = args.n or {"t2d": 1024, "t3d": 92, "chns": 512}[args.workload]

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from feng_b200 import capi
    from feng_b200.linear_system import LinearSystemB200

    strong = args.scaling == "strong" and args.workload == "t3d"
    pb, sol, owned, wl_name, part = build_problem(args.workload, n, rank, world, strong)
    ls = LinearSystemB200(pb, device=local_rank, device_pattern=True, partition=part)
    S = ls.sys
    if args.assembly != "auto":
        S.set_assembly_mode({"scatter": capi.ASSEMBLY_SCATTER, "gather": capi.ASSEMBLY_GATHER}[args.assembly])
    gather = S.has_gather_plan() and args.assembly != "scatter"
    patch = gather and S.gather_plan_kind() == 2
    nE = pb.mesh.n_cells
    host_sol = torch.from_numpy(sol).pin_memory().numpy()        # pinned host buffer of the caller
    chns = pb.chns is not None
    if chns:
        _, dot, c0 = chns_state(pb)
        host_dot = torch.from_numpy(dot).pin_memory().numpy()
        upload = lambda: S.set_solution(host_sol, host_dot, c0, 0.0)
    else:
        upload = lambda: S.set_solution(host_sol)
    upload()

    def barrier():
        S.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------------
    # clocks / throttle reasons are sampled from the warm-up to the end of the end-to-end loop (nvidia-smi needs ~0.2 s to
    # deliver its first sample; the device-timed region alone is a fraction of a second)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        S.set_to_zero(3)
        S.assemble(3)
    barrier()
    capi.reset_kernel_launches()
    kern_ms = []
    barrier()
    S.time_begin()
    for _ in range(args.steps):
        S.set_to_zero(3)
        S.assemble(3)
    total_ms = S.time_end()
    launches = capi.kernel_launches()
    barrier()
    # per-launch duration of the dominant kernel (events recorded around the assembly launch on its own stream)
    for _ in range(args.steps):
        S.set_to_zero(3)
        S.assemble(3)
        kern_ms.append(S.last_assemble_ms())
    # ---- end to end through the C ABI with host buffers ---------------------------------------------------
    for _ in range(2):
        upload()
        S.set_to_zero(3)
        S.assemble(3)
        S.rhs_max_norm()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        upload()
        S.set_to_zero(3)
        S.assemble(3)
        rn = S.rhs_max_norm()
    S.sync()
    e2e_full_s = time.perf_counter() - t0
    # the same step with the state RESIDENT on the device: what a time step / Newton iteration needs from
    # the host is the essential-DOF values of the new level (feSolution::initializeEssentialBC, src/feTimeIntegration.cpp:523-536;
    # pinned host buffer -> b200_set_essential) -- the unknowns are already there, b200_correct_solution keeps them current
    ess_idx = np.arange(pb.n_inc, pb.n_dof, dtype=np.int64)
    ess_val = torch.from_numpy(np.ascontiguousarray(sol[pb.n_inc:])).pin_memory().numpy()
    resident = not chns                      # the CHNS workload supplies a host-side time derivative every step
    if resident:
        for _ in range(2):
            S.set_essential(ess_idx, ess_val)
            S.set_to_zero(3)
            S.assemble(3)
            S.rhs_max_norm()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            S.set_essential(ess_idx, ess_val)
            S.set_to_zero(3)
            S.assemble(3)
            rn = S.rhs_max_norm()
        S.sync()
        e2e_s = time.perf_counter() - t0
    else:
        e2e_s = e2e_full_s
    spmv_ms = S.time_spmv(20)
    clocks = sampler.stop()

    t_all = torch.tensor([total_ms, e2e_s * 1e3, e2e_full_s * 1e3], dtype=torch.float64, device="cuda")
    owned_all = torch.tensor([float(owned)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        dist.all_reduce(owned_all, op=dist.ReduceOp.SUM)
    total_ms, e2e_ms, e2e_full_ms = float(t_all[0]), float(t_all[1]), float(t_all[2])
    tot_owned = float(owned_all[0])

    extra = {}
    if not args.no_solve and not chns:
        extra["newton_step"] = newton_step(ls, sol, pb, args.solve_maxit)
    if not args.no_parity and args.workload == "t3d":
        pc_ = parity_check(rank, world, local_rank)
        if rank == 0:
            extra["parity_check"] = pc_

    if rank == 0:
        peak, peak_src = load_peaks()
        ms_step = total_ms / args.steps
        value = tot_owned / (ms_step * 1e-3) / 1e6
        bpe = algorithmic_bytes_per_element(pb, S.nnz)
        k_ms = statistics.mean(kern_ms)
        achieved = bpe * nE / (k_ms * 1e-3) / 1e9
        spmv_bytes = S.nnz * 12 + (S.n_inc + 1) * 8 + 2 * 8 * S.n_inc
        fp64 = capi.measure_fp64_peak(local_rank)
        dmma = capi.measure_dmma_peak(local_rank)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                per_elem = json.load(f).get(args.workload)      # measured DRAM bytes per element (ncu --set full)
            traffic = per_elem * nE if per_elem else None       # per assembly pass, like `achieved`
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong and world > 1 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name, "elements_per_gpu": int(owned), "ghost_elements_per_gpu": int(nE - owned),
                       "n_dof_per_gpu": int(pb.n_dof), "n_unknowns_per_gpu": int(pb.n_inc), "nnz_per_gpu": int(S.nnz),
                       "assembly": ("chns_kernel: one warp per element, lane j = residual of the state perturbed in local "
                                    "column j (finite-difference Jacobian), red.global.add.f64 into the CSR arrays after "
                                    "a memset") if chns else
                                   ("patch kernel: block-slot owners over Morton patches of elements, element state staged "
                                    "in shared memory, pre-contracted reference tensors; every CSR value written once, "
                                    "no memset, no atomics") if patch else
                                   ("row-owner gather on pre-contracted reference tensors: every CSR row written once, "
                                    "no memset, no atomics") if gather else
                                   "quadrature-loop kernel + atomic (red.global.add.f64) scatter into precomputed CSR slots",
                       "cache": "inputs larger than L2 (CSR values written per pass = %.1f GB)" % (S.nnz * 8 / 1e9),
                       "partition": ("strips / slabs, owner-computes with one ghost layer of elements (no collective in assembly); "
                                     "SpMV input halo over NCCL send/recv, Krylov dots all-reduced") if world > 1
                       else "single GPU"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "kernel": "chns_kernel (+ memsets of val and rhs)" if chns else
                                   ("patch_kernel (one launch per assembly pass) + patch_zero_kernel" if patch else
                                    "gather kernels (one assembly pass = several launches, timed together)"
                                    if gather else "th_kernel fused Jacobian+residual+scatter"),
                         "kernel_ms": k_ms, "algorithmic_bytes_per_element": bpe, "peak_source": peak_src,
                         "kernel_share_of_step": k_ms / ms_step,
                         "fp64": dict(fp64_roofline(args.workload, fp64, nE, k_ms), measured_dmma_peak_tflops=dmma,
                                      tensor_core_gate="mma.sync.m8n8k4.f64 peak measured in the same run: no throughput over the "
                                                       "FP64 CUDA-core path, so the contractions stay on DFMA")},
            "spmv": {"ms": spmv_ms, "achieved_gbs": spmv_bytes / (spmv_ms * 1e-3) / 1e9,
                     "frac": spmv_bytes / (spmv_ms * 1e-3) / 1e9 / peak, "bytes": spmv_bytes},
            "e2e": {"value": tot_owned / (e2e_ms / args.steps * 1e-3) / 1e6, "unit": UNIT,
                    "h2d_bytes_per_step": int((pb.n_dof - pb.n_inc) * 8) if resident else int(pb.n_dof * 8 * 2),
                    "d2h_bytes_per_step": 8, "ms_per_step": e2e_ms / args.steps, "rhs_max_norm": rn,
                    "state": ("device-resident: per step the host sends the essential-DOF values of the new level from a pinned "
                              "buffer (b200_set_essential; index list uploaded once) and reads the rhs max-norm back") if resident
                             else "host-resident: solution and time derivative uploaded every step",
                    "full_upload": {"value": tot_owned / (e2e_full_ms / args.steps * 1e-3) / 1e6,
                                    "h2d_bytes_per_step": int(pb.n_dof * 8 * (2 if chns else 1)),
                                    "ms_per_step": e2e_full_ms / args.steps,
                                    "note": "the unmodified host loop: whole state vector uploaded every step (b200_set_solution)"}},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        line.update(extra)
        if world == 1 and not args.no_cpu:
            ref2d = cpu_reference_baseline(args.cpu_n, 3) if not chns else cpu_reference_baseline(args.cpu_n_chns, 3, chns=True)
            base = ref2d if args.workload != "t3d" else cpu_port_baseline(args.cpu_n3, 3)
            if base is not None:
                per = sum(base["times"]) / len(base["times"])
                line["cpu_baseline"] = {"value": base["n_elm"] / per / 1e6, "unit": UNIT, "cores": base["cores"],
                                        "kind": base.get("kind", "reference"), "sample": base["sample"]}
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                        "sample": "oracle/_ref missing"}
            if args.workload == "t3d":
                anchor = cpu_reference_3d_scalar(12, 3)
                if anchor is not None:
                    line["cpu_reference_3d_scalar"] = anchor
            if args.workload == "t3d" and ref2d is not None:
                # for context: the unmodified reference on its own (2-D) implementation of the same forms
                per = sum(ref2d["times"]) / len(ref2d["times"])
                line["cpu_reference_2d"] = {"value": ref2d["n_elm"] / per / 1e6, "unit": UNIT, "cores": ref2d["cores"],
                                            "kind": "reference", "sample": ref2d["sample"]}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def _seeded_vector(pb):
    """x_i = a smooth function of the DOF's position, field and component: the same vector on every partition"""
    from feng_b200 import problems as PB
    x = np.zeros(pb.n_dof)
    for f, fld in enumerate(pb.num.fields):
        xyz = np.nan_to_num(PB.dof_coordinates(pb.mesh, pb.num, fld, pb.n_dof))
        comp = PB.dof_components(pb.num, fld, pb.n_dof)
        sel = comp >= 0
        x[sel] = np.sin(3.1 * xyz[sel, 0] + 5.3 * xyz[sel, 1] + 7.7 * xyz[sel, 2] + 1.3 * (4 * f + comp[sel]))
    return x[:pb.n_inc]


def parity_check(rank, world, local_rank, n=8):
    """Correctness stamp on the bench line itself.  N = 1: GPU assembly of T3D(n) against the CPU element loops applied
    matrix-free (oracle/port_cpp.cpp), every row.  N > 1: the distributed operator and rhs of `world` slabs (owned rows, summed over
    the ranks) against the undecomposed box assembled by rank 0 on its GPU alone -- sum of (A x)_i^2 and of rhs_i^2 for a seeded,
    partition-independent x."""
    import torch
    import torch.distributed as dist
    from feng_b200 import mesh as M, partition as PT, problems as PB
    from feng_b200.linear_system import LinearSystemB200

    def sums(pb, part):
        ls = LinearSystemB200(pb, device=local_rank, device_pattern=True)     # no communicator: local products on complete inputs
        ls.sys.set_solution(pb.sol)
        ls.sys.set_to_zero(3)
        ls.sys.assemble(3)
        x = _seeded_vector(pb)
        y, r = ls.sys.spmv(x), ls.sys.get_rhs()
        own = np.ones(pb.n_inc, bool) if part is None else part.owned.astype(bool)
        ls.sys.close()
        return float((y[own] ** 2).sum()), float((r[own] ** 2).sum()), x, y, r

    if world == 1:
        from oracle import port
        pb = PB.taylor_hood(M.cube_mesh(n), "ns_div", 6, 3, MU, RHO, build_pattern=False, with_source=False)
        _, _, x, y, r = sums(pb, None)
        if not port.available():
            return {"ok": None, "note": "oracle/_ref/libfeng_port.so missing"}
        cy, cr, _ = port.PortProblem(pb, pattern_free=True).apply(pb.sol, [x])
        dy = float(np.abs(cy[0] - y).max() / np.abs(cy[0]).max())
        dr = float(np.abs(cr - r).max() / np.abs(cr).max())
        return {"against": f"CPU element loops (oracle/port_cpp.cpp, matrix-free) on T3D({n}), every row", "ax_max_rel_diff": dy,
                "rhs_max_rel_diff": dr, "ok": bool(dy <= 1e-12 and dr <= 1e-12)}
    pb, part = PT.slab_problem(n, rank, world, "ns_div", 6, 3, MU, RHO, build_pattern=False, with_source=False)
    sy, sr, *_ = sums(pb, part)
    t = torch.tensor([sy, sr], dtype=torch.float64, device="cuda")
    dist.all_reduce(t)
    out = None
    if rank == 0:
        mg = M.box_mesh(n, n, n * world, float(world))
        mg.point_pressure = 0
        pg = PB.taylor_hood(mg, "ns_div", 6, 3, MU, RHO, build_pattern=False, with_source=False)
        gy, gr, *_ = sums(pg, None)
        dy, dr = abs(float(t[0]) - gy) / gy, abs(float(t[1]) - gr) / gr
        out = {"against": f"the undecomposed {world} x T3D({n}) box assembled on one GPU", "sum_ax2_rel_diff": dy, "sum_rhs2_rel_diff": dr,
               "ok": bool(dy <= 1e-12 and dr <= 1e-12)}
    return out


def newton_step(ls, sol, pb, maxit):
    """One loop body of solveNewtonRaphson (src/feNonLinearSolver.cpp:77-123) timed on the host with syncs; the Krylov
    solve is capped at `maxit` iterations (reported with its convergence flag and relative residual)."""
    from feng_b200 import capi
    S = ls.sys
    s = sol.copy()
    out = {}
    # two consecutive Newton iterations from the same state: the first pays the one-time, pattern-only set-up of the
    # preconditioner (parents, aggregates, coarse patterns); `ms` is the second, i.e. every later iteration of the Newton loop
    for which in ("first", "steady"):
        S.sync()
        t0 = time.perf_counter()
        S.set_solution(s)
        S.set_to_zero(3)
        S.assemble(1)
        S.rhs_max_norm()
        S.assemble(2)
        S.constrain()
        info = S.solve(1e-8, 1e-14, 1e6, maxit, 30, ls.pc, raise_on_fail=False)
        S.correct_solution(None)
        S.sync()
        dt = time.perf_counter() - t0
        out[which] = (dt * 1e3, S.last_solve_ms(), info)
    dt_ms, solve_ms, info = out["steady"]
    its = max(int(info.iterations), 1)
    return {"ms": dt_ms, "gmres_iterations": info.iterations, "gmres_max_iterations": maxit, "gmres_restart": 30,
            "converged": bool(info.converged), "solve_ms": solve_ms, "solve_ms_per_iteration": solve_ms / its,
            "rel_residual": info.rel_residual, "rel_tol": 1e-8, "pc": ls.pc,
            "pc_name": {1: "jacobi", 4: "amg", 5: "schur_amg", 6: "auto (Schur complement + aggregation multigrid for Taylor-Hood)"}.get(ls.pc, str(ls.pc)),
            "assembly_and_update_ms": dt_ms - solve_ms,
            "first_step_ms": out["first"][0], "first_step_solve_ms": out["first"][1],
            "note": "ms = assemble residual + matrix, constrain, GMRES(30) solve to rtol 1e-8, update; first_step_* includes the one-time "
                    "symbolic set-up of the multigrid hierarchy"}


if __name__ == "__main__":
    main()
