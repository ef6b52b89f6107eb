/*
 * feng_b200 -- B200-native (sm_100a, FP64) finite-element assembly + Newton linear solve engine.
 *
 * C ABI of the drop-in boundary.  The reference (arthurbawin/feNG) talks to its linear-system backends only
 * through the 22 pure virtuals of class feLinearSystem (src/feLinearSystem.h:43-169); the adapter
 * adapter/feLinearSystemB200.h implements those virtuals on top of the entry points below, next to
 * feLinearSystemPETSc (src/feLinearSystem.h:174) and feLinearSystemMklPardiso (src/feLinearSystem.h:276).
 * Each entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on failure; b200_last_error() gives the message
 *     (maps onto feStatus / feErrorMsg, src/feMessage.h:44-97).  There is NO CPU fallback: without a CUDA
 *     device every compute entry point fails with B200_ERR_CUDA.
 *   - all pointers are caller-owned HOST memory unless the name ends in _device; set-up tables are copied once.
 *   - indices: DOF numbers int32 (the reference's feInt tables are narrowed by the adapter), CSR row pointers
 *     int64, CSR columns int32.
 *   - sign convention of the reference: Be = -(weak residual), Ae = +d(weak residual)/du, the Newton system is
 *     A du = rhs (src/feLinearSystem.h:128-129).
 *   - calls on one system are not thread-safe (the reference drives a backend from one thread,
 *     src/feNonLinearSolver.cpp:77-123).
 */
#ifndef FENG_B200_H
#define FENG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_system b200_system;

enum {
  B200_OK          = 0,
  B200_ERR_ARG     = -1, /* bad argument / call order */
  B200_ERR_CUDA    = -2, /* CUDA runtime failure or no device */
  B200_ERR_UNSUPP  = -3, /* weak form or space not supported by the fused kernels */
  B200_ERR_SOLVER  = -4  /* Krylov solver diverged or broke down */
};

/* Weak-form kinds: numerically identical to the reference's elementSystemType (src/feSysElm.h:13-77), so the
 * adapter passes feBilinearForm::getID() (src/feBilinearForm.h:165) through unchanged. */
enum {
  B200_FORM_SOURCE                         = 0,  /* feSysElm_Source,                       src/feSysElm.cpp:16-27        */
  B200_FORM_VECTOR_SOURCE                  = 2,  /* feSysElm_VectorSource<dim>,            src/feVectorSysElm.cpp:112-128 */
  B200_FORM_TRANSIENT_MASS                 = 13, /* feSysElm_TransientMass,                src/feSysElm.cpp:426-460      */
  B200_FORM_TRANSIENT_VECTOR_MASS          = 15, /* feSysElm_TransientVectorMass<dim>,     src/feVectorSysElm.cpp:390-425 */
  B200_FORM_DIFFUSION                      = 16, /* feSysElm_Diffusion<dim>,               src/feSysElm.cpp:530-583      */
  B200_FORM_VECTOR_DIFFUSION               = 18, /* feSysElm_VectorDiffusion<dim>,         src/feVectorSysElm.cpp:449-503 */
  B200_FORM_VECTOR_CONVECTIVE_ACCELERATION = 22, /* feSysElm_VectorConvectiveAcceleration, src/feVectorSysElm.cpp:1171-1242 */
  B200_FORM_DIV_NEWTONIAN_STRESS           = 25, /* feSysElm_DivergenceNewtonianStress,    src/feVectorSysElm.cpp:1454-1532 */
  B200_FORM_MIXED_GRADIENT                 = 26, /* feSysElm_MixedGradient<dim>,           src/feVectorSysElm.cpp:528-578 */
  B200_FORM_MIXED_DIVERGENCE               = 31, /* feSysElm_MixedDivergence<dim>,         src/feVectorSysElm.cpp:685-749 */
  B200_FORM_CHNS_ABELS                     = 35, /* CHNS_Abels<2> (volume-averaged CHNS),  src/feSysElmCHNS.cpp:66-273; Jacobian by
                                                    finite differences, src/feBilinearForm.cpp:388-428 */
  B200_FORM_CHNS_MASS_AVERAGED             = 36, /* CHNS_MassAveraged<2>,                  src/feSysElmCHNS.cpp:347-602; needs the Phi
                                                    DOFs of the previous time step (b200_set_solution_n) and the gradient
                                                    table of the pressure space */
  B200_FORM_CHNS_KHANWALE                  = 38  /* CHNS_Khanwale<2> (non-dimensional, time-averaged fields),
                                                    src/feSysElmCHNS.cpp:678-938; needs the whole state of the previous time
                                                    step and the time step (b200_set_solution_n) */
};

/* Parameters of the monolithic Cahn-Hilliard Navier-Stokes weak form.  The reference passes host callbacks
 * (feFunction) for density, viscosity and mobility; CHNS_Solver only ever installs the laws below
 * (src/CHNS_Solver.cpp:124-235), which the device evaluates itself:
 *   rho(phi) = (rho_a - rho_b)/2 phi + (rho_a + rho_b)/2, eta(phi) likewise; limiter != 0 clips phi to [-1, 1] first;
 *   mobility M, or M |1 - phi^2| if degenerate_mobility != 0;  lambda = 3/(2 sqrt 2) sigma epsilon (src/feSysElm.h:1338).
 * Volume force and the four source terms are constants here (the reference's tests use zero sources). */
typedef struct {
  double rho_a, rho_b, visc_a, visc_b, mobility, surface_tension, epsilon;
  double force[3], source_u[3], source_p, source_phi, source_mu;
  int    limiter, degenerate_mobility;
  double mass_alpha; /* CHNS_MassAveraged only: alpha = (rho_2 - rho_1) / (rho_1 + rho_2), CHNSparameters[0] (src/feSysElm.h:1419) */
  double khanwale[7]; /* CHNS_Khanwale only: Re, Pe, Cn, We, Fr, rhoA, rhoB = CHNSparameters[0..6] (src/feSysElm.h:1501-1507); its
                         volume force is the constant (0, -1) of src/feSysElmCHNS.cpp:623, mobility and force[] are unused */
} b200_chns_params;

/* Scatter strategies for the race-free add into the CSR matrix (north-star subsystem 3). */
enum {
  B200_SCATTER_ATOMIC  = 0, /* one launch, red.global.add.f64 into precomputed CSR slots               */
  B200_SCATTER_COLORED = 1  /* one launch per element colour, plain read-modify-write; needs b200_set_colors
                               (feCncGeo::colorElements, src/feCncGeo.cpp:752-794)                        */
};

/* Assembly strategies.  AUTO picks GATHER whenever the registered problem qualifies (fused Taylor-Hood P2/P1 on straight
 * simplices), else SCATTER. */
enum {
  B200_ASSEMBLY_AUTO    = 0,
  B200_ASSEMBLY_SCATTER = 1, /* element-major quadrature-loop kernel + B200_SCATTER_* into precomputed CSR slots          */
  B200_ASSEMBLY_GATHER  = 2  /* row-owner kernel on pre-contracted reference tensors: every CSR row written exactly once,
                                no memset, no atomics, deterministic                                                    */
};

/* Preconditioners of the restarted GMRES (north-star subsystem 4). */
enum {
  B200_PC_NONE         = 0,
  B200_PC_JACOBI       = 1, /* point Jacobi; zero diagonals (pressure rows of Taylor-Hood) are replaced by 1 */
  B200_PC_BLOCK_JACOBI = 2, /* dense inverses of the diagonal blocks given by b200_set_blocks (<= 32 rows each); rows outside
                               every block fall back to point Jacobi; a singular block is reported (B200_ERR_SOLVER)    */
  /* 3 is not used: PETSc's sequential default, ILU(0) (src/feLinearSystem.h:198), is a chain of sparse triangular solves and is
     replaced on this hardware by the multigrid-based preconditioners below                                              */
  B200_PC_AMG          = 4, /* one multigrid cycle on the whole matrix (scalar diffusion-type systems): P2 -> P1 on the same
                               mesh, then aggregation levels (W-cycle, over-corrected), Chebyshev-Jacobi smoothing
                               (csrc/amg.cu)                                                                                */
  B200_PC_SCHUR_AMG    = 5, /* Taylor-Hood saddle-point systems: block upper-triangular preconditioner, velocity block by
                               one multigrid cycle, Schur complement by the scaled pressure mass diagonal with a rank-one
                               term for a pinned pressure (csrc/precond.cu)                                                 */
  B200_PC_AUTO         = 6  /* SCHUR_AMG for Taylor-Hood systems, AMG for scalar systems, JACOBI otherwise (CHNS)           */
};

typedef struct {
  double rel_tol;   /* feLinearSystem::_rel_tol,  src/feLinearSystem.h:66 (default 1e-8)  */
  double abs_tol;   /* feLinearSystem::_abs_tol,  src/feLinearSystem.h:67 (default 1e-14) */
  double div_tol;   /* feLinearSystem::_div_tol,  src/feLinearSystem.h:68 (default 1e6)   */
  int    max_iter;  /* feLinearSystem::_max_iter, src/feLinearSystem.h:69 (default 1e4)   */
  int    restart;   /* GMRES restart length m (PETSc default 30)                           */
  int    pc;        /* B200_PC_*                                                           */
} b200_solver_options;

typedef struct {
  double norm_dx;        /* max-norm of du            (solve(): normDx)        */
  double norm_rhs;       /* max-norm of the rhs       (solve(): normResidual)  */
  double norm_axb;       /* max-norm of A du - rhs    (solve(): normAxb)       */
  int    iterations;     /* Krylov iterations         (solve(): nIter)         */
  int    converged;      /* 1 if the tolerance was met */
  double rel_residual;   /* final preconditioned residual 2-norm / initial     */
} b200_solve_info;

const char *b200_last_error(void);
/* Number of CUDA kernels this library has launched since load / since the last reset (bench.py: gpu_launches). */
int64_t b200_kernel_launches(void);
void    b200_reset_kernel_launches(void);

/* ---- life cycle: feLinearSystem ctor/dtor (src/feLinearSystem.h:81-83, src/feLinearSystemMklPardiso.cpp:174-, :1224-1257) */
int  b200_create(b200_system **out, int device);
void b200_destroy(b200_system *s);

/* ---- set-up tables, copied once (what feBilinearForm::initialize gathers per element on the host,
 *      src/feBilinearForm.cpp:284-367) ---- */
/* vertices and connectivity of the interior ("Domaine") connectivity: feMesh::getVertices (src/feMesh.h:130),
 * feCncGeo::getVerticesConnectivity (src/feCncGeo.h:178).  Straight simplices only (nv = dim+1): the Jacobians
 * feCncGeo::_J (src/feCncGeo.cpp:278-418) and the ElementTransformation (src/feCncGeo.cpp:651-692) are
 * recomputed on the device from the vertices. */
int b200_set_mesh(b200_system *s, int dim, int64_t n_vertices, const double *xyz /* [n_vertices][3] */,
                  int64_t n_elements, int n_vertices_per_element, const int32_t *connectivity);
/* quadrature weights, feSpace::getQuadratureWeights (src/feSpace.h:447) */
int b200_set_quadrature(b200_system *s, int n_quad, const double *weights);
/* one interpolation space: element->DOF table from feSpace::initializeAddressingVector (src/feSpace.h:454) and the
 * SCALAR reference basis at the quadrature nodes, L[k][i], dL[k][i][dim] (feSpace::_L,_dLdr,_dLds,_dLdt,
 * src/feSpace.h:140-147, layout src/feSpace.h:381-384).  For a vector space (n_components = dim) pass the scalar
 * tables of its Lagrange basis: function a*dim+c is phi_a e_c (src/feSpace_2D.cpp:41-55).  Returns the space id. */
int b200_add_space(b200_system *s, int n_scalar_functions, int n_components, const int32_t *adr,
                   const double *L, const double *dL);
/* one weak form (createBilinearForm, src/feBilinearForm.h:23): kind = B200_FORM_*, spaces by id (space_p = -1 if
 * the form has no second field), coeff / param = the constant values of the form's coefficient callbacks
 * (feFunction, src/feFunction.h:78-118), source = host tabulation of the source callback at every (element,
 * quadrature node[, component]) or NULL. */
int b200_add_form(b200_system *s, int kind, int space_u, int space_p, double coeff, double param,
                  const double *source);
/* the monolithic CHNS weak form on the spaces {U, P, Phi, Mu} (createBilinearForm(..., {u, p, phi, mu}, new CHNS_Abels<2> |
 * CHNS_MassAveraged<2> | CHNS_Khanwale<2>), src/CHNS_Solver.cpp:370-446); kind = B200_FORM_CHNS_*.  It must be the only form of the system; both the residual and the finite-difference
 * Jacobian (N+1 residual evaluations per element) run on the device. */
int b200_add_form_chns(b200_system *s, int kind, int space_u, int space_p, int space_phi, int space_mu,
                       const b200_chns_params *params);
/* state vector at the previous time step, nDOF doubles: what feBilinearForm::initialize copies into _solAtTimeN from the
 * global solAtTimeN (src/feBilinearForm.cpp:277,347; set by the time integrators through src/feNonLinearSolver.cpp:35),
 * and the time step feBilinearForm::initialize takes from feSolution::getTimeStep() (src/feBilinearForm.cpp:293).
 * Read by B200_FORM_CHNS_MASS_AVERAGED (Phi) and B200_FORM_CHNS_KHANWALE (every field, dt); sol_n = NULL (or never
 * called) means "equal to the current solution". */
int b200_set_solution_n(b200_system *s, const double *sol_n, double dt);
/* replace the tabulated source of form `form_id` (returned by b200_add_form): time-dependent source callbacks are
 * re-tabulated by the adapter when feSolution::getCurrentTime() changes (the reference evaluates the callback with
 * args.t = tn on every element visit, src/feBilinearForm.cpp:291-295) */
int b200_set_source(b200_system *s, int form_id, const double *source);
/* space- (and time-) dependent coefficient callback of form `form_id`, tabulated by the host at every (element, quadrature
 * node) exactly where the reference evaluates it (feSysElm_Diffusion, src/feSysElm.cpp:538, :566); the form's constant coeff x
 * param multiplies the table.  Built for B200_FORM_DIFFUSION (the quadrature-loop kernel); the fused Taylor-Hood kernels work on
 * pre-contracted tensors and keep constant coefficients (B200_ERR_UNSUPP).  Call before b200_finalize; calling it again later
 * replaces the values (time-dependent coefficients). */
int b200_set_form_coefficient(b200_system *s, int form_id, const double *table);
/* sparsity pattern of feEZCompressedRowStorage (src/feCompressedRowStorage.cpp:15-133), ia[n_inc+1], ja[nnz] */
int b200_set_pattern(b200_system *s, int64_t n_inc, int64_t n_dof, const int64_t *ia, const int32_t *ja);
/* same pattern built on the device from the spaces and forms registered so far; b200_get_pattern to read it back */
int b200_build_pattern(b200_system *s, int64_t n_inc, int64_t n_dof);
int b200_get_pattern_size(b200_system *s, int64_t *n_inc, int64_t *nnz);
int b200_get_pattern(b200_system *s, int64_t *ia, int32_t *ja);
/* element colours, feCncGeo::getColorElm (src/feCncGeo.h:236); only needed for B200_SCATTER_COLORED */
int b200_set_colors(b200_system *s, int n_colors, const int32_t *element_color);
int b200_set_scatter_mode(b200_system *s, int mode);
/* B200_ASSEMBLY_*; may be called before or after b200_finalize.  Returns B200_ERR_UNSUPP if GATHER is requested for a
 * problem that does not qualify. */
int b200_set_assembly_mode(b200_system *s, int mode);
/* 0: no write-once plan (scatter kernels); 1: the last b200_finalize built the row-owner gather plan; 2: it built the
 * patch plan (block-slot owners over Morton patches of elements, the default where the numbering qualifies) */
int b200_has_gather_plan(const b200_system *s);
/* Edge tags of a simplicial mesh, computed on the device: `pairs` lists the vertex pairs (v0, v1) of every local edge in the order
 * the reference's reader sweeps them (boundary triangles first, then the cells; local edges as src/feTriangle.cpp:3-26 /
 * src/feTetrahedron.h:31).  edge_of_pair[i] = tag of pair i = rank of its edge by FIRST appearance (std::set insertion of
 * src/feMeshRead.cpp:1412-1456, :1613-1614); edges[tag] = the pair as first met (NULL to skip; capacity n_pairs pairs);
 * *n_edges = number of distinct edges.  feNumber numbers the mid-edge DOFs of a P2 space in tag order (src/feNumber.cpp:370-483).
 * Needs no b200_system. */
int b200_unique_edges(int device, int64_t n_vertices, int64_t n_pairs, const int32_t *pairs, int32_t *edge_of_pair, int32_t *edges,
                      int64_t *n_edges);
/* which kernels the gather plan of the last b200_finalize launches for the velocity rows: 0 no plan, 1 thread per node, 2 lane
 * groups (10 lanes per node), 3 row lanes (lane per matrix row, P2/P1 tetrahedra whose velocity nodes are three adjacent unknowns;
 * csrc/gather_urow.cuh).  Diagnostic: the choice is made by the engine (B200_GATHER_KERNEL overrides it). */
int b200_gather_kernel(const b200_system *s);
/* Error norms of the state on the device (b200_set_solution / b200_correct_solution) for one space, feNorm of the reference:
 *   kind = B200_NORM_LP:       ( int sum_i |u_i - uh_i|^p )^(1/p)      computeLpNorm / computeVectorLpNorm, src/feNorm.cpp:323-398, :1399-1443
 *   kind = B200_NORM_H1_SEMI:  sqrt( int |grad u - grad uh|^2 )        computeH1SemiNorm / computeVectorH1SemiNorm, :1643-1732, :1794-1846
 * `exact` tabulates the exact field (the reference evaluates its feFunction at the physical quadrature node):
 * [n_elements][n_quad][n_components] values for B200_NORM_LP, [n_elements][n_quad][n_components][dim] gradients for
 * B200_NORM_H1_SEMI; NULL: the norm of uh itself.  Deterministic (fixed summation order).  Single-GPU systems only. */
#define B200_NORM_LP 0
#define B200_NORM_H1_SEMI 1
int b200_error_norm(b200_system *s, int space, int kind, int p, const double *exact, double *out);
/* rows of essential vector components (src/feLinearSystemMklPardiso.cpp:998-1041) and periodic (master, slave)
 * pairs (feMetaNumber::PeriodicDOF, src/feNumber.h:234) */
int b200_set_constraints(b200_system *s, int64_t n_rows, const int64_t *rows, int64_t n_periodic,
                         const int64_t *master, const int64_t *slave);
/* periodic (master, slave) DOF pairs alone; may be called BEFORE the pattern exists, and must be when the pattern is built
 * on the device (b200_build_pattern adds the (slave, master) entries of src/feCompressedRowStorage.cpp:96-107).
 * b200_apply_periodicity fails if a pair has no (slave, master) entry in the pattern. */
int b200_set_periodic(b200_system *s, int64_t n_periodic, const int64_t *master, const int64_t *slave);
/* diagonal blocks for B200_PC_BLOCK_JACOBI: block_ptr[n_blocks+1] into block_rows[] */
int b200_set_blocks(b200_system *s, int64_t n_blocks, const int64_t *block_ptr, const int64_t *block_rows);
/* compile the fused assembly plan (coefficients, CSR slot map); must follow the set-up calls above */
int b200_finalize(b200_system *s);

/* ---- the 22 virtuals of feLinearSystem ---- */
/* getSystemSize (src/feLinearSystem.h:87) */
int64_t b200_system_size(const b200_system *s);
/* state upload: what feBilinearForm::initialize reads from feSolution (sol, solDot, c0, tn),
 * src/feBilinearForm.cpp:291-349; n_dof doubles each, sol_dot may be NULL */
int b200_set_solution(b200_system *s, const double *sol, const double *sol_dot, double c0, double t);
/* ---- device-resident time stepping: after b200_set_solution has been called ONCE, a B200-aware host keeps
 *      the state on the device -- b200_correct_solution updates it in place -- and only these cross the bus per step:
 *      the essential-DOF values of the new time level and a handful of scalars. ---- */
/* essential-BC refresh, feSolution::initializeEssentialBC (src/feTimeIntegration.cpp:523-536): sol[dofs[i]] = values[i] on the
 * device copy.  The index array is uploaded once as long as the same host array is passed again. */
int b200_set_essential(b200_system *s, int64_t n, const int64_t *dofs, const double *values);
/* start of a time step: shift the device history (u_{n-1} <- u_n <- current state; feSolutionContainer::rotate,
 * src/feSolutionContainer.cpp:82-105) and record the state at time n that the time-averaged CHNS forms read
 * (solAtTimeN, src/feNonLinearSolver.cpp:60) */
int b200_state_push(b200_system *s);
/* solDot = sum_j coef[j] u_{n+1-j} with u_{n+1} the current device state (BDFContainer::computeSolTimeDerivative,
 * src/feSolutionContainer.cpp:338-348); c0 = coef[0], current time t and time step dt as b200_set_solution / _n take them */
int b200_state_bdf(b200_system *s, int n_coef, const double *coef, double t, double dt);
/* setToZero / setMatrixToZero / setResidualToZero (src/feLinearSystem.h:110-112): what = 1 rhs, 2 matrix, 3 both */
int b200_set_to_zero(b200_system *s, int what);
/* assemble / assembleMatrices / assembleResiduals (src/feLinearSystem.h:115-120): what = 1 residual, 2 matrix,
 * 3 both in ONE fused pass; only_transient as assembleOnlyTransientMatrices */
int b200_assemble(b200_system *s, int what, int only_transient);
/* getRHSMaxNorm / getResidualMaxNorm (src/feLinearSystem.h:106-107) */
int b200_rhs_max_norm(b200_system *s, double *norm);
int b200_du_max_norm(b200_system *s, double *norm);
/* constrainEssentialComponents + applyPeriodicity (src/feLinearSystem.h:122-124) */
int b200_constrain(b200_system *s);
int b200_apply_periodicity(b200_system *s);
/* solve (src/feLinearSystem.h:137-138): restarted GMRES on the device */
int b200_solve(b200_system *s, const b200_solver_options *opt, b200_solve_info *info);
/* correctSolution (src/feLinearSystem.h:141-142): sol[i] += du[i] (i < nInc) on the device copy, and into the
 * host vector if sol_host != NULL (n_dof doubles, the caller's feSolution::getSolution()) */
int b200_correct_solution(b200_system *s, double *sol_host, int correct_dot);
/* assignResidualToDCResidual / applyCorrectionToResidual (src/feLinearSystem.h:147-153) */
int b200_get_rhs(b200_system *s, double *rhs);
int b200_axpy_rhs(b200_system *s, double coeff, const double *d);
/* viewMatrix / writeMatrix / writeRHS / writeResidual back-ends (src/feLinearSystem.h:156-168): raw downloads */
int b200_get_matrix_values(b200_system *s, double *values);
int b200_get_du(b200_system *s, double *du);
int b200_get_solution(b200_system *s, double *sol);
/* y = A x with the assembled matrix (parity checks of the SpMV kernel); n_inc doubles each */
int b200_spmv(b200_system *s, const double *x, double *y);

/* ---- multi-GPU (one process per GPU).  Elements are partitioned with one ghost layer, so the
 *      assembly needs no exchange; every row is owned by exactly one rank.  The solve exchanges the ghost entries of
 *      the SpMV input with the neighbours (ncclSend/ncclRecv) and all-reduces the Krylov dot products; norms returned by
 *      b200_solve / b200_rhs_max_norm / b200_du_max_norm are global.  Replaces the replicated-mesh MPI scheme of the
 *      reference (src/feLinearSystemPETSc.cpp:626-640, :1055). ---- */
/* 128-byte NCCL unique id created on one rank and broadcast by the host (e.g. torch.distributed) */
int b200_comm_unique_id(char *id128);
int b200_comm_init(b200_system *s, const char *id128, int rank, int world);
/* owned[n_inc]: 1 if this rank owns the row.  For neighbour k: this rank sends x[send_idx[send_ptr[k]..send_ptr[k+1])]
 * and receives into x[recv_idx[recv_ptr[k]..recv_ptr[k+1])] (both sides list the shared rows in the same order). */
int b200_set_halo(b200_system *s, const uint8_t *owned, int n_neighbors, const int32_t *neighbor_rank, const int64_t *send_ptr,
                  const int32_t *send_idx, const int64_t *recv_ptr, const int32_t *recv_idx);
/* ghost entries of a host vector (n_inc doubles) <- owner's values, through the device path (tests) */
int b200_halo_exchange_host(b200_system *s, double *x);

/* ---- measurement helpers (bench.py); device time in milliseconds of the last call, from CUDA events recorded on
 *      the system's own stream ---- */
int b200_last_assemble_ms(const b200_system *s, float *ms);
int b200_last_solve_ms(const b200_system *s, float *ms);
/* time `reps` back-to-back SpMVs (y = A x, device resident) -> average ms per SpMV */
int b200_time_spmv(b200_system *s, int reps, float *ms_per_spmv);
int b200_sync(b200_system *s);
/* bracket an arbitrary sequence of calls with CUDA events on the system's stream */
int b200_time_begin(b200_system *s);
int b200_time_end(b200_system *s, float *ms);
/* DFMA micro-benchmark on `device`: measured FP64 FMA throughput in TFLOP/s (the assembly kernels' second roofline;
 * MEASURED_PEAKS.json only holds HBM and bf16 figures) */
int b200_measure_fp64_peak(int device, double *tflops);
/* same for the FP64 tensor-core path (mma.sync.m8n8k4.f64): the evidence behind the north-star's tensor-core gate */
int b200_measure_dmma_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* FENG_B200_H */
