// Internal state of one b200_system (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/feng_b200.h"

namespace b200 {

void set_error(const std::string &msg);
void count_launch(int n = 1);
// B200_VERBOSE=1: synchronise, report the CUDA error state and the device memory in use after a set-up stage
void log_stage(const char *what);

#define B200_CUDA(call)                                                                                       \
  do {                                                                                                        \
    cudaError_t _e = (call);                                                                                  \
    if(_e != cudaSuccess) {                                                                                   \
      b200::set_error(std::string(#call) + ": " + cudaGetErrorString(_e));                                    \
      return B200_ERR_CUDA;                                                                                   \
    }                                                                                                         \
  } while(0)

struct Space {
  int                 nS = 0, nc = 1; // scalar functions, components
  int32_t            *d_adr = nullptr; // [nElm][nS*nc]
  std::vector<double> L, dL;           // host copies of the reference tables
};

struct Form {
  int     kind = 0, su = 0, sp = -1;
  double  coeff = 1., param = 1.;
  double *d_source = nullptr; // [nElm][nq][ncomp]
  double *d_coeff_table = nullptr; // [nElm][nq] tabulated coefficient callback (b200_set_form_coefficient), or nullptr
};

// Coefficients of the fused Taylor-Hood kernel: sums over the registered forms, see assemble.cu
struct THCoeffs {
  double c_conv = 0.;  // VECTOR_CONVECTIVE_ACCELERATION coeff
  double c_sig = 0.;   // DIV_NEWTONIAN_STRESS coeff
  double sig_mu = 0.;  // DIV_NEWTONIAN_STRESS coeff * viscosity
  double c_div = 0.;   // MIXED_DIVERGENCE coeff
  double diff_k = 0.;  // VECTOR_DIFFUSION coeff * diffusivity
  double c_gradp = 0.; // MIXED_GRADIENT coeff
  double c_mass = 0.;  // TRANSIENT_VECTOR_MASS coeff
  double c_src = 0.;   // 1 if a VECTOR_SOURCE form is present
};

struct ScalarCoeffs {
  double k = 0.;      // DIFFUSION diffusivity (its coefficient callback)
  double c_mass = 0.; // TRANSIENT_MASS coeff
  double c_src = 0.;
};

enum PlanKind { PLAN_NONE = 0, PLAN_SCALAR = 1, PLAN_TAYLOR_HOOD = 2, PLAN_CHNS = 3 };

struct System {
  int          device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t  ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  float        last_assemble_ms = 0.f, last_solve_ms = 0.f;

  // mesh
  int      dim = 0, nv = 0;
  int64_t  nVert = 0, nElm = 0;
  double  *d_xyz = nullptr;  // [nVert][dim]
  int32_t *d_conn = nullptr; // [nElm][nv]
  int      nq = 0;
  std::vector<double> w;

  std::vector<Space> spaces;
  std::vector<Form>  forms;

  // linear system
  int64_t  nInc = 0, nDOF = 0, nnz = 0;
  int64_t *d_ia = nullptr;
  int32_t *d_ja = nullptr;
  double  *d_val = nullptr, *d_rhs = nullptr, *d_du = nullptr, *d_sol = nullptr, *d_soldot = nullptr;
  double   c0 = 0., t = 0.;
  bool     have_soldot = false;
  // device-resident time stepping (capi.cu: b200_state_push / b200_state_bdf / b200_set_essential)
  double        *d_hist[2] = {nullptr, nullptr}; // u_n, u_{n-1}
  int64_t       *d_ess_idx = nullptr;
  double        *d_ess_val = nullptr;
  int64_t        ess_cap = 0, ess_n = 0;
  const int64_t *ess_idx_host = nullptr;

  // fused plan
  PlanKind     plan = PLAN_NONE;
  int          su = -1, sp = -1;       // primary / pressure space ids
  int          M = 0;                  // local rows = local cols of the fused element system
  THCoeffs     th, th_transient;       // all forms / transient-matrix forms only
  ScalarCoeffs sc, sc_transient;
  double      *d_source = nullptr;     // source table of the (single) source form
  const double *d_kcoef = nullptr;     // tabulated diffusivity of the scalar diffusion form (owned by its Form)
  int32_t     *d_slot = nullptr;       // [nElm][M][M] CSR slot or -1
  double      *d_tab = nullptr;        // packed basis tables + weights for the kernels
  int          tab_len = 0;
  bool         has_matrix_block[2][2] = {{false, false}, {false, false}}; // [U|P][U|P]

  // scatter
  int                  scatter_mode = B200_SCATTER_ATOMIC;
  int                  n_colors = 0;
  int32_t             *d_color_elems = nullptr; // elements sorted by colour
  std::vector<int64_t> color_ptr;               // [n_colors+1]

  // constraints
  int64_t  n_crow = 0, n_per = 0;
  int64_t *d_crows = nullptr;
  char    *d_cflag = nullptr; // [nInc]
  int64_t *d_master = nullptr, *d_slave = nullptr;
  std::vector<int64_t> per_master_host, per_slave_host;

  // block Jacobi
  int64_t  n_blocks = 0;
  std::vector<int64_t> block_ptr, block_rows;

  // Krylov workspace (allocated lazily, krylov.cu)
  void *krylov = nullptr;
  // multigrid / Schur-complement preconditioner state (precond.cu, amg.cu)
  void    *precond = nullptr;
  uint64_t val_epoch = 0; // bumped whenever the matrix values change: the numeric part of a preconditioner is redone lazily

  // multi-GPU communicator + halo plan (comm.cu)
  void *comm = nullptr;

  // row-owner gather plan (gather.cu); assembly_mode: B200_ASSEMBLY_*
  void *gather = nullptr;
  int   assembly_mode = 0;
  // patch plan (patch.cu): block-slot owners over spatially compact element patches, built on top of the gather plan
  void *patch = nullptr;
  // setToZero is lazy: the gather kernels overwrite every row, so the memset is only materialised when something
  // else reads or accumulates into the arrays (bit 0 rhs, bit 1 matrix)
  int   pending_zero = 0;

  // monolithic CHNS weak form (chns.cu): spaces U, P, Phi, Mu; concatenated element->DOF table; packed tables
  bool             chns_active = false;
  int              chns_space[4] = {-1, -1, -1, -1};
  b200_chns_params chns_prm = {};
  int              chns_model = 0;       // 0 CHNS_Abels, 1 CHNS_MassAveraged, 2 CHNS_Khanwale
  double           dt = 0.;              // time step (b200_set_solution_n)
  double          *d_soln = nullptr;     // state at the previous time step (b200_set_solution_n)
  bool             have_soln = false;
  int32_t         *chns_adr = nullptr;
  uint16_t        *chns_off = nullptr;   // [nElm][M][M] row-local CSR offsets of the local entries
  double          *chns_tab = nullptr;
  int              chns_tab_len = 0;

  // scratch for reductions
  double *d_scratch = nullptr;
  double *h_scratch = nullptr; // pinned
};

// assemble.cu
int build_plan(System *S);
int launch_assemble(System *S, int what, int only_transient);
// gather.cu
int  build_gather_plan(System *S);
int  launch_gather(System *S, int what, const THCoeffs &c);
void gather_free(System *S);
struct GatherTables {
  const double *d_tab = nullptr, *d_geo = nullptr; // pre-contracted reference tensors (GT layout), per-element inverse map + detJ
  int           tab_len = 0, tab_len_src = 0;
};
bool gather_tables(const System *S, GatherTables *out);
int  gather_kernel_kind(const System *S); // 0 no plan, 1 thread per node, 2 lane groups, 3 row lanes
// chns.cu
int  chns_analyze(System *S);
int  chns_build_plan(System *S);
int  chns_launch(System *S, int what);
void chns_free(System *S);
// capi.cu
int  flush_zero(System *S, int what);
// comm.cu
void          comm_free(System *S);
bool          comm_active(const System *S);
int           comm_rank(const System *S);
int           comm_world(const System *S);
const double *comm_mask(const System *S); // 1/0 per row (owned / ghost) or nullptr on a single GPU
int           comm_halo_exchange(System *S, double *d_x);
int           comm_spmv_overlapped(System *S, double *d_x, double *d_y); // halo update of x hidden behind the interior rows
int           comm_allreduce(System *S, double *d_buf, int count, bool max_op);
int           comm_allgather64(System *S, const void *send, void *recv, size_t count);
void          comm_boundary_rows(const System *S, const int32_t **rows, int64_t *n);
// numbering.cu
int  unique_edges(int device, int64_t nV, int64_t n, const int32_t *pairs, int32_t *edge_of_pair, int32_t *edges, int64_t *n_edges);
// norms.cu
int  error_norm(System *S, int space, int kind, int p, const double *exact, double *out);
// krylov.cu
int  spmv(System *S, const double *d_x, double *d_y);
// rows with skip[row] == 0 (rows == nullptr) or the listed rows (skip == nullptr)
int  spmv_rows(System *S, const double *d_x, double *d_y, const uint8_t *skip, const int32_t *rows, int64_t n_rows);
int  gmres_solve(System *S, const b200_solver_options *opt, b200_solve_info *info);
void krylov_free(System *S);
int  max_abs(System *S, const double *d_x, int64_t n, double *out);

} // namespace b200

struct b200_system : public b200::System {};
