#include <cstdio>
#include <cstdlib>
// C ABI of the engine (include/feng_b200.h): set-up, state transfer, constraints and the feLinearSystem virtuals.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>

#include "amg.h"
#include "system.h"

namespace b200 {

static thread_local std::string g_error;
static std::atomic<int64_t>     g_launches{0};

void set_error(const std::string &msg)
{
  g_error = msg;
  if(getenv("B200_VERBOSE")) fprintf(stderr, "[feng_b200] %s\n", msg.c_str());
}
void count_launch(int n) { g_launches += n; }
void log_stage(const char *what)
{
  if(!getenv("B200_VERBOSE")) return;
  const cudaError_t e = cudaDeviceSynchronize();
  size_t fr = 0, tot = 0;
  cudaMemGetInfo(&fr, &tot);
  fprintf(stderr, "[feng_b200] stage %-28s %s, %.1f GB in use\n", what, cudaGetErrorString(e), (double)(tot - fr) / 1e9);
}

static const int GRID = 148 * 8;

__global__ void fill_kernel(int64_t n, double *x, double v)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}

__global__ void add_kernel(int64_t n, const double *__restrict__ du, double *__restrict__ sol)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) sol[i] += du[i];
}

__global__ void axpy_kernel(int64_t n, double a, const double *__restrict__ x, double *__restrict__ y)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] += a * x[i];
}

// 16 independent DFMA chains per thread
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a)
{
  double x[16];
#pragma unroll
  for(int i = 0; i < 16; ++i) x[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  for(int it = 0; it < iters; ++it) {
#pragma unroll
    for(int i = 0; i < 16; ++i) x[i] = fma(x[i], a, 1e-12);
  }
  double s = 0.;
#pragma unroll
  for(int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP64 tensor-core micro-benchmark: 8 independent accumulator tiles of mma.sync.m8n8k4.f64 per warp (the only FP64 MMA shape of
// sm_100a), so that the north-star's "tensor cores only if they beat the CUDA-core form" gate rests on a measured number
__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters, double a)
{
  double c[8][2];
#pragma unroll
  for(int t = 0; t < 8; ++t) c[t][0] = c[t][1] = 0.;
  const double av = a + 1e-9 * threadIdx.x, bv = 1.0 - 1e-9 * threadIdx.x;
  for(int it = 0; it < iters; ++it) {
#pragma unroll
    for(int t = 0; t < 8; ++t)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(av), "d"(bv));
  }
  double s = 0.;
#pragma unroll
  for(int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// constrainEssentialComponents (src/feLinearSystemMklPardiso.cpp:1092-1114): zero the column, zero the row, unit
// diagonal, zero rhs.  One warp per matrix row.
__global__ void constrain_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja, double *__restrict__ val,
                                 double *__restrict__ rhs, const char *__restrict__ flag)
{
  const int     lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * (int64_t)blockDim.x) >> 5;
  for(int64_t i = w0; i < n; i += nw) {
    const bool rowc = flag[i] != 0;
    for(int64_t k = ia[i] + lane; k < ia[i + 1]; k += 32) {
      const int32_t j = ja[k];
      if(rowc)
        val[k] = (j == i) ? 1. : 0.;
      else if(flag[j])
        val[k] = 0.;
    }
    if(rowc && lane == 0) rhs[i] = 0.;
  }
}

// applyPeriodicity (src/feLinearSystemMklPardiso.cpp:1119-1149)
// *missing is raised when the pattern has no (slave, master) entry: the pattern was built without the periodic pairs
// (src/feCompressedRowStorage.cpp:96-107 adds them) and the row would silently pin du_slave to 0
__global__ void periodic_kernel(int64_t np, const int64_t *__restrict__ master, const int64_t *__restrict__ slave, int64_t nInc,
                                const int64_t *__restrict__ ia, const int32_t *__restrict__ ja, double *__restrict__ val, double *__restrict__ rhs,
                                double *missing)
{
  for(int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < np; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = master[p], s = slave[p];
    if(s < nInc && m < nInc) {
      bool found = false;
      for(int64_t k = ia[s]; k < ia[s + 1]; ++k) {
        double v = 0.;
        if(ja[k] == s) v = 1.;
        if(ja[k] == m) {
          v     = -1.;
          found = true;
        }
        val[k] = v;
      }
      rhs[s] = 0.;
      if(!found) *missing = 1.;
    }
  }
}

static void free_system(System *S)
{
  cudaSetDevice(S->device);
  krylov_free(S);
  precond_free(S);
  gather_free(S);
  chns_free(S);
  comm_free(S);
  for(auto &sp : S->spaces) cudaFree(sp.d_adr);
  for(auto &f : S->forms) {
    cudaFree(f.d_source);
    cudaFree(f.d_coeff_table);
  }
  cudaFree(S->d_xyz);
  cudaFree(S->d_conn);
  cudaFree(S->d_ia);
  cudaFree(S->d_ja);
  cudaFree(S->d_val);
  cudaFree(S->d_rhs);
  cudaFree(S->d_du);
  cudaFree(S->d_sol);
  cudaFree(S->d_soldot);
  cudaFree(S->d_soln);
  cudaFree(S->d_hist[0]);
  cudaFree(S->d_hist[1]);
  cudaFree(S->d_ess_idx);
  cudaFree(S->d_ess_val);
  cudaFree(S->d_slot);
  cudaFree(S->d_tab);
  cudaFree(S->d_color_elems);
  cudaFree(S->d_crows);
  cudaFree(S->d_cflag);
  cudaFree(S->d_master);
  cudaFree(S->d_slave);
  cudaFree(S->d_scratch);
  if(S->h_scratch) cudaFreeHost(S->h_scratch);
  if(S->ev0) cudaEventDestroy(S->ev0);
  if(S->ev1) cudaEventDestroy(S->ev1);
  if(S->ev2) cudaEventDestroy(S->ev2);
  if(S->ev3) cudaEventDestroy(S->ev3);
  if(S->stream) cudaStreamDestroy(S->stream);
}

static int alloc_linear_system(System *S)
{
  const size_t nb = (size_t)S->nInc * sizeof(double);
  cudaFree(S->d_val);
  cudaFree(S->d_rhs);
  cudaFree(S->d_du);
  cudaFree(S->d_sol);
  cudaFree(S->d_soldot);
  cudaFree(S->d_soln);
  S->d_soln    = nullptr;
  S->have_soln = false;
  cudaFree(S->d_hist[0]);
  cudaFree(S->d_hist[1]);
  S->d_hist[0] = S->d_hist[1] = nullptr;
  B200_CUDA(cudaMalloc(&S->d_val, (size_t)S->nnz * sizeof(double)));
  B200_CUDA(cudaMalloc(&S->d_rhs, nb));
  B200_CUDA(cudaMalloc(&S->d_du, nb));
  B200_CUDA(cudaMalloc(&S->d_sol, (size_t)S->nDOF * sizeof(double)));
  B200_CUDA(cudaMalloc(&S->d_soldot, (size_t)S->nDOF * sizeof(double)));
  B200_CUDA(cudaMemsetAsync(S->d_val, 0, (size_t)S->nnz * sizeof(double), S->stream));
  B200_CUDA(cudaMemsetAsync(S->d_rhs, 0, nb, S->stream));
  B200_CUDA(cudaMemsetAsync(S->d_du, 0, nb, S->stream));
  B200_CUDA(cudaMemsetAsync(S->d_sol, 0, (size_t)S->nDOF * sizeof(double), S->stream));
  B200_CUDA(cudaMemsetAsync(S->d_soldot, 0, (size_t)S->nDOF * sizeof(double), S->stream));
  return B200_OK;
}

int alloc_linear_system_public(System *S) { return alloc_linear_system(S); }

// materialise a lazy setToZero (see System::pending_zero)
int flush_zero(System *S, int what)
{
  const int todo = S->pending_zero & what;
  if(todo & 2) B200_CUDA(cudaMemsetAsync(S->d_val, 0, (size_t)S->nnz * sizeof(double), S->stream));
  if(todo & 1) B200_CUDA(cudaMemsetAsync(S->d_rhs, 0, (size_t)S->nInc * sizeof(double), S->stream));
  S->pending_zero &= ~todo;
  return B200_OK;
}
int build_pattern_device(System *S, int64_t n_inc, int64_t n_dof, const std::vector<int64_t> &per_master, const std::vector<int64_t> &per_slave);

} // namespace b200

using namespace b200;

#define CHECK_S(s)                                                                                            \
  do {                                                                                                        \
    if(!(s)) {                                                                                                \
      set_error("null system handle");                                                                        \
      return B200_ERR_ARG;                                                                                    \
    }                                                                                                         \
    if(cudaSetDevice((s)->device) != cudaSuccess) {                                                           \
      set_error("cudaSetDevice failed");                                                                      \
      return B200_ERR_CUDA;                                                                                   \
    }                                                                                                         \
  } while(0)

extern "C" {

const char *b200_last_error(void) { return g_error.c_str(); }
int64_t     b200_kernel_launches(void) { return g_launches.load(); }
void        b200_reset_kernel_launches(void) { g_launches = 0; }

int b200_create(b200_system **out, int device)
{
  if(!out) {
    set_error("b200_create: null output");
    return B200_ERR_ARG;
  }
  int ndev = 0;
  if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("b200_create: no CUDA device (this engine has no CPU fallback)");
    return B200_ERR_CUDA;
  }
  if(device < 0 || device >= ndev) {
    set_error("b200_create: bad device index");
    return B200_ERR_ARG;
  }
  B200_CUDA(cudaSetDevice(device));
  b200_system *s = new b200_system;
  s->device      = device;
  B200_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  B200_CUDA(cudaEventCreate(&s->ev0));
  B200_CUDA(cudaEventCreate(&s->ev1));
  B200_CUDA(cudaMalloc(&s->d_scratch, 256 * sizeof(double)));
  B200_CUDA(cudaMallocHost(&s->h_scratch, 256 * sizeof(double)));
  *out = s;
  return B200_OK;
}

void b200_destroy(b200_system *s)
{
  if(!s) return;
  free_system(s);
  delete s;
}

int b200_set_mesh(b200_system *s, int dim, int64_t n_vertices, const double *xyz, int64_t n_elements, int nv, const int32_t *conn)
{
  CHECK_S(s);
  if((dim != 2 && dim != 3) || nv != dim + 1 || !xyz || !conn || n_vertices <= 0 || n_elements <= 0) {
    set_error("b200_set_mesh: need straight triangles (dim 2, 3 vertices) or tetrahedra (dim 3, 4 vertices)");
    return B200_ERR_ARG;
  }
  s->dim   = dim;
  s->nv    = nv;
  s->nVert = n_vertices;
  s->nElm  = n_elements;
  std::vector<double> packed((size_t)n_vertices * dim);
  for(int64_t i = 0; i < n_vertices; ++i)
    for(int m = 0; m < dim; ++m) packed[i * dim + m] = xyz[3 * i + m];
  cudaFree(s->d_xyz);
  cudaFree(s->d_conn);
  B200_CUDA(cudaMalloc(&s->d_xyz, packed.size() * sizeof(double)));
  B200_CUDA(cudaMalloc(&s->d_conn, (size_t)n_elements * nv * sizeof(int32_t)));
  B200_CUDA(cudaMemcpy(s->d_xyz, packed.data(), packed.size() * sizeof(double), cudaMemcpyHostToDevice));
  B200_CUDA(cudaMemcpy(s->d_conn, conn, (size_t)n_elements * nv * sizeof(int32_t), cudaMemcpyHostToDevice));
  return B200_OK;
}

int b200_set_quadrature(b200_system *s, int n_quad, const double *weights)
{
  CHECK_S(s);
  if(n_quad <= 0 || n_quad > 128 || !weights) {
    set_error("b200_set_quadrature: 1..128 points");
    return B200_ERR_ARG;
  }
  s->nq = n_quad;
  s->w.assign(weights, weights + n_quad);
  return B200_OK;
}

int b200_add_space(b200_system *s, int nS, int nc, const int32_t *adr, const double *L, const double *dL)
{
  CHECK_S(s);
  if(s->nElm == 0 || s->nq == 0) {
    set_error("b200_add_space: set the mesh and the quadrature first");
    return B200_ERR_ARG;
  }
  if(nS <= 0 || nc <= 0 || !adr || !L || !dL) {
    set_error("b200_add_space: bad arguments");
    return B200_ERR_ARG;
  }
  Space sp;
  sp.nS = nS;
  sp.nc = nc;
  sp.L.assign(L, L + (size_t)s->nq * nS);
  sp.dL.assign(dL, dL + (size_t)s->nq * nS * s->dim);
  const size_t bytes = (size_t)s->nElm * nS * nc * sizeof(int32_t);
  B200_CUDA(cudaMalloc(&sp.d_adr, bytes));
  B200_CUDA(cudaMemcpy(sp.d_adr, adr, bytes, cudaMemcpyHostToDevice));
  s->spaces.push_back(sp);
  return (int)s->spaces.size() - 1;
}

int b200_add_form(b200_system *s, int kind, int space_u, int space_p, double coeff, double param, const double *source)
{
  CHECK_S(s);
  const int ns = (int)s->spaces.size();
  if(space_u < 0 || space_u >= ns || space_p >= ns) {
    set_error("b200_add_form: unknown space id");
    return B200_ERR_ARG;
  }
  Form f;
  f.kind  = kind;
  f.su    = space_u;
  f.sp    = space_p;
  f.coeff = coeff;
  f.param = param;
  if(kind == B200_FORM_SOURCE || kind == B200_FORM_VECTOR_SOURCE) {
    if(!source) {
      set_error("b200_add_form: source forms need the tabulated source");
      return B200_ERR_ARG;
    }
    const int    nc    = kind == B200_FORM_VECTOR_SOURCE ? s->dim : 1;
    const size_t count = (size_t)s->nElm * s->nq * nc;
    // the coefficient is folded into the table so that the kernels only ever see one table
    std::vector<double> scaled(source, source + count);
    if(coeff != 1.)
      for(auto &v : scaled) v *= coeff;
    B200_CUDA(cudaMalloc(&f.d_source, count * sizeof(double)));
    B200_CUDA(cudaMemcpy(f.d_source, scaled.data(), count * sizeof(double), cudaMemcpyHostToDevice));
  }
  s->forms.push_back(f);
  s->plan = PLAN_NONE;
  return (int)s->forms.size() - 1;
}

int b200_add_form_chns(b200_system *s, int kind, int space_u, int space_p, int space_phi, int space_mu, const b200_chns_params *params)
{
  CHECK_S(s);
  const int ns = (int)s->spaces.size();
  const int sp[4] = {space_u, space_p, space_phi, space_mu};
  for(int k = 0; k < 4; ++k)
    if(sp[k] < 0 || sp[k] >= ns) {
      set_error("b200_add_form_chns: unknown space id");
      return B200_ERR_ARG;
    }
  if((kind != B200_FORM_CHNS_ABELS && kind != B200_FORM_CHNS_MASS_AVERAGED && kind != B200_FORM_CHNS_KHANWALE) || !params) {
    set_error("b200_add_form_chns: CHNS_Abels, CHNS_MassAveraged and CHNS_Khanwale are built (CHNS_VolumeAveragedGeneric is not)");
    return B200_ERR_UNSUPP;
  }
  Form f;
  f.kind = kind;
  f.su   = space_u;
  f.sp   = space_p;
  s->forms.push_back(f);
  s->chns_active = true;
  for(int k = 0; k < 4; ++k) s->chns_space[k] = sp[k];
  s->chns_prm   = *params;
  s->chns_model = kind == B200_FORM_CHNS_MASS_AVERAGED ? 1 : (kind == B200_FORM_CHNS_KHANWALE ? 2 : 0);
  chns_free(s);
  s->plan = PLAN_NONE;
  return (int)s->forms.size() - 1;
}

int b200_set_source(b200_system *s, int form_id, const double *source)
{
  CHECK_S(s);
  if(form_id < 0 || form_id >= (int)s->forms.size() || !source || !s->forms[form_id].d_source) {
    set_error("b200_set_source: not a source form");
    return B200_ERR_ARG;
  }
  Form        &f     = s->forms[form_id];
  const int    nc    = f.kind == B200_FORM_VECTOR_SOURCE ? s->dim : 1;
  const size_t count = (size_t)s->nElm * s->nq * nc;
  std::vector<double> scaled(source, source + count);
  if(f.coeff != 1.)
    for(auto &v : scaled) v *= f.coeff;
  B200_CUDA(cudaMemcpy(f.d_source, scaled.data(), count * sizeof(double), cudaMemcpyHostToDevice));
  return B200_OK;
}

int b200_set_form_coefficient(b200_system *s, int form_id, const double *table)
{
  CHECK_S(s);
  if(form_id < 0 || form_id >= (int)s->forms.size() || !table) {
    set_error("b200_set_form_coefficient: bad form id / null table");
    return B200_ERR_ARG;
  }
  Form &f = s->forms[form_id];
  if(f.kind != B200_FORM_DIFFUSION) {
    set_error("b200_set_form_coefficient: tabulated coefficients are built for feSysElm_Diffusion (quadrature-loop kernel); the fused "
              "Taylor-Hood kernels work on pre-contracted tensors and need constant coefficients");
    return B200_ERR_UNSUPP;
  }
  const size_t count = (size_t)s->nElm * s->nq;
  const bool   fresh = f.d_coeff_table == nullptr;
  if(fresh) B200_CUDA(cudaMalloc(&f.d_coeff_table, count * sizeof(double)));
  B200_CUDA(cudaMemcpy(f.d_coeff_table, table, count * sizeof(double), cudaMemcpyHostToDevice));
  if(fresh) s->plan = PLAN_NONE; // picked up by the next b200_finalize
  ++s->val_epoch;
  return B200_OK;
}

int b200_set_pattern(b200_system *s, int64_t n_inc, int64_t n_dof, const int64_t *ia, const int32_t *ja)
{
  CHECK_S(s);
  if(n_inc <= 0 || n_dof < n_inc || !ia || !ja) {
    set_error("b200_set_pattern: bad arguments");
    return B200_ERR_ARG;
  }
  s->nInc = n_inc;
  s->nDOF = n_dof;
  s->nnz  = ia[n_inc];
  cudaFree(s->d_ia);
  cudaFree(s->d_ja);
  B200_CUDA(cudaMalloc(&s->d_ia, (size_t)(n_inc + 1) * sizeof(int64_t)));
  B200_CUDA(cudaMalloc(&s->d_ja, (size_t)s->nnz * sizeof(int32_t)));
  B200_CUDA(cudaMemcpy(s->d_ia, ia, (size_t)(n_inc + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
  B200_CUDA(cudaMemcpy(s->d_ja, ja, (size_t)s->nnz * sizeof(int32_t), cudaMemcpyHostToDevice));
  s->plan = PLAN_NONE;
  krylov_free(s);
  return alloc_linear_system(s);
}

int b200_build_pattern(b200_system *s, int64_t n_inc, int64_t n_dof)
{
  CHECK_S(s);
  if(n_inc <= 0 || n_dof < n_inc) {
    set_error("b200_build_pattern: bad sizes");
    return B200_ERR_ARG;
  }
  return build_pattern_device(s, n_inc, n_dof, s->per_master_host, s->per_slave_host);
}

int b200_get_pattern_size(b200_system *s, int64_t *n_inc, int64_t *nnz)
{
  CHECK_S(s);
  *n_inc = s->nInc;
  *nnz   = s->nnz;
  return B200_OK;
}

int b200_get_pattern(b200_system *s, int64_t *ia, int32_t *ja)
{
  CHECK_S(s);
  if(!s->d_ia) {
    set_error("b200_get_pattern: no pattern");
    return B200_ERR_ARG;
  }
  B200_CUDA(cudaMemcpy(ia, s->d_ia, (size_t)(s->nInc + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  B200_CUDA(cudaMemcpy(ja, s->d_ja, (size_t)s->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
  return B200_OK;
}

int b200_set_colors(b200_system *s, int n_colors, const int32_t *element_color)
{
  CHECK_S(s);
  if(n_colors <= 0 || !element_color || s->nElm == 0) {
    set_error("b200_set_colors: bad arguments");
    return B200_ERR_ARG;
  }
  std::vector<int64_t> ptr(n_colors + 1, 0);
  for(int64_t e = 0; e < s->nElm; ++e) {
    const int c = element_color[e];
    if(c < 0 || c >= n_colors) {
      set_error("b200_set_colors: colour out of range");
      return B200_ERR_ARG;
    }
    ptr[c + 1]++;
  }
  for(int c = 0; c < n_colors; ++c) ptr[c + 1] += ptr[c];
  std::vector<int32_t> list(s->nElm);
  std::vector<int64_t> cur(ptr.begin(), ptr.end() - 1);
  for(int64_t e = 0; e < s->nElm; ++e) list[cur[element_color[e]]++] = (int32_t)e; // element order inside a colour, as the reference
  cudaFree(s->d_color_elems);
  B200_CUDA(cudaMalloc(&s->d_color_elems, (size_t)s->nElm * sizeof(int32_t)));
  B200_CUDA(cudaMemcpy(s->d_color_elems, list.data(), (size_t)s->nElm * sizeof(int32_t), cudaMemcpyHostToDevice));
  s->n_colors  = n_colors;
  s->color_ptr = ptr;
  return B200_OK;
}

int b200_set_scatter_mode(b200_system *s, int mode)
{
  CHECK_S(s);
  if(mode != B200_SCATTER_ATOMIC && mode != B200_SCATTER_COLORED) {
    set_error("b200_set_scatter_mode: unknown mode");
    return B200_ERR_ARG;
  }
  s->scatter_mode = mode;
  return B200_OK;
}

int b200_set_constraints(b200_system *s, int64_t n_rows, const int64_t *rows, int64_t n_periodic, const int64_t *master, const int64_t *slave)
{
  CHECK_S(s);
  if(s->nInc == 0) {
    set_error("b200_set_constraints: set the pattern first");
    return B200_ERR_ARG;
  }
  std::vector<char> flag(s->nInc, 0);
  for(int64_t i = 0; i < n_rows; ++i) {
    if(rows[i] < 0 || rows[i] >= s->nInc) {
      set_error("b200_set_constraints: row out of range");
      return B200_ERR_ARG;
    }
    flag[rows[i]] = 1;
  }
  s->n_crow = n_rows;
  cudaFree(s->d_cflag);
  B200_CUDA(cudaMalloc(&s->d_cflag, (size_t)s->nInc));
  B200_CUDA(cudaMemcpy(s->d_cflag, flag.data(), (size_t)s->nInc, cudaMemcpyHostToDevice));
  if(n_periodic > 0 && master && slave) return b200_set_periodic(s, n_periodic, master, slave);
  return B200_OK;
}

int b200_set_periodic(b200_system *s, int64_t n_periodic, const int64_t *master, const int64_t *slave)
{
  CHECK_S(s);
  if(n_periodic < 0 || (n_periodic > 0 && (!master || !slave))) {
    set_error("b200_set_periodic: bad arguments");
    return B200_ERR_ARG;
  }
  s->n_per = n_periodic;
  s->per_master_host.assign(master, master + n_periodic);
  s->per_slave_host.assign(slave, slave + n_periodic);
  cudaFree(s->d_master);
  cudaFree(s->d_slave);
  s->d_master = s->d_slave = nullptr;
  if(n_periodic > 0) {
    B200_CUDA(cudaMalloc(&s->d_master, (size_t)n_periodic * sizeof(int64_t)));
    B200_CUDA(cudaMalloc(&s->d_slave, (size_t)n_periodic * sizeof(int64_t)));
    B200_CUDA(cudaMemcpy(s->d_master, master, (size_t)n_periodic * sizeof(int64_t), cudaMemcpyHostToDevice));
    B200_CUDA(cudaMemcpy(s->d_slave, slave, (size_t)n_periodic * sizeof(int64_t), cudaMemcpyHostToDevice));
  }
  return B200_OK;
}

int b200_set_blocks(b200_system *s, int64_t n_blocks, const int64_t *block_ptr, const int64_t *block_rows)
{
  CHECK_S(s);
  if(n_blocks <= 0 || !block_ptr || !block_rows) {
    set_error("b200_set_blocks: bad arguments");
    return B200_ERR_ARG;
  }
  if(s->nInc == 0) {
    set_error("b200_set_blocks: set the pattern first");
    return B200_ERR_ARG;
  }
  if(block_ptr[0] != 0) {
    set_error("b200_set_blocks: block_ptr[0] must be 0");
    return B200_ERR_ARG;
  }
  {
    std::vector<char> seen(s->nInc, 0);
    for(int64_t b = 0; b < n_blocks; ++b) {
      if(block_ptr[b + 1] <= block_ptr[b] || block_ptr[b + 1] - block_ptr[b] > 32) {
        set_error("b200_set_blocks: every block needs 1..32 rows");
        return B200_ERR_ARG;
      }
      for(int64_t k = block_ptr[b]; k < block_ptr[b + 1]; ++k) {
        const int64_t r = block_rows[k];
        if(r < 0 || r >= s->nInc || seen[r]) {
          set_error("b200_set_blocks: row out of range or listed twice");
          return B200_ERR_ARG;
        }
        seen[r] = 1;
      }
    }
  }
  s->n_blocks = n_blocks;
  s->block_ptr.assign(block_ptr, block_ptr + n_blocks + 1);
  s->block_rows.assign(block_rows, block_rows + block_ptr[n_blocks]);
  krylov_free(s);
  return B200_OK;
}

int b200_finalize(b200_system *s)
{
  CHECK_S(s);
  return build_plan(s);
}

int b200_set_assembly_mode(b200_system *s, int mode)
{
  CHECK_S(s);
  if(mode < B200_ASSEMBLY_AUTO || mode > B200_ASSEMBLY_GATHER) {
    set_error("b200_set_assembly_mode: unknown mode");
    return B200_ERR_ARG;
  }
  if(mode == B200_ASSEMBLY_GATHER && s->plan != PLAN_NONE && s->gather == nullptr) {
    set_error("b200_set_assembly_mode: no gather plan for this problem");
    return B200_ERR_UNSUPP;
  }
  const int rc = flush_zero(s, 3);
  if(rc != B200_OK) return rc;
  s->assembly_mode = mode;
  return B200_OK;
}

int b200_has_gather_plan(const b200_system *s) { return s && s->gather != nullptr ? (s->patch != nullptr ? 2 : 1) : 0; }

int b200_unique_edges(int device, int64_t n_vertices, int64_t n_pairs, const int32_t *pairs, int32_t *edge_of_pair, int32_t *edges,
                      int64_t *n_edges)
{
  return unique_edges(device, n_vertices, n_pairs, pairs, edge_of_pair, edges, n_edges);
}

int b200_gather_kernel(const b200_system *s) { return s ? gather_kernel_kind(s) : 0; }

int b200_error_norm(b200_system *s, int space, int kind, int p, const double *exact, double *out)
{
  CHECK_S(s);
  return error_norm(s, space, kind, p, exact, out);
}

int64_t b200_system_size(const b200_system *s) { return s ? s->nInc : 0; }

int b200_set_solution(b200_system *s, const double *sol, const double *sol_dot, double c0, double t)
{
  CHECK_S(s);
  if(!s->d_sol || !sol) {
    set_error("b200_set_solution: set the pattern first / null solution");
    return B200_ERR_ARG;
  }
  B200_CUDA(cudaMemcpyAsync(s->d_sol, sol, (size_t)s->nDOF * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  s->have_soldot = sol_dot != nullptr;
  if(sol_dot) B200_CUDA(cudaMemcpyAsync(s->d_soldot, sol_dot, (size_t)s->nDOF * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  s->c0 = c0;
  s->t  = t;
  B200_CUDA(cudaStreamSynchronize(s->stream)); // the host buffers may be pageable and reused by the caller
  return B200_OK;
}

int b200_set_solution_n(b200_system *s, const double *sol_n, double dt)
{
  CHECK_S(s);
  if(!s->d_sol) {
    set_error("b200_set_solution_n: set the pattern first");
    return B200_ERR_ARG;
  }
  s->dt        = dt;
  s->have_soln = sol_n != nullptr;
  if(!sol_n) return B200_OK;
  if(!s->d_soln) B200_CUDA(cudaMalloc(&s->d_soln, (size_t)s->nDOF * sizeof(double)));
  B200_CUDA(cudaMemcpyAsync(s->d_soln, sol_n, (size_t)s->nDOF * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  B200_CUDA(cudaStreamSynchronize(s->stream));
  return B200_OK;
}

// ---- device-resident time stepping: the state never leaves the GPU between Newton iterations / time steps ----
__global__ void set_entries_kernel(int64_t n, const int64_t *__restrict__ idx, const double *__restrict__ v, int64_t nDOF, double *x)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if(idx[i] >= 0 && idx[i] < nDOF) x[idx[i]] = v[i];
}

// solDot = sum_j c[j] hist_j (hist_0 = the current state)
__global__ void bdf_kernel(int64_t n, int nc, double c0, double c1, double c2, const double *__restrict__ u0, const double *__restrict__ u1,
                           const double *__restrict__ u2, double *__restrict__ dot)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double s = c0 * u0[i];
    if(nc > 1) s += c1 * u1[i];
    if(nc > 2) s += c2 * u2[i];
    dot[i] = s;
  }
}

int b200_set_essential(b200_system *s, int64_t n, const int64_t *dofs, const double *values)
{
  CHECK_S(s);
  if(!s->d_sol || n < 0 || (n > 0 && (!dofs || !values))) {
    set_error("b200_set_essential: bad arguments / no state");
    return B200_ERR_ARG;
  }
  if(n == 0) return B200_OK;
  if(s->ess_cap < n) {
    cudaFree(s->d_ess_idx);
    cudaFree(s->d_ess_val);
    B200_CUDA(cudaMalloc(&s->d_ess_idx, (size_t)n * sizeof(int64_t)));
    B200_CUDA(cudaMalloc(&s->d_ess_val, (size_t)n * sizeof(double)));
    s->ess_cap = n;
    s->ess_idx_host = nullptr;
  }
  // the index list is usually the same array at every time step: it crosses the bus once
  if(s->ess_idx_host != dofs || s->ess_n != n) {
    B200_CUDA(cudaMemcpyAsync(s->d_ess_idx, dofs, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, s->stream));
    s->ess_idx_host = dofs;
    s->ess_n        = n;
  }
  B200_CUDA(cudaMemcpyAsync(s->d_ess_val, values, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  set_entries_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, GRID), 256, 0, s->stream>>>(n, s->d_ess_idx, s->d_ess_val, s->nDOF, s->d_sol);
  count_launch();
  B200_CUDA(cudaStreamSynchronize(s->stream)); // the caller may reuse `values`
  return B200_OK;
}

int b200_state_push(b200_system *s)
{
  CHECK_S(s);
  if(!s->d_sol) {
    set_error("b200_state_push: no state");
    return B200_ERR_ARG;
  }
  const size_t nb = (size_t)s->nDOF * sizeof(double);
  for(int k = 0; k < 2; ++k)
    if(!s->d_hist[k]) {
      B200_CUDA(cudaMalloc(&s->d_hist[k], nb));
      B200_CUDA(cudaMemsetAsync(s->d_hist[k], 0, nb, s->stream));
    }
  std::swap(s->d_hist[0], s->d_hist[1]); // u_{n-1} <- u_n
  B200_CUDA(cudaMemcpyAsync(s->d_hist[0], s->d_sol, nb, cudaMemcpyDeviceToDevice, s->stream));
  // the state at time n of the time-averaged CHNS forms (the global solAtTimeN of src/feNonLinearSolver.cpp:60)
  if(!s->d_soln) B200_CUDA(cudaMalloc(&s->d_soln, nb));
  B200_CUDA(cudaMemcpyAsync(s->d_soln, s->d_sol, nb, cudaMemcpyDeviceToDevice, s->stream));
  s->have_soln = true;
  return B200_OK;
}

int b200_state_bdf(b200_system *s, int n_coef, const double *coef, double t, double dt)
{
  CHECK_S(s);
  if(!s->d_sol || n_coef < 1 || n_coef > 3 || !coef || (n_coef > 1 && !s->d_hist[0]) || (n_coef > 2 && !s->d_hist[1])) {
    set_error("b200_state_bdf: 1..3 coefficients, and b200_state_push once per past level");
    return B200_ERR_ARG;
  }
  bdf_kernel<<<GRID, 256, 0, s->stream>>>(s->nDOF, n_coef, coef[0], n_coef > 1 ? coef[1] : 0., n_coef > 2 ? coef[2] : 0., s->d_sol, s->d_hist[0],
                                         s->d_hist[1], s->d_soldot);
  count_launch();
  s->have_soldot = true;
  s->c0          = coef[0];
  s->t           = t;
  s->dt          = dt;
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int b200_set_to_zero(b200_system *s, int what)
{
  CHECK_S(s);
  if(!s->d_val) {
    set_error("b200_set_to_zero: set the pattern first");
    return B200_ERR_ARG;
  }
  if(what & 2) ++s->val_epoch;
  if(s->gather != nullptr && s->assembly_mode != B200_ASSEMBLY_SCATTER) {
    s->pending_zero |= (what & 3); // the gather kernels overwrite: memset only if something else touches the arrays
    return B200_OK;
  }
  if(what & 2) B200_CUDA(cudaMemsetAsync(s->d_val, 0, (size_t)s->nnz * sizeof(double), s->stream));
  if(what & 1) B200_CUDA(cudaMemsetAsync(s->d_rhs, 0, (size_t)s->nInc * sizeof(double), s->stream));
  s->pending_zero &= ~(what & 3);
  return B200_OK;
}

int b200_assemble(b200_system *s, int what, int only_transient)
{
  CHECK_S(s);
  B200_CUDA(cudaEventRecord(s->ev0, s->stream));
  if(what & 2) ++s->val_epoch;
  const int rc = launch_assemble(s, what, only_transient);
  if(rc != B200_OK) return rc;
  B200_CUDA(cudaEventRecord(s->ev1, s->stream));
  return B200_OK;
}

int b200_rhs_max_norm(b200_system *s, double *norm)
{
  CHECK_S(s);
  if(flush_zero(s, 1) != B200_OK) return B200_ERR_CUDA;
  return max_abs(s, s->d_rhs, s->nInc, norm);
}

int b200_du_max_norm(b200_system *s, double *norm)
{
  CHECK_S(s);
  return max_abs(s, s->d_du, s->nInc, norm);
}

int b200_constrain(b200_system *s)
{
  CHECK_S(s);
  if(flush_zero(s, 3) != B200_OK) return B200_ERR_CUDA;
  if(s->n_crow == 0) return B200_OK;
  ++s->val_epoch;
  constrain_kernel<<<GRID, 256, 0, s->stream>>>(s->nInc, s->d_ia, s->d_ja, s->d_val, s->d_rhs, s->d_cflag);
  count_launch();
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int b200_apply_periodicity(b200_system *s)
{
  CHECK_S(s);
  if(flush_zero(s, 3) != B200_OK) return B200_ERR_CUDA;
  if(s->n_per == 0) return B200_OK;
  ++s->val_epoch;
  B200_CUDA(cudaMemsetAsync(s->d_scratch, 0, sizeof(double), s->stream));
  periodic_kernel<<<(unsigned)std::min<int64_t>((s->n_per + 127) / 128, GRID), 128, 0, s->stream>>>(s->n_per, s->d_master, s->d_slave, s->nInc,
                                                                                                 s->d_ia, s->d_ja, s->d_val, s->d_rhs, s->d_scratch);
  count_launch();
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(s->h_scratch, s->d_scratch, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  B200_CUDA(cudaStreamSynchronize(s->stream));
  if(s->h_scratch[0] != 0.) {
    set_error("b200_apply_periodicity: the pattern has no (slave, master) entry -- give the periodic pairs (b200_set_periodic) before "
              "b200_set_pattern / b200_build_pattern");
    return B200_ERR_ARG;
  }
  return B200_OK;
}

int b200_solve(b200_system *s, const b200_solver_options *opt, b200_solve_info *info)
{
  CHECK_S(s);
  if(!opt || !info || !s->d_val) {
    set_error("b200_solve: bad arguments");
    return B200_ERR_ARG;
  }
  if(flush_zero(s, 3) != B200_OK) return B200_ERR_CUDA;
  B200_CUDA(cudaEventRecord(s->ev0, s->stream));
  const int rc = gmres_solve(s, opt, info);
  cudaEventRecord(s->ev1, s->stream);
  cudaEventSynchronize(s->ev1);
  cudaEventElapsedTime(&s->last_solve_ms, s->ev0, s->ev1);
  return rc;
}

int b200_correct_solution(b200_system *s, double *sol_host, int correct_dot)
{
  CHECK_S(s);
  double *target = correct_dot ? s->d_soldot : s->d_sol;
  add_kernel<<<GRID, 256, 0, s->stream>>>(s->nInc, s->d_du, target);
  count_launch();
  B200_CUDA(cudaGetLastError());
  if(sol_host) {
    // the essential entries (>= nInc) stay the caller's: only the unknowns are written back
    B200_CUDA(cudaMemcpyAsync(sol_host, target, (size_t)s->nInc * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    B200_CUDA(cudaStreamSynchronize(s->stream));
  }
  return B200_OK;
}

int b200_get_rhs(b200_system *s, double *rhs)
{
  CHECK_S(s);
  if(flush_zero(s, 1) != B200_OK) return B200_ERR_CUDA;
  B200_CUDA(cudaMemcpyAsync(rhs, s->d_rhs, (size_t)s->nInc * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  B200_CUDA(cudaStreamSynchronize(s->stream));
  return B200_OK;
}

int b200_axpy_rhs(b200_system *s, double coeff, const double *d)
{
  CHECK_S(s);
  if(flush_zero(s, 1) != B200_OK) return B200_ERR_CUDA;
  double *tmp = s->d_du; // du is rewritten by the next solve
  B200_CUDA(cudaMemcpyAsync(tmp, d, (size_t)s->nInc * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  axpy_kernel<<<GRID, 256, 0, s->stream>>>(s->nInc, coeff, tmp, s->d_rhs);
  count_launch();
  B200_CUDA(cudaStreamSynchronize(s->stream));
  return B200_OK;
}

int b200_get_matrix_values(b200_system *s, double *values)
{
  CHECK_S(s);
  if(flush_zero(s, 2) != B200_OK) return B200_ERR_CUDA;
  B200_CUDA(cudaMemcpyAsync(values, s->d_val, (size_t)s->nnz * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  B200_CUDA(cudaStreamSynchronize(s->stream));
  return B200_OK;
}

int b200_get_du(b200_system *s, double *du)
{
  CHECK_S(s);
  B200_CUDA(cudaMemcpyAsync(du, s->d_du, (size_t)s->nInc * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  B200_CUDA(cudaStreamSynchronize(s->stream));
  return B200_OK;
}

int b200_get_solution(b200_system *s, double *sol)
{
  CHECK_S(s);
  B200_CUDA(cudaMemcpyAsync(sol, s->d_sol, (size_t)s->nDOF * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  B200_CUDA(cudaStreamSynchronize(s->stream));
  return B200_OK;
}

int b200_spmv(b200_system *s, const double *x, double *y)
{
  CHECK_S(s);
  if(flush_zero(s, 2) != B200_OK) return B200_ERR_CUDA;
  double *dx = nullptr, *dy = nullptr;
  B200_CUDA(cudaMalloc(&dx, (size_t)s->nInc * sizeof(double)));
  B200_CUDA(cudaMalloc(&dy, (size_t)s->nInc * sizeof(double)));
  B200_CUDA(cudaMemcpyAsync(dx, x, (size_t)s->nInc * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  int rc = spmv(s, dx, dy);
  if(rc == B200_OK) {
    cudaMemcpyAsync(y, dy, (size_t)s->nInc * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    if(cudaStreamSynchronize(s->stream) != cudaSuccess) {
      set_error("b200_spmv: stream sync failed");
      rc = B200_ERR_CUDA;
    }
  }
  cudaFree(dx);
  cudaFree(dy);
  return rc;
}

int b200_last_assemble_ms(const b200_system *s, float *ms)
{
  if(!s || !ms) return B200_ERR_ARG;
  if(cudaEventSynchronize(s->ev1) != cudaSuccess) return B200_ERR_CUDA;
  if(cudaEventElapsedTime(ms, s->ev0, s->ev1) != cudaSuccess) return B200_ERR_CUDA;
  return B200_OK;
}

int b200_last_solve_ms(const b200_system *s, float *ms)
{
  if(!s || !ms) return B200_ERR_ARG;
  *ms = s->last_solve_ms;
  return B200_OK;
}

int b200_time_spmv(b200_system *s, int reps, float *ms_per_spmv)
{
  CHECK_S(s);
  if(flush_zero(s, 2) != B200_OK) return B200_ERR_CUDA;
  if(reps <= 0 || !s->d_val) {
    set_error("b200_time_spmv: bad arguments");
    return B200_ERR_ARG;
  }
  double *dx = nullptr, *dy = nullptr;
  B200_CUDA(cudaMalloc(&dx, (size_t)s->nInc * sizeof(double)));
  B200_CUDA(cudaMalloc(&dy, (size_t)s->nInc * sizeof(double)));
  fill_kernel<<<GRID, 256, 0, s->stream>>>(s->nInc, dx, 1.0);
  count_launch();
  int rc = B200_OK;
  // multi-GPU: the product includes the halo update of its input (the exchange step of the path)
  for(int i = 0; i < 3 && rc == B200_OK; ++i) {
    rc = comm_halo_exchange(s, dx);
    if(rc == B200_OK) rc = spmv(s, dx, dy);
  }
  cudaEventRecord(s->ev0, s->stream);
  for(int i = 0; i < reps && rc == B200_OK; ++i) rc = comm_spmv_overlapped(s, dx, dy);
  cudaEventRecord(s->ev1, s->stream);
  cudaEventSynchronize(s->ev1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, s->ev0, s->ev1);
  *ms_per_spmv = ms / reps;
  cudaFree(dx);
  cudaFree(dy);
  return rc;
}

int b200_time_begin(b200_system *s)
{
  CHECK_S(s);
  if(!s->ev2) {
    B200_CUDA(cudaEventCreate(&s->ev2));
    B200_CUDA(cudaEventCreate(&s->ev3));
  }
  B200_CUDA(cudaEventRecord(s->ev2, s->stream));
  return B200_OK;
}

int b200_time_end(b200_system *s, float *ms)
{
  CHECK_S(s);
  if(!s->ev2 || !ms) {
    set_error("b200_time_end: call b200_time_begin first");
    return B200_ERR_ARG;
  }
  B200_CUDA(cudaEventRecord(s->ev3, s->stream));
  B200_CUDA(cudaEventSynchronize(s->ev3));
  B200_CUDA(cudaEventElapsedTime(ms, s->ev2, s->ev3));
  return B200_OK;
}

int b200_measure_fp64_peak(int device, double *tflops)
{
  if(!tflops || cudaSetDevice(device) != cudaSuccess) {
    set_error("b200_measure_fp64_peak: bad device");
    return B200_ERR_CUDA;
  }
  double *d = nullptr;
  B200_CUDA(cudaMalloc(&d, 148 * 8 * 256 * sizeof(double)));
  cudaEvent_t e0, e1;
  B200_CUDA(cudaEventCreate(&e0));
  B200_CUDA(cudaEventCreate(&e1));
  const int iters = 4096;
  dfma_peak_kernel<<<148 * 8, 256>>>(d, iters, 1.000001);
  float best = 1e30f;
  for(int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    dfma_peak_kernel<<<148 * 8, 256>>>(d, iters, 1.000001);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  count_launch(6);
  const double flops = 2.0 * 16.0 * iters * 148.0 * 8.0 * 256.0;
  *tflops = flops / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int b200_measure_dmma_peak(int device, double *tflops)
{
  if(!tflops || cudaSetDevice(device) != cudaSuccess) {
    set_error("b200_measure_dmma_peak: bad device");
    return B200_ERR_CUDA;
  }
  double *d = nullptr;
  B200_CUDA(cudaMalloc(&d, 148 * 8 * 256 * sizeof(double)));
  cudaEvent_t e0, e1;
  B200_CUDA(cudaEventCreate(&e0));
  B200_CUDA(cudaEventCreate(&e1));
  const int iters = 2048;
  dmma_peak_kernel<<<148 * 8, 256>>>(d, iters, 1.000001);
  float best = 1e30f;
  for(int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    dmma_peak_kernel<<<148 * 8, 256>>>(d, iters, 1.000001);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  count_launch(6);
  // one m8n8k4 MMA = 8*8*4 FMA = 512 flop per warp
  const double flops = 512.0 * 8.0 * iters * 148.0 * 8.0 * (256.0 / 32.0);
  *tflops = flops / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int b200_sync(b200_system *s)
{
  CHECK_S(s);
  B200_CUDA(cudaStreamSynchronize(s->stream));
  return B200_OK;
}

} // extern "C"
