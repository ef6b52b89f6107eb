// Multigrid-based preconditioners of the restarted GMRES (north-star subsystem 4).
//
//   B200_PC_AMG        z = V-cycle(r) on the whole matrix (scalar problems: feSysElm_Diffusion / TransientMass systems)
//   B200_PC_SCHUR_AMG  Taylor-Hood saddle-point systems  A = [F G; D 0]  (velocity rows U, pressure rows P; the P-P block is
//                      structurally a forced diagonal with value zero, src/feCompressedRowStorage.cpp:33).  Block upper
//                      triangular preconditioner
//                          z_p = S^-1 r_p,   z_u = F^-1 (r_u - G z_p)
//                      with F^-1 ~ one V-cycle of the velocity hierarchy (amg.cu) and the Schur complement S = -D F^-1 G
//                      replaced by the scaled diagonal of the pressure mass matrix: with F ~ f K (K the vector Laplacian,
//                      f = VectorDiffusion coeff*k - DivergenceNewtonianStress coeff*mu), G = g B^T (g = stress coeff - MixedGradient
//                      coeff), D = d B (d = MixedDivergence coeff) the inf-sup property gives S ~ -(d g / f_S) M_p, where the
//                      stress form enters f_S twice (its symbol is mu (|k|^2 I + k k^T): k^T F^-1 k = 1 / (2 mu)).
//                      When a single pressure unknown is pinned (the reference's `PointPression` essential space,
//                      tests/withLinearSolver/navier_stokes.cpp:63-81) the constant pressure mode survives in S with an O(h^dim)
//                      eigenvalue; it is treated by a rank-one term: z_p += beta sum(r_p).
//
// What the reference does at this point: PETSc's default KSP preconditioner, ILU(0) on one rank / block-Jacobi on n ranks
// (src/feLinearSystem.h:196-199, src/feLinearSystemPETSc.cpp:337), or a sparse direct solve
// (src/feLinearSystemMklPardiso.cpp:893-965).  On several GPUs the velocity hierarchy is rank-local (block-Jacobi across
// ranks); the pressure part needs one halo update of z_p and one all-reduce of sum(r_p) per application.
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/extrema.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/scan.h>
#include <thrust/transform_reduce.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "amg.h"

namespace b200 {

static const int GRID = 148 * 8;

struct Precond {
  int      kind = 0;
  uint8_t *d_fld = nullptr, *d_fld_all = nullptr; // field id per row: ghost rows masked out / still labelled
  Amg      amg;
  // Schur part
  double  *d_alpha = nullptr;  // [n] S^-1 scaling of the pressure rows (0 elsewhere)
  double  *d_pmass = nullptr;  // [n] diagonal of the pressure mass matrix at the pressure rows
  int64_t *d_pstart = nullptr; // [n] first entry of row i whose column is a pressure unknown (velocity rows), or -1: no table
  double   beta = 0.;          // rank-one coefficient of the pinned-pressure mode
  double  *d_sum = nullptr;
  double  *d_zp = nullptr, *d_b = nullptr;
  bool     tail = false;       // pressure columns are the trailing entries of every velocity row
  int64_t  pattern_nnz = -1;
  const void *pattern_ia = nullptr;
  uint64_t numeric_epoch = ~0ull;
  double   schur_scale = 0.;
  int64_t  pmin = 0, umax = 0; // first pressure unknown / last velocity unknown of the local numbering (ghost rows included)
  // compact single-precision copy of the block G = A[velocity rows, pressure columns] (the product G z_p of every application)
  int64_t *d_iag = nullptr;
  int32_t *d_jag = nullptr;
  float   *d_valg = nullptr;
  int64_t  nnz_g = 0;
};

__global__ void pc_pmass_kernel(int64_t nElm, int dim, const double *__restrict__ xyz, const int32_t *__restrict__ conn,
                                const int32_t *__restrict__ adrP, int nP, const double *__restrict__ mloc, int64_t nDOF, double *pm)
{
  for(int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nElm; e += (int64_t)gridDim.x * blockDim.x) {
    const int nv = dim + 1;
    double    X[4][3];
    for(int v = 0; v < nv; ++v)
      for(int m = 0; m < dim; ++m) X[v][m] = xyz[(int64_t)conn[e * nv + v] * dim + m];
    double J;
    if(dim == 2)
      J = (X[1][0] - X[0][0]) * (X[2][1] - X[0][1]) - (X[2][0] - X[0][0]) * (X[1][1] - X[0][1]);
    else {
      const double a0 = X[1][0] - X[0][0], a1 = X[1][1] - X[0][1], a2 = X[1][2] - X[0][2];
      const double b0 = X[2][0] - X[0][0], b1 = X[2][1] - X[0][1], b2 = X[2][2] - X[0][2];
      const double c0 = X[3][0] - X[0][0], c1 = X[3][1] - X[0][1], c2 = X[3][2] - X[0][2];
      J = a0 * (b1 * c2 - b2 * c1) - a1 * (b0 * c2 - b2 * c0) + a2 * (b0 * c1 - b1 * c0);
    }
    J = fabs(J);
    for(int i = 0; i < nP; ++i) {
      const int64_t d = adrP[e * nP + i];
      if(d < nDOF) atomicAdd(pm + d, J * mloc[i]);
    }
  }
}

// alpha[i] = scale / pmass[i] on owned pressure rows, 0 elsewhere
__global__ void pc_alpha_kernel(int64_t n, const uint8_t *__restrict__ fld, const double *__restrict__ pm, double scale, double *alpha)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    alpha[i] = (fld[i] == AMG_FLD_P && pm[i] > 0.) ? scale / pm[i] : 0.;
}

__global__ void pc_psum_kernel(int64_t n, const double *__restrict__ alpha, const double *__restrict__ r, double *sum)
{
  double s = 0.;
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if(alpha[i] != 0.) s += r[i];
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if((threadIdx.x & 31) == 0 && s != 0.) atomicAdd(sum, s);
}

// zp = alpha r + beta sum on pressure rows, 0 elsewhere
__global__ void pc_zp_kernel(int64_t n, const double *__restrict__ alpha, const double *__restrict__ r, double beta, const double *sum,
                             double *__restrict__ zp)
{
  const double bs = beta != 0. ? beta * *sum : 0.;
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    zp[i] = alpha[i] != 0. ? alpha[i] * r[i] + bs : 0.;
}

__global__ void pc_pstart_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja, const uint8_t *__restrict__ fld,
                                 int64_t pmin, int64_t *pstart)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = ia[i], hi = ia[i + 1];
    if(fld[i] >= AMG_FLD_P) { // not a velocity row
      pstart[i] = hi;
      continue;
    }
    while(lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if(ja[mid] < pmin)
        lo = mid + 1;
      else
        hi = mid;
    }
    pstart[i] = lo;
  }
}

// b = r - G zp on the rows of the velocity field (tail = entries pstart[i] .. ia[i+1]); other rows: b = 0
template <int LPR>
__global__ void __launch_bounds__(256) pc_rhs_tail_kernel(int64_t n, const int64_t *__restrict__ ia, const int64_t *__restrict__ pstart,
                                                          const int32_t *__restrict__ ja, const double *__restrict__ val,
                                                          const uint8_t *__restrict__ fld, const double *__restrict__ r,
                                                          const double *__restrict__ zp, double *__restrict__ b)
{
  const int     lane = threadIdx.x % LPR;
  const int64_t g0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR, ng = (gridDim.x * (int64_t)blockDim.x) / LPR;
  for(int64_t i = g0; i < n; i += ng) {
    double s = 0.;
    if(fld[i] < AMG_FLD_P)
      for(int64_t k = pstart[i] + lane; k < ia[i + 1]; k += LPR) s += val[k] * zp[ja[k]];
#pragma unroll
    for(int o = LPR / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, LPR);
    if(lane == 0) b[i] = fld[i] < AMG_FLD_P ? r[i] - s : 0.;
  }
}

// compact single-precision copy of the velocity-row x pressure-column block G (any numbering): count / fill like the level-0 copy of
// amg.cu, one warp per row
__global__ void __launch_bounds__(256) pc_g_count_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja,
                                                         const uint8_t *__restrict__ fld, const uint8_t *__restrict__ fld_col, int64_t *cnt)
{
  const int     lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * (int64_t)blockDim.x) >> 5;
  for(int64_t i = w0; i < n; i += nw) {
    int c = 0;
    if(fld[i] < AMG_FLD_P)
      for(int64_t k = ia[i] + lane; k < ia[i + 1]; k += 32) c += fld_col[ja[k]] == AMG_FLD_P ? 1 : 0;
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if(lane == 0) cnt[i] = c;
  }
}

template <bool WITH_JA>
__global__ void __launch_bounds__(256) pc_g_fill_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja,
                                                        const double *__restrict__ val, const uint8_t *__restrict__ fld,
                                                        const uint8_t *__restrict__ fld_col, const int64_t *__restrict__ iag, int32_t *jag,
                                                        float *valg)
{
  const int     lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * (int64_t)blockDim.x) >> 5;
  for(int64_t i = w0; i < n; i += nw) {
    if(fld[i] >= AMG_FLD_P) continue;
    int64_t       out = iag[i];
    const int64_t beg = ia[i], end = ia[i + 1];
    for(int64_t k0 = beg; k0 < end; k0 += 32) {
      const int64_t  k = k0 + lane;
      const int32_t  j = k < end ? ja[k] : 0;
      const bool     keep = k < end && fld_col[j] == AMG_FLD_P;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if(keep) {
        const int64_t o = out + __popc(m & ((1u << lane) - 1u));
        if(WITH_JA) jag[o] = j;
        valg[o] = (float)val[k];
      }
      out += __popc(m);
    }
  }
}

// b = r - G zp on velocity rows (compact G), 0 elsewhere
template <int LPR>
__global__ void __launch_bounds__(256) pc_rhs_g_kernel(int64_t n, const int64_t *__restrict__ iag, const int32_t *__restrict__ jag,
                                                       const float *__restrict__ valg, const uint8_t *__restrict__ fld,
                                                       const double *__restrict__ r, const double *__restrict__ zp, double *__restrict__ b)
{
  const int     lane = threadIdx.x % LPR;
  const int64_t g0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR, ng = (gridDim.x * (int64_t)blockDim.x) / LPR;
  for(int64_t i = g0; i < n; i += ng) {
    double s = 0.;
    for(int64_t k = iag[i] + lane; k < iag[i + 1]; k += LPR) s += (double)valg[k] * zp[jag[k]];
#pragma unroll
    for(int o = LPR / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, LPR);
    if(lane == 0) b[i] = fld[i] < AMG_FLD_P ? r[i] - s : 0.;
  }
}

// b = r - t on velocity rows, 0 elsewhere (general numbering: t = A [0; zp] from a full product)
__global__ void pc_rhs_full_kernel(int64_t n, const uint8_t *__restrict__ fld, const double *__restrict__ r, const double *__restrict__ t,
                                   double *__restrict__ b)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    b[i] = fld[i] < AMG_FLD_P ? r[i] - t[i] : 0.;
}

// z = x on velocity rows (x = V-cycle output, 0 on inactive rows), zp on pressure rows
__global__ void pc_combine_kernel(int64_t n, const uint8_t *__restrict__ fld, const double *__restrict__ zp, double *__restrict__ z)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if(fld[i] >= AMG_FLD_P) z[i] = zp[i];
}

void precond_free(System *S)
{
  Precond *P = static_cast<Precond *>(S->precond);
  if(!P) return;
  amg_free(&P->amg);
  cudaFree(P->d_fld);
  cudaFree(P->d_fld_all);
  cudaFree(P->d_alpha);
  cudaFree(P->d_pmass);
  cudaFree(P->d_pstart);
  cudaFree(P->d_iag);
  cudaFree(P->d_jag);
  cudaFree(P->d_valg);
  cudaFree(P->d_sum);
  cudaFree(P->d_zp);
  cudaFree(P->d_b);
  delete P;
  S->precond = nullptr;
}

bool precond_fallback(System *S)
{
  Precond *P = static_cast<Precond *>(S->precond);
  if(!P || P->amg.cheb_degree <= 1) return false;
  P->amg.cheb_degree = 1; // sticky: the following Newton iterations of this system start in the safe mode
  if(P->amg.G) P->amg.G->cheb_degree = 1; // the replicated global levels (several GPUs) follow
  if(P->amg.verbose) fprintf(stderr, "[feng_b200] GMRES stagnates: multigrid smoother degraded to damped Jacobi\n");
  return true;
}

struct FldIs {
  const uint8_t *f;
  int            lo, hi;
  int64_t        miss;
  bool           want_min;
  __host__ __device__ int64_t operator()(int64_t i) const { return (f[i] >= lo && f[i] < hi) ? i : miss; }
};
struct MinOp {
  __host__ __device__ int64_t operator()(int64_t a, int64_t b) const { return a < b ? a : b; }
};
struct MaxOp {
  __host__ __device__ int64_t operator()(int64_t a, int64_t b) const { return a > b ? a : b; }
};

static int schur_symbolic(System *S, Precond *P)
{
  const int64_t n = S->nInc;
  const Space  &PS = S->spaces[S->sp];
  // diagonal of the pressure mass matrix from the pressure tabulation: m_i = |J| sum_k w_k psi_i(k)^2
  std::vector<double> mloc(PS.nS, 0.);
  for(int k = 0; k < S->nq; ++k)
    for(int i = 0; i < PS.nS; ++i) mloc[i] += S->w[k] * PS.L[(size_t)k * PS.nS + i] * PS.L[(size_t)k * PS.nS + i];
  double *d_mloc = nullptr, *pm_all = nullptr;
  B200_CUDA(cudaMalloc(&d_mloc, mloc.size() * sizeof(double)));
  B200_CUDA(cudaMemcpyAsync(d_mloc, mloc.data(), mloc.size() * sizeof(double), cudaMemcpyHostToDevice, S->stream));
  B200_CUDA(cudaMalloc(&pm_all, (size_t)S->nDOF * sizeof(double)));
  B200_CUDA(cudaMemsetAsync(pm_all, 0, (size_t)S->nDOF * sizeof(double), S->stream));
  pc_pmass_kernel<<<GRID, 128, 0, S->stream>>>(S->nElm, S->dim, S->d_xyz, S->d_conn, PS.d_adr, PS.nS, d_mloc, S->nDOF, pm_all);
  count_launch();
  B200_CUDA(cudaMalloc(&P->d_pmass, (size_t)n * sizeof(double)));
  B200_CUDA(cudaMemcpyAsync(P->d_pmass, pm_all, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, S->stream));
  // essential pressure DOFs (>= nInc): exactly one (over all ranks) = pinned enclosed-flow pressure -> rank-one term
  std::vector<double> tail((size_t)(S->nDOF - n));
  if(!tail.empty()) B200_CUDA(cudaMemcpyAsync(tail.data(), pm_all + n, tail.size() * sizeof(double), cudaMemcpyDeviceToHost, S->stream));
  B200_CUDA(cudaStreamSynchronize(S->stream));
  int    npin = 0;
  double mpin = 0.;
  for(double v : tail)
    if(v > 0.) {
      ++npin;
      mpin += v;
    }
  // on several GPUs an essential pressure DOF of the ghost layer shows up on more than one rank with a partial mass: the
  // decision needs the global picture
  if(comm_active(S)) {
    double h[2] = {(double)npin, mpin};
    // ghost copies would be double counted: only count a pinned DOF whose full element star is local, i.e. take the MAX of
    // the counts and of the masses (the owner sees the whole star)
    B200_CUDA(cudaMemcpyAsync(S->d_scratch, h, 2 * sizeof(double), cudaMemcpyHostToDevice, S->stream));
    int rc = comm_allreduce(S, S->d_scratch, 2, true);
    if(rc != B200_OK) return rc;
    B200_CUDA(cudaMemcpyAsync(h, S->d_scratch, 2 * sizeof(double), cudaMemcpyDeviceToHost, S->stream));
    B200_CUDA(cudaStreamSynchronize(S->stream));
    npin = (int)h[0];
    mpin = h[1];
  }
  P->beta = 0.;
  if(npin == 1 && mpin > 0.) P->beta = 1. / mpin; // scaled by schur_scale in the numeric phase
  cudaFree(pm_all);
  cudaFree(d_mloc);
  B200_CUDA(cudaMalloc(&P->d_alpha, (size_t)n * sizeof(double)));
  B200_CUDA(cudaMalloc(&P->d_sum, sizeof(double)));
  B200_CUDA(cudaMalloc(&P->d_zp, (size_t)n * sizeof(double)));
  B200_CUDA(cudaMalloc(&P->d_b, (size_t)n * sizeof(double)));
  // pressure unknowns numbered after every velocity unknown (the reference numbers field by field, src/feNumber.cpp:370-483)?
  // then G z_p only touches the tail of the velocity rows
  P->tail = false;
  if(P->pmin > P->umax && P->pmin < n) {
    B200_CUDA(cudaMalloc(&P->d_pstart, (size_t)n * sizeof(int64_t)));
    pc_pstart_kernel<<<GRID, 256, 0, S->stream>>>(n, S->d_ia, S->d_ja, P->d_fld, P->pmin, P->d_pstart);
    count_launch();
    P->tail = true;
  }
  if(!getenv("B200_PC_NO_G")) {
    // pattern of the compact G block; columns are labelled by the UNMASKED field map (ghost pressure columns count)
    const uint8_t *fcol = P->d_fld_all ? P->d_fld_all : P->d_fld;
    int64_t       *cnt = nullptr;
    B200_CUDA(cudaMalloc(&cnt, (size_t)(n + 1) * sizeof(int64_t)));
    B200_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(n + 1) * sizeof(int64_t), S->stream));
    pc_g_count_kernel<<<GRID * 4, 256, 0, S->stream>>>(n, S->d_ia, S->d_ja, P->d_fld, fcol, cnt);
    count_launch();
    thrust::device_ptr<int64_t> cp(cnt);
    thrust::exclusive_scan(thrust::cuda::par.on(S->stream), cp, cp + n + 1, cp);
    B200_CUDA(cudaMemcpyAsync(&P->nnz_g, cnt + n, sizeof(int64_t), cudaMemcpyDeviceToHost, S->stream));
    B200_CUDA(cudaStreamSynchronize(S->stream));
    P->d_iag = cnt;
    B200_CUDA(cudaMalloc(&P->d_jag, (size_t)std::max<int64_t>(P->nnz_g, 1) * sizeof(int32_t)));
    B200_CUDA(cudaMalloc(&P->d_valg, (size_t)std::max<int64_t>(P->nnz_g, 1) * sizeof(float)));
    pc_g_fill_kernel<true><<<GRID * 4, 256, 0, S->stream>>>(n, S->d_ia, S->d_ja, S->d_val, P->d_fld, fcol, P->d_iag, P->d_jag, P->d_valg);
    count_launch();
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int precond_setup(System *S, int pc)
{
  Precond *P = static_cast<Precond *>(S->precond);
  if(P && (P->kind != pc || P->pattern_nnz != S->nnz || P->pattern_ia != (const void *)S->d_ia)) {
    precond_free(S);
    P = nullptr;
  }
  if(pc == B200_PC_SCHUR_AMG && (S->plan != PLAN_TAYLOR_HOOD || S->sp < 0)) {
    set_error("B200_PC_SCHUR_AMG needs a Taylor-Hood velocity/pressure system");
    return B200_ERR_UNSUPP;
  }
  if(pc == B200_PC_AMG && S->plan == PLAN_CHNS) {
    set_error("B200_PC_AMG: not available for the monolithic CHNS system");
    return B200_ERR_UNSUPP;
  }
  if(!P) {
    P = new Precond;
    P->kind = pc;
    P->pattern_nnz = S->nnz;
    P->pattern_ia = S->d_ia;
    S->precond = P;
    int rc = amg_build_field_map(S, &P->d_fld);
    if(rc != B200_OK) return rc;
    {
      auto                               pol = thrust::cuda::par.on(S->stream);
      const int64_t                      n = S->nInc;
      thrust::counting_iterator<int64_t> c0(0);
      P->pmin = thrust::transform_reduce(pol, c0, c0 + n, FldIs{P->d_fld, AMG_FLD_P, AMG_FLD_P + 1, n, true}, n, MinOp());
      P->umax = thrust::transform_reduce(pol, c0, c0 + n, FldIs{P->d_fld, 0, AMG_FLD_P, -1, false}, (int64_t)-1, MaxOp());
    }
    if(comm_active(S)) {
      B200_CUDA(cudaMalloc(&P->d_fld_all, (size_t)S->nInc));
      B200_CUDA(cudaMemcpyAsync(P->d_fld_all, P->d_fld, (size_t)S->nInc, cudaMemcpyDeviceToDevice, S->stream));
    }
    rc = amg_mask_ghosts(S, P->d_fld);
    if(rc != B200_OK) return rc;
    // B200_PC_AMG on a Taylor-Hood system would treat the pressure rows as part of the elliptic field: refuse
    const int fld_hi = pc == B200_PC_SCHUR_AMG ? S->dim : S->spaces[S->su].nc;
    rc = amg_setup_symbolic(S, &P->amg, P->d_fld, 0, fld_hi, S->su, P->d_fld_all);
    if(rc != B200_OK) return rc;
    if(pc == B200_PC_SCHUR_AMG) {
      rc = schur_symbolic(S, P);
      if(rc != B200_OK) return rc;
    }
    P->numeric_epoch = ~0ull;
  }
  if(P->numeric_epoch != S->val_epoch) {
    int rc = amg_setup_numeric(S, &P->amg);
    if(rc != B200_OK) return rc;
    if(pc == B200_PC_SCHUR_AMG) {
      const THCoeffs &c = S->th;
      const double    f = c.diff_k - c.sig_mu, g = c.c_sig - c.c_gradp, d = c.c_div;
      if(f == 0. || g == 0. || d == 0.) {
        set_error("B200_PC_SCHUR_AMG: the system has no viscous / pressure-gradient / divergence form");
        return B200_ERR_UNSUPP;
      }
      // S^-1 ~ -(f_S / (d g)) M_p^-1.  The stress-divergence form counts twice in f_S: for F = mu (|k|^2 I + k k^T) (Fourier symbol of
      // -div(mu (grad u + grad u^T))) k^T F^-1 k = 1 / (2 mu), against 1 / mu for the Laplacian form.  Measured at T3D(92), divergence
      // form: factor 1 -> 103 iterations, 1.5 -> 98, 2 -> 92, 2.5 -> 90, 4 -> 91, 0.7 -> 113 (B200_PC_SCHUR_SCALE multiplies on top).
      const double f_schur = c.diff_k - 2. * c.sig_mu;
      P->schur_scale = -f_schur / (d * g);
      if(const char *e = getenv("B200_PC_SCHUR_SCALE")) P->schur_scale *= atof(e);
      pc_alpha_kernel<<<GRID, 256, 0, S->stream>>>(S->nInc, P->d_fld, P->d_pmass, P->schur_scale, P->d_alpha);
      count_launch();
      if(P->d_valg) {
        pc_g_fill_kernel<false><<<GRID * 4, 256, 0, S->stream>>>(S->nInc, S->d_ia, S->d_ja, S->d_val, P->d_fld,
                                                                 P->d_fld_all ? P->d_fld_all : P->d_fld, P->d_iag, nullptr, P->d_valg);
        count_launch();
      }
    }
    P->numeric_epoch = S->val_epoch;
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int precond_apply(System *S, int pc, const double *r, double *z)
{
  Precond *P = static_cast<Precond *>(S->precond);
  if(!P) {
    set_error("precond_apply: no set-up");
    return B200_ERR_ARG;
  }
  const int64_t n = S->nInc;
  if(pc == B200_PC_AMG) return amg_vcycle(S, &P->amg, r, z);
  // pressure: z_p = alpha r_p + beta sum(r_p)
  const double beta = P->beta * P->schur_scale;
  if(beta != 0.) {
    B200_CUDA(cudaMemsetAsync(P->d_sum, 0, sizeof(double), S->stream));
    pc_psum_kernel<<<GRID, 256, 0, S->stream>>>(n, P->d_alpha, r, P->d_sum);
    count_launch();
    int rc = comm_allreduce(S, P->d_sum, 1, false);
    if(rc != B200_OK) return rc;
  }
  pc_zp_kernel<<<GRID, 256, 0, S->stream>>>(n, P->d_alpha, r, beta, P->d_sum, P->d_zp);
  count_launch();
  int rc = comm_halo_exchange(S, P->d_zp);
  if(rc != B200_OK) return rc;
  // velocity right-hand side b = r_u - G z_p
  if(P->d_valg) {
    const int64_t blocks = (n * 4 + 255) / 256;
    pc_rhs_g_kernel<4><<<(unsigned)std::min<int64_t>(blocks, 148 * 32), 256, 0, S->stream>>>(n, P->d_iag, P->d_jag, P->d_valg, P->d_fld, r, P->d_zp,
                                                                                            P->d_b);
    count_launch();
  } else if(P->tail) {
    const int64_t blocks = (n * 4 + 255) / 256;
    pc_rhs_tail_kernel<4><<<(unsigned)std::min<int64_t>(blocks, 148 * 32), 256, 0, S->stream>>>(n, S->d_ia, P->d_pstart, S->d_ja, S->d_val, P->d_fld,
                                                                                               r, P->d_zp, P->d_b);
    count_launch();
  } else {
    rc = spmv(S, P->d_zp, z); // z as scratch: A [0; zp]
    if(rc != B200_OK) return rc;
    pc_rhs_full_kernel<<<GRID, 256, 0, S->stream>>>(n, P->d_fld, r, z, P->d_b);
    count_launch();
  }
  rc = amg_vcycle(S, &P->amg, P->d_b, z);
  if(rc != B200_OK) return rc;
  pc_combine_kernel<<<GRID, 256, 0, S->stream>>>(n, P->d_fld, P->d_zp, z);
  count_launch();
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

} // namespace b200
