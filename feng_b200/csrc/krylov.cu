// CSR SpMV and left-preconditioned restarted GMRES(m) on the device.
//
// Replaces the linear solve of the reference backends (feLinearSystemPETSc::solve, KSP GMRES(30) with rtol on the
// preconditioned residual, src/feLinearSystemPETSc.cpp:900-1031; Pardiso LU, src/feLinearSystemMklPardiso.cpp:893-965)
// and reports the same four numbers back to solveNewtonRaphson (src/feNonLinearSolver.cpp:98).
//
// Orthogonalisation: classical Gram-Schmidt applied twice (CGS2).  Each pass is two kernels
// (all k+1 projections in one batched reduction, then one fused multi-axpy), so an iteration costs one SpMV,
// one preconditioner application, two batched dot kernels, two multi-axpys and ONE host synchronisation for the
// (k+2) Hessenberg entries; the Givens rotations run on the host.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "amg.h"
#include "system.h"

namespace b200 {

// ----------------------------------------------------------------------------------------------------------
// kernels
// ----------------------------------------------------------------------------------------------------------
template <int LPR> // lanes per row
__global__ void __launch_bounds__(256) spmv_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja,
                                                   const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y)
{
  const int     lane  = threadIdx.x % LPR;
  const int64_t row0  = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR;
  const int64_t nrows = (gridDim.x * (int64_t)blockDim.x) / LPR;
  for(int64_t row = row0; row < n; row += nrows) {
    const int64_t beg = ia[row], end = ia[row + 1];
    double        s = 0.;
    for(int64_t k = beg + lane; k < end; k += LPR) s += val[k] * x[ja[k]];
#pragma unroll
    for(int o = LPR / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, LPR);
    if(lane == 0) y[row] = s;
  }
}

// Same product with RPG rows per lane group in flight: the value / column streams are read with streaming loads
// (ld.global.cs, no reuse) and every lane keeps RPG independent load chains (val, ja -> x[ja]) outstanding, which is what
// the r01a profile asked for (latency-bound at 39 % of DRAM peak with one chain per lane).
template <int LPR, int RPG, bool CS>
__global__ void __launch_bounds__(256) spmv_rpg_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja,
                                                       const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y,
                                                       const uint8_t *__restrict__ skip = nullptr)
{
  const int     lane = threadIdx.x % LPR;
  const int64_t g0   = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR;
  const int64_t ng   = (gridDim.x * (int64_t)blockDim.x) / LPR;
  for(int64_t row0 = g0 * RPG; row0 < n; row0 += ng * RPG) {
    int64_t beg[RPG], end[RPG];
#pragma unroll
    for(int r = 0; r < RPG; ++r) {
      const int64_t row = row0 + r;
      const bool    on  = row < n && !(skip && skip[row]);
      beg[r] = on ? ia[row] : 0;
      end[r] = on ? ia[row + 1] : 0;
    }
    double s[RPG];
#pragma unroll
    for(int r = 0; r < RPG; ++r) s[r] = 0.;
    int64_t longest = 0;
#pragma unroll
    for(int r = 0; r < RPG; ++r) longest = max(longest, end[r] - beg[r]);
    for(int64_t off = lane; off < longest; off += LPR) {
      double  v[RPG];
      int32_t c[RPG];
#pragma unroll
      for(int r = 0; r < RPG; ++r) {
        const int64_t k  = beg[r] + off;
        const bool    ok = k < end[r];
        v[r] = ok ? (CS ? __ldcs(val + k) : val[k]) : 0.;
        c[r] = ok ? (CS ? __ldcs(ja + k) : ja[k]) : 0;
      }
#pragma unroll
      for(int r = 0; r < RPG; ++r) s[r] += v[r] * x[c[r]];
    }
#pragma unroll
    for(int r = 0; r < RPG; ++r) {
#pragma unroll
      for(int o = LPR / 2; o > 0; o >>= 1) s[r] += __shfl_down_sync(0xffffffffu, s[r], o, LPR);
    }
    if(lane == 0) {
#pragma unroll
      for(int r = 0; r < RPG; ++r)
        if(row0 + r < n && !(skip && skip[row0 + r])) y[row0 + r] = s[r];
    }
  }
}

// the same product restricted to a row subset: rows == nullptr: every row with skip[row] == 0; else the listed rows
template <int LPR>
__global__ void __launch_bounds__(256) spmv_subset_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja,
                                                          const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y,
                                                          const uint8_t *__restrict__ skip, const int32_t *__restrict__ rows, int64_t n_rows)
{
  const int     lane = threadIdx.x % LPR;
  const int64_t g0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR, ng = (gridDim.x * (int64_t)blockDim.x) / LPR;
  const int64_t cnt = rows ? n_rows : n;
  for(int64_t t = g0; t < cnt; t += ng) {
    const int64_t row = rows ? (int64_t)rows[t] : t;
    if(skip && skip[row]) continue;
    double s = 0.;
    for(int64_t k = ia[row] + lane; k < ia[row + 1]; k += LPR) s += val[k] * x[ja[k]];
#pragma unroll
    for(int o = LPR / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, LPR);
    if(lane == 0) y[row] = s;
  }
}

// out[j] += sum_i V[j*ld + i] * w[i]   for j < nv (grid.y = ceil(nv / JB)); out must be zeroed
template <int JB>
__global__ void __launch_bounds__(256) multi_dot_kernel(int64_t n, const double *__restrict__ V, int64_t ld, int nv, const double *__restrict__ w,
                                                        double *__restrict__ out, const double *__restrict__ mask)
{
  const int j0 = blockIdx.y * JB;
  double    s[JB];
#pragma unroll
  for(int j = 0; j < JB; ++j) s[j] = 0.;
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double wi = mask ? w[i] * mask[i] : w[i]; // ghost rows are owned (and counted) by another rank
#pragma unroll
    for(int j = 0; j < JB; ++j)
      if(j0 + j < nv) s[j] += V[(int64_t)(j0 + j) * ld + i] * wi;
  }
  __shared__ double red[JB][8];
  const int         lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for(int j = 0; j < JB; ++j) {
    double v = s[j];
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if(lane == 0) red[j][wid] = v;
  }
  __syncthreads();
  if(threadIdx.x < JB) {
    double v = 0.;
#pragma unroll
    for(int k = 0; k < 8; ++k) v += red[threadIdx.x][k];
    if(j0 + threadIdx.x < nv) atomicAdd(out + j0 + threadIdx.x, v);
  }
}

// w[i] += sign * sum_j h[j] V[j*ld + i];  optionally hacc[j] += h[j] (thread 0 of block 0)
__global__ void __launch_bounds__(256) multi_axpy_kernel(int64_t n, const double *__restrict__ V, int64_t ld, int nv, const double *__restrict__ h,
                                                         double sign, double *__restrict__ w, double *hacc)
{
  extern __shared__ double sh[];
  for(int j = threadIdx.x; j < nv; j += blockDim.x) sh[j] = h[j];
  __syncthreads();
  if(hacc && blockIdx.x == 0)
    for(int j = threadIdx.x; j < nv; j += blockDim.x) hacc[j] += sh[j];
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.;
    for(int j = 0; j < nv; ++j) s += sh[j] * V[(int64_t)j * ld + i];
    w[i] += sign * s;
  }
}

// y = alpha * x (alpha read from device memory as 1/sqrt(*nrm2) when inv_sqrt is set)
__global__ void scale_copy_kernel(int64_t n, const double *__restrict__ x, const double *nrm2, double *__restrict__ y)
{
  const double a = 1. / sqrt(*nrm2);
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = a * x[i];
}

__global__ void axpby_kernel(int64_t n, double a, const double *__restrict__ x, double b, double *__restrict__ y)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = a * x[i] + b * y[i];
}

__global__ void max_abs_kernel(int64_t n, const double *__restrict__ x, double *out, const double *__restrict__ mask)
{
  double m = 0.;
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmax(m, mask ? fabs(x[i]) * mask[i] : fabs(x[i]));
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  __shared__ double red[32];
  const int         lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if(lane == 0) red[wid] = m;
  __syncthreads();
  if(wid == 0) {
    m = lane < (blockDim.x >> 5) ? red[lane] : 0.;
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    // non-negative doubles order like their bit patterns
    if(lane == 0) atomicMax(reinterpret_cast<unsigned long long *>(out), (unsigned long long)__double_as_longlong(m));
  }
}

// Point-Jacobi: dinv[i] = 1/A_ii, 1 where the diagonal vanishes (structurally present, numerically zero P-P block
// of Taylor-Hood, src/feCompressedRowStorage.cpp:33)
__global__ void jacobi_setup_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja, const double *__restrict__ val,
                                    double *__restrict__ dinv)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double d = 0.;
    for(int64_t k = ia[i]; k < ia[i + 1]; ++k)
      if(ja[k] == i) d = val[k];
    dinv[i] = (d != 0.) ? 1. / d : 1.;
  }
}

__global__ void diag_scale_kernel(int64_t n, const double *__restrict__ dinv, const double *__restrict__ r, double *__restrict__ z)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) z[i] = dinv[i] * r[i];
}

// Block Jacobi: one warp per block.  Gathers A[rows, rows] from the CSR matrix, inverts it by Gauss-Jordan with
// partial pivoting in shared memory and stores the dense inverse (row-major, bs x bs) at inv + off[b].
constexpr int MAXB = 32;
__global__ void __launch_bounds__(64) block_setup_kernel(int64_t nb, const int64_t *__restrict__ bptr, const int64_t *__restrict__ brow,
                                                          const int64_t *__restrict__ boff, const int64_t *__restrict__ ia,
                                                          const int32_t *__restrict__ ja, const double *__restrict__ val, double *__restrict__ inv,
                                                          int *singular)
{
  __shared__ double A[2][MAXB][2 * MAXB + 1];
  const int         w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t     b = blockIdx.x * 2 + w;
  if(b >= nb) return;
  const int64_t r0 = bptr[b];
  const int     bs = (int)(bptr[b + 1] - r0);
  // gather: lane = local row
  if(lane < bs) {
    const int64_t I = brow[r0 + lane];
    for(int j = 0; j < bs; ++j) {
      const int64_t J = brow[r0 + j];
      double        v = 0.;
      int64_t       lo = ia[I], hi = ia[I + 1] - 1;
      while(lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if(ja[mid] < J)
          lo = mid + 1;
        else
          hi = mid;
      }
      if(lo < ia[I + 1] && ja[lo] == J) v = val[lo];
      A[w][lane][j]      = v;
      A[w][lane][bs + j] = (j == lane) ? 1. : 0.;
    }
  }
  __syncwarp();
  for(int p = 0; p < bs; ++p) {
    // pivot search (all lanes redundantly)
    int    piv = p;
    double best = fabs(A[w][p][p]);
    for(int r = p + 1; r < bs; ++r) {
      const double v = fabs(A[w][r][p]);
      if(v > best) {
        best = v;
        piv = r;
      }
    }
    if(best == 0.) {
      if(lane == 0) atomicExch(singular, 1);
      // regularise: treat as identity row
      if(lane == 0) A[w][p][p] = 1.;
      __syncwarp();
      piv = p;
    }
    if(piv != p) {
      for(int j = lane; j < 2 * bs; j += 32) {
        const double t = A[w][p][j];
        A[w][p][j]     = A[w][piv][j];
        A[w][piv][j]   = t;
      }
    }
    __syncwarp();
    const double ip = 1. / A[w][p][p];
    __syncwarp();
    for(int j = lane; j < 2 * bs; j += 32) A[w][p][j] *= ip;
    __syncwarp();
    if(lane < bs && lane != p) {
      const double f = A[w][lane][p];
      if(f != 0.)
        for(int j = 0; j < 2 * bs; ++j) A[w][lane][j] -= f * A[w][p][j];
    }
    __syncwarp();
  }
  double *out = inv + boff[b];
  if(lane < bs)
    for(int j = 0; j < bs; ++j) out[lane * bs + j] = A[w][lane][bs + j];
}

// z[rows] = inv_b * r[rows]; rows not covered by any block are handled by the caller (identity)
__global__ void __launch_bounds__(128) block_apply_kernel(int64_t nb, const int64_t *__restrict__ bptr, const int64_t *__restrict__ brow,
                                                          const int64_t *__restrict__ boff, const double *__restrict__ inv,
                                                          const double *__restrict__ r, double *__restrict__ z)
{
  const int     w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = blockIdx.x * 4 + w;
  if(b >= nb) return;
  const int64_t r0 = bptr[b];
  const int     bs = (int)(bptr[b + 1] - r0);
  double        rv = 0.;
  int64_t       I  = 0;
  if(lane < bs) {
    I  = brow[r0 + lane];
    rv = r[I];
  }
  const double *Mi = inv + boff[b];
  double        s  = 0.;
  for(int j = 0; j < bs; ++j) {
    const double rj = __shfl_sync(0xffffffffu, rv, j);
    if(lane < bs) s += Mi[lane * bs + j] * rj;
  }
  if(lane < bs) z[I] = s;
}

// ----------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------
struct Krylov {
  int64_t n = 0;
  int     m = 0;
  double *V = nullptr;   // (m+1) x n
  double *w = nullptr, *z = nullptr, *r = nullptr, *x = nullptr;
  double *h = nullptr;   // device: [0..m] pass buffer, [m+1..2m+1] accumulated h, [2m+2] norm^2
  double *dinv = nullptr;
  // block Jacobi
  int64_t  nb = 0, inv_len = 0;
  int64_t *bptr = nullptr, *brow = nullptr, *boff = nullptr;
  double  *binv = nullptr;
  int     *d_flag = nullptr;
};

static const int GRID = 148 * 8;
// the Hessenberg column travels through the pinned scratch buffer of b200_create (256 doubles) and multi_axpy_kernel keeps the
// coefficients in shared memory
constexpr int GMRES_MAX_RESTART = 250;

void krylov_free(System *S)
{
  Krylov *K = static_cast<Krylov *>(S->krylov);
  if(!K) return;
  cudaFree(K->V);
  cudaFree(K->w);
  cudaFree(K->z);
  cudaFree(K->r);
  cudaFree(K->x);
  cudaFree(K->h);
  cudaFree(K->dinv);
  cudaFree(K->bptr);
  cudaFree(K->brow);
  cudaFree(K->boff);
  cudaFree(K->binv);
  cudaFree(K->d_flag);
  delete K;
  S->krylov = nullptr;
}

static int spmv_skip(System *S, const double *d_x, double *d_y, const uint8_t *skip)
{
  const int64_t n   = S->nInc;
  const double  avg = n > 0 ? (double)S->nnz / (double)n : 0.;
  // lanes per row / rows in flight per lane group picked on the B200 (profiles/README.md, r01c SpMV sweep): few lanes
  // per row and two rows per group keep the most independent loads outstanding: 71 % (2-D) / 68 % (3-D) of the
  // measured copy bandwidth against 49 % / 59 % for one 16- / 32-lane group per row
#define B200_SPMV_V(L, R)                                                                                     \
  {                                                                                                           \
    const int64_t blocks = (n * L / R + 255) / 256;                                                           \
    spmv_rpg_kernel<L, R, false><<<(unsigned)(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, S->stream>>>(   \
      n, S->d_ia, S->d_ja, S->d_val, d_x, d_y, skip);                                                         \
  }
  if(avg > 48.)
    B200_SPMV_V(8, 2)
  else if(avg > 12.)
    B200_SPMV_V(4, 2)
  else
    B200_SPMV_V(2, 2)
#undef B200_SPMV_V
  count_launch();
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int spmv(System *S, const double *d_x, double *d_y) { return spmv_skip(S, d_x, d_y, nullptr); }

int spmv_rows(System *S, const double *d_x, double *d_y, const uint8_t *skip, const int32_t *rows, int64_t n_rows)
{
  if(!rows) return spmv_skip(S, d_x, d_y, skip); // all rows but the flagged ones: the tuned two-rows-per-group kernel
  const int64_t cnt = rows ? n_rows : S->nInc;
  if(cnt <= 0) return B200_OK;
  const double  avg = S->nInc > 0 ? (double)S->nnz / (double)S->nInc : 0.;
  if(avg > 48.) {
    const int64_t blocks = (cnt * 8 + 255) / 256;
    spmv_subset_kernel<8><<<(unsigned)(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, S->stream>>>(S->nInc, S->d_ia, S->d_ja, S->d_val, d_x, d_y,
                                                                                                  skip, rows, n_rows);
  } else {
    const int64_t blocks = (cnt * 4 + 255) / 256;
    spmv_subset_kernel<4><<<(unsigned)(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, S->stream>>>(S->nInc, S->d_ia, S->d_ja, S->d_val, d_x, d_y,
                                                                                                  skip, rows, n_rows);
  }
  count_launch();
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int max_abs(System *S, const double *d_x, int64_t n, double *out)
{
  B200_CUDA(cudaMemsetAsync(S->d_scratch, 0, sizeof(double), S->stream));
  max_abs_kernel<<<GRID, 256, 0, S->stream>>>(n, d_x, S->d_scratch, comm_mask(S));
  count_launch();
  {
    const int crc = comm_allreduce(S, S->d_scratch, 1, true);
    if(crc != B200_OK) return crc;
  }
  B200_CUDA(cudaMemcpyAsync(S->h_scratch, S->d_scratch, sizeof(double), cudaMemcpyDeviceToHost, S->stream));
  B200_CUDA(cudaStreamSynchronize(S->stream));
  *out = S->h_scratch[0];
  return B200_OK;
}

static int ensure_workspace(System *S, int m)
{
  Krylov *K = static_cast<Krylov *>(S->krylov);
  if(K && (K->n != S->nInc || K->m != m)) {
    krylov_free(S);
    K = nullptr;
  }
  if(!K) {
    K    = new Krylov;
    K->n = S->nInc;
    K->m = m;
    S->krylov = K;
    const size_t nb = (size_t)K->n * sizeof(double);
    B200_CUDA(cudaMalloc(&K->V, (size_t)(m + 1) * nb));
    B200_CUDA(cudaMalloc(&K->w, nb));
    B200_CUDA(cudaMalloc(&K->z, nb));
    B200_CUDA(cudaMalloc(&K->r, nb));
    B200_CUDA(cudaMalloc(&K->x, nb));
    B200_CUDA(cudaMalloc(&K->h, (size_t)(2 * m + 4) * sizeof(double)));
    B200_CUDA(cudaMalloc(&K->dinv, nb));
    B200_CUDA(cudaMalloc(&K->d_flag, sizeof(int)));
  }
  return B200_OK;
}

static int setup_blocks(System *S, Krylov *K)
{
  if(K->bptr == nullptr) {
    if(S->n_blocks <= 0) {
      set_error("block-Jacobi requested but b200_set_blocks was not called");
      return B200_ERR_ARG;
    }
    K->nb = S->n_blocks;
    std::vector<int64_t> off(K->nb + 1, 0);
    for(int64_t b = 0; b < K->nb; ++b) {
      const int64_t bs = S->block_ptr[b + 1] - S->block_ptr[b];
      if(bs > MAXB || bs < 1) {
        set_error("block-Jacobi blocks must have 1..32 rows");
        return B200_ERR_ARG;
      }
      off[b + 1] = off[b] + bs * bs;
    }
    K->inv_len = off[K->nb];
    B200_CUDA(cudaMalloc(&K->bptr, (K->nb + 1) * sizeof(int64_t)));
    B200_CUDA(cudaMalloc(&K->boff, (K->nb + 1) * sizeof(int64_t)));
    B200_CUDA(cudaMalloc(&K->brow, S->block_rows.size() * sizeof(int64_t)));
    B200_CUDA(cudaMalloc(&K->binv, (size_t)K->inv_len * sizeof(double)));
    B200_CUDA(cudaMemcpy(K->bptr, S->block_ptr.data(), (K->nb + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    B200_CUDA(cudaMemcpy(K->boff, off.data(), (K->nb + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    B200_CUDA(cudaMemcpy(K->brow, S->block_rows.data(), S->block_rows.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
  }
  B200_CUDA(cudaMemsetAsync(K->d_flag, 0, sizeof(int), S->stream));
  block_setup_kernel<<<(unsigned)((K->nb + 1) / 2), 64, 0, S->stream>>>(K->nb, K->bptr, K->brow, K->boff, S->d_ia, S->d_ja, S->d_val, K->binv,
                                                                        K->d_flag);
  count_launch();
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

static int apply_pc(System *S, Krylov *K, int pc, const double *r, double *z)
{
  const int64_t n = K->n;
  if(pc == B200_PC_JACOBI) {
    diag_scale_kernel<<<GRID, 256, 0, S->stream>>>(n, K->dinv, r, z);
    count_launch();
  } else if(pc == B200_PC_BLOCK_JACOBI) {
    // rows outside every block fall back to point Jacobi
    diag_scale_kernel<<<GRID, 256, 0, S->stream>>>(n, K->dinv, r, z);
    block_apply_kernel<<<(unsigned)((K->nb + 3) / 4), 128, 0, S->stream>>>(K->nb, K->bptr, K->brow, K->boff, K->binv, r, z);
    count_launch(2);
  } else if(pc == B200_PC_AMG || pc == B200_PC_SCHUR_AMG) {
    return precond_apply(S, pc, r, z);
  } else {
    B200_CUDA(cudaMemcpyAsync(z, r, n * sizeof(double), cudaMemcpyDeviceToDevice, S->stream));
  }
  return B200_OK;
}

// Left-preconditioned GMRES(m):  M^-1 A x = M^-1 b, x0 = 0, convergence on the preconditioned residual 2-norm
// (KSP defaults: rnorm <= max(rtol * rnorm0, atol); divergence if rnorm > dtol * rnorm0).
int gmres_solve(System *S, const b200_solver_options *opt, b200_solve_info *info)
{
  const int m = opt->restart > 0 ? opt->restart : 30;
  if(m > GMRES_MAX_RESTART) {
    set_error("b200_solve: restart length above " + std::to_string(GMRES_MAX_RESTART));
    return B200_ERR_ARG;
  }
  int rc = ensure_workspace(S, m);
  if(rc != B200_OK) return rc;
  Krylov       *K = static_cast<Krylov *>(S->krylov);
  const int64_t n = K->n;
  int           pc = opt->pc;
  if(pc == B200_PC_AUTO) pc = S->plan == PLAN_TAYLOR_HOOD && S->sp >= 0 ? B200_PC_SCHUR_AMG : (S->plan == PLAN_SCALAR ? B200_PC_AMG : B200_PC_JACOBI);
  if(pc == B200_PC_AMG || pc == B200_PC_SCHUR_AMG) {
    rc = precond_setup(S, pc);
    if(rc != B200_OK) return rc;
  }
  if(pc < B200_PC_NONE || pc > B200_PC_SCHUR_AMG || pc == 3) {
    set_error("b200_solve: unknown preconditioner");
    return B200_ERR_ARG;
  }
  if(pc == B200_PC_JACOBI || pc == B200_PC_BLOCK_JACOBI) {
    jacobi_setup_kernel<<<GRID, 256, 0, S->stream>>>(n, S->d_ia, S->d_ja, S->d_val, K->dinv);
    count_launch();
  }
  if(pc == B200_PC_BLOCK_JACOBI) {
    rc = setup_blocks(S, K);
    if(rc != B200_OK) return rc;
    int flag = 0;
    B200_CUDA(cudaMemcpyAsync(&flag, K->d_flag, sizeof(int), cudaMemcpyDeviceToHost, S->stream));
    B200_CUDA(cudaStreamSynchronize(S->stream));
    if(flag) {
      set_error("B200_PC_BLOCK_JACOBI: a diagonal block is singular (e.g. a block of pressure rows only: the P-P block of a Taylor-Hood "
                "system is zero, src/feCompressedRowStorage.cpp:33) -- put every pressure row in a block with velocity rows");
      return B200_ERR_SOLVER;
    }
  }

  double *hbuf = K->h, *hacc = K->h + (m + 1), *nrm = K->h + (2 * m + 2);
  const double *mask = comm_mask(S); // multi-GPU: dot products over owned rows + all-reduce, halo update before every SpMV
  std::vector<double> H((size_t)(m + 1) * m, 0.), cs(m, 0.), sn(m, 0.), g(m + 1, 0.), y(m, 0.);
  std::vector<double> hh(m + 3, 0.);

  B200_CUDA(cudaMemsetAsync(K->x, 0, n * sizeof(double), S->stream));
  // r = M^-1 b
  rc = apply_pc(S, K, pc, S->d_rhs, K->r);
  if(rc != B200_OK) return rc;

  auto norm2_of = [&](const double *v, double *out) -> int {
    B200_CUDA(cudaMemsetAsync(nrm, 0, sizeof(double), S->stream));
    multi_dot_kernel<1><<<dim3(GRID, 1), 256, 0, S->stream>>>(n, v, n, 1, v, nrm, mask);
    count_launch();
    {
      const int crc = comm_allreduce(S, nrm, 1, false);
      if(crc != B200_OK) return crc;
    }
    B200_CUDA(cudaMemcpyAsync(S->h_scratch, nrm, sizeof(double), cudaMemcpyDeviceToHost, S->stream));
    B200_CUDA(cudaStreamSynchronize(S->stream));
    *out = sqrt(S->h_scratch[0]);
    return B200_OK;
  };

  double rnorm0 = 0.;
  rc = norm2_of(K->r, &rnorm0);
  if(rc != B200_OK) return rc;
  double    rnorm = rnorm0;
  double    tol = fmax(opt->rel_tol * rnorm0, opt->abs_tol), cycle_start = rnorm0;
  int       slow_cycles = 0;
  int       its = 0;
  bool      converged = rnorm0 <= opt->abs_tol, diverged = false;

  while(!converged && !diverged && its < opt->max_iter) {
    // v0 = r / ||r||  (nrm holds ||r||^2 on the device)
    scale_copy_kernel<<<GRID, 256, 0, S->stream>>>(n, K->r, nrm, K->V);
    count_launch();
    std::fill(g.begin(), g.end(), 0.);
    g[0] = rnorm;
    int k = 0;
    for(; k < m && its < opt->max_iter; ++k) {
      ++its;
      double *vk = K->V + (int64_t)k * n, *vk1 = K->V + (int64_t)(k + 1) * n;
      // w = M^-1 A v_k (several GPUs: the halo update of v_k is hidden behind the interior rows of the product)
      rc = comm_spmv_overlapped(S, vk, K->z);
      if(rc != B200_OK) return rc;
      rc = apply_pc(S, K, pc, K->z, K->w);
      if(rc != B200_OK) return rc;
      // CGS2
      B200_CUDA(cudaMemsetAsync(hbuf, 0, (size_t)(2 * m + 3) * sizeof(double), S->stream));
      for(int pass = 0; pass < 2; ++pass) {
        if(pass == 1) B200_CUDA(cudaMemsetAsync(hbuf, 0, (size_t)(m + 1) * sizeof(double), S->stream));
        multi_dot_kernel<4><<<dim3(GRID / 4, (k + 1 + 3) / 4), 256, 0, S->stream>>>(n, K->V, n, k + 1, K->w, hbuf, mask);
        rc = comm_allreduce(S, hbuf, k + 1, false);
        if(rc != B200_OK) return rc;
        multi_axpy_kernel<<<GRID, 256, (k + 1) * sizeof(double), S->stream>>>(n, K->V, n, k + 1, hbuf, -1., K->w, hacc);
        count_launch(2);
      }
      multi_dot_kernel<1><<<dim3(GRID, 1), 256, 0, S->stream>>>(n, K->w, n, 1, K->w, nrm, mask);
      count_launch();
      rc = comm_allreduce(S, nrm, 1, false);
      if(rc != B200_OK) return rc;
      B200_CUDA(cudaMemcpyAsync(S->h_scratch, hacc, (size_t)(m + 2) * sizeof(double), cudaMemcpyDeviceToHost, S->stream));
      B200_CUDA(cudaStreamSynchronize(S->stream));
      for(int j = 0; j <= k; ++j) hh[j] = S->h_scratch[j];
      const double hk1 = sqrt(S->h_scratch[m + 1]); // nrm sits right after hacc[0..m]
      hh[k + 1] = hk1;
      // v_{k+1} = w / h_{k+1,k}
      if(hk1 > 0.) {
        scale_copy_kernel<<<GRID, 256, 0, S->stream>>>(n, K->w, nrm, vk1);
        count_launch();
      }
      // Givens
      for(int j = 0; j < k; ++j) {
        const double t = cs[j] * hh[j] + sn[j] * hh[j + 1];
        hh[j + 1]      = -sn[j] * hh[j] + cs[j] * hh[j + 1];
        hh[j]          = t;
      }
      const double den = hypot(hh[k], hh[k + 1]);
      cs[k] = den > 0. ? hh[k] / den : 1.;
      sn[k] = den > 0. ? hh[k + 1] / den : 0.;
      hh[k]     = den;
      hh[k + 1] = 0.;
      for(int j = 0; j <= k; ++j) H[(size_t)j * m + k] = hh[j];
      g[k + 1] = -sn[k] * g[k];
      g[k]     = cs[k] * g[k];
      rnorm    = fabs(g[k + 1]);
      if(rnorm <= tol) {
        converged = true;
        ++k;
        break;
      }
      if(rnorm > opt->div_tol * rnorm0 || !std::isfinite(rnorm)) {
        diverged = true;
        ++k;
        break;
      }
      if(hk1 == 0.) { // happy breakdown
        converged = true;
        ++k;
        break;
      }
    }
    // solve H y = g (k x k upper triangular), x += V y
    for(int i = k - 1; i >= 0; --i) {
      double s = g[i];
      for(int j = i + 1; j < k; ++j) s -= H[(size_t)i * m + j] * y[j];
      y[i] = s / H[(size_t)i * m + i];
    }
    if(k > 0) {
      B200_CUDA(cudaMemcpyAsync(hbuf, y.data(), (size_t)k * sizeof(double), cudaMemcpyHostToDevice, S->stream));
      multi_axpy_kernel<<<GRID, 256, k * sizeof(double), S->stream>>>(n, K->V, n, k, hbuf, 1., K->x, nullptr);
      count_launch();
      B200_CUDA(cudaStreamSynchronize(S->stream)); // y is reused by the next cycle
    }
    if(converged || diverged) break;
    // true preconditioned residual for the restart: r = M^-1 (b - A x)
    rc = comm_halo_exchange(S, K->x);
    if(rc != B200_OK) return rc;
    rc = spmv(S, K->x, K->z);
    if(rc != B200_OK) return rc;
    axpby_kernel<<<GRID, 256, 0, S->stream>>>(n, 1., S->d_rhs, -1., K->z);
    count_launch();
    rc = apply_pc(S, K, pc, K->z, K->r);
    if(rc != B200_OK) return rc;
    rc = norm2_of(K->r, &rnorm);
    if(rc != B200_OK) return rc;
    if(rnorm <= tol) converged = true;
    // stagnation guard of the multigrid preconditioners: two restart cycles in a row that gain less than 30 % mean the
    // Chebyshev smoother is amplifying a convection-dominated level (cell Peclet number above one: coarse meshes at high
    // Reynolds number); fall back to plain damped Jacobi smoothing, which is slower per cycle but safe, and re-base the
    // convergence test on the new preconditioned norm
    if(!converged && (pc == B200_PC_AMG || pc == B200_PC_SCHUR_AMG)) {
      slow_cycles = rnorm > 0.7 * cycle_start ? slow_cycles + 1 : 0;
      cycle_start = rnorm;
      if(slow_cycles >= 2 && precond_fallback(S)) {
        slow_cycles = 0;
        rc = apply_pc(S, K, pc, S->d_rhs, K->r);
        if(rc != B200_OK) return rc;
        rc = norm2_of(K->r, &rnorm0);
        if(rc != B200_OK) return rc;
        tol = fmax(opt->rel_tol * rnorm0, opt->abs_tol);
        rc = apply_pc(S, K, pc, K->z, K->r);
        if(rc != B200_OK) return rc;
        rc = norm2_of(K->r, &rnorm);
        if(rc != B200_OK) return rc;
        cycle_start = rnorm;
        if(rnorm <= tol) converged = true;
      }
    }
  }

  // du = x; norms reported to the Newton loop are max-norms (src/feLinearSystemMklPardiso.cpp:960-961)
  rc = comm_halo_exchange(S, K->x); // ghost entries of du <- owner's values: every rank corrects its whole local state
  if(rc != B200_OK) return rc;
  B200_CUDA(cudaMemcpyAsync(S->d_du, K->x, n * sizeof(double), cudaMemcpyDeviceToDevice, S->stream));
  rc = spmv(S, K->x, K->z);
  if(rc != B200_OK) return rc;
  axpby_kernel<<<GRID, 256, 0, S->stream>>>(n, -1., S->d_rhs, 1., K->z); // z = A x - b
  count_launch();
  double ndx, nrhs, naxb;
  if((rc = max_abs(S, K->x, n, &ndx)) != B200_OK) return rc;
  if((rc = max_abs(S, S->d_rhs, n, &nrhs)) != B200_OK) return rc;
  if((rc = max_abs(S, K->z, n, &naxb)) != B200_OK) return rc;
  info->norm_dx      = ndx;
  info->norm_rhs     = nrhs;
  info->norm_axb     = naxb;
  info->iterations   = its;
  info->converged    = converged ? 1 : 0;
  info->rel_residual = rnorm0 > 0. ? rnorm / rnorm0 : 0.;
  if(diverged) {
    set_error("GMRES diverged (preconditioned residual exceeded div_tol)");
    return B200_ERR_SOLVER;
  }
  return B200_OK;
}

} // namespace b200
