// Aggregation multigrid on the device: the approximate inverse of an elliptic field block inside the preconditioners
// of the Newton linear solve (north-star subsystem 4).
//
// The reference delegates the solve to PETSc's KSP (GMRES(30) + ILU(0) on one rank, block-Jacobi on n ranks,
// src/feLinearSystem.h:196-199, src/feLinearSystemPETSc.cpp:337) or to Pardiso LU
// (src/feLinearSystemMklPardiso.cpp:893-965).  A triangular-solve-based ILU is a poor fit for the GPU and a
// single-level method needs O(1/h) iterations; this hierarchy is built from sort / scan / SpMV primitives only:
//
//   level 0   the assembled CSR matrix itself (a row mask selects the field: velocity rows of a Taylor-Hood system, or
//             every row of a scalar problem); nothing is copied
//   level 1   for P2 spaces: the P1 sub-space on the same mesh (vertex unknowns keep their value, a mid-edge unknown
//             is the mean of its two end vertices) -- Galerkin product P^T A P of the component-diagonal blocks
//   level 2+  plain aggregation around a maximal independent set of the matrix graph (Bell, Dalton, Olson, SIAM J. Sci.
//             Comput. 34 (2012)): roots = MIS(1) by default (MIS(2): B200_AMG_MIS=2), every unknown joins the nearest root,
//             Galerkin product = sum of the entries of an aggregate pair
//   coarsest  dense inverse (Gauss-Jordan in one CTA)
//
// Cycle: Chebyshev-accelerated Jacobi smoothing, 2 + 2 sweeps (spectral radius of D^-1 A from a few power iterations); V on the
// P2 -> P1 step, W below it (B200_AMG_GAMMA), corrections of the aggregation levels scaled by 1.5 (B200_AMG_OVERCORRECT:
// piecewise-constant prolongation under-estimates them; a fixed factor keeps the cycle a linear operator).  Measured at
// T3D(92): 197 -> 103 Krylov iterations against MIS(2) / V / 1 (profiles/README.md r02y).  The symbolic part (parents,
// aggregates, coarse patterns) depends on the mesh and the pattern only and is built once; the numeric part (Galerkin values,
// inverse diagonals, spectral radii, coarse inverse) is redone whenever the matrix changed.
// On several GPUs level 0 is rank-local but keeps its couplings to ghost columns (halo update before every product), and the
// hierarchy from level 1 on is global and replicated: the level-1 triplets of all ranks (couplings across the cuts included)
// are all-gathered, merged and coarsened identically everywhere; B200_AMG_GLOBAL_LEVEL=2 keeps the P1 level rank-local.
#include <thrust/binary_search.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "amg.h"

namespace b200 {

static const int GRID = 148 * 8;

// ----------------------------------------------------------------------------------------------------------
// field map / P2 -> P1 parents
// ----------------------------------------------------------------------------------------------------------
// fld[dof] = base + component of the local function (function a*nc + c is phi_a e_c, src/feSpace_2D.cpp:41-55)
__global__ void amg_field_kernel(int64_t nElm, const int32_t *__restrict__ adr, int nloc, int nc, int64_t nInc, int base, uint8_t *fld)
{
  const int64_t tot = nElm * nloc;
  for(int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    const int     k = (int)(idx % nloc);
    const int64_t dof = adr[idx];
    if(dof < nInc) fld[dof] = (uint8_t)(base + k % nc);
  }
}

__global__ void amg_mask_ghost_kernel(int64_t n, const double *__restrict__ owned, uint8_t *fld)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if(owned[i] == 0.) fld[i] = AMG_FLD_NONE;
}

__global__ void amg_vertex_flag_kernel(int64_t nElm, const int32_t *__restrict__ adr, int nloc, int nvc, int64_t nInc,
                                       const uint8_t *__restrict__ fld, int fld_lo, int fld_hi, int32_t *isvert)
{
  const int64_t tot = nElm * nvc;
  for(int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx / nvc;
    const int     k = (int)(idx - e * nvc);
    const int64_t dof = adr[e * nloc + k];
    if(dof < nInc && fld[dof] >= fld_lo && fld[dof] < fld_hi) isvert[dof] = 1;
  }
}

// local edge -> end vertices of the P2 Lagrange bases of the reference: triangle edges (0,1),(1,2),(2,0)
// (src/feSpace_2D.cpp:536-544), tetrahedron edges in the order of _edgesOrder (src/feTetrahedron.h:31).  The host
// checks the table against the basis tabulation it was given (amg_check_p2_edges).
__constant__ int c_edge_tri[3][2] = {{0, 1}, {1, 2}, {2, 0}};
__constant__ int c_edge_tet[6][2] = {{0, 2}, {2, 1}, {1, 0}, {1, 3}, {3, 0}, {3, 2}};
static const int h_edge_tri[3][2] = {{0, 1}, {1, 2}, {2, 0}};
static const int h_edge_tet[6][2] = {{0, 2}, {2, 1}, {1, 0}, {1, 3}, {3, 0}, {3, 2}};

// kind: 0 inactive, 1 one parent with weight 1, 2 mid-edge unknown: up to two parents with weight 1/2 each (a missing
// parent is an essential / foreign vertex: its coarse function does not exist)
__global__ void amg_p2_parent_kernel(int64_t nElm, const int32_t *__restrict__ adr, int nloc, int nv, int nc, int dim, int64_t nInc,
                                     const uint8_t *__restrict__ fld, int fld_lo, int fld_hi, const int32_t *__restrict__ isvert,
                                     const int32_t *__restrict__ cid, int32_t *par0, int32_t *par1, uint8_t *pkind)
{
  const int     nS = nloc / nc;
  const int64_t tot = nElm * nloc;
  for(int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx / nloc;
    const int     k = (int)(idx - e * nloc);
    const int     a = k / nc, c = k - a * nc;
    const int64_t dof = adr[idx];
    if(dof >= nInc || fld[dof] < fld_lo || fld[dof] >= fld_hi) continue;
    if(a < nv) {
      par0[dof]  = cid[dof];
      par1[dof]  = -1;
      pkind[dof] = 1;
    } else if(a < nS) {
      const int     ed = a - nv;
      const int     va = dim == 2 ? c_edge_tri[ed][0] : c_edge_tet[ed][0], vb = dim == 2 ? c_edge_tri[ed][1] : c_edge_tet[ed][1];
      const int64_t da = adr[e * nloc + va * nc + c], db = adr[e * nloc + vb * nc + c];
      int32_t       pa = (da < nInc && isvert[da]) ? cid[da] : -1;
      int32_t       pb = (db < nInc && isvert[db]) ? cid[db] : -1;
      // canonical order so that every element sharing the edge writes the same pair
      if(pa < pb) {
        const int32_t t = pa;
        pa = pb;
        pb = t;
      }
      par0[dof]  = pa;
      par1[dof]  = pb;
      pkind[dof] = 2;
    }
  }
}

// keys of the P1 element pattern of the coarse unknowns: (I, J) for every pair of vertex functions of an element
// (same component only when decoupled)
__global__ void amg_p1_keys_kernel(int64_t nElm, const int32_t *__restrict__ adr, int nloc, int nvc, int nc, int64_t nInc,
                                   const int32_t *__restrict__ isvert, const int32_t *__restrict__ cid, int64_t ncoarse, int decoupled,
                                   uint64_t *keys)
{
  const int64_t per = (int64_t)nvc * nvc, tot = nElm * per;
  for(int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx / per;
    const int     r = (int)(idx - e * per);
    const int     i = r / nvc, j = r - i * nvc;
    uint64_t      key = ~0ull;
    if(!decoupled || (i % nc) == (j % nc)) {
      const int64_t di = adr[e * nloc + i], dj = adr[e * nloc + j];
      if(di < nInc && dj < nInc && isvert[di] && isvert[dj]) key = (uint64_t)cid[di] * (uint64_t)ncoarse + (uint64_t)cid[dj];
    }
    keys[idx] = key;
  }
}

// keys of the aggregated pattern: (parent(i), parent(j)) for every stored entry
__global__ void amg_agg_keys_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja, const int32_t *__restrict__ par0,
                                    int64_t ncoarse, uint64_t *keys)
{
  const int     lane = threadIdx.x & 7;
  const int64_t g0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3, ng = (gridDim.x * (int64_t)blockDim.x) >> 3;
  for(int64_t i = g0; i < n; i += ng) {
    const int32_t I = par0[i];
    for(int64_t k = ia[i] + lane; k < ia[i + 1]; k += 8) {
      const int32_t J = par0[ja[k]];
      keys[k] = (I >= 0 && J >= 0) ? (uint64_t)I * (uint64_t)ncoarse + (uint64_t)J : ~0ull;
    }
  }
}

__global__ void amg_row_start_keys_kernel(int64_t n, uint64_t *q)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) q[i] = (uint64_t)i * (uint64_t)n;
}

__global__ void amg_split_keys_kernel(int64_t nnz, int64_t n, const uint64_t *keys, int32_t *ja)
{
  for(int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x)
    ja[k] = (int32_t)(keys[k] % (uint64_t)n);
}


// ----------------------------------------------------------------------------------------------------------
// level-0 single-precision working copy (same-field entries of the active rows)
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool amg_keep(const uint8_t *__restrict__ fld_col, const uint8_t *__restrict__ kind_col, int fi, int32_t j)
{
  return kind_col[j] != 0 && fld_col[j] == fi;
}

__global__ void __launch_bounds__(256) amg_dec_count_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja,
                                                            const uint8_t *__restrict__ fld, const uint8_t *__restrict__ pkind,
                                                            const uint8_t *__restrict__ fld_col, const uint8_t *__restrict__ kind_col, int64_t *cnt)
{
  const int     lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * (int64_t)blockDim.x) >> 5;
  for(int64_t i = w0; i < n; i += nw) {
    int c = 0;
    if(pkind[i]) {
      const int fi = fld[i];
      for(int64_t k = ia[i] + lane; k < ia[i + 1]; k += 32) c += amg_keep(fld_col, kind_col, fi, ja[k]) ? 1 : 0;
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if(lane == 0) cnt[i] = c;
  }
}

template <bool WITH_JA>
__global__ void __launch_bounds__(256) amg_dec_fill_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja,
                                                           const double *__restrict__ val, const uint8_t *__restrict__ fld,
                                                           const uint8_t *__restrict__ pkind, const uint8_t *__restrict__ fld_col,
                                                           const uint8_t *__restrict__ kind_col, const int64_t *__restrict__ ias, int32_t *jas,
                                                           float *vals)
{
  const int     lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * (int64_t)blockDim.x) >> 5;
  for(int64_t i = w0; i < n; i += nw) {
    if(!pkind[i]) continue;
    const int     fi = fld[i];
    int64_t       out = ias[i];
    const int64_t beg = ia[i], end = ia[i + 1];
    for(int64_t k0 = beg; k0 < end; k0 += 32) {
      const int64_t  k = k0 + lane;
      const int32_t  j = k < end ? ja[k] : 0;
      const bool     keep = k < end && amg_keep(fld_col, kind_col, fi, j);
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if(keep) {
        const int64_t o = out + __popc(m & ((1u << lane) - 1u));
        if(WITH_JA) jas[o] = j;
        vals[o] = (float)val[k];
      }
      out += __popc(m);
    }
  }
}

template <int LPR, int RPG>
__global__ void __launch_bounds__(256) amg_spmv_f32_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja,
                                                           const float *__restrict__ val, const double *__restrict__ x, double *__restrict__ y)
{
  const int     lane = threadIdx.x % LPR;
  const int64_t g0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR, ng = (gridDim.x * (int64_t)blockDim.x) / LPR;
  for(int64_t row0 = g0 * RPG; row0 < n; row0 += ng * RPG) {
    int64_t beg[RPG], end[RPG];
#pragma unroll
    for(int r = 0; r < RPG; ++r) {
      const int64_t row = row0 + r;
      beg[r] = row < n ? ia[row] : 0;
      end[r] = row < n ? ia[row + 1] : 0;
    }
    double s[RPG];
#pragma unroll
    for(int r = 0; r < RPG; ++r) s[r] = 0.;
    int64_t longest = 0;
#pragma unroll
    for(int r = 0; r < RPG; ++r) longest = max(longest, end[r] - beg[r]);
    for(int64_t off = lane; off < longest; off += LPR) {
      float   v[RPG];
      int32_t c[RPG];
#pragma unroll
      for(int r = 0; r < RPG; ++r) {
        const int64_t k = beg[r] + off;
        const bool    ok = k < end[r];
        v[r] = ok ? val[k] : 0.f;
        c[r] = ok ? ja[k] : 0;
      }
#pragma unroll
      for(int r = 0; r < RPG; ++r) s[r] += (double)v[r] * x[c[r]];
    }
#pragma unroll
    for(int r = 0; r < RPG; ++r) {
#pragma unroll
      for(int o = LPR / 2; o > 0; o >>= 1) s[r] += __shfl_down_sync(0xffffffffu, s[r], o, LPR);
    }
    if(lane == 0) {
#pragma unroll
      for(int r = 0; r < RPG; ++r)
        if(row0 + r < n) y[row0 + r] = s[r];
    }
  }
}

// ----------------------------------------------------------------------------------------------------------
// Galerkin product  Ac(I, J) += w_iI w_jJ A(i, j)
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void amg_add_coarse(const int64_t *__restrict__ iac, const int32_t *__restrict__ jac, double *valc, int32_t I, int32_t J,
                                               double v)
{
  int64_t lo = iac[I], hi = iac[I + 1] - 1;
  while(lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if(jac[mid] < J)
      lo = mid + 1;
    else
      hi = mid;
  }
  if(jac[lo] == J) atomicAdd(valc + lo, v);
}

template <int LPR, typename VT>
__global__ void __launch_bounds__(256) amg_galerkin_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja,
                                                           const VT *__restrict__ val, const uint8_t *__restrict__ fld,
                                                           const int32_t *__restrict__ par0, const int32_t *__restrict__ par1,
                                                           const uint8_t *__restrict__ pkind, const int64_t *__restrict__ iac,
                                                           const int32_t *__restrict__ jac, double *valc)
{
  const int     lane = threadIdx.x % LPR;
  const int64_t g0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR, ng = (gridDim.x * (int64_t)blockDim.x) / LPR;
  for(int64_t i = g0; i < n; i += ng) {
    const int ki = pkind[i];
    if(ki == 0) continue;
    const int32_t I0 = par0[i], I1 = par1 ? par1[i] : -1;
    const double  wi = ki == 2 ? 0.5 : 1.;
    const int     fi = fld ? fld[i] : 0;
    for(int64_t k = ia[i] + lane; k < ia[i + 1]; k += LPR) {
      const int32_t j = ja[k];
      const int     kj = pkind[j];
      if(kj == 0 || (fld && fld[j] != fi)) continue;
      const double  v = (double)val[k] * wi * (kj == 2 ? 0.5 : 1.);
      const int32_t J0 = par0[j], J1 = par1 ? par1[j] : -1;
      if(I0 >= 0) {
        if(J0 >= 0) amg_add_coarse(iac, jac, valc, I0, J0, v);
        if(J1 >= 0) amg_add_coarse(iac, jac, valc, I0, J1, v);
      }
      if(I1 >= 0) {
        if(J0 >= 0) amg_add_coarse(iac, jac, valc, I1, J0, v);
        if(J1 >= 0) amg_add_coarse(iac, jac, valc, I1, J1, v);
      }
    }
  }
}

// inverse diagonal of the active rows (0 elsewhere: inactive rows never change)
__global__ void amg_dinv_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja, const double *__restrict__ val,
                                const uint8_t *__restrict__ pkind, double *__restrict__ dinv)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double d = 0.;
    if(!pkind || pkind[i]) {
      for(int64_t k = ia[i]; k < ia[i + 1]; ++k)
        if(ja[k] == i) d = val[k];
    }
    dinv[i] = d != 0. ? 1. / d : 0.;
  }
}

// ----------------------------------------------------------------------------------------------------------
// MIS aggregation (distance 2: two propagation steps per round; distance 1: one)
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t amg_hash(uint32_t x)
{
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

// key = state (2 bits: 0 decided-out, 1 undecided, 2 root) | hash (30 bits) | index (32 bits); `active` (nullable) is
// the row mask of the graph: inactive rows neither take part nor relay
__global__ void amg_mis_init_kernel(int64_t n, const uint8_t *__restrict__ active, uint64_t *key)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    key[i] = (active && !active[i]) ? 0ull : ((1ull << 62) | ((uint64_t)(amg_hash((uint32_t)i) & 0x3fffffffu) << 32) | (uint64_t)i);
}

__global__ void amg_mis_prop_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja, const uint8_t *__restrict__ active,
                                    const uint64_t *__restrict__ in, uint64_t *__restrict__ out)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t m = 0ull;
    if(!active || active[i]) {
      m = in[i];
      for(int64_t k = ia[i]; k < ia[i + 1]; ++k) {
        const uint64_t v = in[ja[k]];
        m = v > m ? v : m;
      }
    }
    out[i] = m;
  }
}

__global__ void amg_mis_update_kernel(int64_t n, uint64_t *key, const uint64_t *__restrict__ m2, int *undecided)
{
  int cnt = 0;
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = key[i];
    if((k >> 62) == 1ull) {
      if(m2[i] == k)
        key[i] = (k & ~(3ull << 62)) | (2ull << 62); // largest undecided key of its 2-ring: new root
      else if((m2[i] >> 62) == 2ull)
        key[i] = 0ull;                               // a root within distance 2: covered
      else
        ++cnt;
    }
  }
  if(cnt) atomicAdd(undecided, cnt);
}

__global__ void amg_root_flag_kernel(int64_t n, const uint64_t *__restrict__ key, int32_t *isroot)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    isroot[i] = (key[i] >> 62) == 2ull ? 1 : 0;
}

__global__ void amg_agg_root_kernel(int64_t n, const uint64_t *__restrict__ key, const int32_t *__restrict__ rid, int32_t *agg)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    agg[i] = (key[i] >> 62) == 2ull ? rid[i] : -1;
}

// one ring of growth: an unassigned active unknown joins the aggregate of its most strongly connected assigned neighbour
__global__ void amg_agg_grow_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja, const double *__restrict__ val,
                                    const uint8_t *__restrict__ active, const int32_t *__restrict__ in, int32_t *__restrict__ out)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int32_t a = in[i];
    if(a < 0 && (!active || active[i])) {
      double best = -1.;
      for(int64_t k = ia[i]; k < ia[i + 1]; ++k) {
        const int32_t b = in[ja[k]];
        const double  w = val ? fabs(val[k]) : 1.;
        if(b >= 0 && (w > best || (w == best && b > a))) {
          best = w;
          a    = b;
        }
      }
    }
    out[i] = a;
  }
}

__global__ void amg_orphan_flag_kernel(int64_t n, const uint8_t *__restrict__ active, const int32_t *__restrict__ agg, int32_t *flag)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flag[i] = ((!active || active[i]) && agg[i] < 0) ? 1 : 0;
}

__global__ void amg_orphan_assign_kernel(int64_t n, const int32_t *__restrict__ flag, const int32_t *__restrict__ oid, int32_t base, int32_t *agg)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if(flag[i]) agg[i] = base + oid[i];
}

__global__ void amg_kind_from_agg_kernel(int64_t n, const int32_t *__restrict__ agg, uint8_t *pkind)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) pkind[i] = agg[i] >= 0 ? 1 : 0;
}

__global__ void amg_in_range_kernel(int64_t n, const uint8_t *__restrict__ fld, int lo, int hi, uint8_t *act)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    act[i] = (fld[i] >= lo && fld[i] < hi) ? 1 : 0;
}

__global__ void amg_norm2_kernel(int64_t n, const double *__restrict__ x, double *nrm2)
{
  double s = 0.;
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += x[i] * x[i];
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if((threadIdx.x & 31) == 0 && s != 0.) atomicAdd(nrm2, s);
}

// ----------------------------------------------------------------------------------------------------------
// cycle kernels
// ----------------------------------------------------------------------------------------------------------
// generic CSR product (coarse levels), LPR lanes per row
template <int LPR>
__global__ void __launch_bounds__(256) amg_spmv_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja,
                                                       const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y)
{
  const int     lane = threadIdx.x % LPR;
  const int64_t g0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR, ng = (gridDim.x * (int64_t)blockDim.x) / LPR;
  for(int64_t i = g0; i < n; i += ng) {
    double s = 0.;
    for(int64_t k = ia[i] + lane; k < ia[i + 1]; k += LPR) s += val[k] * x[ja[k]];
#pragma unroll
    for(int o = LPR / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, LPR);
    if(lane == 0) y[i] = s;
  }
}

// Chebyshev step: r = b - t (t = A x, or 0 when first); d = c1 d + c2 dinv r; x (+)= d
__global__ void amg_cheb_kernel(int64_t n, const double *__restrict__ b, const double *__restrict__ t, const double *__restrict__ dinv, double c1,
                                double c2, double *__restrict__ d, double *__restrict__ x, int first_zero)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double r  = t ? b[i] - t[i] : b[i];
    const double dn = (c1 != 0. ? c1 * d[i] : 0.) + c2 * dinv[i] * r;
    d[i] = dn;
    x[i] = first_zero ? dn : x[i] + dn;
  }
}

// rc[parent] += w (b - t)
__global__ void amg_restrict_kernel(int64_t n, const double *__restrict__ b, const double *__restrict__ t, const int32_t *__restrict__ par0,
                                    const int32_t *__restrict__ par1, const uint8_t *__restrict__ pkind, double *rc)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = pkind[i];
    if(k == 0) continue;
    const double  r = (b[i] - t[i]) * (k == 2 ? 0.5 : 1.);
    const int32_t p0 = par0[i], p1 = par1 ? par1[i] : -1;
    if(p0 >= 0) atomicAdd(rc + p0, r);
    if(p1 >= 0) atomicAdd(rc + p1, r);
  }
}

__global__ void amg_prolong_kernel(int64_t n, const double *__restrict__ xc, const int32_t *__restrict__ par0, const int32_t *__restrict__ par1,
                                   const uint8_t *__restrict__ pkind, double w, double *__restrict__ x)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = pkind[i];
    if(k == 0) continue;
    const int32_t p0 = par0[i], p1 = par1 ? par1[i] : -1;
    double        s = 0.;
    if(p0 >= 0) s += xc[p0];
    if(p1 >= 0) s += xc[p1];
    x[i] += (k == 2 ? 0.5 : 1.) * w * s;
  }
}

// power iteration helpers: y = dinv * t, accumulate |y|^2
__global__ void amg_scale_norm_kernel(int64_t n, const double *__restrict__ dinv, const double *__restrict__ t, double *__restrict__ y, double *nrm2)
{
  double s = 0.;
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = dinv[i] * t[i];
    y[i] = v;
    s += v * v;
  }
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if((threadIdx.x & 31) == 0 && s != 0.) atomicAdd(nrm2, s);
}

__global__ void amg_normalize_kernel(int64_t n, const double *nrm2, double *__restrict__ y)
{
  const double a = *nrm2 > 0. ? 1. / sqrt(*nrm2) : 0.;
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] *= a;
}

__global__ void amg_seed_kernel(int64_t n, const double *__restrict__ dinv, double *__restrict__ y)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = dinv[i] != 0. ? 0.5 + (double)(amg_hash((uint32_t)i) & 0xffffu) / 65536. : 0.;
}

// coarsest level: dense copy [A | I] and Gauss-Jordan with partial pivoting in one CTA (n <= AMG_MAX_DENSE)
__global__ void amg_dense_fill_kernel(int n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja, const double *__restrict__ val, double *M)
{
  const int ld = 2 * n;
  for(int i = blockIdx.x; i < n; i += gridDim.x) {
    for(int j = threadIdx.x; j < ld; j += blockDim.x) M[(size_t)i * ld + j] = (j == n + i) ? 1. : 0.;
    __syncthreads();
    for(int64_t k = ia[i] + threadIdx.x; k < ia[i + 1]; k += blockDim.x) M[(size_t)i * ld + ja[k]] = val[k];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) amg_gauss_jordan_kernel(int n, double *M, double *inv)
{
  const int   ld = 2 * n;
  __shared__ int    s_piv;
  __shared__ double s_red[32];
  __shared__ int    s_idx[32];
  for(int p = 0; p < n; ++p) {
    // pivot search in column p, rows p..n-1
    double best = -1.;
    int    bi = p;
    for(int r = p + threadIdx.x; r < n; r += blockDim.x) {
      const double v = fabs(M[(size_t)r * ld + p]);
      if(v > best) {
        best = v;
        bi   = r;
      }
    }
    for(int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_down_sync(0xffffffffu, best, o);
      const int    oi = __shfl_down_sync(0xffffffffu, bi, o);
      if(ob > best) {
        best = ob;
        bi   = oi;
      }
    }
    if((threadIdx.x & 31) == 0) {
      s_red[threadIdx.x >> 5] = best;
      s_idx[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    if(threadIdx.x == 0) {
      double b = -1.;
      int    i = p;
      for(int w = 0; w < (int)(blockDim.x >> 5); ++w)
        if(s_red[w] > b) {
          b = s_red[w];
          i = s_idx[w];
        }
      s_piv = i;
      if(b <= 0.) { // empty column (an aggregate of decoupled zero rows): identity
        M[(size_t)p * ld + p] = 1.;
        s_piv = p;
      }
    }
    __syncthreads();
    const int piv = s_piv;
    if(piv != p) {
      for(int j = threadIdx.x; j < ld; j += blockDim.x) {
        const double t = M[(size_t)p * ld + j];
        M[(size_t)p * ld + j]   = M[(size_t)piv * ld + j];
        M[(size_t)piv * ld + j] = t;
      }
    }
    __syncthreads();
    const double ip = 1. / M[(size_t)p * ld + p];
    __syncthreads();
    for(int j = threadIdx.x; j < ld; j += blockDim.x) M[(size_t)p * ld + j] *= ip;
    __syncthreads();
    // eliminate column p from every other row: thread tile over (row, col)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for(int r = warp; r < n; r += nw) {
      if(r == p) continue;
      const double f = M[(size_t)r * ld + p];
      if(f == 0.) continue;
      for(int j = lane; j < ld; j += 32)
        if(j != p) M[(size_t)r * ld + j] -= f * M[(size_t)p * ld + j];
    }
    __syncthreads();
    for(int r = threadIdx.x; r < n; r += blockDim.x)
      if(r != p) M[(size_t)r * ld + p] = 0.;
    __syncthreads();
  }
  for(int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx - i * n;
    inv[idx] = M[(size_t)i * ld + n + j];
  }
}

__global__ void amg_dense_apply_kernel(int n, const double *__restrict__ inv, const double *__restrict__ b, double *__restrict__ x)
{
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
  for(int i = warp; i < n; i += nw) {
    double s = 0.;
    for(int j = lane; j < n; j += 32) s += inv[(size_t)i * n + j] * b[j];
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if(lane == 0) x[i] = s;
  }
}


// ----------------------------------------------------------------------------------------------------------
// global coarsest level (several GPUs)
// ----------------------------------------------------------------------------------------------------------
struct AmgChain {
  const int32_t *par[AMG_MAX_LEVELS];
  int            n;
};

// gid[v] = off + (coarsest aggregate of the level-1 unknown of vertex DOF v) for the owned vertex unknowns, -1 elsewhere
__global__ void amg_gc_gid_kernel(int64_t n, const uint8_t *__restrict__ pkind0, const int32_t *__restrict__ par00, AmgChain ch, int off,
                                  double *gid)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double g = -1.;
    if(pkind0[i] == 1) {
      int32_t cur = par00[i];
      for(int l = 0; l < ch.n && cur >= 0; ++l) cur = ch.par[l][cur];
      if(cur >= 0) g = (double)(off + cur);
    }
    gid[i] = g;
  }
}

// (global coarse id, weight) of every unknown of the field, owned or ghost, from the element tables
__global__ void amg_gc_parent_kernel(int64_t nElm, const int32_t *__restrict__ adr, int nloc, int nv, int nc, int dim, int64_t nInc,
                                     const uint8_t *__restrict__ fld_all, int fld_lo, int fld_hi, const double *__restrict__ gid, int32_t *p0,
                                     int32_t *p1, uint8_t *kind)
{
  const int     nS = nloc / nc;
  const int64_t tot = nElm * nloc;
  for(int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx / nloc;
    const int     k = (int)(idx - e * nloc);
    const int     a = k / nc, c = k - a * nc;
    const int64_t dof = adr[idx];
    if(dof >= nInc || fld_all[dof] < fld_lo || fld_all[dof] >= fld_hi) continue;
    if(a < nv) {
      p0[dof]   = (int32_t)gid[dof];
      p1[dof]   = -1;
      kind[dof] = 1;
    } else if(a < nS) {
      const int     ed = a - nv;
      const int     va = dim == 2 ? c_edge_tri[ed][0] : c_edge_tet[ed][0], vb = dim == 2 ? c_edge_tri[ed][1] : c_edge_tet[ed][1];
      const int64_t da = adr[e * nloc + va * nc + c], db = adr[e * nloc + vb * nc + c];
      int32_t       pa = da < nInc ? (int32_t)gid[da] : -1, pb = db < nInc ? (int32_t)gid[db] : -1;
      if(pa < pb) {
        const int32_t t = pa;
        pa = pb;
        pb = t;
      }
      p0[dof]   = pa;
      p1[dof]   = pb;
      kind[dof] = 2;
    }
  }
}

// triplets (global row, global column, value) of the rank-local level-gl operator
__global__ void amg_gl_local_triplets_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja,
                                             const double *__restrict__ val, int64_t off, int64_t ng, uint64_t *key, double *tv)
{
  const int     lane = threadIdx.x & 7;
  const int64_t g0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3, nn = (gridDim.x * (int64_t)blockDim.x) >> 3;
  for(int64_t i = g0; i < n; i += nn)
    for(int64_t k = ia[i] + lane; k < ia[i + 1]; k += 8) {
      key[k] = (uint64_t)(off + i) * (uint64_t)ng + (uint64_t)(off + ja[k]);
      tv[k]  = val[k];
    }
}

// Galerkin contributions w_iI w_jJ a_ij of the rows that read ghost columns, for the index pairs (I, J) with at least one
// REMOTE member (the pairs with two local members are in the rank-local operator already).  count != nullptr: count only.
__global__ void amg_gl_cut_triplets_kernel(int64_t n_rows, const int32_t *__restrict__ rows, const int64_t *__restrict__ ia,
                                           const int32_t *__restrict__ ja, const double *__restrict__ val, const uint8_t *__restrict__ fld_all,
                                           const uint8_t *__restrict__ pkind0, const int32_t *__restrict__ p0, const int32_t *__restrict__ p1,
                                           const uint8_t *__restrict__ kind, int64_t off, int64_t nloc, int64_t ng, unsigned long long *cursor,
                                           uint64_t *key, double *tv, int64_t cap, int count_only)
{
  const int     lane = threadIdx.x & 7;
  const int64_t g0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3, nn = (gridDim.x * (int64_t)blockDim.x) >> 3;
  for(int64_t t = g0; t < n_rows; t += nn) {
    const int64_t i = rows[t];
    if(pkind0[i] == 0) continue; // ghost row or other field
    const int32_t Ii[2] = {p0[i], p1[i]};
    const double  wi = kind[i] == 2 ? 0.5 : 1.;
    const int     fi = fld_all[i];
    for(int64_t k = ia[i] + lane; k < ia[i + 1]; k += 8) {
      const int32_t j = ja[k];
      if(kind[j] == 0 || fld_all[j] != fi) continue;
      const double  v = val[k] * wi * (kind[j] == 2 ? 0.5 : 1.);
      const int32_t Jj[2] = {p0[j], p1[j]};
#pragma unroll
      for(int a = 0; a < 2; ++a)
#pragma unroll
        for(int b = 0; b < 2; ++b) {
          const int64_t I = Ii[a], J = Jj[b];
          if(I < 0 || J < 0) continue;
          const bool Iloc = I >= off && I < off + nloc, Jloc = J >= off && J < off + nloc;
          if(Iloc && Jloc) continue;
          const unsigned long long pos = atomicAdd(cursor, 1ull);
          if(!count_only && (int64_t)pos < cap) {
            key[pos] = (uint64_t)I * (uint64_t)ng + (uint64_t)J;
            tv[pos]  = v;
          }
        }
    }
  }
}

__global__ void amg_fill_u64_kernel(int64_t n, uint64_t v, uint64_t *x)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}

// ----------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------
static void free_level(AmgLevel &L)
{
  cudaFree(L.ia_own);
  cudaFree(L.ja_own);
  cudaFree(L.val_own);
  cudaFree(L.dinv);
  cudaFree(L.par0);
  cudaFree(L.par1);
  cudaFree(L.pkind);
  cudaFree(L.x);
  cudaFree(L.b);
  cudaFree(L.t);
  cudaFree(L.d);
  cudaFree(L.ev);
  cudaFree(L.ia_s);
  cudaFree(L.ja_s);
  cudaFree(L.val_s);
  L = AmgLevel();
}

void amg_free(Amg *A)
{
  if(!A) return;
  for(auto &L : A->L) free_level(L);
  A->L.clear();
  cudaFree(A->dense);
  cudaFree(A->cinv);
  cudaFree(A->d_nrm);
  cudaFree(A->d_active0);
  cudaFree(A->gc_p0);
  cudaFree(A->gc_p1);
  cudaFree(A->gc_kind);
  cudaFree(A->gc_b);
  cudaFree(A->gc_x);
  cudaFree(A->g_ia);
  cudaFree(A->g_ja);
  cudaFree(A->g_val);
  cudaFree(A->t_key);
  cudaFree(A->t_keyall);
  cudaFree(A->t_val);
  cudaFree(A->t_valall);
  if(A->G) {
    amg_free(A->G);
    delete A->G;
    A->G = nullptr;
  }
  A->gc_p0 = A->gc_p1 = nullptr;
  A->gc_kind = nullptr;
  A->gc_b = A->gc_x = nullptr;
  A->g_ia = nullptr;
  A->g_ja = nullptr;
  A->g_val = nullptr;
  A->t_key = A->t_keyall = nullptr;
  A->t_val = A->t_valall = nullptr;
  A->t_cap = 0;
  A->gc_active = false;
  A->dense = A->cinv = A->d_nrm = nullptr;
  A->d_active0 = nullptr;
  A->symbolic = false;
}

static int csr_product(System *S, const AmgLevel &L, const double *x, double *y)
{
  if(L.n <= 0) return B200_OK;
  if(L.halo) {
    const int rc = comm_halo_exchange(S, const_cast<double *>(x));
    if(rc != B200_OK) return rc;
  }
  if(L.val_s) {
    // lanes per row / rows in flight per lane group: B200_AMG_SPMV=LR (e.g. 42, 82, 44) overrides the default
    static const int cfg = [] {
      const char *e = getenv("B200_AMG_SPMV");
      return e ? atoi(e) : 0;
    }();
    const double avg = (double)L.nnz_s / (double)L.n;
    const int    sel = cfg ? cfg : (avg > 40. ? 82 : 42);
#define B200_F32_SPMV(LL, RR)                                                                                                     \
  {                                                                                                                               \
    const int64_t blocks = (L.n * LL / RR + 255) / 256;                                                                           \
    amg_spmv_f32_kernel<LL, RR><<<(unsigned)std::min<int64_t>(blocks, 148 * 32), 256, 0, S->stream>>>(L.n, L.ia_s, L.ja_s, L.val_s, x, y); \
  }
    switch(sel) {
      case 82: B200_F32_SPMV(8, 2) break;
      case 84: B200_F32_SPMV(8, 4) break;
      case 44: B200_F32_SPMV(4, 4) break;
      case 24: B200_F32_SPMV(2, 4) break;
      case 22: B200_F32_SPMV(2, 2) break;
      case 162: B200_F32_SPMV(16, 2) break;
      default: B200_F32_SPMV(4, 2) break;
    }
#undef B200_F32_SPMV
    count_launch();
    return B200_OK;
  }
  if(L.is_system) return spmv(S, x, y);
  const double avg = (double)L.nnz / (double)L.n;
  const int64_t blocks8 = (L.n * 8 + 255) / 256, blocks2 = (L.n * 2 + 255) / 256;
  if(avg > 12.)
    amg_spmv_kernel<8><<<(unsigned)std::min<int64_t>(blocks8, 148 * 32), 256, 0, S->stream>>>(L.n, L.ia, L.ja, L.val, x, y);
  else
    amg_spmv_kernel<2><<<(unsigned)std::min<int64_t>(blocks2, 148 * 32), 256, 0, S->stream>>>(L.n, L.ia, L.ja, L.val, x, y);
  count_launch();
  return B200_OK;
}

static inline unsigned grid_for(int64_t n, int per_block = 256)
{
  const int64_t b = (n + per_block - 1) / per_block;
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(b, GRID));
}

// sorted unique keys (sentinel ~0 dropped) -> CSR of an n x n matrix
static int keys_to_csr(System *S, uint64_t *keys, int64_t nkeys, int64_t n, AmgLevel &C)
{
  auto pol = thrust::cuda::par.on(S->stream);
  thrust::device_ptr<uint64_t> kp(keys);
  // thrust::sort on 64-bit keys is a radix sort; chunking is not needed below 2^31 keys
  thrust::sort(pol, kp, kp + nkeys);
  int64_t nu = thrust::unique(pol, kp, kp + nkeys) - kp;
  if(nu > 0) {
    uint64_t last = 0;
    cudaMemcpyAsync(&last, keys + nu - 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, S->stream);
    cudaStreamSynchronize(S->stream);
    if(last == ~0ull) --nu;
  }
  C.n   = n;
  C.nnz = nu;
  B200_CUDA(cudaMalloc(&C.ia_own, (size_t)(n + 1) * sizeof(int64_t)));
  B200_CUDA(cudaMalloc(&C.ja_own, (size_t)std::max<int64_t>(nu, 1) * sizeof(int32_t)));
  B200_CUDA(cudaMalloc(&C.val_own, (size_t)std::max<int64_t>(nu, 1) * sizeof(double)));
  uint64_t *q = nullptr;
  B200_CUDA(cudaMalloc(&q, (size_t)(n + 1) * sizeof(uint64_t)));
  amg_row_start_keys_kernel<<<grid_for(n + 1), 256, 0, S->stream>>>(n, q);
  thrust::device_ptr<uint64_t> qp(q);
  thrust::lower_bound(pol, kp, kp + nu, qp, qp + n + 1, thrust::device_pointer_cast(C.ia_own));
  amg_split_keys_kernel<<<grid_for(nu), 256, 0, S->stream>>>(nu, n, keys, C.ja_own);
  count_launch(2);
  B200_CUDA(cudaStreamSynchronize(S->stream));
  cudaFree(q);
  C.ia  = C.ia_own;
  C.ja  = C.ja_own;
  C.val = C.val_own;
  return B200_OK;
}

static int alloc_vectors(AmgLevel &L, bool own_xb)
{
  const size_t nb = (size_t)std::max<int64_t>(L.n, 1) * sizeof(double);
  if(own_xb) {
    B200_CUDA(cudaMalloc(&L.x, nb));
    B200_CUDA(cudaMalloc(&L.b, nb));
  }
  B200_CUDA(cudaMalloc(&L.t, nb));
  B200_CUDA(cudaMalloc(&L.d, nb));
  B200_CUDA(cudaMalloc(&L.dinv, nb));
  return B200_OK;
}

// checks the hard-wired local edge tables against the P2 tabulation: with lambda_v = phi_v + 1/2 sum_{e contains v} phi_e
// every mid-edge function must equal 4 lambda_a lambda_b at every quadrature node
static bool check_p2_edges(const Space &sp, int dim, int nq)
{
  const int nv = dim + 1, ne = dim == 2 ? 3 : 6;
  if(sp.nS != nv + ne) return false;
  for(int k = 0; k < nq; ++k) {
    double lam[4] = {0., 0., 0., 0.};
    for(int v = 0; v < nv; ++v) lam[v] = sp.L[(size_t)k * sp.nS + v];
    for(int e = 0; e < ne; ++e) {
      const int a = dim == 2 ? h_edge_tri[e][0] : h_edge_tet[e][0], b = dim == 2 ? h_edge_tri[e][1] : h_edge_tet[e][1];
      lam[a] += 0.5 * sp.L[(size_t)k * sp.nS + nv + e];
      lam[b] += 0.5 * sp.L[(size_t)k * sp.nS + nv + e];
    }
    for(int e = 0; e < ne; ++e) {
      const int a = dim == 2 ? h_edge_tri[e][0] : h_edge_tet[e][0], b = dim == 2 ? h_edge_tri[e][1] : h_edge_tet[e][1];
      if(fabs(sp.L[(size_t)k * sp.nS + nv + e] - 4. * lam[a] * lam[b]) > 1e-10) return false;
    }
  }
  return true;
}

// aggregation of level `l` (symbolic): MIS roots, two rings of growth (the second finds nothing after MIS(1)), orphans as singletons
static int aggregate_level(System *S, Amg *A, int l, int64_t *ncoarse)
{
  AmgLevel     &L = A->L[l];
  const int64_t n = L.n;
  auto          pol = thrust::cuda::par.on(S->stream);
  uint64_t     *key = nullptr, *m1 = nullptr, *m2 = nullptr;
  int32_t      *tmp = nullptr, *tmp2 = nullptr;
  int          *d_cnt = nullptr;
  B200_CUDA(cudaMalloc(&key, n * sizeof(uint64_t)));
  B200_CUDA(cudaMalloc(&m1, n * sizeof(uint64_t)));
  B200_CUDA(cudaMalloc(&m2, n * sizeof(uint64_t)));
  B200_CUDA(cudaMalloc(&tmp, n * sizeof(int32_t)));
  B200_CUDA(cudaMalloc(&tmp2, n * sizeof(int32_t)));
  B200_CUDA(cudaMalloc(&d_cnt, sizeof(int)));
  B200_CUDA(cudaMalloc(&L.par0, n * sizeof(int32_t)));
  B200_CUDA(cudaMalloc(&L.pkind, n));
  const uint8_t *active = l == 0 ? A->d_active0 : nullptr;
  const double  *gval = l == 0 ? L.val : nullptr; // coarse values do not exist yet at symbolic time
  amg_mis_init_kernel<<<grid_for(n), 256, 0, S->stream>>>(n, active, key);
  count_launch();
  for(int round = 0; round < 200; ++round) {
    amg_mis_prop_kernel<<<grid_for(n), 256, 0, S->stream>>>(n, L.ia, L.ja, active, key, m1);
    if(A->mis_distance >= 2) amg_mis_prop_kernel<<<grid_for(n), 256, 0, S->stream>>>(n, L.ia, L.ja, active, m1, m2);
    B200_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int), S->stream));
    amg_mis_update_kernel<<<grid_for(n), 256, 0, S->stream>>>(n, key, A->mis_distance >= 2 ? m2 : m1, d_cnt);
    count_launch(3);
    int cnt = 0;
    B200_CUDA(cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, S->stream));
    B200_CUDA(cudaStreamSynchronize(S->stream));
    if(cnt == 0) break;
  }
  // number the roots
  amg_root_flag_kernel<<<grid_for(n), 256, 0, S->stream>>>(n, key, tmp);
  thrust::device_ptr<int32_t> tp(tmp), t2(tmp2);
  int32_t last_flag = 0, last_id = 0;
  thrust::exclusive_scan(pol, tp, tp + n, t2);
  B200_CUDA(cudaMemcpyAsync(&last_flag, tmp + n - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, S->stream));
  B200_CUDA(cudaMemcpyAsync(&last_id, tmp2 + n - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, S->stream));
  B200_CUDA(cudaStreamSynchronize(S->stream));
  int32_t nroots = last_flag + last_id;
  amg_agg_root_kernel<<<grid_for(n), 256, 0, S->stream>>>(n, key, tmp2, L.par0);
  // two rings (every unknown is within distance 2 of a root)
  amg_agg_grow_kernel<<<grid_for(n), 256, 0, S->stream>>>(n, L.ia, L.ja, gval, active, L.par0, tmp);
  amg_agg_grow_kernel<<<grid_for(n), 256, 0, S->stream>>>(n, L.ia, L.ja, gval, active, tmp, L.par0);
  count_launch(4);
  // orphans (cannot happen with a symmetric pattern; a non-symmetric one may leave some): singletons
  amg_orphan_flag_kernel<<<grid_for(n), 256, 0, S->stream>>>(n, active, L.par0, tmp);
  thrust::exclusive_scan(pol, tp, tp + n, t2);
  B200_CUDA(cudaMemcpyAsync(&last_flag, tmp + n - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, S->stream));
  B200_CUDA(cudaMemcpyAsync(&last_id, tmp2 + n - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, S->stream));
  B200_CUDA(cudaStreamSynchronize(S->stream));
  const int32_t norph = last_flag + last_id;
  if(norph > 0) amg_orphan_assign_kernel<<<grid_for(n), 256, 0, S->stream>>>(n, tmp, tmp2, nroots, L.par0);
  amg_kind_from_agg_kernel<<<grid_for(n), 256, 0, S->stream>>>(n, L.par0, L.pkind);
  count_launch(3);
  B200_CUDA(cudaStreamSynchronize(S->stream));
  *ncoarse = (int64_t)nroots + norph;
  cudaFree(key);
  cudaFree(m1);
  cudaFree(m2);
  cudaFree(tmp);
  cudaFree(tmp2);
  cudaFree(d_cnt);
  return B200_OK;
}

// aggregation levels below the last level built so far
static int build_coarse_levels(System *S, Amg *A)
{
  for(int l = (int)A->L.size() - 1; l < AMG_MAX_LEVELS - 1; ++l) {
    if(l > 0 && A->L[l].n <= AMG_MAX_DENSE) break;
    int64_t nc2 = 0;
    int     rc = aggregate_level(S, A, l, &nc2);
    if(rc != B200_OK) return rc;
    if(nc2 <= 0 || nc2 * 10 > A->L[l].n * 9) { // coarsening stalled: stop here
      cudaFree(A->L[l].par0);
      cudaFree(A->L[l].pkind);
      A->L[l].par0  = nullptr;
      A->L[l].pkind = nullptr;
      break;
    }
    A->L.emplace_back();
    AmgLevel &F = A->L[l];
    uint64_t *keys = nullptr;
    B200_CUDA(cudaMalloc(&keys, (size_t)std::max<int64_t>(F.nnz, 1) * sizeof(uint64_t)));
    amg_agg_keys_kernel<<<grid_for(F.n * 8), 256, 0, S->stream>>>(F.n, F.ia, F.ja, F.par0, nc2, keys);
    count_launch();
    rc = keys_to_csr(S, keys, F.nnz, nc2, A->L[l + 1]);
    cudaFree(keys);
    if(rc != B200_OK) return rc;
    rc = alloc_vectors(A->L[l + 1], true);
    if(rc != B200_OK) return rc;
  }
  return B200_OK;
}

// Symbolic set-up.  fld_lo..fld_hi-1 = the field ids (AmgFieldMap) of the rows the hierarchy acts on; space = the
// interpolation space of that field (P2 -> P1 level when it has mid-edge functions).
int amg_setup_symbolic(System *S, Amg *A, const uint8_t *d_fld, int fld_lo, int fld_hi, int space, const uint8_t *d_fld_all)
{
  amg_free(A);
  const int64_t n = S->nInc;
  const Space  &sp = S->spaces[space];
  const int     nv = S->nv, nloc = sp.nS * sp.nc;
  auto          pol = thrust::cuda::par.on(S->stream);
  A->verbose = getenv("B200_VERBOSE") != nullptr;
  if(const char *e = getenv("B200_AMG_DEGREE")) A->cheb_degree = std::max(1, atoi(e));
  if(const char *e = getenv("B200_AMG_PRE")) A->pre_degree = std::max(0, atoi(e));
  if(const char *e = getenv("B200_AMG_CYCLES")) A->cycles = std::max(1, atoi(e));
  if(const char *e = getenv("B200_AMG_RATIO")) A->cheb_ratio = std::max(1.5, atof(e));
  if(const char *e = getenv("B200_AMG_F32")) A->use_f32 = atoi(e) != 0;
  if(const char *e = getenv("B200_AMG_MIS")) A->mis_distance = atoi(e) == 1 ? 1 : 2;
  if(const char *e = getenv("B200_AMG_GAMMA")) A->gamma = std::max(1, atoi(e));
  if(const char *e = getenv("B200_AMG_OVERCORRECT")) A->overcorrect = std::max(0.1, atof(e));
  A->L.emplace_back();
  {
    AmgLevel &L0 = A->L[0];
    L0.n = n;
    L0.nnz = S->nnz;
    L0.ia = S->d_ia;
    L0.ja = S->d_ja;
    L0.val = S->d_val;
    L0.is_system = true;
    int rc = alloc_vectors(L0, false);
    if(rc != B200_OK) return rc;
  }
  A->d_fld = d_fld;
  B200_CUDA(cudaMalloc(&A->d_nrm, 4 * sizeof(double)));
  int64_t ncoarse = 0;
  const bool p2 = sp.nS > nv;
  if(p2) {
    if(!check_p2_edges(sp, S->dim, S->nq)) {
      set_error("amg: the basis tabulation is not the P2 Lagrange basis with the reference's local edge order");
      return B200_ERR_UNSUPP;
    }
    AmgLevel &L0 = A->L[0];
    int32_t  *isvert = nullptr, *cid = nullptr;
    B200_CUDA(cudaMalloc(&isvert, n * sizeof(int32_t)));
    B200_CUDA(cudaMalloc(&cid, n * sizeof(int32_t)));
    B200_CUDA(cudaMemsetAsync(isvert, 0, n * sizeof(int32_t), S->stream));
    amg_vertex_flag_kernel<<<GRID, 256, 0, S->stream>>>(S->nElm, sp.d_adr, nloc, nv * sp.nc, n, d_fld, fld_lo, fld_hi, isvert);
    thrust::device_ptr<int32_t> ip(isvert), cp(cid);
    thrust::exclusive_scan(pol, ip, ip + n, cp);
    int32_t lf = 0, li = 0;
    B200_CUDA(cudaMemcpyAsync(&lf, isvert + n - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, S->stream));
    B200_CUDA(cudaMemcpyAsync(&li, cid + n - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, S->stream));
    B200_CUDA(cudaStreamSynchronize(S->stream));
    ncoarse = (int64_t)lf + li;
    B200_CUDA(cudaMalloc(&L0.par0, n * sizeof(int32_t)));
    B200_CUDA(cudaMalloc(&L0.par1, n * sizeof(int32_t)));
    B200_CUDA(cudaMalloc(&L0.pkind, n));
    B200_CUDA(cudaMemsetAsync(L0.pkind, 0, n, S->stream));
    B200_CUDA(cudaMemsetAsync(L0.par0, 0xff, n * sizeof(int32_t), S->stream));
    B200_CUDA(cudaMemsetAsync(L0.par1, 0xff, n * sizeof(int32_t), S->stream));
    amg_p2_parent_kernel<<<GRID, 256, 0, S->stream>>>(S->nElm, sp.d_adr, nloc, nv, sp.nc, S->dim, n, d_fld, fld_lo, fld_hi, isvert, cid, L0.par0,
                                                      L0.par1, L0.pkind);
    count_launch(2);
    // coarse pattern = P1 element pattern of the vertex unknowns
    A->L.emplace_back();
    if(ncoarse > 0) {
      const int     nvc = nv * sp.nc;
      const int64_t nkeys = S->nElm * (int64_t)nvc * nvc;
      uint64_t     *keys = nullptr;
      B200_CUDA(cudaMalloc(&keys, (size_t)nkeys * sizeof(uint64_t)));
      amg_p1_keys_kernel<<<GRID * 2, 256, 0, S->stream>>>(S->nElm, sp.d_adr, nloc, nvc, sp.nc, n, isvert, cid, ncoarse, 1, keys);
      count_launch();
      int rc = keys_to_csr(S, keys, nkeys, ncoarse, A->L[1]);
      cudaFree(keys);
      if(rc != B200_OK) return rc;
    } else
      A->L[1].n = 0;
    A->L[0].decoupled = true;
    cudaFree(isvert);
    cudaFree(cid);
    int rc = alloc_vectors(A->L[1], true);
    if(rc != B200_OK) return rc;
  } else {
    // P1 space: level 0 is aggregated directly; the row mask of the graph is the field range
    B200_CUDA(cudaMalloc(&A->d_active0, n));
    amg_in_range_kernel<<<grid_for(n), 256, 0, S->stream>>>(n, d_fld, fld_lo, fld_hi, A->d_active0);
    count_launch();
    A->L[0].decoupled = true;
  }
  {
    const int rc = build_coarse_levels(S, A);
    if(rc != B200_OK) return rc;
  }
  // coarsest level
  const AmgLevel &C = A->L.back();
  A->dense_n = (C.n <= AMG_MAX_DENSE && !C.is_system) ? (int)C.n : 0;
  if(A->dense_n > 0) {
    B200_CUDA(cudaMalloc(&A->dense, (size_t)A->dense_n * 2 * A->dense_n * sizeof(double)));
    B200_CUDA(cudaMalloc(&A->cinv, (size_t)A->dense_n * A->dense_n * sizeof(double)));
  }
  // several GPUs: global hierarchy from level gl on (P2 hierarchies only; every rank must have that level)
  if(comm_active(S) && d_fld_all && p2 && !getenv("B200_AMG_LOCAL_COARSE")) {
    const int world = comm_world(S), rank = comm_rank(S);
    // first global level: 1 by default -- the P1 level is replicated too, so the hierarchy is that of the undecomposed operator at N
    // times the level-1 work per rank (measured, T3D(92) on 2 GPUs: 240 iterations / 4.24 s against 270 / 4.62 s with
    // B200_AMG_GLOBAL_LEVEL=2; T3D(48) on 4 GPUs: 224 / 0.71 s against 280 / 0.79 s)
    int gl_first = 1;
    if(const char *e = getenv("B200_AMG_GLOBAL_LEVEL")) gl_first = std::max(1, atoi(e));
    const int gl = std::min(gl_first, (int)A->L.size() - 1);
    std::vector<double> cnt(world + 2, 0.);
    cnt[rank]      = gl >= 1 ? (double)A->L[gl].n : 0.;
    cnt[world]     = gl >= 1 ? 0. : 1.; // somebody without a coarse level: no global level
    cnt[world + 1] = (double)gl;        // (summed: must be world * gl)
    double *d_cnt = nullptr;
    B200_CUDA(cudaMalloc(&d_cnt, (world + 2) * sizeof(double)));
    B200_CUDA(cudaMemcpyAsync(d_cnt, cnt.data(), (world + 2) * sizeof(double), cudaMemcpyHostToDevice, S->stream));
    int rc = comm_allreduce(S, d_cnt, world + 2, false);
    if(rc != B200_OK) return rc;
    B200_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, (world + 2) * sizeof(double), cudaMemcpyDeviceToHost, S->stream));
    B200_CUDA(cudaStreamSynchronize(S->stream));
    cudaFree(d_cnt);
    int64_t tot = 0, off = 0;
    for(int r = 0; r < world; ++r) {
      if(r == rank) off = tot;
      tot += (int64_t)cnt[r];
    }
    if(cnt[world] == 0. && cnt[world + 1] == (double)world * gl && tot > 0 && tot < (int64_t)2000000000) {
      A->gl      = gl;
      A->gc_n    = tot;
      A->gc_off  = off;
      A->gc_nloc = A->L[gl].n;
      // global level-gl id of every owned vertex unknown, then of the ghost ones through the halo plan of the fine vectors
      double *gid = nullptr;
      B200_CUDA(cudaMalloc(&gid, (size_t)n * sizeof(double)));
      AmgChain ch;
      ch.n = 0;
      for(int l = 1; l < gl; ++l) ch.par[ch.n++] = A->L[l].par0;
      amg_gc_gid_kernel<<<grid_for(n), 256, 0, S->stream>>>(n, A->L[0].pkind, A->L[0].par0, ch, (int)off, gid);
      count_launch();
      rc = comm_halo_exchange(S, gid);
      if(rc != B200_OK) return rc;
      B200_CUDA(cudaMalloc(&A->gc_p0, (size_t)n * sizeof(int32_t)));
      B200_CUDA(cudaMalloc(&A->gc_p1, (size_t)n * sizeof(int32_t)));
      B200_CUDA(cudaMalloc(&A->gc_kind, (size_t)n));
      B200_CUDA(cudaMemsetAsync(A->gc_p0, 0xff, (size_t)n * sizeof(int32_t), S->stream));
      B200_CUDA(cudaMemsetAsync(A->gc_p1, 0xff, (size_t)n * sizeof(int32_t), S->stream));
      B200_CUDA(cudaMemsetAsync(A->gc_kind, 0, (size_t)n, S->stream));
      amg_gc_parent_kernel<<<GRID, 256, 0, S->stream>>>(S->nElm, sp.d_adr, nloc, nv, sp.nc, S->dim, n, d_fld_all, fld_lo, fld_hi, gid, A->gc_p0,
                                                        A->gc_p1, A->gc_kind);
      count_launch();
      B200_CUDA(cudaStreamSynchronize(S->stream));
      cudaFree(gid);
      B200_CUDA(cudaMalloc(&A->gc_b, (size_t)tot * sizeof(double)));
      B200_CUDA(cudaMalloc(&A->gc_x, (size_t)tot * sizeof(double)));
      A->gc_active = true;
      A->d_fld_all = d_fld_all;
    }
  }
  // single-precision working copy of level 0 (pattern now, values in the numeric phase)
  if(A->use_f32 && A->L[0].pkind) {
    AmgLevel &L0 = A->L[0];
    // several GPUs with the global hierarchy: the rows of level 0 keep their couplings to ghost columns and every level-0 product
    // is preceded by the halo update of its input -- the fine-level smoother and residual are those of the undecomposed operator
    const uint8_t *fcol = A->gc_active ? A->d_fld_all : d_fld, *kcol = A->gc_active ? A->gc_kind : L0.pkind;
    L0.halo = A->gc_active && !getenv("B200_AMG_LOCAL_SMOOTHER");
    if(!L0.halo) {
      fcol = d_fld;
      kcol = L0.pkind;
    }
    A->d_fcol0 = fcol;
    A->d_kcol0 = kcol;
    int64_t  *cnt = nullptr;
    B200_CUDA(cudaMalloc(&cnt, (size_t)(n + 1) * sizeof(int64_t)));
    B200_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(n + 1) * sizeof(int64_t), S->stream));
    amg_dec_count_kernel<<<GRID * 4, 256, 0, S->stream>>>(n, L0.ia, L0.ja, d_fld, L0.pkind, fcol, kcol, cnt);
    count_launch();
    thrust::device_ptr<int64_t> cp(cnt);
    thrust::exclusive_scan(pol, cp, cp + n + 1, cp);
    B200_CUDA(cudaMemcpyAsync(&L0.nnz_s, cnt + n, sizeof(int64_t), cudaMemcpyDeviceToHost, S->stream));
    B200_CUDA(cudaStreamSynchronize(S->stream));
    L0.ia_s = cnt;
    B200_CUDA(cudaMalloc(&L0.ja_s, (size_t)std::max<int64_t>(L0.nnz_s, 1) * sizeof(int32_t)));
    B200_CUDA(cudaMalloc(&L0.val_s, (size_t)std::max<int64_t>(L0.nnz_s, 1) * sizeof(float)));
    amg_dec_fill_kernel<true><<<GRID * 4, 256, 0, S->stream>>>(n, L0.ia, L0.ja, L0.val, d_fld, L0.pkind, fcol, kcol, L0.ia_s, L0.ja_s, L0.val_s);
    count_launch();
  }
  if(A->verbose) {
    if(A->gc_active)
      fprintf(stderr, "[feng_b200] amg: global hierarchy from level %d on, %lld unknowns (%lld local at offset %lld)\n", A->gl, (long long)A->gc_n,
              (long long)A->gc_nloc, (long long)A->gc_off);
    fprintf(stderr, "[feng_b200] amg hierarchy:");
    for(auto &L : A->L) fprintf(stderr, " (%lld rows, %lld nnz)", (long long)L.n, (long long)L.nnz);
    fprintf(stderr, " dense %d\n", A->dense_n);
  }
  A->symbolic = true;
  return B200_OK;
}

// spectral radius of D^-1 A by power iteration (warm-started from the previous set-up's vector)
static int power_iteration(System *S, Amg *A, int l, int iters)
{
  AmgLevel &L = A->L[l];
  L.lam = 1.;
  if(L.n <= 0) return B200_OK;
  if(!L.ev) B200_CUDA(cudaMalloc(&L.ev, (size_t)L.n * sizeof(double)));
  if(!L.have_eig) {
    amg_seed_kernel<<<grid_for(L.n), 256, 0, S->stream>>>(L.n, L.dinv, L.ev);
    B200_CUDA(cudaMemsetAsync(A->d_nrm, 0, sizeof(double), S->stream));
    amg_norm2_kernel<<<grid_for(L.n), 256, 0, S->stream>>>(L.n, L.ev, A->d_nrm);
    amg_normalize_kernel<<<grid_for(L.n), 256, 0, S->stream>>>(L.n, A->d_nrm, L.ev);
    count_launch(3);
  }
  for(int it = 0; it < iters; ++it) {
    int rc = csr_product(S, L, L.ev, L.t);
    if(rc != B200_OK) return rc;
    B200_CUDA(cudaMemsetAsync(A->d_nrm, 0, sizeof(double), S->stream));
    amg_scale_norm_kernel<<<grid_for(L.n), 256, 0, S->stream>>>(L.n, L.dinv, L.t, L.ev, A->d_nrm); // ev = D^-1 A ev, |ev|^2
    if(it == iters - 1) B200_CUDA(cudaMemcpyAsync(S->h_scratch, A->d_nrm, sizeof(double), cudaMemcpyDeviceToHost, S->stream));
    amg_normalize_kernel<<<grid_for(L.n), 256, 0, S->stream>>>(L.n, A->d_nrm, L.ev);
    count_launch(2);
  }
  B200_CUDA(cudaStreamSynchronize(S->stream));
  const double lam = sqrt(S->h_scratch[0]);
  L.have_eig = true;
  L.lam = (lam > 0. && std::isfinite(lam)) ? 1.1 * lam : 1.;
  return B200_OK;
}

int amg_setup_numeric(System *S, Amg *A);

// hierarchy on a CSR matrix given as is (every row active): level 0 aliases the arrays
static int amg_setup_symbolic_csr(System *S, Amg *G, int64_t n, int64_t nnz, const int64_t *ia, const int32_t *ja, const double *val)
{
  amg_free(G);
  G->verbose = getenv("B200_VERBOSE") != nullptr;
  G->use_f32 = false;
  G->L.emplace_back();
  AmgLevel &L0 = G->L[0];
  L0.n = n;
  L0.nnz = nnz;
  L0.ia = ia;
  L0.ja = ja;
  L0.val = val;
  int rc = alloc_vectors(L0, false);
  if(rc != B200_OK) return rc;
  B200_CUDA(cudaMalloc(&G->d_nrm, 4 * sizeof(double)));
  rc = build_coarse_levels(S, G);
  if(rc != B200_OK) return rc;
  const AmgLevel &C = G->L.back();
  G->dense_n = (C.n <= AMG_MAX_DENSE && G->L.size() > 1) ? (int)C.n : 0;
  if(G->dense_n > 0) {
    B200_CUDA(cudaMalloc(&G->dense, (size_t)G->dense_n * 2 * G->dense_n * sizeof(double)));
    B200_CUDA(cudaMalloc(&G->cinv, (size_t)G->dense_n * G->dense_n * sizeof(double)));
  }
  if(G->verbose) {
    fprintf(stderr, "[feng_b200] amg global hierarchy:");
    for(auto &L : G->L) fprintf(stderr, " (%lld rows, %lld nnz)", (long long)L.n, (long long)L.nnz);
    fprintf(stderr, " dense %d\n", G->dense_n);
  }
  G->symbolic = true;
  return B200_OK;
}

// Merged level-gl operator of all ranks (values), its hierarchy (symbolic once, numeric every time)
static int global_level_numeric(System *S, Amg *A)
{
  auto            pol = thrust::cuda::par.on(S->stream);
  const int       world = comm_world(S);
  const AmgLevel &L0 = A->L[0], &Lg = A->L[A->gl];
  const int32_t  *brows = nullptr;
  int64_t         nb = 0;
  comm_boundary_rows(S, &brows, &nb);
  unsigned long long *cursor = reinterpret_cast<unsigned long long *>(A->d_nrm + 2);
  const unsigned      gb = (unsigned)std::max<int64_t>(1, std::min<int64_t>((nb * 8 + 255) / 256, 148 * 16));
  if(A->t_cap == 0) {
    // capacity: own level-gl entries + cut contributions (counted), the maximum over the ranks
    B200_CUDA(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), S->stream));
    if(nb > 0) {
      amg_gl_cut_triplets_kernel<<<gb, 256, 0, S->stream>>>(nb, brows, L0.ia, L0.ja, L0.val, A->d_fld_all, L0.pkind, A->gc_p0, A->gc_p1, A->gc_kind,
                                                            A->gc_off, A->gc_nloc, A->gc_n, cursor, nullptr, nullptr, 0, 1);
      count_launch();
    }
    unsigned long long ncut = 0;
    B200_CUDA(cudaMemcpyAsync(&ncut, cursor, sizeof(unsigned long long), cudaMemcpyDeviceToHost, S->stream));
    B200_CUDA(cudaStreamSynchronize(S->stream));
    double h = (double)(Lg.nnz + (int64_t)ncut);
    B200_CUDA(cudaMemcpyAsync(S->d_scratch, &h, sizeof(double), cudaMemcpyHostToDevice, S->stream));
    int rc = comm_allreduce(S, S->d_scratch, 1, true);
    if(rc != B200_OK) return rc;
    B200_CUDA(cudaMemcpyAsync(&h, S->d_scratch, sizeof(double), cudaMemcpyDeviceToHost, S->stream));
    B200_CUDA(cudaStreamSynchronize(S->stream));
    A->t_cap = (int64_t)h;
    B200_CUDA(cudaMalloc(&A->t_key, (size_t)A->t_cap * sizeof(uint64_t)));
    B200_CUDA(cudaMalloc(&A->t_val, (size_t)A->t_cap * sizeof(double)));
    B200_CUDA(cudaMalloc(&A->t_keyall, (size_t)A->t_cap * world * sizeof(uint64_t)));
    B200_CUDA(cudaMalloc(&A->t_valall, (size_t)A->t_cap * world * sizeof(double)));
  }
  // own triplets: [local level-gl entries | cut contributions | padding with the sentinel key]
  amg_fill_u64_kernel<<<grid_for(A->t_cap), 256, 0, S->stream>>>(A->t_cap, ~0ull, A->t_key);
  B200_CUDA(cudaMemsetAsync(A->t_val, 0, (size_t)A->t_cap * sizeof(double), S->stream));
  if(Lg.n > 0) amg_gl_local_triplets_kernel<<<grid_for(Lg.n * 8), 256, 0, S->stream>>>(Lg.n, Lg.ia, Lg.ja, Lg.val, A->gc_off, A->gc_n, A->t_key, A->t_val);
  B200_CUDA(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), S->stream));
  if(nb > 0)
    amg_gl_cut_triplets_kernel<<<gb, 256, 0, S->stream>>>(nb, brows, L0.ia, L0.ja, L0.val, A->d_fld_all, L0.pkind, A->gc_p0, A->gc_p1, A->gc_kind,
                                                          A->gc_off, A->gc_nloc, A->gc_n, cursor, A->t_key + Lg.nnz, A->t_val + Lg.nnz,
                                                          A->t_cap - Lg.nnz, 0);
  count_launch(3);
  int rc = comm_allgather64(S, A->t_key, A->t_keyall, (size_t)A->t_cap);
  if(rc != B200_OK) return rc;
  rc = comm_allgather64(S, A->t_val, A->t_valall, (size_t)A->t_cap);
  if(rc != B200_OK) return rc;
  // merge: sort by key, sum duplicates
  const int64_t nt = A->t_cap * world;
  thrust::device_ptr<uint64_t> kp(A->t_keyall);
  thrust::device_ptr<double>   vp(A->t_valall);
  thrust::stable_sort_by_key(pol, kp, kp + nt, vp);
  uint64_t *uk = nullptr;
  double   *uv = nullptr;
  B200_CUDA(cudaMalloc(&uk, (size_t)nt * sizeof(uint64_t)));
  B200_CUDA(cudaMalloc(&uv, (size_t)nt * sizeof(double)));
  auto    ends = thrust::reduce_by_key(pol, kp, kp + nt, vp, thrust::device_pointer_cast(uk), thrust::device_pointer_cast(uv));
  int64_t nu = ends.first - thrust::device_pointer_cast(uk);
  if(nu > 0) {
    uint64_t last = 0;
    B200_CUDA(cudaMemcpyAsync(&last, uk + nu - 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, S->stream));
    B200_CUDA(cudaStreamSynchronize(S->stream));
    if(last == ~0ull) --nu;
  }
  const bool first = A->G == nullptr;
  if(first) {
    A->g_nnz = nu;
    B200_CUDA(cudaMalloc(&A->g_ia, (size_t)(A->gc_n + 1) * sizeof(int64_t)));
    B200_CUDA(cudaMalloc(&A->g_ja, (size_t)std::max<int64_t>(nu, 1) * sizeof(int32_t)));
    B200_CUDA(cudaMalloc(&A->g_val, (size_t)std::max<int64_t>(nu, 1) * sizeof(double)));
    uint64_t *q = nullptr;
    B200_CUDA(cudaMalloc(&q, (size_t)(A->gc_n + 1) * sizeof(uint64_t)));
    amg_row_start_keys_kernel<<<grid_for(A->gc_n + 1), 256, 0, S->stream>>>(A->gc_n, q);
    thrust::device_ptr<uint64_t> qp(q), ukp(uk);
    thrust::lower_bound(pol, ukp, ukp + nu, qp, qp + A->gc_n + 1, thrust::device_pointer_cast(A->g_ia));
    amg_split_keys_kernel<<<grid_for(nu), 256, 0, S->stream>>>(nu, A->gc_n, uk, A->g_ja);
    count_launch(2);
    B200_CUDA(cudaStreamSynchronize(S->stream));
    cudaFree(q);
  } else if(nu != A->g_nnz) {
    cudaFree(uk);
    cudaFree(uv);
    set_error("amg: the merged global level changed its pattern between two numeric set-ups");
    return B200_ERR_ARG;
  }
  B200_CUDA(cudaMemcpyAsync(A->g_val, uv, (size_t)nu * sizeof(double), cudaMemcpyDeviceToDevice, S->stream));
  B200_CUDA(cudaStreamSynchronize(S->stream));
  cudaFree(uk);
  cudaFree(uv);
  if(first) {
    A->G = new Amg;
    A->G->mis_distance = A->mis_distance;
    rc = amg_setup_symbolic_csr(S, A->G, A->gc_n, A->g_nnz, A->g_ia, A->g_ja, A->g_val);
    if(rc != B200_OK) return rc;
    A->G->cheb_degree = A->cheb_degree;
    A->G->gamma       = A->gamma;
    A->G->gamma_from  = 0;
    A->G->overcorrect = A->overcorrect;
    A->G->cheb_ratio  = A->cheb_ratio;
  }
  return amg_setup_numeric(S, A->G);
}

// Numeric set-up: Galerkin values level by level, inverse diagonals, spectral radii, coarsest inverse.
int amg_setup_numeric(System *S, Amg *A)
{
  if(!A->symbolic) {
    set_error("amg: symbolic set-up missing");
    return B200_ERR_ARG;
  }
  const int nl = (int)A->L.size();
  if(A->L[0].val_s) {
    AmgLevel &L0 = A->L[0];
    amg_dec_fill_kernel<false><<<GRID * 4, 256, 0, S->stream>>>(L0.n, L0.ia, L0.ja, L0.val, A->d_fld, L0.pkind, A->d_fcol0, A->d_kcol0, L0.ia_s, nullptr,
                                                               L0.val_s);
    count_launch();
  }
  for(int l = 0; l < nl; ++l) {
    AmgLevel &L = A->L[l];
    if(L.n <= 0) continue;
    // inverse diagonal of the active rows
    const uint8_t *mask = L.pkind;
    if(l == 0 && !L.pkind) mask = A->d_active0; // single-level corner case
    amg_dinv_kernel<<<grid_for(L.n), 256, 0, S->stream>>>(L.n, L.ia, L.ja, L.val, l == 0 ? mask : nullptr, L.dinv);
    count_launch();
    if(l + 1 < nl) {
      AmgLevel &C = A->L[l + 1];
      if(C.n > 0) {
        B200_CUDA(cudaMemsetAsync(C.val_own, 0, (size_t)C.nnz * sizeof(double), S->stream));
        const int64_t blocks = (L.n * 8 + 255) / 256;
        const unsigned gb = (unsigned)std::min<int64_t>(blocks, 148 * 32);
        if(L.val_s)
          amg_galerkin_kernel<8, float><<<gb, 256, 0, S->stream>>>(L.n, L.ia_s, L.ja_s, L.val_s, nullptr, L.par0, L.par1, L.pkind, C.ia, C.ja,
                                                                   C.val_own);
        else
          amg_galerkin_kernel<8, double><<<gb, 256, 0, S->stream>>>(L.n, L.ia, L.ja, L.val, (l == 0 && L.decoupled) ? A->d_fld : nullptr, L.par0,
                                                                    L.par1, L.pkind, C.ia, C.ja, C.val_own);
        count_launch();
      }
    }
    B200_CUDA(cudaGetLastError());
  }
  for(int l = 0; l < nl; ++l) {
    if(l == nl - 1 && A->dense_n > 0) break;
    const bool warm = A->L[l].have_eig;
    int        rc = power_iteration(S, A, l, warm ? 3 : (l == 0 ? 8 : 12));
    if(rc != B200_OK) return rc;
  }
  if(A->gc_active) {
    const int rc = global_level_numeric(S, A);
    if(rc != B200_OK) return rc;
  }
  if(A->dense_n > 0) {
    const AmgLevel &C = A->L.back();
    amg_dense_fill_kernel<<<std::min(A->dense_n, 148 * 4), 128, 0, S->stream>>>(A->dense_n, C.ia, C.ja, C.val, A->dense);
    amg_gauss_jordan_kernel<<<1, 1024, 0, S->stream>>>(A->dense_n, A->dense, A->cinv);
    count_launch(2);
  }
  B200_CUDA(cudaGetLastError());
  if(A->verbose) {
    cudaStreamSynchronize(S->stream);
    fprintf(stderr, "[feng_b200] amg numeric: lambda_max(D^-1 A) =");
    for(auto &L : A->L) fprintf(stderr, " %.3f", L.lam);
    fprintf(stderr, "\n");
  }
  return B200_OK;
}

// Chebyshev smoother of degree AMG_CHEB_DEGREE on [lam / AMG_CHEB_RATIO, lam]; zero_guess: x is not read
static int smooth(System *S, Amg *A, int l, const double *b, double *x, bool zero_guess, int degree = 0)
{
  if(degree <= 0) degree = A->cheb_degree;
  AmgLevel    &L = A->L[l];
  const double lmax = L.lam, lmin = L.lam / A->cheb_ratio;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  double       rho = 1. / sigma;
  for(int k = 0; k < degree; ++k) {
    const bool first = k == 0;
    const double *t = nullptr;
    if(!(first && zero_guess)) {
      int rc = csr_product(S, L, x, L.t);
      if(rc != B200_OK) return rc;
      t = L.t;
    }
    double c1, c2;
    if(first) {
      c1 = 0.;
      c2 = 1. / theta;
    } else {
      const double rho_n = 1. / (2. * sigma - rho);
      c1 = rho_n * rho;
      c2 = 2. * rho_n / delta;
      rho = rho_n;
    }
    amg_cheb_kernel<<<grid_for(L.n), 256, 0, S->stream>>>(L.n, b, t, L.dinv, c1, c2, L.d, x, (first && zero_guess) ? 1 : 0);
    count_launch();
  }
  return B200_OK;
}

int amg_vcycle(System *S, Amg *A, const double *b, double *x);

static int cycle(System *S, Amg *A, int l, const double *b, double *x, bool zero_guess = true)
{
  AmgLevel &L = A->L[l];
  const int nl = (int)A->L.size();
  if(A->gc_active && l == A->gl) {
    // global level: every rank contributes its slice of the right-hand side, runs the replicated hierarchy, keeps its slice
    B200_CUDA(cudaMemsetAsync(A->gc_b, 0, (size_t)A->gc_n * sizeof(double), S->stream));
    if(A->gc_nloc > 0)
      B200_CUDA(cudaMemcpyAsync(A->gc_b + A->gc_off, b, (size_t)A->gc_nloc * sizeof(double), cudaMemcpyDeviceToDevice, S->stream));
    int rc = comm_allreduce(S, A->gc_b, (int)A->gc_n, false);
    if(rc != B200_OK) return rc;
    rc = amg_vcycle(S, A->G, A->gc_b, A->gc_x);
    if(rc != B200_OK) return rc;
    if(A->gc_nloc > 0)
      B200_CUDA(cudaMemcpyAsync(x, A->gc_x + A->gc_off, (size_t)A->gc_nloc * sizeof(double), cudaMemcpyDeviceToDevice, S->stream));
    return B200_OK;
  }
  if(L.n <= 0) return B200_OK;
  if(l == nl - 1) {
    if(A->dense_n > 0) {
      amg_dense_apply_kernel<<<std::max(1, std::min((A->dense_n + 7) / 8, 148 * 4)), 256, 0, S->stream>>>(A->dense_n, A->cinv, b, x);
      count_launch();
      return B200_OK;
    }
    // no dense solve (single level, or coarsening stalled above the dense limit): a few smoothing sweeps
    int rc = smooth(S, A, l, b, x, true);
    for(int s = 0; s < 2 && rc == B200_OK; ++s) rc = smooth(S, A, l, b, x, false);
    return rc;
  }
  AmgLevel &C = A->L[l + 1];
  int       rc = smooth(S, A, l, b, x, zero_guess, A->pre_degree);
  if(rc != B200_OK) return rc;
  if(C.n > 0) {
    rc = csr_product(S, L, x, L.t);
    if(rc != B200_OK) return rc;
    B200_CUDA(cudaMemsetAsync(C.b, 0, (size_t)C.n * sizeof(double), S->stream));
    amg_restrict_kernel<<<grid_for(L.n), 256, 0, S->stream>>>(L.n, b, L.t, L.par0, L.par1, L.pkind, C.b);
    count_launch();
    rc = cycle(S, A, l + 1, C.b, C.x);
    // cycle index gamma (2 = W-cycle) below the finest level; the global replicated levels and the dense level are visited once
    for(int v = 1; v < A->gamma && rc == B200_OK && l >= A->gamma_from && l + 1 < nl - 1 && !(A->gc_active && l + 1 == A->gl); ++v) rc = cycle(S, A, l + 1, C.b, C.x, false);
    if(rc != B200_OK) return rc;
    // plain aggregation under-estimates the coarse correction: over-correction factor on the aggregation levels (not on the P2 -> P1 step)
    const double w = (l >= 1 || !L.par1) ? A->overcorrect : 1.;
    amg_prolong_kernel<<<grid_for(L.n), 256, 0, S->stream>>>(L.n, C.x, L.par0, L.par1, L.pkind, w, x);
    count_launch();
  }
  return smooth(S, A, l, b, x, false);
}

// x = one multigrid cycle applied to b (both n-vectors of the system; inactive rows of x come out 0)
int amg_vcycle(System *S, Amg *A, const double *b, double *x)
{
  int rc = cycle(S, A, 0, b, x, true);
  for(int c = 1; c < A->cycles && rc == B200_OK; ++c) rc = cycle(S, A, 0, b, x, false);
  if(rc != B200_OK) return rc;
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int amg_build_field_map(System *S, uint8_t **out)
{
  const int64_t n = S->nInc;
  uint8_t      *fld = nullptr;
  B200_CUDA(cudaMalloc(&fld, n));
  B200_CUDA(cudaMemsetAsync(fld, AMG_FLD_NONE, n, S->stream));
  const Space &U = S->spaces[S->su];
  amg_field_kernel<<<GRID, 256, 0, S->stream>>>(S->nElm, U.d_adr, U.nS * U.nc, U.nc, n, 0, fld);
  count_launch();
  if(S->sp >= 0) {
    const Space &P = S->spaces[S->sp];
    amg_field_kernel<<<GRID, 256, 0, S->stream>>>(S->nElm, P.d_adr, P.nS * P.nc, P.nc, n, AMG_FLD_P, fld);
    count_launch();
  }
  B200_CUDA(cudaGetLastError());
  *out = fld;
  return B200_OK;
}

// ghost rows (owned by another rank) leave every field: the hierarchy and the pressure scaling are rank-local
int amg_mask_ghosts(System *S, uint8_t *fld)
{
  const double *owned = comm_mask(S);
  if(owned) {
    amg_mask_ghost_kernel<<<GRID, 256, 0, S->stream>>>(S->nInc, owned, fld);
    count_launch();
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

} // namespace b200
