// Row-owner ("gather") assembly with pre-contracted reference tensors.
//
// Same numbers as the reference's per-form quadrature loops (feSysElm_*::computeAe/computeBe, src/feVectorSysElm.cpp,
// restated in oracle/fe_oracle.py) and the colour-ordered scatter (src/feLinearSystemMklPardiso.cpp:501-749), but
// organised for the GPU memory system:
//
//  * Straight simplices + constant coefficients: every entry of the fused element system is
//        sum_k w_k (products of reference-basis tables at node k) x (geometry factors constant on the element)
//        x (local DOFs for the convective terms).
//    The k-sums are moved into constant tables built ONCE from the tables the host passes (any quadrature rule):
//        Kref[a][b][al][be] = sum_k w_k dphi_a/dxi_al dphi_b/dxi_be        (viscous / diffusion blocks)
//        Mref[a][b]         = sum_k w_k phi_a phi_b                        (transient mass)
//        Bref[q][a][al]     = sum_k w_k psi_q dphi_a/dxi_al                (pressure gradient / divergence blocks)
//        T3[a][b][v]        = sum_k w_k phi_a phi_b psi_v                  (convection)
//        E[c][al][v]        : dphi_c/dxi_al (xi_k) = sum_v psi_v(xi_k) E[c][al][v]  (gradients of P2 functions are
//                             P1: verified on the host at plan time, otherwise the quadrature kernel is used)
//    This is an exact re-association of the reference's sums (differences are rounding, ~1e-15 relative).
//  * Each CSR row is produced by ONE thread from the elements adjacent to its DOF ("row owner"): contributions are
//    accumulated in a shared-memory copy of the row and written to HBM exactly once with coalesced stores -- no
//    memset, no atomics, deterministic summation order (ascending element index), 16-bit row-local column offsets
//    instead of a 32-bit CSR slot per local entry.
#include <thrust/device_vector.h>
#include <thrust/execution_policy.h>
#include <thrust/extrema.h>
#include <thrust/scan.h>
#include <thrust/sort.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

#include "device_common.cuh"
#include "system.h"

namespace b200 {


struct NodeSet {
  int32_t   nNodes = 0, nCta = 0;
  int64_t   nPairs = 0;
  int32_t  *pair = nullptr;     // [nPairs] e * nLoc + local index, sorted by node then element
  int2     *range = nullptr;    // [nNodes] (first pair, number of pairs)
  int32_t  *row = nullptr;      // [nNodes][nRow] global rows (>= nInc: essential, not assembled)
  uint32_t *smoff = nullptr;    // [nNodes] offset (doubles) of the node's row buffer inside its CTA's shared memory
  uint32_t *cta_size = nullptr; // [nCta] doubles of row buffer per CTA
  int64_t  *cta_g0 = nullptr;   // [nCta] first CSR slot of the CTA's rows if they are consecutive in memory, else -1
  int32_t  *cta_perm = nullptr; // [nCta] launch order inside every segment: CTAs sorted along a Morton curve through the mesh
  std::vector<int32_t>  seg_begin; // launch segments: CTAs with similar row-buffer sizes share one launch
  std::vector<uint32_t> seg_smem;  // doubles of row buffer for the segment
  std::vector<double>   seg_pairs; // average number of adjacent elements per node in the segment
  uint16_t *off = nullptr;      // [nPairs][OFFW] row-local column offsets, 0xFFFF = not assembled
  uint32_t  max_cta = 0;
  void release()
  {
    cudaFree(pair);
    cudaFree(range);
    cudaFree(row);
    cudaFree(smoff);
    cudaFree(cta_size);
    cudaFree(cta_g0);
    cudaFree(cta_perm);
    cudaFree(off);
    *this = NodeSet();
  }
};

struct GatherPlan {
  NodeSet  U, P;
  double  *d_tab = nullptr;
  double  *d_geo = nullptr; // [nElm][GW] inverse affine map + detJ of every element (the mesh is static)
  double   E[120] = {0.};
  int      tab_len = 0, tab_len_src = 0;
  int      npbU = 0, npbP = 0;
  int      lanes = 1;       // lanes per node of the lane-group kernels
  bool     patch = true;    // patch kernel (patch.cu) instead of the row-owner kernels of this file
  bool     lane = true;     // lane-per-column kernels (gather_lane.cuh) or thread-per-node kernels (B200_GATHER_KERNEL=node)
  double  *d_es = nullptr;  // [nElm][ES::W] per-element convective state, rewritten by every assembly pass
  // row-lane velocity kernels (gather_urow.cuh): P2/P1 tetrahedra
  bool      urow = false;
  double   *d_geo4 = nullptr;  // [nElm][4][4] gradients of the barycentric coordinates + detJ
  int32_t  *d_order = nullptr; // velocity nodes sorted by (row length, signature)
  uint64_t *d_lacnt = nullptr; // [node] pairs per local index, 6 bits each
  void     *d_rec = nullptr;   // [nPairs] URowPair
  int32_t  *d_wstep = nullptr; // [warp + 1] first schedule step of every warp (10 nodes)
  int32_t  *d_sched = nullptr; // [step][10] pair of group g at this step, or -1
  uint64_t *d_sla = nullptr;   // [warp] steps the warp spends on every local index, 6 bits each
  std::vector<double>   utab;  // [10][URowTab::LEN] reference tensors per local row node
  std::vector<double>   estab; // [ESTab::LEN] reference tensors of the pre-pass
  std::vector<int32_t>  useg_begin; // launch segments (in warps of 10 nodes)
  std::vector<uint32_t> useg_lmax;  // longest row of the segment
};

struct GatherArgs {
  const double   *xyz;
  const int32_t  *conn, *adrU, *adrP;
  const double   *sol, *soldot, *source, *tab, *geo;
  const int64_t  *ia;
  double         *val, *rhs;
  const double   *es; // per-element state (lane kernels)
  const int32_t  *pair;
  const int2     *range;
  const int32_t  *row;
  const uint32_t *smoff, *cta_size;
  const int64_t  *cta_g0;
  int32_t         cta0; // first CTA of this launch segment
  const int32_t  *cta_perm; // launch order (block index -> CTA), or nullptr
  const uint16_t *off;
  int32_t         nNodes;
  int64_t         nInc;
  int             nq, ntab;
  THCoeffs        c;
  double          c0;
  // E[c][al][v] lives in the kernel-parameter constant bank: its indices are compile-time constants after unrolling,
  // so every use is a constant operand of a DFMA instead of a shared-memory load (sized for P2/P1 tetrahedra)
  double          E[120];
};

// ----------------------------------------------------------------------------------------------------------
// U rows: one thread per velocity node (D rows), loop over the adjacent elements
// ----------------------------------------------------------------------------------------------------------
// per (node, element) pair: indices (stage 1) and values (stage 2) of the software pipeline
template <int D, int NS, int NP> struct PairIdx {
  int     e, la;
  int32_t au[NS * D], ap[NP];
};
template <int D, int NS, int NP> struct PairVal {
  double G[D * D], J, U[NS][D], P[NP];
  uint4  ow[GT<D, NS, NP>::OFFW_U / 8];
};

template <int D, int NS, int NP> __device__ __forceinline__ void load_pair_idx(const GatherArgs &a, int ea, PairIdx<D, NS, NP> &I)
{
  constexpr int NU = NS * D;
  I.e  = ea / NS;
  I.la = ea - I.e * NS;
  const int32_t *au = a.adrU + (int64_t)I.e * NU;
  if(NU % 4 == 0) {
    const int4 *a4 = reinterpret_cast<const int4 *>(au);
#pragma unroll
    for(int k = 0; k < NU / 4; ++k) {
      const int4 v    = a4[k];
      I.au[4 * k + 0] = v.x;
      I.au[4 * k + 1] = v.y;
      I.au[4 * k + 2] = v.z;
      I.au[4 * k + 3] = v.w;
    }
  } else {
#pragma unroll
    for(int k = 0; k < NU; ++k) I.au[k] = au[k];
  }
  const int32_t *ap = a.adrP + (int64_t)I.e * NP;
#pragma unroll
  for(int q = 0; q < NP; ++q) I.ap[q] = ap[q];
}

template <int D, int NS, int NP, bool MAT, bool RES>
__device__ __forceinline__ void load_pair_val(const GatherArgs &a, const PairIdx<D, NS, NP> &I, int p, PairVal<D, NS, NP> &V)
{
  {
    constexpr int GW = GT<D, NS, NP>::GW;
    const double2 *ge = reinterpret_cast<const double2 *>(a.geo + (int64_t)I.e * GW);
    double         g[GW];
#pragma unroll
    for(int i = 0; i < GW / 2; ++i) {
      const double2 v = ge[i];
      g[2 * i]     = v.x;
      g[2 * i + 1] = v.y;
    }
#pragma unroll
    for(int i = 0; i < D * D; ++i) V.G[i] = g[i];
    V.J = g[D * D];
  }
  if(D == 2) {
    // the two components of a node are consecutive DOFs in the reference numbering (src/feNumber.cpp:370-483): one
    // 16-byte load per node when the pair is aligned, two 8-byte loads otherwise
#pragma unroll
    for(int b = 0; b < NS; ++b) {
      const int32_t d0 = I.au[b * D], d1 = I.au[b * D + 1];
      if(((d0 & 1) == 0) && d1 == d0 + 1) {
        const double2 v = *reinterpret_cast<const double2 *>(a.sol + d0);
        V.U[b][0] = v.x;
        V.U[b][1] = v.y;
      } else {
        V.U[b][0] = a.sol[d0];
        V.U[b][1] = a.sol[d1];
      }
    }
  } else {
#pragma unroll
    for(int b = 0; b < NS; ++b)
#pragma unroll
      for(int m = 0; m < D; ++m) V.U[b][m] = a.sol[I.au[b * D + m]];
  }
  if(RES) {
#pragma unroll
    for(int q = 0; q < NP; ++q) V.P[q] = a.sol[I.ap[q]];
  }
  if(MAT) {
    const uint4 *src = reinterpret_cast<const uint4 *>(a.off + (int64_t)p * GT<D, NS, NP>::OFFW_U);
#pragma unroll
    for(int w = 0; w < GT<D, NS, NP>::OFFW_U / 8; ++w) V.ow[w] = src[w];
  }
}

__device__ __forceinline__ uint32_t off16(const uint4 *ow, int j)
{
  const uint4    v = ow[j >> 3];
  const int      k = j & 7;
  const uint32_t w = (k >> 1) == 0 ? v.x : (k >> 1) == 1 ? v.y : (k >> 1) == 2 ? v.z : v.w;
  return (k & 1) ? (w >> 16) : (w & 0xffffu);
}

template <int D, int NS, int NP, int NPB, bool MAT, bool RES, bool PIPE>
__global__ void __launch_bounds__(NPB, (D == 2 && !PIPE) ? 8 : 1) gather_u_kernel(const GatherArgs a)
{
  using T = GT<D, NS, NP>;
  constexpr int NU = T::NU;
  extern __shared__ double sm[];
  double *s_tab = sm;
  double *s_buf = sm + a.ntab;
  __shared__ int32_t  s_row[NPB * D];
  __shared__ uint32_t s_base[NPB];
  __shared__ int32_t  s_len[NPB];

  const int tid = threadIdx.x;
  for(int i = tid; i < a.ntab; i += NPB) s_tab[i] = a.tab[i];
  const int32_t  cta  = a.cta_perm ? a.cta_perm[a.cta0 + blockIdx.x] : a.cta0 + (int32_t)blockIdx.x;
  const int32_t  n    = cta * NPB + tid;
  const bool     live = n < a.nNodes;
  const uint32_t tot  = MAT ? a.cta_size[cta] : 0;
  int32_t        row[D];
  int            len = 0;
  uint32_t       base = 0;
  if(live) {
#pragma unroll
    for(int c = 0; c < D; ++c) row[c] = a.row[n * D + c];
#pragma unroll
    for(int c = D - 1; c >= 0; --c)
      if(row[c] < a.nInc) len = (int)(a.ia[row[c] + 1] - a.ia[row[c]]);
    base = a.smoff[n];
  } else {
#pragma unroll
    for(int c = 0; c < D; ++c) row[c] = 0x7fffffff;
  }
  if(MAT) {
#pragma unroll
    for(int c = 0; c < D; ++c) s_row[tid * D + c] = row[c];
    s_base[tid] = base;
    s_len[tid]  = len;
    for(uint32_t i = tid; i < tot; i += NPB) s_buf[i] = 0.;
  }
  __syncthreads();

  if(live) {
    // shared-memory index of the row buffer of component c (unknown rows only, packed); entries that are not
    // assembled (essential row or column) are redirected to a per-thread trash slot behind the CTA's buffers, so the
    // read-modify-write sequences below are branch-free and can be issued as independent batches
    const int trash = (int)tot + tid;
    int       bi[D];
    {
      uint32_t o = base;
#pragma unroll
      for(int c = 0; c < D; ++c) {
        bi[c] = row[c] < a.nInc ? (int)o : -1;
        if(row[c] < a.nInc) o += len;
      }
    }
    double res[D];
#pragma unroll
    for(int c = 0; c < D; ++c) res[c] = 0.;
    const THCoeffs c      = a.c;
    const double   mass0  = c.c_mass * a.c0;
    const int2     rg     = a.range[n];
    const int      pend   = rg.x + rg.y;
    const bool     domass = (c.c_mass != 0.) && (a.soldot != nullptr);

    // software pipeline over the adjacent elements: while pair p is computed, the values of pair p+1 and the
    // indices of pair p+2 are in flight (the chain pair -> element DOF table -> solution entries is three dependent
    // global loads deep and only ~8 warps per SM are resident)
    PairIdx<D, NS, NP> I1, I2;
    PairVal<D, NS, NP> V1;
    int                ea3 = 0;
    if(PIPE) {
      load_pair_idx<D, NS, NP>(a, a.pair[rg.x], I1);
      load_pair_idx<D, NS, NP>(a, a.pair[min(rg.x + 1, pend - 1)], I2);
      ea3 = a.pair[min(rg.x + 2, pend - 1)];
      load_pair_val<D, NS, NP, MAT, RES>(a, I1, rg.x, V1);
    }

    for(int p = rg.x; p < pend; ++p) {
      if(!PIPE) { // plain variant: fewer live registers, higher occupancy
        load_pair_idx<D, NS, NP>(a, a.pair[p], I1);
        load_pair_val<D, NS, NP, MAT, RES>(a, I1, p, V1);
      }
      const PairVal<D, NS, NP> V  = V1;
      const int                e  = I1.e, la = I1.la;
      int32_t                  aud[NU];
      if(domass) {
#pragma unroll
        for(int k = 0; k < NU; ++k) aud[k] = I1.au[k];
      }
      if(PIPE) {
        load_pair_val<D, NS, NP, MAT, RES>(a, I2, min(p + 1, pend - 1), V1);
        I1 = I2;
        load_pair_idx<D, NS, NP>(a, ea3, I2);
        ea3 = a.pair[min(p + 3, pend - 1)];
      }
      const double *G = V.G;
      const double  J = V.J;
      // contravariant velocity DOFs Ut[c][al] = sum_m U[c][m] dxi_al/dx_m and Z[al][v] = sum_c Ut[c][al] T3[la][c][v]
      const double *T3a = s_tab + T::O_T3 + la * NS * NP;
      double        Z[D][NP];
#pragma unroll
      for(int al = 0; al < D; ++al)
#pragma unroll
        for(int v = 0; v < NP; ++v) Z[al][v] = 0.;
#pragma unroll
      for(int cc = 0; cc < NS; ++cc) {
        double ut[D];
#pragma unroll
        for(int al = 0; al < D; ++al) {
          double s = 0.;
#pragma unroll
          for(int m = 0; m < D; ++m) s += V.U[cc][m] * G[al * D + m];
          ut[al] = s;
        }
#pragma unroll
        for(int v = 0; v < NP; ++v) {
          const double t = T3a[cc * NP + v];
#pragma unroll
          for(int al = 0; al < D; ++al) Z[al][v] += ut[al] * t;
        }
      }
      // velocity gradient at the vertices: Dv[v][j][i] = d_j u_i (v) = sum_al dxi_al/dx_j sum_c U[c][i] E[c][al][v]
      double Dv[NP][D][D];
#pragma unroll
      for(int i = 0; i < D; ++i) {
        double X[D][NP];
#pragma unroll
        for(int al = 0; al < D; ++al)
#pragma unroll
          for(int v = 0; v < NP; ++v) X[al][v] = 0.;
#pragma unroll
        for(int cc = 0; cc < NS; ++cc)
#pragma unroll
          for(int al = 0; al < D; ++al)
#pragma unroll
            for(int v = 0; v < NP; ++v) X[al][v] += V.U[cc][i] * a.E[(cc * D + al) * NP + v];
#pragma unroll
        for(int v = 0; v < NP; ++v)
#pragma unroll
          for(int j = 0; j < D; ++j) {
            double s = 0.;
#pragma unroll
            for(int al = 0; al < D; ++al) s += G[al * D + j] * X[al][v];
            Dv[v][j][i] = s;
          }
      }

#pragma unroll
      for(int b = 0; b < NS; ++b) {
        const double *Kr = s_tab + T::O_K + (la * NS + b) * D * D;
        double        K[D][D]; // K[m][n] = int d_m phi_a d_n phi_b
        {
          double H[D][D];
#pragma unroll
          for(int al = 0; al < D; ++al)
#pragma unroll
            for(int nn = 0; nn < D; ++nn) {
              double s = 0.;
#pragma unroll
              for(int be = 0; be < D; ++be) s += Kr[al * D + be] * G[be * D + nn];
              H[al][nn] = s;
            }
#pragma unroll
          for(int m = 0; m < D; ++m)
#pragma unroll
            for(int nn = 0; nn < D; ++nn) {
              double s = 0.;
#pragma unroll
              for(int al = 0; al < D; ++al) s += G[al * D + m] * H[al][nn];
              K[m][nn] = J * s;
            }
        }
        double trK = 0.;
#pragma unroll
        for(int m = 0; m < D; ++m) trK += K[m][m];
        double        C1 = 0.; // int phi_a (u . grad phi_b)
#pragma unroll
        for(int al = 0; al < D; ++al)
#pragma unroll
          for(int v = 0; v < NP; ++v) C1 += a.E[(b * D + al) * NP + v] * Z[al][v];
        C1 *= J;
        const double *T3ab = T3a + b * NP;
        double        t3[NP];
#pragma unroll
        for(int v = 0; v < NP; ++v) t3[v] = J * T3ab[v];
        const double Mab = J * s_tab[T::O_M + la * NS + b];
        const double s   = c.c_conv * C1 + (c.diff_k - c.sig_mu) * trK + mass0 * Mab;
        if(MAT) {
          int    idx[D][D];
          double A[D][D], old[D][D];
#pragma unroll
          for(int j = 0; j < D; ++j) {
            const uint32_t o = off16(V.ow, b * D + j);
#pragma unroll
            for(int i = 0; i < D; ++i) {
              idx[i][j] = (bi[i] >= 0 && o != 0xFFFFu) ? bi[i] + (int)o : trash;
              double C2 = 0.; // int phi_a phi_b d_j u_i
#pragma unroll
              for(int v = 0; v < NP; ++v) C2 += Dv[v][j][i] * t3[v];
              A[i][j] = (i == j ? s : 0.) - c.sig_mu * K[j][i] + c.c_conv * C2;
            }
          }
#pragma unroll
          for(int i = 0; i < D; ++i)
#pragma unroll
            for(int j = 0; j < D; ++j) old[i][j] = s_buf[idx[i][j]];
#pragma unroll
          for(int i = 0; i < D; ++i)
#pragma unroll
            for(int j = 0; j < D; ++j) s_buf[idx[i][j]] = old[i][j] + A[i][j];
        }
        if(RES) {
          const double r1 = (c.sig_mu - c.diff_k) * trK - c.c_conv * C1;
#pragma unroll
          for(int i = 0; i < D; ++i) {
            double r = r1 * V.U[b][i];
#pragma unroll
            for(int m = 0; m < D; ++m) r += c.sig_mu * K[m][i] * V.U[b][m];
            if(domass) r -= c.c_mass * Mab * a.soldot[aud[b * D + i]];
            res[i] += r;
          }
        }
      }
      // pressure columns
      {
        int    idx[NP][D];
        double Bp[NP][D];
#pragma unroll
        for(int q = 0; q < NP; ++q) {
          const double  *Br = s_tab + T::O_B + (q * NS + la) * D;
          const uint32_t o  = MAT ? off16(V.ow, NU + q) : 0xFFFFu;
#pragma unroll
          for(int i = 0; i < D; ++i) {
            double s = 0.;
#pragma unroll
            for(int al = 0; al < D; ++al) s += G[al * D + i] * Br[al];
            Bp[q][i]  = J * s;
            idx[q][i] = (bi[i] >= 0 && o != 0xFFFFu) ? bi[i] + (int)o : trash;
            if(RES) res[i] += (c.c_gradp - c.c_sig) * Bp[q][i] * V.P[q];
          }
        }
        if(MAT) {
          double old[NP][D];
#pragma unroll
          for(int q = 0; q < NP; ++q)
#pragma unroll
            for(int i = 0; i < D; ++i) old[q][i] = s_buf[idx[q][i]];
#pragma unroll
          for(int q = 0; q < NP; ++q)
#pragma unroll
            for(int i = 0; i < D; ++i) s_buf[idx[q][i]] = old[q][i] + (c.c_sig - c.c_gradp) * Bp[q][i];
        }
      }
      if(RES && a.source != nullptr) {
        const double *W   = s_tab + T::O_W;
        const double *src = a.source + (int64_t)e * a.nq * D;
        for(int k = 0; k < a.nq; ++k) {
          const double wj = J * W[k * NS + la];
#pragma unroll
          for(int i = 0; i < D; ++i) res[i] -= wj * src[k * D + i];
        }
      }
    }
    if(RES) {
#pragma unroll
      for(int i = 0; i < D; ++i)
        if(row[i] < a.nInc) a.rhs[row[i]] = res[i];
    }
  }
  if(MAT) {
    __syncthreads();
    const int64_t g0 = a.cta_g0[cta];
    if(g0 >= 0) {
      // the CTA's rows are consecutive in the CSR arrays: the row buffers are a contiguous image of val[g0 ...]
      double *dst = a.val + g0;
      for(uint32_t i = tid; i < tot; i += NPB) dst[i] = s_buf[i];
    } else {
      // general numbering: one warp per row segment
      const int lane = tid & 31, wid = tid >> 5, nw = NPB / 32;
      for(int t = wid; t < NPB; t += nw) {
        const int ln = s_len[t];
        uint32_t  o  = s_base[t];
#pragma unroll
        for(int c = 0; c < D; ++c) {
          const int32_t r = s_row[t * D + c];
          if(r < a.nInc) {
            double *dst = a.val + a.ia[r];
            for(int k = lane; k < ln; k += 32) dst[k] = s_buf[o + k];
            o += ln;
          }
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------------------
// P rows: one thread per pressure node (1 row)
// ----------------------------------------------------------------------------------------------------------
template <int D, int NS, int NP> struct PPairVal {
  double G[D * D], J, U[NS][D];
  uint4  ow[GT<D, NS, NP>::OFFW_P / 8];
};

template <int D, int NS, int NP, bool MAT, bool RES>
__device__ __forceinline__ void load_ppair(const GatherArgs &a, int eq, int p, int &q, PPairVal<D, NS, NP> &V)
{
  constexpr int NU = NS * D, GW = GT<D, NS, NP>::GW;
  const int     e = eq / NP;
  q               = eq - e * NP;
  int32_t        au[NU];
  const int32_t *pu = a.adrU + (int64_t)e * NU;
  if(RES) {
    if(NU % 4 == 0) {
      const int4 *a4 = reinterpret_cast<const int4 *>(pu);
#pragma unroll
      for(int k = 0; k < NU / 4; ++k) {
        const int4 v  = a4[k];
        au[4 * k + 0] = v.x;
        au[4 * k + 1] = v.y;
        au[4 * k + 2] = v.z;
        au[4 * k + 3] = v.w;
      }
    } else {
#pragma unroll
      for(int k = 0; k < NU; ++k) au[k] = pu[k];
    }
  }
  const double2 *ge = reinterpret_cast<const double2 *>(a.geo + (int64_t)e * GW);
  double         g[GW];
#pragma unroll
  for(int i = 0; i < GW / 2; ++i) {
    const double2 v = ge[i];
    g[2 * i]     = v.x;
    g[2 * i + 1] = v.y;
  }
#pragma unroll
  for(int i = 0; i < D * D; ++i) V.G[i] = g[i];
  V.J = g[D * D];
  if(MAT) {
    const uint4 *src = reinterpret_cast<const uint4 *>(a.off + (int64_t)p * GT<D, NS, NP>::OFFW_P);
#pragma unroll
    for(int w = 0; w < GT<D, NS, NP>::OFFW_P / 8; ++w) V.ow[w] = src[w];
  }
  if(RES) {
    if(D == 2) {
#pragma unroll
      for(int b = 0; b < NS; ++b) {
        const int32_t d0 = au[b * D], d1 = au[b * D + 1];
        if(((d0 & 1) == 0) && d1 == d0 + 1) {
          const double2 v = *reinterpret_cast<const double2 *>(a.sol + d0);
          V.U[b][0] = v.x;
          V.U[b][1] = v.y;
        } else {
          V.U[b][0] = a.sol[d0];
          V.U[b][1] = a.sol[d1];
        }
      }
    } else {
#pragma unroll
      for(int b = 0; b < NS; ++b)
#pragma unroll
        for(int m = 0; m < D; ++m) V.U[b][m] = a.sol[au[b * D + m]];
    }
  }
}

template <int D, int NS, int NP, int NPB, bool MAT, bool RES>
__global__ void __launch_bounds__(NPB) gather_p_kernel(const GatherArgs a)
{
  using T = GT<D, NS, NP>;
  extern __shared__ double sm[];
  double *s_tab = sm;
  double *s_buf = sm + a.ntab;
  __shared__ int32_t  s_row[NPB];
  __shared__ uint32_t s_base[NPB];
  __shared__ int32_t  s_len[NPB];
  const int tid = threadIdx.x;
  for(int i = tid; i < a.ntab; i += NPB) s_tab[i] = a.tab[i];
  const int32_t  cta  = a.cta_perm ? a.cta_perm[a.cta0 + blockIdx.x] : a.cta0 + (int32_t)blockIdx.x;
  const int32_t  n    = cta * NPB + tid;
  const bool     live = n < a.nNodes;
  const uint32_t tot  = MAT ? a.cta_size[cta] : 0;
  int32_t        row  = 0x7fffffff;
  int            len  = 0;
  uint32_t       base = 0;
  if(live) {
    row  = a.row[n];
    len  = row < a.nInc ? (int)(a.ia[row + 1] - a.ia[row]) : 0;
    base = a.smoff[n];
  }
  if(MAT) {
    s_row[tid]  = row;
    s_base[tid] = base;
    s_len[tid]  = len;
    for(uint32_t i = tid; i < tot; i += NPB) s_buf[i] = 0.;
  }
  __syncthreads();
  if(live && row < a.nInc) {
    const int    trash = (int)tot + tid;
    double       res   = 0.;
    const double cdiv  = a.c.c_div;
    const int2   rg    = a.range[n];
    const int    pend  = rg.x + rg.y;
    PPairVal<D, NS, NP> V1;
    int                 q1;
    load_ppair<D, NS, NP, MAT, RES>(a, a.pair[rg.x], rg.x, q1, V1);
    int eq2 = a.pair[min(rg.x + 1, pend - 1)];
    for(int p = rg.x; p < pend; ++p) {
      const PPairVal<D, NS, NP> V = V1;
      const int                 q = q1;
      // values of the next pair and the pair index after it are in flight while this one is computed
      load_ppair<D, NS, NP, MAT, RES>(a, eq2, min(p + 1, pend - 1), q1, V1);
      eq2 = a.pair[min(p + 2, pend - 1)];
      double Bp[NS][D];
      int    idx[NS][D];
#pragma unroll
      for(int b = 0; b < NS; ++b) {
        const double *Br = s_tab + T::O_B + (q * NS + b) * D;
#pragma unroll
        for(int j = 0; j < D; ++j) {
          double s = 0.;
#pragma unroll
          for(int al = 0; al < D; ++al) s += V.G[al * D + j] * Br[al];
          Bp[b][j] = cdiv * V.J * s;
          if(RES) res -= Bp[b][j] * V.U[b][j];
          if(MAT) {
            const uint32_t o = off16(V.ow, b * D + j);
            idx[b][j] = o != 0xFFFFu ? (int)base + (int)o : trash;
          }
        }
      }
      if(MAT) {
        // distinct columns of one row: the read-modify-writes are independent and issued as one batch
        double old[NS][D];
#pragma unroll
        for(int b = 0; b < NS; ++b)
#pragma unroll
          for(int j = 0; j < D; ++j) old[b][j] = s_buf[idx[b][j]];
#pragma unroll
        for(int b = 0; b < NS; ++b)
#pragma unroll
          for(int j = 0; j < D; ++j) s_buf[idx[b][j]] = old[b][j] + Bp[b][j];
      }
    }
    if(RES) a.rhs[row] = res;
  }
  if(MAT) {
    __syncthreads();
    const int64_t g0 = a.cta_g0[cta];
    if(g0 >= 0) {
      double *dst = a.val + g0;
      for(uint32_t i = tid; i < tot; i += NPB) dst[i] = s_buf[i];
    } else {
      const int lane = tid & 31, wid = tid >> 5, nw = NPB / 32;
      for(int t = wid; t < NPB; t += nw) {
        const int32_t r = s_row[t];
        if(r < a.nInc) {
          double        *dst = a.val + a.ia[r];
          const uint32_t o   = s_base[t];
          const int      ln  = s_len[t];
          for(int k = lane; k < ln; k += 32) dst[k] = s_buf[o + k];
        }
      }
    }
  }
}

} // namespace b200
#include "gather_lane.cuh"
#include "gather_urow.cuh"
namespace b200 {

// lanes per node -> (warps per CTA, register budget); the plan's nodes-per-CTA follows from it
struct LaneCfg {
  int L, NW;
};
static LaneCfg lane_cfg(int D, int L)
{
  if(D == 2) {
    switch(L) {
    case 1: return {1, 2};
    case 2: return {2, 4};
    case 3: return {3, 4};
    default: return {6, 8};
    }
  }
  switch(L) {
  case 1: return {1, 1};
  case 2: return {2, 2};
  case 5: return {5, 4};
  default: return {10, 4};
  }
}
static int lane_default(int D)
{
  const char *k = getenv("B200_GATHER_LANES");
  if(k) return atoi(k);
  return D == 2 ? 1 : 10;
}


// ----------------------------------------------------------------------------------------------------------
// plan construction (set-up)
// ----------------------------------------------------------------------------------------------------------
// element geometry table: the reference tabulates the same quantities once per mesh (feCncGeo::_J, src/feCncGeo.cpp:278-418,
// and the constant ElementTransformation of P1 geometry, :489-509)
template <int D> __global__ void geometry_kernel(int64_t nElm, const double *xyz, const int32_t *conn, double *geo)
{
  constexpr int GW = (D * D + 1 + 1) / 2 * 2;
  for(int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nElm; e += (int64_t)gridDim.x * blockDim.x) {
    int32_t vtx[D + 1];
#pragma unroll
    for(int v = 0; v <= D; ++v) vtx[v] = conn[e * (D + 1) + v];
    double G[D * D], J;
    element_geometry<D>(xyz, vtx, G, &J);
#pragma unroll
    for(int i = 0; i < D * D; ++i) geo[e * GW + i] = G[i];
    geo[e * GW + D * D] = J;
    if(GW > D * D + 1) geo[e * GW + D * D + 1] = 0.;
  }
}

__global__ void node_keys_kernel(int64_t nElm, int nLoc, int nF, int stride, const int32_t *adr, int32_t *keys, int32_t *vals)
{
  const int64_t tot = nElm * nLoc;
  for(int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx / nLoc;
    const int     l = (int)(idx - e * nLoc);
    keys[idx] = adr[e * nF + l * stride]; // DOF of component 0 identifies the node
    vals[idx] = (int32_t)idx;
  }
}

__global__ void node_flag_kernel(int64_t n, const int32_t *keys, int32_t *flag)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one thread per pair position: node start -> (range, rows)
__global__ void node_fill_kernel(int64_t n, const int32_t *keys, const int32_t *flag, const int32_t *rank, const int32_t *vals, int nLoc, int nF,
                                 int nRow, const int32_t *adr, int2 *range, int32_t *row)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if(!flag[i]) continue;
    const int32_t node = rank[i] - 1;
    int64_t       j = i + 1;
    while(j < n && keys[j] == keys[i]) ++j;
    range[node] = make_int2((int)i, (int)(j - i));
    const int64_t e = vals[i] / nLoc;
    const int     l = vals[i] - (int32_t)(e * nLoc);
    for(int c = 0; c < nRow; ++c) row[(int64_t)node * nRow + c] = adr[e * nF + l * nRow + c];
  }
}

// nodes with at least one unknown row are kept (compaction flag)
__global__ void node_keep_kernel(int32_t nNodes, int nRow, int64_t nInc, const int32_t *row, int32_t *keep)
{
  for(int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nNodes; i += gridDim.x * blockDim.x) {
    int k = 0;
    for(int c = 0; c < nRow; ++c) k |= row[(int64_t)i * nRow + c] < nInc;
    keep[i] = k;
  }
}

__global__ void node_compact_kernel(int32_t nNodes, int nRow, const int32_t *keep, const int32_t *pos, const int2 *range, const int32_t *row,
                                    int2 *range2, int32_t *row2)
{
  for(int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nNodes; i += gridDim.x * blockDim.x) {
    if(!keep[i]) continue;
    const int32_t o = pos[i] - 1;
    range2[o] = range[i];
    for(int c = 0; c < nRow; ++c) row2[(int64_t)o * nRow + c] = row[(int64_t)i * nRow + c];
  }
}

// shared-memory layout of the row buffers: one thread per CTA
__global__ void node_smem_kernel(int32_t nNodes, int npb, int nRow, int64_t nInc, const int64_t *ia, const int32_t *row, const int2 *range,
                                 uint32_t *smoff, uint32_t *cta_size, int64_t *cta_g0, int32_t *cta_pairs, int *err)
{
  const int32_t cta = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t n0 = cta * npb;
  if(n0 >= nNodes) return;
  uint32_t o = 0;
  int64_t  prev = -1, g0 = -1;
  bool     contiguous = true;
  int32_t  npairs = 0;
  for(int32_t n = n0; n < n0 + npb && n < nNodes; ++n) {
    smoff[n] = o;
    npairs += range[n].y;
    int64_t len = -1;
    int     nun = 0;
    for(int c = 0; c < nRow; ++c) {
      const int32_t r = row[(int64_t)n * nRow + c];
      if(r < nInc) {
        const int64_t l = ia[r + 1] - ia[r];
        if(len >= 0 && l != len) atomicExch(err, 2); // rows of one node must share their column structure
        len = l;
        ++nun;
        if(prev < 0)
          g0 = ia[r];
        else if(r != prev + 1)
          contiguous = false;
        prev = r;
      }
    }
    if(len >= 65535) atomicExch(err, 3);
    o += (uint32_t)(nun * (len > 0 ? len : 0));
  }
  cta_size[cta]  = o;
  cta_g0[cta]    = contiguous ? g0 : -1;
  cta_pairs[cta] = npairs;
}

// row-local offsets of the columns of every (node, element) pair; one thread per (pair, local column)
__global__ void node_offsets_kernel(int32_t nNodes, int nRow, int nLoc, int offw, int ncol, int NU, int NP, const int2 *range, const int32_t *row,
                                    const int32_t *pair, const int32_t *adrU, const int32_t *adrP, const int64_t *ia, const int32_t *ja,
                                    int64_t nInc, int colmaskU, int colmaskP, uint16_t *off, int *err)
{
  // grid-stride over nodes; threads of a block cooperate over (pair, column) of the node
  for(int32_t n = blockIdx.x; n < nNodes; n += gridDim.x) {
    const int2 rg = range[n];
    int32_t    r0 = -1;
    for(int c = nRow - 1; c >= 0; --c)
      if(row[(int64_t)n * nRow + c] < nInc) r0 = row[(int64_t)n * nRow + c];
    const int64_t beg = ia[r0], end = ia[r0 + 1];
    for(int idx = threadIdx.x; idx < rg.y * offw; idx += blockDim.x) {
      const int     pp = idx / offw, j = idx - pp * offw;
      const int64_t p  = rg.x + pp;
      uint16_t      o  = 0xFFFF;
      if(j < ncol) {
        const int64_t e   = pair[p] / nLoc;
        const bool    isU = j < NU;
        if(isU ? colmaskU : colmaskP) {
          const int32_t col = isU ? adrU[e * NU + j] : adrP[e * NP + (j - NU)];
          if(col < nInc) {
            int64_t lo = beg, hi = end - 1;
            while(lo < hi) {
              const int64_t mid = (lo + hi) >> 1;
              if(ja[mid] < col)
                lo = mid + 1;
              else
                hi = mid;
            }
            if(lo < end && ja[lo] == col) {
              o = (uint16_t)(lo - beg);
              // every unknown row of the node must hold this column at the same offset
              for(int c = 0; c < nRow; ++c) {
                const int32_t r = row[(int64_t)n * nRow + c];
                if(r < nInc && r != r0 && ja[ia[r] + (lo - beg)] != col) atomicExch(err, 4);
              }
            } else
              atomicExch(err, 1);
          }
        }
      }
      off[p * offw + j] = o;
    }
  }
}

// 30-bit Morton code of the centroid of the first element adjacent to the first node of every CTA
__device__ __forceinline__ uint32_t morton_spread(uint32_t v)
{
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void cta_key_kernel(int32_t nCta, int npb, int32_t nNodes, const int2 *__restrict__ range, const int32_t *__restrict__ pair, int nLoc,
                               const int32_t *__restrict__ conn, const double *__restrict__ xyz, int dim, double x0, double y0, double z0,
                               double sx, double sy, double sz, uint32_t *key)
{
  for(int32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nCta; c += gridDim.x * blockDim.x) {
    const int32_t n = min(c * npb + npb / 2, nNodes - 1); // a node in the middle of the CTA
    const int2    rg = range[n];
    uint32_t      k = 0;
    if(rg.y > 0) {
      const int e = pair[rg.x] / nLoc, nv = dim + 1;
      double    ctr[3] = {0., 0., 0.};
      for(int v = 0; v < nv; ++v)
        for(int m = 0; m < dim; ++m) ctr[m] += xyz[(int64_t)conn[e * nv + v] * dim + m] / nv;
      const uint32_t qx = (uint32_t)fmin(1023., fmax(0., (ctr[0] - x0) * sx));
      const uint32_t qy = (uint32_t)fmin(1023., fmax(0., (ctr[1] - y0) * sy));
      const uint32_t qz = dim == 3 ? (uint32_t)fmin(1023., fmax(0., (ctr[2] - z0) * sz)) : 0u;
      k = morton_spread(qx) | (morton_spread(qy) << 1) | (morton_spread(qz) << 2);
    }
    key[c] = k;
  }
}

// Launch order of the CTAs inside every segment: nodes are numbered like the unknowns (the reference numbers vertices in file order
// and mid-edge nodes in order of first appearance, src/feNumber.cpp:370-483), so consecutive CTAs sweep the mesh line by line and the
// per-element records shared by neighbouring lines / planes have left the L2 cache when they are needed again.  Sorting the CTAs of
// a segment along a Morton curve keeps the elements of one neighbourhood in flight together (opt-in: B200_GATHER_ORDER=morton).
static int build_cta_order(System *S, NodeSet &N, int npb, int nLoc)
{
  // measured on T3D(92) / T2D(1024) (profiles/README.md, r02f): 173.0 vs 173.7 and 1 042 vs 1 065 Melem/s -- the row kernels are bound
  // by the L1 / shared-memory pipe, not by where the element records come from, so the linear order stays the default
  const char *e = getenv("B200_GATHER_ORDER");
  if(!(e && std::string(e) == "morton") || N.nCta == 0) return B200_OK;
  const int dim = S->dim;
  std::vector<double> xyz((size_t)S->nVert * dim);
  B200_CUDA(cudaMemcpy(xyz.data(), S->d_xyz, xyz.size() * sizeof(double), cudaMemcpyDeviceToHost));
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for(int64_t i = 0; i < S->nVert; ++i)
    for(int m = 0; m < dim; ++m) {
      lo[m] = std::min(lo[m], xyz[i * dim + m]);
      hi[m] = std::max(hi[m], xyz[i * dim + m]);
    }
  double sc[3] = {0., 0., 0.};
  for(int m = 0; m < dim; ++m) sc[m] = hi[m] > lo[m] ? 1023.999 / (hi[m] - lo[m]) : 0.;
  uint32_t *d_key = nullptr;
  B200_CUDA(cudaMalloc(&d_key, (size_t)N.nCta * sizeof(uint32_t)));
  cta_key_kernel<<<148 * 4, 256, 0, S->stream>>>(N.nCta, npb, N.nNodes, N.range, N.pair, nLoc, S->d_conn, S->d_xyz, dim, lo[0], lo[1], dim == 3 ? lo[2] : 0.,
                                                 sc[0], sc[1], sc[2], d_key);
  count_launch();
  std::vector<uint32_t> key(N.nCta);
  B200_CUDA(cudaMemcpyAsync(key.data(), d_key, (size_t)N.nCta * sizeof(uint32_t), cudaMemcpyDeviceToHost, S->stream));
  B200_CUDA(cudaStreamSynchronize(S->stream));
  cudaFree(d_key);
  std::vector<int32_t> perm(N.nCta);
  for(int32_t i = 0; i < N.nCta; ++i) perm[i] = i;
  for(size_t sg = 0; sg + 1 < N.seg_begin.size(); ++sg)
    std::stable_sort(perm.begin() + N.seg_begin[sg], perm.begin() + N.seg_begin[sg + 1], [&](int32_t a, int32_t b) { return key[a] < key[b]; });
  B200_CUDA(cudaMalloc(&N.cta_perm, (size_t)N.nCta * sizeof(int32_t)));
  B200_CUDA(cudaMemcpy(N.cta_perm, perm.data(), (size_t)N.nCta * sizeof(int32_t), cudaMemcpyHostToDevice));
  return B200_OK;
}

static int build_node_set(System *S, NodeSet &N, const int32_t *d_adr, int nLoc, int nF, int nRow, int npb, int offw, int ncol, int NU, int NP,
                          int colmaskU, int colmaskP, double band = 1.25, int min_div = 50, bool want_off = true)
{
  auto          pol = thrust::cuda::par.on(S->stream);
  const int64_t np  = S->nElm * (int64_t)nLoc;
  if(np >= (int64_t)2147483647) {
    set_error("gather plan: more than 2^31 (element, node) pairs");
    return B200_ERR_UNSUPP;
  }
  N.release();
  N.nPairs = np;
  thrust::device_vector<int32_t> keys(np), flag(np), rank(np);
  B200_CUDA(cudaMalloc(&N.pair, np * sizeof(int32_t)));
  node_keys_kernel<<<148 * 8, 256, 0, S->stream>>>(S->nElm, nLoc, nF, nRow, d_adr, thrust::raw_pointer_cast(keys.data()), N.pair);
  thrust::stable_sort_by_key(pol, keys.begin(), keys.end(), thrust::device_pointer_cast(N.pair));
  node_flag_kernel<<<148 * 8, 256, 0, S->stream>>>(np, thrust::raw_pointer_cast(keys.data()), thrust::raw_pointer_cast(flag.data()));
  thrust::inclusive_scan(pol, flag.begin(), flag.end(), rank.begin());
  const int32_t nAll = rank[np - 1];
  thrust::device_vector<int2>    range(nAll);
  thrust::device_vector<int32_t> row((size_t)nAll * nRow), keep(nAll), pos(nAll);
  node_fill_kernel<<<148 * 8, 256, 0, S->stream>>>(np, thrust::raw_pointer_cast(keys.data()), thrust::raw_pointer_cast(flag.data()),
                                                  thrust::raw_pointer_cast(rank.data()), N.pair, nLoc, nF, nRow, d_adr,
                                                  thrust::raw_pointer_cast(range.data()), thrust::raw_pointer_cast(row.data()));
  node_keep_kernel<<<148 * 8, 256, 0, S->stream>>>(nAll, nRow, S->nInc, thrust::raw_pointer_cast(row.data()), thrust::raw_pointer_cast(keep.data()));
  thrust::inclusive_scan(pol, keep.begin(), keep.end(), pos.begin());
  N.nNodes = pos[nAll - 1];
  count_launch(4);
  if(N.nNodes == 0) return B200_OK;
  B200_CUDA(cudaMalloc(&N.range, (size_t)N.nNodes * sizeof(int2)));
  B200_CUDA(cudaMalloc(&N.row, (size_t)N.nNodes * nRow * sizeof(int32_t)));
  node_compact_kernel<<<148 * 8, 256, 0, S->stream>>>(nAll, nRow, thrust::raw_pointer_cast(keep.data()), thrust::raw_pointer_cast(pos.data()),
                                                     thrust::raw_pointer_cast(range.data()), thrust::raw_pointer_cast(row.data()), N.range, N.row);
  N.nCta = (N.nNodes + npb - 1) / npb;
  B200_CUDA(cudaMalloc(&N.smoff, (size_t)N.nNodes * sizeof(uint32_t)));
  B200_CUDA(cudaMalloc(&N.cta_size, (size_t)N.nCta * sizeof(uint32_t)));
  B200_CUDA(cudaMalloc(&N.cta_g0, (size_t)N.nCta * sizeof(int64_t)));
  if(want_off) {
    B200_CUDA(cudaMalloc(&N.off, (size_t)np * offw * sizeof(uint16_t)));
    B200_CUDA(cudaMemsetAsync(N.off, 0xFF, (size_t)np * offw * sizeof(uint16_t), S->stream));
  }
  thrust::device_vector<int32_t> cta_pairs(N.nCta);
  int *d_err;
  B200_CUDA(cudaMalloc(&d_err, sizeof(int)));
  B200_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), S->stream));
  node_smem_kernel<<<(N.nCta + 127) / 128, 128, 0, S->stream>>>(N.nNodes, npb, nRow, S->nInc, S->d_ia, N.row, N.range, N.smoff, N.cta_size, N.cta_g0,
                                                            thrust::raw_pointer_cast(cta_pairs.data()), d_err);
  if(want_off)
    node_offsets_kernel<<<148 * 16, 64, 0, S->stream>>>(N.nNodes, nRow, nLoc, offw, ncol, NU, NP, N.range, N.row, N.pair, S->spaces[S->su].d_adr,
                                                       S->sp >= 0 ? S->spaces[S->sp].d_adr : nullptr, S->d_ia, S->d_ja, S->nInc, colmaskU, colmaskP,
                                                       N.off, d_err);
  count_launch(3);
  int h_err = 0;
  B200_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, S->stream));
  B200_CUDA(cudaStreamSynchronize(S->stream));
  cudaFree(d_err);
  if(h_err) {
    set_error("gather plan: CSR pattern does not have the regular node-row structure (code " + std::to_string(h_err) + ")");
    return B200_ERR_UNSUPP;
  }
  // launch segments: consecutive CTAs whose row buffers have similar sizes share one launch (and its shared-memory
  // request), so that CTAs of short rows (edge nodes) are not limited by the occupancy of the longest rows
  std::vector<uint32_t> h_size(N.nCta);
  B200_CUDA(cudaMemcpy(h_size.data(), N.cta_size, (size_t)N.nCta * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  std::vector<int32_t> h_pairs(N.nCta);
  B200_CUDA(cudaMemcpy(h_pairs.data(), thrust::raw_pointer_cast(cta_pairs.data()), (size_t)N.nCta * sizeof(int32_t), cudaMemcpyDeviceToHost));
  N.seg_begin.clear();
  N.seg_smem.clear();
  N.seg_pairs.clear();
  N.max_cta = 0;
  {
    int32_t  b = 0;
    uint32_t mx = 0, mn = 0xffffffffu;
    for(int32_t i = 0; i <= N.nCta; ++i) {
      const bool last = i == N.nCta;
      if(!last) {
        const uint32_t v = h_size[i];
        // cut where the size leaves the +-25 % band of the running segment (at most 8 segments, at least 64 CTAs each)
        const bool cut = i > b + std::max(64, N.nCta / min_div) && N.nCta - i > std::max(64, N.nCta / min_div) && N.seg_smem.size() < 7 &&
                         ((double)v > band * (double)std::max<uint32_t>(mn, 1u) || band * (double)v < (double)mx);
        if(!cut) {
          mx = std::max(mx, v);
          mn = std::min(mn, v);
          continue;
        }
      }
      if(i > b) {
        N.seg_begin.push_back(b);
        N.seg_smem.push_back(mx);
        N.max_cta = std::max(N.max_cta, mx);
      }
      if(!last) {
        b  = i;
        mx = mn = h_size[i];
      }
    }
    N.seg_begin.push_back(N.nCta);
    for(size_t sg = 0; sg + 1 < N.seg_begin.size(); ++sg) {
      double np = 0.;
      for(int32_t i = N.seg_begin[sg]; i < N.seg_begin[sg + 1]; ++i) np += h_pairs[i];
      const double nn = std::min<double>((double)(N.seg_begin[sg + 1] - N.seg_begin[sg]) * npb, (double)N.nNodes);
      N.seg_pairs.push_back(np / std::max(1., nn));
    }
  }
  return build_cta_order(S, N, npb, nLoc);
}

// reference tensors from the host tables
static bool build_tables(const System *S, int D, int NS, int NP, std::vector<double> &tab, int &len_nosrc)
{
  const Space &U = S->spaces[S->su], &P = S->spaces[S->sp];
  const int    nq = S->nq;
  const std::vector<double> &w = S->w, &L = U.L, &dL = U.dL, &LP = P.L;
  std::vector<double> K((size_t)NS * NS * D * D, 0.), T3((size_t)NS * NS * NP, 0.), Mr((size_t)NS * NS, 0.), E((size_t)NS * D * NP, 0.),
    B((size_t)NP * NS * D, 0.), W((size_t)nq * NS, 0.);
  for(int k = 0; k < nq; ++k)
    for(int a = 0; a < NS; ++a) {
      W[(size_t)k * NS + a] = w[k] * L[(size_t)k * NS + a];
      for(int b = 0; b < NS; ++b) {
        Mr[(size_t)a * NS + b] += w[k] * L[(size_t)k * NS + a] * L[(size_t)k * NS + b];
        for(int v = 0; v < NP; ++v) T3[((size_t)a * NS + b) * NP + v] += w[k] * L[(size_t)k * NS + a] * L[(size_t)k * NS + b] * LP[(size_t)k * NP + v];
        for(int al = 0; al < D; ++al)
          for(int be = 0; be < D; ++be)
            K[(((size_t)a * NS + b) * D + al) * D + be] += w[k] * dL[((size_t)k * NS + a) * D + al] * dL[((size_t)k * NS + b) * D + be];
      }
      for(int q = 0; q < NP; ++q)
        for(int al = 0; al < D; ++al) B[((size_t)q * NS + a) * D + al] += w[k] * LP[(size_t)k * NP + q] * dL[((size_t)k * NS + a) * D + al];
    }
  // E: least squares dL[k][c][al] = sum_v LP[k][v] E[c][al][v]  (normal equations, NP x NP, Gauss with pivoting)
  std::vector<double> N((size_t)NP * NP, 0.);
  for(int k = 0; k < nq; ++k)
    for(int v = 0; v < NP; ++v)
      for(int u = 0; u < NP; ++u) N[(size_t)v * NP + u] += LP[(size_t)k * NP + v] * LP[(size_t)k * NP + u];
  // invert N
  std::vector<double> A(N), Inv((size_t)NP * NP, 0.);
  for(int i = 0; i < NP; ++i) Inv[(size_t)i * NP + i] = 1.;
  for(int p = 0; p < NP; ++p) {
    int piv = p;
    for(int r = p + 1; r < NP; ++r)
      if(std::fabs(A[(size_t)r * NP + p]) > std::fabs(A[(size_t)piv * NP + p])) piv = r;
    if(std::fabs(A[(size_t)piv * NP + p]) < 1e-14) return false;
    for(int j = 0; j < NP; ++j) {
      std::swap(A[(size_t)p * NP + j], A[(size_t)piv * NP + j]);
      std::swap(Inv[(size_t)p * NP + j], Inv[(size_t)piv * NP + j]);
    }
    const double ip = 1. / A[(size_t)p * NP + p];
    for(int j = 0; j < NP; ++j) {
      A[(size_t)p * NP + j] *= ip;
      Inv[(size_t)p * NP + j] *= ip;
    }
    for(int r = 0; r < NP; ++r) {
      if(r == p) continue;
      const double f = A[(size_t)r * NP + p];
      for(int j = 0; j < NP; ++j) {
        A[(size_t)r * NP + j] -= f * A[(size_t)p * NP + j];
        Inv[(size_t)r * NP + j] -= f * Inv[(size_t)p * NP + j];
      }
    }
  }
  double worst = 0.;
  for(int c = 0; c < NS; ++c)
    for(int al = 0; al < D; ++al) {
      double rhs[8] = {0.};
      for(int k = 0; k < nq; ++k)
        for(int v = 0; v < NP; ++v) rhs[v] += LP[(size_t)k * NP + v] * dL[((size_t)k * NS + c) * D + al];
      for(int v = 0; v < NP; ++v) {
        double s = 0.;
        for(int u = 0; u < NP; ++u) s += Inv[(size_t)v * NP + u] * rhs[u];
        // the exact coefficients of P2-gradient-in-P1 are small integers: snap away the least-squares rounding
        const double r = std::nearbyint(s);
        E[((size_t)c * D + al) * NP + v] = (std::fabs(s - r) < 1e-10) ? r : s;
      }
      for(int k = 0; k < nq; ++k) {
        double s = 0.;
        for(int v = 0; v < NP; ++v) s += LP[(size_t)k * NP + v] * E[((size_t)c * D + al) * NP + v];
        worst = std::fmax(worst, std::fabs(s - dL[((size_t)k * NS + c) * D + al]));
      }
    }
  if(worst > 1e-12) return false; // gradients of the velocity basis are not in the span of the pressure basis
  tab.clear();
  tab.insert(tab.end(), K.begin(), K.end());
  tab.insert(tab.end(), T3.begin(), T3.end());
  tab.insert(tab.end(), Mr.begin(), Mr.end());
  tab.insert(tab.end(), E.begin(), E.end());
  tab.insert(tab.end(), B.begin(), B.end());
  len_nosrc = (int)tab.size();
  tab.insert(tab.end(), W.begin(), W.end());
  return true;
}

static const void *g_urow_owner = nullptr; // plan whose reference tensors are in the __constant__ bank (gather_urow.cuh)

bool gather_tables(const System *S, GatherTables *out)
{
  const GatherPlan *G = static_cast<const GatherPlan *>(S->gather);
  if(!G || !G->d_tab || !G->d_geo) return false;
  out->d_tab       = G->d_tab;
  out->d_geo       = G->d_geo;
  out->tab_len     = G->tab_len;
  out->tab_len_src = G->tab_len_src;
  return true;
}

int gather_kernel_kind(const System *S)
{
  const GatherPlan *G = static_cast<const GatherPlan *>(S->gather);
  return !G ? 0 : G->urow ? 3 : G->lane ? 2 : 1;
}

void gather_free(System *S)
{
  GatherPlan *G = static_cast<GatherPlan *>(S->gather);
  if(!G) return;
  if(g_urow_owner == G) g_urow_owner = nullptr; // a later plan may be allocated at the same address
  G->U.release();
  G->P.release();
  cudaFree(G->d_tab);
  cudaFree(G->d_geo);
  cudaFree(G->d_es);
  cudaFree(G->d_geo4);
  cudaFree(G->d_order);
  cudaFree(G->d_lacnt);
  cudaFree(G->d_rec);
  cudaFree(G->d_wstep);
  cudaFree(G->d_sched);
  cudaFree(G->d_sla);
  delete G;
  S->gather = nullptr;
}


// ----------------------------------------------------------------------------------------------------------
// row-lane plan (gather_urow.cuh)
// ----------------------------------------------------------------------------------------------------------
static void urow_tables(const std::vector<double> &tab, std::vector<double> &ut)
{
  using T = GT<3, 10, 4>;
  using CT = URowTab;
  const double *K = tab.data() + T::O_K, *T3 = tab.data() + T::O_T3, *Mr = tab.data() + T::O_M, *B = tab.data() + T::O_B;
  ut.assign((size_t)10 * CT::LEN, 0.);
  for(int la = 0; la < 10; ++la) {
    double *o = ut.data() + (size_t)la * CT::LEN;
    for(int b = 0; b < 10; ++b) {
      const double *Kb = K + (la * 10 + b) * 9;
      for(int k = 0; k < 9; ++k) o[CT::O_K + b * 9 + k] = Kb[k];
      const double ks[6] = {Kb[0], Kb[4], Kb[8], Kb[1] + Kb[3], Kb[2] + Kb[6], Kb[5] + Kb[7]};
      for(int k = 0; k < 6; ++k) o[CT::O_KS + b * 6 + k] = ks[k];
      for(int v = 0; v < 4; ++v) o[CT::O_T3 + b * 4 + v] = T3[(la * 10 + b) * 4 + v];
      o[CT::O_M + b] = Mr[la * 10 + b];
    }
    for(int q = 0; q < 4; ++q)
      for(int al = 0; al < 3; ++al) o[CT::O_B + q * 3 + al] = B[(q * 10 + la) * 3 + al];
  }
}

// B200_OK, or B200_ERR_UNSUPP when the system does not have the structure the row-lane kernels assume (the lane-group kernels take over)
static int build_urow_plan(System *S, GatherPlan *G, const std::vector<double> &tab)
{
  urow_tables(tab, G->utab);
  {
    using T = GT<3, 10, 4>;
    G->estab.assign(ESTab::LEN, 0.);
    for(int k = 0; k < 400; ++k) G->estab[ESTab::O_T3 + k] = tab[T::O_T3 + k];
    for(int k = 0; k < 120; ++k) G->estab[ESTab::O_B + k] = tab[T::O_B + k];
    for(int k = 0; k < 100; ++k) G->estab[ESTab::O_M + k] = tab[T::O_M + k];
  }
  NodeSet &N = G->U;
  if(N.nNodes == 0) return B200_ERR_UNSUPP;
  auto pol = thrust::cuda::par.on(S->stream);
  B200_CUDA(cudaMalloc(&G->d_geo4, (size_t)S->nElm * 16 * sizeof(double)));
  urow_geo4_kernel<<<148 * 8, 256, 0, S->stream>>>(S->nElm, G->d_geo, GT<3, 10, 4>::GW, G->d_geo4);
  int *d_err;
  B200_CUDA(cudaMalloc(&d_err, sizeof(int)));
  B200_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), S->stream));
  {
    // pairs of every node: ascending local index, then ascending element (residual-only passes read the same order)
    thrust::device_vector<uint64_t> pkey(N.nPairs);
    urow_pair_key_init_kernel<<<148 * 8, 256, 0, S->stream>>>(N.nPairs, thrust::raw_pointer_cast(pkey.data()));
    urow_pair_key_kernel<<<148 * 16, 64, 0, S->stream>>>(N.nNodes, N.range, N.pair, thrust::raw_pointer_cast(pkey.data()));
    thrust::stable_sort_by_key(pol, pkey.begin(), pkey.end(), thrust::device_pointer_cast(N.pair));
  }
  // bounding box of the mesh (Morton cells of ~16 K nodes order the launch)
  double lo[3] = {1e300, 1e300, 1e300}, sc[3] = {0., 0., 0.};
  {
    std::vector<double> xyz((size_t)S->nVert * 3);
    B200_CUDA(cudaMemcpyAsync(xyz.data(), S->d_xyz, xyz.size() * sizeof(double), cudaMemcpyDeviceToHost, S->stream));
    B200_CUDA(cudaStreamSynchronize(S->stream));
    double hi[3] = {-1e300, -1e300, -1e300};
    for(int64_t v = 0; v < S->nVert; ++v)
      for(int m = 0; m < 3; ++m) {
        lo[m] = std::min(lo[m], xyz[v * 3 + m]);
        hi[m] = std::max(hi[m], xyz[v * 3 + m]);
      }
    for(int m = 0; m < 3; ++m) sc[m] = hi[m] > lo[m] ? 1023.999 / (hi[m] - lo[m]) : 0.;
  }
  int cellbits = 0;
  {
    const char *ev = getenv("B200_UROW_CELL");
    const double per = ev ? atof(ev) : 16384.;
    while(cellbits < 20 && (double)N.nNodes / (double)(1u << cellbits) > per) ++cellbits;
  }
  thrust::device_vector<uint64_t> key(N.nNodes);
  B200_CUDA(cudaMalloc(&G->d_order, (size_t)N.nNodes * sizeof(int32_t)));
  B200_CUDA(cudaMalloc(&G->d_lacnt, (size_t)N.nNodes * sizeof(uint64_t)));
  urow_node_key_kernel<<<148 * 8, 256, 0, S->stream>>>(N.nNodes, N.range, N.pair, N.row, S->d_ia, S->nInc, S->d_conn, S->d_xyz, lo[0], lo[1], lo[2], sc[0],
                                                      sc[1], sc[2], 30 - cellbits, thrust::raw_pointer_cast(key.data()), G->d_order, G->d_lacnt, d_err);
  thrust::stable_sort_by_key(pol, key.begin(), key.end(), thrust::device_pointer_cast(G->d_order));
  count_launch(5);
  std::vector<uint64_t> h_key(N.nNodes);
  B200_CUDA(cudaMemcpyAsync(h_key.data(), thrust::raw_pointer_cast(key.data()), (size_t)N.nNodes * sizeof(uint64_t), cudaMemcpyDeviceToHost, S->stream));
  B200_CUDA(cudaMalloc(&G->d_rec, (size_t)N.nPairs * sizeof(URowPair)));
  B200_CUDA(cudaMemsetAsync(G->d_rec, 0, (size_t)N.nPairs * sizeof(URowPair), S->stream)); // record 0 serves the idle lanes: must be readable
  urow_pairs_kernel<<<148 * 16, 64, 0, S->stream>>>(N.nNodes, N.range, N.row, N.pair, S->spaces[S->su].d_adr, S->spaces[S->sp].d_adr, S->d_ia, S->d_ja,
                                                   S->nInc, S->has_matrix_block[0][0] ? 1 : 0, S->has_matrix_block[0][1] ? 1 : 0,
                                                   static_cast<URowPair *>(G->d_rec), d_err);
  count_launch();
  // schedule of every warp of 10 consecutive nodes
  const int32_t cnt = N.nNodes, nw = (cnt + 9) / 10;
  {
    thrust::device_vector<int32_t> nsteps(nw + 1, 0);
    urow_sched_count_kernel<<<(nw + 255) / 256, 256, 0, S->stream>>>(nw, cnt, G->d_order, G->d_lacnt, thrust::raw_pointer_cast(nsteps.data()));
    B200_CUDA(cudaMalloc(&G->d_wstep, (size_t)(nw + 1) * sizeof(int32_t)));
    thrust::exclusive_scan(pol, nsteps.begin(), nsteps.end(), thrust::device_pointer_cast(G->d_wstep));
    int32_t tot = 0;
    B200_CUDA(cudaMemcpyAsync(&tot, G->d_wstep + nw, sizeof(int32_t), cudaMemcpyDeviceToHost, S->stream));
    B200_CUDA(cudaStreamSynchronize(S->stream));
    B200_CUDA(cudaMalloc(&G->d_sched, (size_t)std::max(tot, 1) * 10 * sizeof(int32_t)));
    B200_CUDA(cudaMalloc(&G->d_sla, (size_t)nw * sizeof(uint64_t)));
    urow_sched_fill_kernel<<<(nw + 255) / 256, 256, 0, S->stream>>>(nw, cnt, G->d_order, G->d_lacnt, N.range, G->d_wstep, G->d_sched, G->d_sla);
    count_launch(3);
    if(getenv("B200_VERBOSE"))
      fprintf(stderr, "[b200] row-lane plan: %d nodes in %d groups of 10, %d schedule steps (%.2f per node), %d Morton cells\n", cnt, nw, tot,
              (double)tot * 10. / std::max(cnt, 1), 1 << cellbits);
  }
  int h_err = 0;
  B200_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, S->stream));
  B200_CUDA(cudaStreamSynchronize(S->stream));
  cudaFree(d_err);
  if(h_err) {
    set_error("row-lane plan: columns of a velocity node are not three adjacent unknowns, or > 63 elements share a local index (code " +
              std::to_string(h_err) + ")");
    return B200_ERR_UNSUPP;
  }
  // launch segments: runs of warps of one shared-memory class (the class of a warp is that of its last, longest-class node)
  G->useg_begin.clear();
  G->useg_lmax.clear();
  {
    int      cur = -1;
    uint32_t mx = 0;
    for(int32_t w = 0; w < nw; ++w) {
      const int32_t k0 = w * 10, k1 = std::min(cnt, k0 + 10);
      const int     cls = (int)(h_key[k1 - 1] >> 62);
      if(cls != cur) {
        if(cur >= 0) G->useg_lmax.push_back(mx);
        G->useg_begin.push_back(w);
        cur = cls;
        mx  = 0;
      }
      for(int32_t k = k0; k < k1; ++k) mx = std::max(mx, (uint32_t)((h_key[k] >> 26) & 0xffffu));
    }
    G->useg_lmax.push_back(mx);
    G->useg_begin.push_back(nw);
  }
  if((size_t)(*std::max_element(G->useg_lmax.begin(), G->useg_lmax.end()) + 3) * 32 * 8 + 4096 > 200 * 1024) {
    set_error("row-lane plan: row images exceed shared memory");
    return B200_ERR_UNSUPP;
  }
  return B200_OK;
}

// Builds the gather plan if the registered problem qualifies (Taylor-Hood P2/P1, no P-P block); B200_ERR_UNSUPP otherwise.
int build_gather_plan(System *S)
{
  gather_free(S);
  if(S->plan != PLAN_TAYLOR_HOOD || S->has_matrix_block[1][1]) {
    set_error("gather plan: only the fused Taylor-Hood system is supported");
    return B200_ERR_UNSUPP;
  }
  const int D = S->dim, NS = S->spaces[S->su].nS, NP = S->spaces[S->sp].nS, NU = NS * D, M = NU + NP;
  std::vector<double> tab;
  int                 len_nosrc = 0;
  if(!build_tables(S, D, NS, NP, tab, len_nosrc)) {
    set_error("gather plan: velocity-basis gradients are not in the span of the pressure basis at the quadrature nodes");
    return B200_ERR_UNSUPP;
  }
  GatherPlan *G = new GatherPlan;
  S->gather     = G;
  {
    const int oE = NS * NS * D * D + NS * NS * NP + NS * NS; // GT<D,NS,NP>::O_E
    if(NS * D * NP > 120) {
      set_error("gather plan: E table too large");
      gather_free(S);
      return B200_ERR_UNSUPP;
    }
    for(int i = 0; i < NS * D * NP; ++i) G->E[i] = tab[oE + i];
  }
  G->tab_len     = len_nosrc;
  G->tab_len_src = (int)tab.size();
  {
    // measured on the B200 (profiles/README.md, r01d): in 2-D the thread-per-node kernels that re-gather the solution
    // from the L2-resident state vector win (the per-element state is 512 B/element of extra HBM traffic); in 3-D the
    // lane-group kernels win 2.3x (the row images of one node are 2-5 KB, one thread per node leaves the SM empty)
    const char *k = getenv("B200_GATHER_KERNEL");
    G->lane       = k ? (std::string(k) == "lane" || std::string(k) == "urow") : D == 3;
    G->patch      = false; // the round-1 patch / row-slice kernels were measured slower and moved to experiments/
    // row-lane velocity rows (gather_urow.cuh): P2/P1 tetrahedra, lane-group pressure rows
    G->urow       = D == 3 && NS == 10 && NP == 4 && (!k || std::string(k) == "urow");
    if(G->urow && lane_default(D) != 10) G->urow = false;
  }
  if(G->lane) {
    const LaneCfg lc = lane_cfg(D, lane_default(D));
    G->lanes         = lc.L;
    G->npbU = G->npbP = lc.NW * (32 / lc.L);
  } else {
    G->npbU = D == 2 ? 64 : 32;
    G->npbP = D == 2 ? 64 : 32;
  }
  try {
    B200_CUDA(cudaMalloc(&G->d_tab, tab.size() * sizeof(double)));
    B200_CUDA(cudaMemcpyAsync(G->d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, S->stream));
    B200_CUDA(cudaMalloc(&G->d_geo, (size_t)S->nElm * ((D * D + 2) / 2 * 2) * sizeof(double)));
    if(D == 2)
      geometry_kernel<2><<<148 * 8, 256, 0, S->stream>>>(S->nElm, S->d_xyz, S->d_conn, G->d_geo);
    else
      geometry_kernel<3><<<148 * 8, 256, 0, S->stream>>>(S->nElm, S->d_xyz, S->d_conn, G->d_geo);
    count_launch();
    B200_CUDA(cudaStreamSynchronize(S->stream));
    log_stage("gather plan: geometry");
    const int offwU = (M + 7) / 8 * 8, offwP = (NU + 7) / 8 * 8;
    // lane-group kernels are not limited by the shared memory of the longest rows: few, large launch segments
    const double band    = (G->lane && G->lanes > 1) ? 1.6 : 1.25;
    const int    min_div = (G->lane && G->lanes > 1) ? 12 : 50;
    int       rc = build_node_set(S, G->U, S->spaces[S->su].d_adr, NS, NU, D, G->npbU, offwU, M, NU, NP, S->has_matrix_block[0][0] ? 1 : 0,
                                  S->has_matrix_block[0][1] ? 1 : 0, band, min_div, !G->urow);
    log_stage("gather plan: U node set");
    if(rc == B200_OK && G->urow) {
      const int crc = build_urow_plan(S, G, tab);
      log_stage("gather plan: row-lane pair records");
      if(crc == B200_ERR_UNSUPP) {
        // not the structure the row-lane kernels assume: the lane-group kernels need the per-column offsets after all
        G->urow = false;
        cudaFree(G->d_geo4);
        cudaFree(G->d_order);
        cudaFree(G->d_lacnt);
        cudaFree(G->d_rec);
        cudaFree(G->d_wstep);
        cudaFree(G->d_sched);
        cudaFree(G->d_sla);
        G->d_wstep = nullptr;
        G->d_sched = nullptr;
        G->d_sla   = nullptr;
        G->d_geo4  = nullptr;
        G->d_order = nullptr;
        G->d_lacnt = nullptr;
        G->d_rec   = nullptr;
        if(getenv("B200_VERBOSE")) fprintf(stderr, "[b200] row-lane plan not applicable: %s\n", b200_last_error());
        rc = build_node_set(S, G->U, S->spaces[S->su].d_adr, NS, NU, D, G->npbU, offwU, M, NU, NP, S->has_matrix_block[0][0] ? 1 : 0,
                            S->has_matrix_block[0][1] ? 1 : 0, band, min_div, true);
      } else if(crc != B200_OK)
        rc = crc;
    }
    if(rc == B200_OK && G->lane) {
      const size_t esw = G->urow ? (size_t)ESC::W : (size_t)(NP * D * D + NS * NS + NS * D + NP + 1) / 2 * 2; // ES<D,NS,NP>::W
      B200_CUDA(cudaMalloc(&G->d_es, (size_t)S->nElm * esw * sizeof(double)));
    }
    if(rc == B200_OK)
      rc = build_node_set(S, G->P, S->spaces[S->sp].d_adr, NP, NP, 1, G->npbP, offwP, NU, NU, NP, S->has_matrix_block[1][0] ? 1 : 0, 0, band,
                          min_div);
    log_stage("gather plan: P node set");
    if(rc != B200_OK) {
      gather_free(S);
      return rc;
    }
  } catch(const std::exception &ex) {
    set_error(std::string("gather plan: ") + ex.what());
    gather_free(S);
    return B200_ERR_CUDA;
  }
  const size_t smem = ((size_t)G->tab_len_src + std::max(G->U.max_cta, G->P.max_cta)) * sizeof(double);
  if(smem > 220 * 1024) {
    set_error("gather plan: row buffers exceed shared memory");
    gather_free(S);
    return B200_ERR_UNSUPP;
  }
  return B200_OK;
}

template <int D, int NS, int NP, int NPB> static int launch_gather_t(System *S, int what, const THCoeffs &c)
{
  GatherPlan *G = static_cast<GatherPlan *>(S->gather);
  GatherArgs  a;
  a.xyz    = S->d_xyz;
  a.conn   = S->d_conn;
  a.adrU   = S->spaces[S->su].d_adr;
  a.adrP   = S->spaces[S->sp].d_adr;
  a.sol    = S->d_sol;
  a.soldot = S->have_soldot ? S->d_soldot : nullptr;
  a.source = (c.c_src != 0.) ? S->d_source : nullptr;
  a.tab    = G->d_tab;
  a.geo    = G->d_geo;
  a.ia     = S->d_ia;
  a.val    = S->d_val;
  a.rhs    = S->d_rhs;
  a.nInc   = S->nInc;
  a.nq     = S->nq;
  a.ntab   = a.source ? G->tab_len_src : G->tab_len;
  a.c      = c;
  a.c0     = S->c0;
  for(int i = 0; i < 120; ++i) a.E[i] = G->E[i];
  const bool mat = what & 2;
  for(int pass = 0; pass < 2; ++pass) {
    const NodeSet &N = pass == 0 ? G->U : G->P;
    if(N.nNodes == 0) continue;
    a.pair     = N.pair;
    a.range    = N.range;
    a.row      = N.row;
    a.smoff    = N.smoff;
    a.cta_perm = N.cta_perm;
    a.cta_size = N.cta_size;
    a.off      = N.off;
    a.nNodes   = N.nNodes;
    a.cta_g0   = N.cta_g0;
    const size_t smem_max = ((size_t)a.ntab + (mat ? N.max_cta + NPB : 0)) * sizeof(double);
    const int    nseg     = mat ? (int)N.seg_smem.size() : 1;
#define B200_LAUNCH_G(KERN)                                                                                                              \
  do {                                                                                                                                   \
    B200_CUDA(cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));                                   \
    for(int sg = 0; sg < nseg; ++sg) {                                                                                                   \
      if(pass == 0 && ((D == 2) && N.seg_pairs[mat ? sg : 0] >= 3.5) != pipe) continue;                                                  \
      a.cta0            = mat ? N.seg_begin[sg] : 0;                                                                                     \
      const int    nc   = mat ? N.seg_begin[sg + 1] - N.seg_begin[sg] : N.nCta;                                                          \
      const size_t smem = ((size_t)a.ntab + (mat ? N.seg_smem[sg] + NPB : 0)) * sizeof(double);                                          \
      KERN<<<nc, NPB, smem, S->stream>>>(a);                                                                                             \
      count_launch();                                                                                                                    \
    }                                                                                                                                    \
  } while(0)
    bool pipe = false;
    if(pass == 0) {
      // software-pipelined variant (more registers) for segments of nodes with many adjacent elements, plain variant
      // (higher occupancy) for the others; 3-D always plain (register budget)
      for(int v = 0; v < 2; ++v) {
        pipe = v == 1;
        if(pipe && D != 2) continue;
        if(what == 3) {
          if(pipe)
            B200_LAUNCH_G((gather_u_kernel<D, NS, NP, NPB, true, true, (D == 2)>));
          else
            B200_LAUNCH_G((gather_u_kernel<D, NS, NP, NPB, true, true, false>));
        } else if(what == 2) {
          if(pipe)
            B200_LAUNCH_G((gather_u_kernel<D, NS, NP, NPB, true, false, (D == 2)>));
          else
            B200_LAUNCH_G((gather_u_kernel<D, NS, NP, NPB, true, false, false>));
        } else {
          if(pipe)
            B200_LAUNCH_G((gather_u_kernel<D, NS, NP, NPB, false, true, (D == 2)>));
          else
            B200_LAUNCH_G((gather_u_kernel<D, NS, NP, NPB, false, true, false>));
        }
      }
    } else {
      if(what == 3)
        B200_LAUNCH_G((gather_p_kernel<D, NS, NP, NPB, true, true>));
      else if(what == 2)
        B200_LAUNCH_G((gather_p_kernel<D, NS, NP, NPB, true, false>));
      else
        B200_LAUNCH_G((gather_p_kernel<D, NS, NP, NPB, false, true>));
    }
#undef B200_LAUNCH_G
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

// lane-group kernels: element-state pre-pass, then one launch per (node set, launch segment)
template <int D, int NS, int NP, int NW, int L, int REGS> static int launch_gather_lane_t(System *S, int what, const THCoeffs &c)
{
  GatherPlan *G = static_cast<GatherPlan *>(S->gather);
  using X = ES<D, NS, NP>;
  using T = GT<D, NS, NP>;
  constexpr int NT = NW * 32, MINB = 65536 / (NT * REGS) > 0 ? 65536 / (NT * REGS) : 1;
  const double *d_source = (c.c_src != 0.) ? S->d_source : nullptr;
  const int     ntab     = d_source ? G->tab_len_src : G->tab_len;
  {
    ElementStateArgs ea;
    ea.nElm   = S->nElm;
    ea.adrU   = S->spaces[S->su].d_adr;
    ea.adrP   = S->spaces[S->sp].d_adr;
    ea.sol    = S->d_sol;
    ea.soldot = S->have_soldot ? S->d_soldot : nullptr;
    ea.source = d_source;
    ea.geo    = G->d_geo;
    ea.tab    = G->d_tab;
    ea.es     = G->d_es;
    ea.nq     = S->nq;
    ea.ntab   = ntab;
    ea.c      = c;
    for(int i = 0; i < 120; ++i) ea.E[i] = G->E[i];
    static_assert(X::W % 2 == 0, "element state records must be 16-byte aligned");
    const size_t smem = (size_t)(ntab - T::O_T3) * sizeof(double);
    B200_CUDA(cudaFuncSetAttribute(element_state_kernel<D, NS, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    element_state_kernel<D, NS, NP><<<(unsigned)((S->nElm + 127) / 128), 128, smem, S->stream>>>(ea);
    count_launch();
  }
  GatherArgs a;
  a.xyz    = S->d_xyz;
  a.conn   = S->d_conn;
  a.adrU   = S->spaces[S->su].d_adr;
  a.adrP   = S->spaces[S->sp].d_adr;
  a.sol    = S->d_sol;
  a.soldot = S->have_soldot ? S->d_soldot : nullptr;
  a.source = d_source;
  a.tab    = G->d_tab;
  a.geo    = G->d_geo;
  a.es     = G->d_es;
  a.ia     = S->d_ia;
  a.val    = S->d_val;
  a.rhs    = S->d_rhs;
  a.nInc   = S->nInc;
  a.nq     = S->nq;
  a.ntab   = G->tab_len; // the W table (sources) is only used by the pre-pass
  a.c      = c;
  a.c0     = S->c0;
  for(int i = 0; i < 120; ++i) a.E[i] = G->E[i];
  const bool mat = what & 2;
  for(int pass = 0; pass < 2; ++pass) {
    const NodeSet &N = pass == 0 ? G->U : G->P;
    if(N.nNodes == 0) continue;
    a.pair     = N.pair;
    a.range    = N.range;
    a.row      = N.row;
    a.smoff    = N.smoff;
    a.cta_perm = N.cta_perm;
    a.cta_size = N.cta_size;
    a.off      = N.off;
    a.nNodes   = N.nNodes;
    a.cta_g0   = N.cta_g0;
    if(!mat) {
      // residual only: sum of the per-element contributions, one thread per node
      a.cta0 = 0;
      const int nc = (N.nNodes + 255) / 256;
      if(pass == 0)
        gather_lane_kernel<D, NS, NP, 8, 1, 1, false, true, false><<<nc, 256, 0, S->stream>>>(a);
      else
        gather_lane_kernel<D, NS, NP, 8, 1, 1, false, true, true><<<nc, 256, 0, S->stream>>>(a);
      count_launch();
      continue;
    }
    const size_t smem_max = ((size_t)a.ntab + N.max_cta + NT) * sizeof(double);
    const int    nseg     = (int)N.seg_smem.size();
#define B200_LAUNCH_L(KERN)                                                                                                              \
  do {                                                                                                                                   \
    B200_CUDA(cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));                                   \
    for(int sg = 0; sg < nseg; ++sg) {                                                                                                   \
      a.cta0            = N.seg_begin[sg];                                                                                               \
      const int    nc   = N.seg_begin[sg + 1] - N.seg_begin[sg];                                                                         \
      const size_t smem = ((size_t)a.ntab + N.seg_smem[sg] + NT) * sizeof(double);                                                       \
      KERN<<<nc, NT, smem, S->stream>>>(a);                                                                                              \
      count_launch();                                                                                                                    \
    }                                                                                                                                    \
  } while(0)
    if(pass == 0) {
      if(what == 3)
        B200_LAUNCH_L((gather_lane_kernel<D, NS, NP, NW, L, MINB, true, true, false>));
      else
        B200_LAUNCH_L((gather_lane_kernel<D, NS, NP, NW, L, MINB, true, false, false>));
    } else {
      if(what == 3)
        B200_LAUNCH_L((gather_lane_kernel<D, NS, NP, NW, L, MINB, true, true, true>));
      else
        B200_LAUNCH_L((gather_lane_kernel<D, NS, NP, NW, L, MINB, true, false, true>));
    }
#undef B200_LAUNCH_L
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}


// row-lane kernels: pre-pass, velocity rows per launch segment, lane-group pressure rows
static int launch_gather_urow(System *S, int what, const THCoeffs &c)
{
  constexpr int D = 3, NS = 10, NP = 4, NW = 4, L = 10, NT = NW * 32, MINB = 4;
  GatherPlan *G = static_cast<GatherPlan *>(S->gather);
  using T = GT<D, NS, NP>;
  const double *d_source = (c.c_src != 0.) ? S->d_source : nullptr;
  const int     ntab     = d_source ? G->tab_len_src : G->tab_len;
  if(g_urow_owner != G) {
    // the tensors of another system (another quadrature rule) are in the constant bank: drain the device before replacing them
    if(g_urow_owner != nullptr) B200_CUDA(cudaDeviceSynchronize());
    B200_CUDA(cudaMemcpyToSymbolAsync(b200_urow_tab, G->utab.data(), G->utab.size() * sizeof(double), 0, cudaMemcpyHostToDevice, S->stream));
    B200_CUDA(cudaMemcpyToSymbolAsync(b200_es_tab, G->estab.data(), G->estab.size() * sizeof(double), 0, cudaMemcpyHostToDevice, S->stream));
    g_urow_owner = G;
  }
  {
    ElementStateArgs ea;
    ea.nElm   = S->nElm;
    ea.adrU   = S->spaces[S->su].d_adr;
    ea.adrP   = S->spaces[S->sp].d_adr;
    ea.sol    = S->d_sol;
    ea.soldot = S->have_soldot ? S->d_soldot : nullptr;
    ea.source = d_source;
    ea.geo    = G->d_geo;
    ea.tab    = G->d_tab;
    ea.es     = G->d_es;
    ea.nq     = S->nq;
    ea.ntab   = ntab;
    ea.c      = c;
    for(int i = 0; i < 120; ++i) ea.E[i] = G->E[i];
    const size_t smem = d_source ? (size_t)S->nq * NS * sizeof(double) : 0;
    element_state_urow_kernel<<<(unsigned)((S->nElm + 127) / 128), 128, smem, S->stream>>>(ea);
    count_launch();
  }
  const bool mat = what & 2, res = what & 1;
  GatherArgs a;
  a.xyz    = S->d_xyz;
  a.conn   = S->d_conn;
  a.adrU   = S->spaces[S->su].d_adr;
  a.adrP   = S->spaces[S->sp].d_adr;
  a.sol    = S->d_sol;
  a.soldot = S->have_soldot ? S->d_soldot : nullptr;
  a.source = d_source;
  a.tab    = G->d_tab;
  a.geo    = G->d_geo;
  a.es     = G->d_es;
  a.ia     = S->d_ia;
  a.val    = S->d_val;
  a.rhs    = S->d_rhs;
  a.nInc   = S->nInc;
  a.nq     = S->nq;
  a.ntab   = G->tab_len;
  a.c      = c;
  a.c0     = S->c0;
  for(int i = 0; i < 120; ++i) a.E[i] = G->E[i];
  if(mat) {
    URowArgs ua;
    ua.geo4  = G->d_geo4;
    ua.es    = G->d_es;
    ua.rec   = static_cast<const URowPair *>(G->d_rec);
    ua.order = G->d_order;
    ua.wstep = G->d_wstep;
    ua.sched = G->d_sched;
    ua.wmax  = G->d_sla;
    ua.range = G->U.range;
    ua.row   = G->U.row;
    ua.ia    = S->d_ia;
    ua.val   = S->d_val;
    ua.rhs   = S->d_rhs;
    ua.count = G->U.nNodes;
    ua.nInc  = S->nInc;
    ua.nsm   = -c.sig_mu;
    ua.cdk   = c.diff_k - c.sig_mu;
    ua.mass0 = c.c_mass * S->c0;
    ua.cpr   = c.c_sig - c.c_gradp;
    const int nseg = (int)G->useg_lmax.size();
    for(int sg = 0; sg < nseg; ++sg) {
      constexpr int NGRP = 4; // node groups per CTA
      ua.warp0          = G->useg_begin[sg];
      ua.warp_end       = G->useg_begin[sg + 1];
      const int    nc   = (ua.warp_end - ua.warp0 + NGRP - 1) / NGRP;
      const bool   big  = (size_t)(G->useg_lmax[sg] + 3) * 32 * sizeof(double) > 24 * 1024;
      const size_t smem = ((size_t)(G->useg_lmax[sg] + 3) * 32 + 32 * (big ? 8 : 16)) * sizeof(double);
#define B200_LAUNCH_C(KERN)                                                                                         \
  do {                                                                                                              \
    B200_CUDA(cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                  \
    KERN<<<nc, 32, smem, S->stream>>>(ua);                                                                          \
    count_launch();                                                                                                 \
  } while(0)
      // long rows (vertex nodes): 4 warps per SM fit anyway; short rows: up to 8 warps per SM, 255 registers for the software pipeline
      if(big) {
        if(res)
          B200_LAUNCH_C((gather_urow_kernel<true, 4, NGRP, 8>));
        else
          B200_LAUNCH_C((gather_urow_kernel<false, 4, NGRP, 8>));
      } else {
        if(res)
          B200_LAUNCH_C((gather_urow_kernel<true, 8, NGRP, 16>));
        else
          B200_LAUNCH_C((gather_urow_kernel<false, 8, NGRP, 16>));
      }
#undef B200_LAUNCH_C
    }
  }
  for(int pass = mat ? 1 : 0; pass < 2; ++pass) {
    const NodeSet &N = pass == 0 ? G->U : G->P;
    if(N.nNodes == 0) continue;
    a.pair     = N.pair;
    a.range    = N.range;
    a.row      = N.row;
    a.smoff    = N.smoff;
    a.cta_perm = N.cta_perm;
    a.cta_size = N.cta_size;
    a.off      = N.off;
    a.nNodes   = N.nNodes;
    a.cta_g0   = N.cta_g0;
    if(!mat) {
      a.cta0 = 0;
      const int nc = (N.nNodes + 255) / 256;
      if(pass == 0)
        gather_lane_kernel<D, NS, NP, 8, 1, 1, false, true, false, true><<<nc, 256, 0, S->stream>>>(a);
      else
        gather_lane_kernel<D, NS, NP, 8, 1, 1, false, true, true, true><<<nc, 256, 0, S->stream>>>(a);
      count_launch();
      continue;
    }
    const size_t smem_max = ((size_t)a.ntab + N.max_cta + NT) * sizeof(double);
    const int    nseg     = (int)N.seg_smem.size();
#define B200_LAUNCH_L(KERN)                                                                                                              \
  do {                                                                                                                                   \
    B200_CUDA(cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));                                   \
    for(int sg = 0; sg < nseg; ++sg) {                                                                                                   \
      a.cta0            = N.seg_begin[sg];                                                                                               \
      const int    nc   = N.seg_begin[sg + 1] - N.seg_begin[sg];                                                                         \
      const size_t smem = ((size_t)a.ntab + N.seg_smem[sg] + NT) * sizeof(double);                                                       \
      KERN<<<nc, NT, smem, S->stream>>>(a);                                                                                              \
      count_launch();                                                                                                                    \
    }                                                                                                                                    \
  } while(0)
    if(res)
      B200_LAUNCH_L((gather_lane_kernel<D, NS, NP, NW, L, MINB, true, true, true, true>));
    else
      B200_LAUNCH_L((gather_lane_kernel<D, NS, NP, NW, L, MINB, true, false, true, true>));
#undef B200_LAUNCH_L
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

// what: bit 0 residual, bit 1 matrix; OVERWRITES val / rhs (every row is written exactly once)
int launch_gather(System *S, int what, const THCoeffs &c)
{
  const GatherPlan *G = static_cast<const GatherPlan *>(S->gather);
  if(G->urow) return launch_gather_urow(S, what, c);
  if(G->lane) {
    if(S->dim == 2) {
      switch(G->lanes) {
      case 1: return launch_gather_lane_t<2, 6, 3, 2, 1, 96>(S, what, c);
      case 2: return launch_gather_lane_t<2, 6, 3, 4, 2, 96>(S, what, c);
      case 3: return launch_gather_lane_t<2, 6, 3, 4, 3, 96>(S, what, c);
      default: return launch_gather_lane_t<2, 6, 3, 8, 6, 80>(S, what, c);
      }
    }
    switch(G->lanes) {
    case 1: return launch_gather_lane_t<3, 10, 4, 1, 1, 168>(S, what, c);
    case 2: return launch_gather_lane_t<3, 10, 4, 2, 2, 128>(S, what, c);
    case 5: return launch_gather_lane_t<3, 10, 4, 4, 5, 128>(S, what, c);
    default: return launch_gather_lane_t<3, 10, 4, 4, 10, 128>(S, what, c);
    }
  }
  if(S->dim == 2) return launch_gather_t<2, 6, 3, 64>(S, what, c);
  return launch_gather_t<3, 10, 4, 32>(S, what, c);
}

} // namespace b200
