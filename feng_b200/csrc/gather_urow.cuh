// Row-lane velocity-row kernels for P2/P1 tetrahedra (included by gather.cu after gather_lane.cuh).
//
// What the r02h capture of gather_lane_kernel says (profiles/README.md): the velocity rows run at the speed of the L1 /
// shared-memory data stage -- per (row node, element) pair 55 shared wavefronts (14 of them loads of the reference
// tensors K[la][b], T3[la][b], M[la][b] indexed by the LANE, 46 % bank-conflict replays of the scattered row-image
// updates) and 38 global sectors (three lane groups per warp read three element records) for 17 cycles of FP64 work.
// This file removes those three costs with the same sums as gather_lane_kernel (feSysElm_*::computeAe,
// src/feVectorSysElm.cpp:1171-1204, :1454-1494, :449-503, :528-578, re-associated as in gather.cu):
//
//  1. One lane per MATRIX ROW: lane (g, i) owns row i of node g of its warp (10 nodes x 3 components = 30 lanes) and
//     computes row i of every 3 x 3 block of that node's pairs.  The pairs of every node are sorted by the LOCAL index
//     `la` the node has in the element, the nodes are sorted by the signature (number of pairs per la), and the warp
//     walks la = 0 .. 9 together: at any time all lanes of a warp work on the same (la, column b).  The reference
//     tensors are therefore indexed by warp-uniform values: they live in __constant__ memory and reach the DFMAs through
//     uniform registers (LDCU) -- no shared-memory or L1 traffic at all.
//  2. The row image of lane l is INTERLEAVED in shared memory: entry k of lane l lives at word k * 32 + l, so a warp-wide
//     read-modify-write touches 32 different bank pairs whatever the column offsets are -- zero bank conflicts by
//     construction, and no __syncwarp between elements (images are lane-private).
//  3. 256-bit global loads (ld.global.nc.v4.f64, sm_100) of 32-byte-aligned records: a pair needs 12 load instructions
//     instead of 39 sectors spread over 25.
//  The finished images are transposed through a swizzled 2 KB staging tile and written once with 64-byte runs.
//  A lane whose node has no pair for the (la, step) its warp is at recomputes one of its pairs into its trash entries
//  (the loop stays convergent, which keeps the constant loads on the uniform datapath); with nodes sorted by signature
//  that only happens where the sort changes signature.
#pragma once

namespace b200 {

struct __align__(32) D4 {
  double x, y, z, w;
};
__device__ __forceinline__ D4 ldg256(const void *p)
{
  D4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
struct U8 {
  uint32_t w[8];
};
__device__ __forceinline__ U8 ldg256u(const void *p)
{
  U8 r;
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
               : "l"(p));
  return r;
}
// 16-bit field t of a pair record held in registers (word 0 = ea, fields start at word 1); t is a compile-time constant after unrolling
__device__ __forceinline__ uint32_t rec_off(const U8 &r, int t) { return (t & 1) ? (r.w[1 + (t >> 1)] >> 16) : (r.w[1 + (t >> 1)] & 0xffffu); }
__device__ __forceinline__ void stg256(void *p, const D4 &v)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}

// Reference tensors per local row node la, unscaled (the coefficients of the registered forms are folded into per-pair geometry
// factors).  They are indexed with warp-uniform values: la, plus an offset that depends on the loop counter and is always 0,
// which the compiler cannot prove.  FP64 instructions of sm_100 take constants through UNIFORM REGISTERS (DFMA R, R, UR, R fed
// by LDCU); without that dependence ptxas hoists all 212 constants of a row out of the pair loop, runs out of uniform registers,
// copies them to vector registers and spills (measured: 255 registers + 260 bytes of spill traffic per pair).
struct URowTab {
  static constexpr int O_K = 0;    // [10][3][3] Kref[la][b][al][be]
  static constexpr int O_KS = 90;  // [10][6]    {K00, K11, K22, K01 + K10, K02 + K20, K12 + K21}
  static constexpr int O_T3 = 150; // [10][4]    T3[la][b][v]
  static constexpr int O_M = 190;  // [10]       Mref[la][b]
  static constexpr int O_B = 200;  // [4][3]     Bref[q][la][al]
  static constexpr int LEN = 212;
};
__constant__ double b200_urow_tab[10][URowTab::LEN];
// local edge -> end vertices of the reference's P2 tetrahedron (src/feTetrahedron.h:31; amg.cu checks the same table against the basis
// tabulation); used only to place mid-edge nodes in space for the launch order
__constant__ int c_edge_tet_u[6][2] = {{0, 2}, {2, 1}, {1, 0}, {1, 3}, {3, 0}, {3, 2}};

// ----------------------------------------------------------------------------------------------------------------------
// pre-pass: everything that depends on the element only (as element_state_kernel), in the record layout ESC
// ----------------------------------------------------------------------------------------------------------------------
// Reference tensors of the pre-pass in __constant__ memory, indexed by the row counter aa of its (not unrolled) loop: they reach the
// DFMAs through uniform registers; the shared-memory copy of element_state_kernel cost 22 of the 126 L1 cycles per element (r02p).
struct ESTab {
  static constexpr int O_T3 = 0;   // [10][10][4] T3[a][c][v]
  static constexpr int O_B = 400;  // [4][10][3]  Bref[q][a][al]
  static constexpr int O_M = 520;  // [10][10]    Mref[a][b]
  static constexpr int LEN = 620;
};
__constant__ double b200_es_tab[ESTab::LEN];

__global__ void __launch_bounds__(128) element_state_urow_kernel(const ElementStateArgs a)
{
  constexpr int D = 3, NS = 10, NP = 4;
  using T = GT<D, NS, NP>;
  using X = ESC;
  using ET = ESTab;
  constexpr int NU = NS * D, GW = T::GW;
  extern __shared__ double s_w[]; // W[k][a] = w_k phi_a(k): source forms only
  if(a.source != nullptr) {
    for(int i = threadIdx.x; i < a.nq * NS; i += blockDim.x) s_w[i] = a.tab[T::O_W + i];
    __syncthreads();
  }
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if(e >= a.nElm) return;
  double G[D * D], J;
  {
    const double2 *ge = reinterpret_cast<const double2 *>(a.geo + e * GW);
    double         g[GW];
#pragma unroll
    for(int i = 0; i < GW / 2; ++i) {
      const double2 v = ge[i];
      g[2 * i]     = v.x;
      g[2 * i + 1] = v.y;
    }
#pragma unroll
    for(int i = 0; i < D * D; ++i) G[i] = g[i];
    J = g[D * D];
  }
  double U[NS][D];
  {
    const int2 *a2 = reinterpret_cast<const int2 *>(a.adrU + e * NU);
#pragma unroll
    for(int k = 0; k < NU / 2; ++k) {
      const int2 v = a2[k];
      U[(2 * k) / D][(2 * k) % D]         = a.sol[v.x];
      U[(2 * k + 1) / D][(2 * k + 1) % D] = a.sol[v.y];
    }
  }
  double P[NP];
#pragma unroll
  for(int q = 0; q < NP; ++q) P[q] = a.sol[a.adrP[e * NP + q]];
  const THCoeffs c   = a.c;
  double        *out = a.es + e * X::W;
  // velocity gradient at the vertices: gu[v][j][i] = d_j u_i (v)   (src/feSpace.cpp:1352-1405)
  double gu[NP * D * D];
#pragma unroll
  for(int i = 0; i < D; ++i) {
    double Xi[D][NP];
#pragma unroll
    for(int al = 0; al < D; ++al)
#pragma unroll
      for(int v = 0; v < NP; ++v) Xi[al][v] = 0.;
#pragma unroll
    for(int cc = 0; cc < NS; ++cc)
#pragma unroll
      for(int al = 0; al < D; ++al)
#pragma unroll
        for(int v = 0; v < NP; ++v) Xi[al][v] += U[cc][i] * a.E[(cc * D + al) * NP + v];
#pragma unroll
    for(int v = 0; v < NP; ++v)
#pragma unroll
      for(int j = 0; j < D; ++j) {
        double s = 0.;
#pragma unroll
        for(int al = 0; al < D; ++al) s += G[al * D + j] * Xi[al][v];
        gu[(v * D + j) * D + i] = s;
      }
  }
  const double sJ = c.c_conv * J;
#pragma unroll
  for(int i = 0; i < D; ++i) {
    double t[12]; // DvT[i][v][j]
#pragma unroll
    for(int v = 0; v < NP; ++v)
#pragma unroll
      for(int j = 0; j < D; ++j) t[v * 3 + j] = sJ * gu[(v * D + j) * D + i];
#pragma unroll
    for(int k = 0; k < 3; ++k) {
      D4 o;
      o.x = t[4 * k];
      o.y = t[4 * k + 1];
      o.z = t[4 * k + 2];
      o.w = t[4 * k + 3];
      stg256(out + X::O_DVT + i * 12 + k * 4, o);
    }
  }
  // divergence part of the residual, P rows: rp[q] = -c_div J sum_{b, j, al} G[al][j] Bref[q][b][al] U[b][j]
  double rp[NP];
  {
    double GU[NS][D]; // GU[b][al] = sum_j G[al][j] U[b][j]
#pragma unroll
    for(int b = 0; b < NS; ++b)
#pragma unroll
      for(int al = 0; al < D; ++al) GU[b][al] = G[al * D + 0] * U[b][0] + G[al * D + 1] * U[b][1] + G[al * D + 2] * U[b][2];
#pragma unroll
    for(int q = 0; q < NP; ++q) {
      double s = 0.;
#pragma unroll
      for(int b = 0; b < NS; ++b)
#pragma unroll
        for(int al = 0; al < D; ++al) s += b200_es_tab[ET::O_B + (q * NS + b) * D + al] * GU[b][al];
      rp[q] = -c.c_div * J * s;
    }
  }
  const bool   domass = (c.c_mass != 0.) && (a.soldot != nullptr);
  const double cvis1 = c.sig_mu - c.diff_k, cvis2 = c.sig_mu, cpre = c.c_gradp - c.c_sig;
  // the loop counter aa indexes the constant tables ONLY (uniform datapath); everything per-thread that depends on the row goes
  // through `orow` / `av`, which the compiler cannot merge with aa
  double *orow = out + X::O_ROW;
  int     av   = 0;
#pragma unroll 1
  for(int aa = 0; aa < NS; ++aa) {
    const double *t3 = &b200_es_tab[ET::O_T3 + aa * NS * NP];
    // C1[aa][b] = c_conv J sum_{al, v} E[b][al][v] Z[al][v],  Z[al][v] = sum_m G[al][m] Y[m][v],  Y[m][v] = sum_c U[c][m] T3[aa][c][v]
    double Y[D * NP];
#pragma unroll
    for(int i = 0; i < D * NP; ++i) Y[i] = 0.;
#pragma unroll
    for(int cc = 0; cc < NS; ++cc)
#pragma unroll
      for(int v = 0; v < NP; ++v) {
        const double t = t3[cc * NP + v];
#pragma unroll
        for(int m = 0; m < D; ++m) Y[m * NP + v] += U[cc][m] * t;
      }
    double Z[D * NP];
#pragma unroll
    for(int al = 0; al < D; ++al)
#pragma unroll
      for(int v = 0; v < NP; ++v) Z[al * NP + v] = sJ * (G[al * D + 0] * Y[0 * NP + v] + G[al * D + 1] * Y[1 * NP + v] + G[al * D + 2] * Y[2 * NP + v]);
    double r[D];
#pragma unroll
    for(int i = 0; i < D; ++i) r[i] = 0.;
    {
      double c1[NS];
#pragma unroll
      for(int b = 0; b < NS; ++b) {
        double s = 0.;
#pragma unroll
        for(int i = 0; i < D * NP; ++i) s += a.E[b * D * NP + i] * Z[i];
        c1[b] = s;
#pragma unroll
        for(int i = 0; i < D; ++i) r[i] -= s * U[b][i];
      }
      D4 o;
      o.x = c1[0], o.y = c1[1], o.z = c1[2], o.w = c1[3];
      stg256(orow, o);
      o.x = c1[4], o.y = c1[5], o.z = c1[6], o.w = c1[7];
      stg256(orow + 4, o);
      orow[8] = c1[8];
      orow[9] = c1[9];
    }
    // viscous and pressure parts through Bp[v][aa][m] = J sum_al G[al][m] Bref[v][aa][al]
#pragma unroll
    for(int v = 0; v < NP; ++v) {
      const double *Br = &b200_es_tab[ET::O_B + v * NS * D + aa * D];
#pragma unroll
      for(int m = 0; m < D; ++m) {
        const double bp = J * (G[0 * D + m] * Br[0] + G[1 * D + m] * Br[1] + G[2 * D + m] * Br[2]);
        r[m] += cpre * bp * P[v];
#pragma unroll
        for(int i = 0; i < D; ++i) r[i] += bp * (cvis1 * gu[(v * D + m) * D + i] + cvis2 * gu[(v * D + i) * D + m]);
      }
    }
    if(domass) {
      const int32_t *ad = a.adrU + e * NU;
#pragma unroll
      for(int b = 0; b < NS; ++b) {
        const double mab = c.c_mass * J * b200_es_tab[ET::O_M + aa * NS + b];
#pragma unroll
        for(int i = 0; i < D; ++i) r[i] -= mab * a.soldot[ad[b * D + i]];
      }
    }
    if(a.source != nullptr) {
      const double *src = a.source + e * a.nq * D;
      for(int k = 0; k < a.nq; ++k) {
        const double wj = J * s_w[k * NS + av];
#pragma unroll
        for(int i = 0; i < D; ++i) r[i] -= wj * src[k * D + i];
      }
    }
    // slots 10..12: the velocity-row residual; slot 13: pressure-row residual of local pressure node aa (rows 0..3)
    double p13 = 0.;
#pragma unroll
    for(int q = 0; q < NP; ++q) p13 = (av == q) ? rp[q] : p13;
    orow[10] = r[0];
    orow[11] = r[1];
    {
      D4 o;
      o.x = r[2], o.y = p13, o.z = 0., o.w = 0.;
      stg256(orow + 12, o);
    }
    orow += 16;
    av = __shfl_sync(0xffffffffu, av + 1, threadIdx.x & 31); // same value, opaque to the induction-variable optimiser
  }
}

// ----------------------------------------------------------------------------------------------------------------------
// plan kernels
// ----------------------------------------------------------------------------------------------------------------------
// geo4[e][v][0..2] = physical gradient of the barycentric coordinate lambda_v (rows 1..3 = the inverse affine map), [v][3] = detJ
__global__ void urow_geo4_kernel(int64_t nElm, const double *__restrict__ geo, int gw, double *__restrict__ geo4)
{
  for(int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nElm; e += (int64_t)gridDim.x * blockDim.x) {
    const double *g = geo + e * gw;
    const double  J = g[9];
    double       *o = geo4 + e * 16;
    double        s[3] = {0., 0., 0.};
    for(int al = 0; al < 3; ++al) {
      for(int m = 0; m < 3; ++m) {
        o[(al + 1) * 4 + m] = g[al * 3 + m];
        s[m] -= g[al * 3 + m];
      }
      o[(al + 1) * 4 + 3] = J;
    }
    for(int m = 0; m < 3; ++m) o[m] = s[m];
    o[3] = J;
  }
}

// sort key of a pair: (first pair of its node, local index of the node in the element); the stable sort keeps the elements of one
// (node, la) ascending.  Pairs of nodes without unknown rows are not covered by any range: their key is their own position, which
// keeps them where they are (urow_pair_key_init_kernel runs first).
__global__ void urow_pair_key_init_kernel(int64_t nPairs, uint64_t *key)
{
  for(int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nPairs; p += (int64_t)gridDim.x * blockDim.x) key[p] = (uint64_t)p << 4;
}
__global__ void urow_pair_key_kernel(int32_t nNodes, const int2 *__restrict__ range, const int32_t *__restrict__ pair, uint64_t *key)
{
  for(int32_t n = blockIdx.x; n < nNodes; n += gridDim.x) {
    const int2 rg = range[n];
    for(int k = threadIdx.x; k < rg.y; k += blockDim.x) key[rg.x + k] = ((uint64_t)rg.x << 4) | (uint64_t)(pair[rg.x + k] % 10);
  }
}

// per node: pairs per local index (6 bits each) and the sort key
//   [shared-memory class : 2][Morton cell of the node : 20][row length : 16][hash of the signature : 26]
// class: rows of similar length share a launch (its shared-memory request); cell: nodes of one neighbourhood are processed together,
// so the ten (node, element) pairs that read the record of an element find it in L2; signature (pairs per local index): the nodes of
// a warp walk the local indices together, equal signatures leave no lane idle.
__device__ __forceinline__ uint32_t urow_spread3(uint32_t v)
{
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__host__ __device__ inline int urow_class(int64_t len) { return len <= 96 ? 0 : len <= 256 ? 1 : 2; }

__global__ void urow_node_key_kernel(int32_t nNodes, const int2 *__restrict__ range, const int32_t *__restrict__ pair, const int32_t *__restrict__ row,
                                     const int64_t *__restrict__ ia, int64_t nInc, const int32_t *__restrict__ conn, const double *__restrict__ xyz,
                                     double x0, double y0, double z0, double sx, double sy, double sz, int cellshift, uint64_t *key, int32_t *idx,
                                     uint64_t *lacnt, int *err)
{
  for(int32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < nNodes; n += gridDim.x * blockDim.x) {
    const int2 rg = range[n];
    int        c[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for(int k = 0; k < rg.y; ++k) {
      const int la = pair[rg.x + k] % 10;
#pragma unroll
      for(int l = 0; l < 10; ++l) c[l] += la == l;
    }
    int64_t len = 0;
    for(int cc = 2; cc >= 0; --cc) {
      const int32_t r = row[(int64_t)n * 3 + cc];
      if(r < nInc) len = ia[r + 1] - ia[r];
    }
    uint64_t pk = 0, sig = 0;
#pragma unroll
    for(int l = 0; l < 10; ++l) {
      if(c[l] > 63) atomicExch(err, 6);
      pk |= (uint64_t)(c[l] & 63) << (6 * l);
      sig |= (uint64_t)min(c[l], 15) << (4 * l);
    }
    // position of the node: its corner / mid-edge point in the first adjacent element
    uint32_t cell = 0;
    {
      const int     ea = pair[rg.x];
      const int64_t e  = ea / 10;
      const int     la = ea - (int)e * 10;
      const int     va = la < 4 ? la : c_edge_tet_u[la - 4][0], vb = la < 4 ? la : c_edge_tet_u[la - 4][1];
      const double *pa = xyz + 3 * (int64_t)conn[e * 4 + va], *pb = xyz + 3 * (int64_t)conn[e * 4 + vb];
      const uint32_t qx = (uint32_t)fmin(1023., fmax(0., (0.5 * (pa[0] + pb[0]) - x0) * sx));
      const uint32_t qy = (uint32_t)fmin(1023., fmax(0., (0.5 * (pa[1] + pb[1]) - y0) * sy));
      const uint32_t qz = (uint32_t)fmin(1023., fmax(0., (0.5 * (pa[2] + pb[2]) - z0) * sz));
      cell = (urow_spread3(qx) | (urow_spread3(qy) << 1) | (urow_spread3(qz) << 2)) >> cellshift;
    }
    const uint64_t h = (sig * 0x9E3779B97F4A7C15ull) >> 38; // 26 bits
    lacnt[n] = pk;
    key[n]   = ((uint64_t)urow_class(len) << 62) | ((uint64_t)(cell & 0xfffffu) << 42) | ((uint64_t)(len & 0xffff) << 26) | h;
    idx[n]   = n;
  }
}

// schedule of a warp (10 consecutive nodes of the sorted order): the warp walks la = 0 .. 9, max_g cnt_g(la) steps each
__global__ void urow_sched_count_kernel(int32_t nWarps, int32_t nNodes, const int32_t *__restrict__ order, const uint64_t *__restrict__ lacnt, int32_t *nsteps)
{
  for(int32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < nWarps; w += gridDim.x * blockDim.x) {
    int tot = 0;
    for(int la = 0; la < 10; ++la) {
      int mx = 0;
      for(int g = 0; g < 10; ++g) {
        const int32_t k = w * 10 + g;
        if(k < nNodes) mx = max(mx, (int)((lacnt[order[k]] >> (6 * la)) & 63u));
      }
      tot += mx;
    }
    nsteps[w] = tot;
  }
}

// sched[step][g] = pair of group g at this step or -1 (the node has no pair left with this local index); wmax[warp] = steps the
// warp spends on every local index la, 6 bits each
__global__ void urow_sched_fill_kernel(int32_t nWarps, int32_t nNodes, const int32_t *__restrict__ order, const uint64_t *__restrict__ lacnt,
                                       const int2 *__restrict__ range, const int32_t *__restrict__ wstep, int32_t *sched, uint64_t *wmax)
{
  for(int32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < nWarps; w += gridDim.x * blockDim.x) {
    uint64_t cn[10];
    int      base[10];
    for(int g = 0; g < 10; ++g) {
      const int32_t k = w * 10 + g;
      cn[g]   = k < nNodes ? lacnt[order[k]] : 0ull;
      base[g] = k < nNodes ? range[order[k]].x : 0;
    }
    int64_t  s = wstep[w];
    uint64_t pk = 0;
    for(int la = 0; la < 10; ++la) {
      int mx = 0;
      for(int g = 0; g < 10; ++g) mx = max(mx, (int)((cn[g] >> (6 * la)) & 63u));
      pk |= (uint64_t)mx << (6 * la);
      for(int t = 0; t < mx; ++t, ++s)
        for(int g = 0; g < 10; ++g) sched[s * 10 + g] = t < (int)((cn[g] >> (6 * la)) & 63u) ? base[g] + t : -1;
      for(int g = 0; g < 10; ++g) base[g] += (int)((cn[g] >> (6 * la)) & 63u);
    }
    wmax[w] = pk;
  }
}

struct __align__(32) URowPair {
  int32_t  ea;      // (element << 4) | local row node
  uint16_t off[14]; // row-local offset of component 0 of the column nodes 0..9, then of the pressure columns 0..3; 0xFFFF: not assembled
};

// one block per node: records of its pairs.  err: 1 column missing from the pattern, 5 the three components of a column node are
// not adjacent columns / only partly unknown (the general lane kernels handle those systems)
__global__ void urow_pairs_kernel(int32_t nNodes, const int2 *__restrict__ range, const int32_t *__restrict__ row, const int32_t *__restrict__ pair,
                                  const int32_t *__restrict__ adrU, const int32_t *__restrict__ adrP, const int64_t *__restrict__ ia,
                                  const int32_t *__restrict__ ja, int64_t nInc, int colmaskU, int colmaskP, URowPair *rec, int *err)
{
  for(int32_t n = blockIdx.x; n < nNodes; n += gridDim.x) {
    const int2 rg = range[n];
    int32_t    r0 = -1;
    for(int c = 2; c >= 0; --c)
      if(row[(int64_t)n * 3 + c] < nInc) r0 = row[(int64_t)n * 3 + c];
    const int64_t beg = ia[r0], end = ia[r0 + 1];
    for(int idx = threadIdx.x; idx < rg.y * 16; idx += blockDim.x) {
      const int     pp = idx >> 4, t = idx & 15;
      const int64_t p  = rg.x + pp;
      const int     ea = pair[p];
      const int64_t e  = ea / 10;
      const int     la = ea - (int)e * 10;
      if(t == 15) {
        rec[p].ea = (int32_t)((e << 4) | la);
        continue;
      }
      if(t == 14) continue;
      uint16_t   o = 0xFFFF;
      int32_t    col = -1;
      const bool isU = t < 10;
      if(isU) {
        if(colmaskU) {
          const int32_t c0 = adrU[e * 30 + t * 3], c1 = adrU[e * 30 + t * 3 + 1], c2 = adrU[e * 30 + t * 3 + 2];
          if(c0 < nInc || c1 < nInc || c2 < nInc) {
            if(c0 < nInc && c1 == c0 + 1 && c2 == c0 + 2 && c2 < nInc)
              col = c0;
            else
              atomicExch(err, 5);
          }
        }
      } else if(colmaskP) {
        const int32_t cq = adrP[e * 4 + (t - 10)];
        if(cq < nInc) col = cq;
      }
      if(col >= 0) {
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
          if(isU && (lo + 2 >= end || ja[lo + 1] != col + 1 || ja[lo + 2] != col + 2)) atomicExch(err, 5);
        } else
          atomicExch(err, 1);
      }
      rec[p].off[t] = o;
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------------
// velocity rows
// ----------------------------------------------------------------------------------------------------------------------
struct URowArgs {
  const double   *geo4, *es;
  const URowPair *rec;
  const int32_t  *order; // nodes sorted by (class, cell, row length, signature)
  const int32_t  *wstep; // [warp + 1] first step of the warp
  const int32_t  *sched; // [step][10] pair of group g or -1
  const uint64_t *wmax;  // [warp] steps per local index la, 6 bits each
  const int2     *range;
  const int32_t  *row;
  const int64_t  *ia;
  double         *val, *rhs;
  int32_t         count, warp0, warp_end; // nodes, first and one-past-last node group (10 nodes) of this launch
  int64_t         nInc;
  double          nsm, cdk, mass0, cpr; // -sig_mu, diff_k - sig_mu, c_mass c0, c_sig - c_gradp
};

// One warp per CTA; the CTA works through NGRP consecutive node groups (10 nodes each) of its launch segment.  The schedule steps of
// consecutive groups are consecutive in `sched`, so the software pipeline of the global loads runs across the group boundaries: the
// first steps of the next group are in flight while the finished images of this group are written out (a warp that handles a single
// group of mid-edge nodes lives for ~5 steps and spends a third of them waiting for its first loads, r02m).
template <bool RES, int MINB, int NGRP, int TW> __global__ void __launch_bounds__(32, MINB) gather_urow_kernel(const URowArgs a)
{
  using X = ESC;
  using CT = URowTab;
  extern __shared__ double sm[];
  const int lane = threadIdx.x;
  const int g = lane / 3, i = lane - 3 * g;
  const int gc = min(g, 9);
  const int wbeg = a.warp0 + (int)blockIdx.x * NGRP, wend = min(wbeg + NGRP, a.warp_end);
  // component selectors as 0 / 1 factors: a per-lane `i == j ? x : y` inside the step loop may be compiled into a divergent branch,
  // and a divergent branch in the loop body makes ptxas give up the uniform datapath for the constant loads (products with 0 / 1 and
  // sums with +-0 are exact)
  const double mi[3] = {i == 0 ? 1. : 0., i == 1 ? 1. : 0., i == 2 ? 1. : 0.};

  // per-group data of lane (g, i): its row, the length of the rows of its node, the first CSR slot of its row
  struct Meta {
    int32_t row;
    int     len;
    int64_t ia;
  };
  auto load_meta = [&](int w) -> Meta {
    Meta      m = {0x7fffffff, 0, -1};
    const int k = w * 10 + g;
    if(g < 10 && w < wend && k < a.count) {
      const int32_t n = a.order[k];
      m.row           = a.row[(int64_t)n * 3 + i];
      // the unknown rows of a node share their length
#pragma unroll
      for(int c = 2; c >= 0; --c) {
        const int32_t r = a.row[(int64_t)n * 3 + c];
        if(r < a.nInc) m.len = (int)(a.ia[r + 1] - a.ia[r]);
      }
      if(m.row < a.nInc) m.ia = a.ia[m.row];
    }
    return m;
  };

  // The warp follows the precomputed schedule: per group it walks the local indices la = 0 .. 9, wmax[w] steps each; at flat step s
  // group g works on pair sched[s][g] (idle: record 0 into the trash entries).  Counted loops with uniform bounds and NO branch in the
  // body: ptxas keeps the constant loads on the uniform datapath only when the table index is a plain loop counter (a REDUX result or
  // a loaded value ends up in vector LDCs, measured) and the control flow is provably convergent.  The global loads are software-
  // pipelined over the flat step index: schedule entries three steps ahead, pair records two, element data one.
  const int s_beg = a.wstep[wbeg], s_end = a.wstep[wend];
  auto sched_at = [&](int s) -> int { // branch-free (clamped load + select)
    const int v = a.sched[(size_t)max(min(s, s_end - 1), s_beg) * 10 + gc];
    return (s < s_end && g < 10) ? v : -1;
  };
  struct PData {
    D4 g1, g2, g3, dv[3], r0, r1, r2, r3;
  };
  auto load_data = [&](const U8 &rec, PData &d) {
    const int64_t e  = (int64_t)(rec.w[0] >> 4);
    const double *ge = a.geo4 + e * 16;
    const double *es = a.es + e * X::W;
    d.g1 = ldg256(ge + 4);
    d.g2 = ldg256(ge + 8);
    d.g3 = ldg256(ge + 12);
#pragma unroll
    for(int v = 0; v < 3; ++v) d.dv[v] = ldg256(es + X::O_DVT + i * 12 + v * 4); // DvT[i][0..11]: 96 contiguous bytes per lane
    // the row index comes from the record, NOT from the step's la (equal for active lanes): a use of the warp-uniform la in per-lane
    // address arithmetic makes ptxas keep it in a vector register and turn the 212 uniform constant loads into vector LDCs
    const double *er = es + X::O_ROW + (int)(rec.w[0] & 15u) * 16;
    d.r0 = ldg256(er);
    d.r1 = ldg256(er + 4);
    d.r2 = ldg256(er + 8);
    d.r3 = ldg256(er + 12);
  };
  int   fs = s_beg; // flat step
  int   iA = sched_at(fs), iB = sched_at(fs + 1), iC = sched_at(fs + 2);
  U8    rA = ldg256u(a.rec + max(iA, 0)), rB = ldg256u(a.rec + max(iB, 0));
  PData dA;
  load_data(rA, dA);
  Meta     mt = load_meta(wbeg);
  uint64_t wm = a.wmax[wbeg];

#pragma unroll 1
  for(int w = wbeg; w < wend; ++w) {
    const Meta     m   = mt;
    const uint64_t wmc = wm;
    mt = load_meta(w + 1); // next group: in flight during this group's steps
    wm = a.wmax[min(w + 1, wend - 1)];
    const int Lmax = __reduce_max_sync(0xffffffffu, m.len);
    const int uz   = __reduce_max_sync(0xffffffffu, (int)(wmc >> 60)); // 0 (the ten 6-bit counts end at bit 59)
    double   *S    = sm;                           // [Lmax + 3][32] interleaved row images, rows Lmax .. Lmax + 2 = trash
    double   *T    = sm + (size_t)(Lmax + 3) * 32; // [32][TW] staging tile of the write-out
    {
      double2 *z = reinterpret_cast<double2 *>(S);
      for(int k = lane; k < (Lmax + 3) * 16; k += 32) z[k] = make_double2(0., 0.);
    }
    __syncwarp();
    double  res = 0.;
    double *Sl  = S + lane;
#pragma unroll 1
    for(int la_c = 0; la_c < 10; ++la_c) {
      const int maxc = __reduce_max_sync(0xffffffffu, (int)((wmc >> (6 * la_c)) & 63u));
#pragma unroll 1
      for(int it = 0; it < maxc; ++it) {
        // ptxas sinks the register loads of step fs + 1 to the middle of this step whatever the source order says (r02o: a third of
        // the stall samples sit on their first use), so the lines they will read are first pulled into L2 here, at the top of the
        // step: lane i of a node asks for the i-th line of the vertex gradients, then for the geometry / the row / the next record.
        // The prefetches sit in a basic block of their own (the scheduler does not move instructions across the always-taken uniform
        // branch `uz == 0` below), otherwise they are sunk next to the loads.
        {
          const int64_t e  = (int64_t)(rB.w[0] >> 4);
          const double *es = a.es + e * X::W;
          prefetch_l2(es + i * 16);
          const void *p2 = i == 0 ? (const void *)(a.geo4 + e * 16) : i == 1 ? (const void *)(es + X::O_ROW + (int)(rB.w[0] & 15u) * 16)
                                                                              : (const void *)(a.rec + max(iC, 0));
          prefetch_l2(p2);
        }
        if(uz != 0) continue; // never: uz is 0 in a uniform register, which the compiler cannot know
        PData dB;
        load_data(rB, dB);                          // step fs + 1
        const U8  rC = ldg256u(a.rec + max(iC, 0)); // step fs + 2
        const int iD = sched_at(fs + 3);
        {
          const bool   act = iA >= 0;
          const U8    &rec = rA;
          const PData &d   = dA;
          // it >> 24 == 0, which the compiler cannot prove: keeps the constant loads inside the step loop (otherwise ptxas hoists the
          // 212 constants of the row into registers and spills); la_c and it are used for nothing else
          const double *ct = &b200_urow_tab[la_c][it >> 24];
          const double  J = d.g1.w;
          const double  Gp[3][3] = {{d.g1.x, d.g1.y, d.g1.z}, {d.g2.x, d.g2.y, d.g2.z}, {d.g3.x, d.g3.y, d.g3.z}};
          const double  c1[10] = {d.r0.x, d.r0.y, d.r0.z, d.r0.w, d.r1.x, d.r1.y, d.r1.z, d.r1.w, d.r2.x, d.r2.y};
          if(RES) res += (act ? 1. : 0.) * (mi[0] * d.r2.z + mi[1] * d.r2.w + mi[2] * d.r3.x);
          // column i of the inverse map; -sig_mu J G; (diff_k - sig_mu) J G G^T (symmetric: 00, 11, 22, 01, 02, 12)
          double gi[3], GJ[3][3], GG[6];
#pragma unroll
          for(int be = 0; be < 3; ++be) gi[be] = mi[0] * Gp[be][0] + mi[1] * Gp[be][1] + mi[2] * Gp[be][2];
          {
            const double nJ = a.nsm * J;
#pragma unroll
            for(int al = 0; al < 3; ++al)
#pragma unroll
              for(int mm = 0; mm < 3; ++mm) GJ[al][mm] = nJ * Gp[al][mm];
            constexpr int pa[6] = {0, 1, 2, 0, 0, 1}, pb[6] = {0, 1, 2, 1, 2, 2};
            const double  cJ = a.cdk * J;
#pragma unroll
            for(int k = 0; k < 6; ++k) GG[k] = cJ * (Gp[pa[k]][0] * Gp[pb[k]][0] + Gp[pa[k]][1] * Gp[pb[k]][1] + Gp[pa[k]][2] * Gp[pb[k]][2]);
          }
          // DvT[i][v][j] = c_conv J d_j u_i (vertex v)
          const double dvv[4][3] = {{d.dv[0].x, d.dv[0].y, d.dv[0].z}, {d.dv[0].w, d.dv[1].x, d.dv[1].y}, {d.dv[1].z, d.dv[1].w, d.dv[2].x},
                                    {d.dv[2].y, d.dv[2].z, d.dv[2].w}};
          const double mJ = a.mass0 * J;

          // NB column nodes per batch: 3 NB independent read-modify-writes in flight
          constexpr int NB = 5;
#pragma unroll
          for(int bb = 0; bb < 10; bb += NB) {
            double  A[NB][3];
            double *q[NB];
#pragma unroll
            for(int h = 0; h < NB; ++h) {
              const int     b  = bb + h;
              const double *Kr = ct + CT::O_K + b * 9;
              const double *Ks = ct + CT::O_KS + b * 6;
              const double *t3 = ct + CT::O_T3 + b * 4;
              double        H[3];
#pragma unroll
              for(int al = 0; al < 3; ++al) H[al] = Kr[al * 3 + 0] * gi[0] + Kr[al * 3 + 1] * gi[1] + Kr[al * 3 + 2] * gi[2];
              double s = c1[b] + mJ * ct[CT::O_M + b];
#pragma unroll
              for(int k = 0; k < 6; ++k) s += Ks[k] * GG[k];
              double tv[4];
#pragma unroll
              for(int v = 0; v < 4; ++v) tv[v] = t3[v];
#pragma unroll
              for(int j = 0; j < 3; ++j) {
                // one DFMA chain per entry: (i == j ? s : 0) - sig_mu K[j][i] + c_conv int phi_a phi_b d_j u_i
                double x = mi[j] * s;
#pragma unroll
                for(int al = 0; al < 3; ++al) x = fma(GJ[al][j], H[al], x);
#pragma unroll
                for(int v = 0; v < 4; ++v) x = fma(dvv[v][j], tv[v], x);
                A[h][j] = x;
              }
              const uint32_t o = rec_off(rec, b);
              q[h]             = Sl + (size_t)((act && o != 0xFFFFu) ? (int)o : Lmax) * 32;
            }
            double old[NB][3];
#pragma unroll
            for(int h = 0; h < NB; ++h)
#pragma unroll
              for(int j = 0; j < 3; ++j) old[h][j] = q[h][j * 32];
#pragma unroll
            for(int h = 0; h < NB; ++h)
#pragma unroll
              for(int j = 0; j < 3; ++j) q[h][j * 32] = old[h][j] + A[h][j];
          }
          // pressure columns: A[(a, i)][q] = (c_sig - c_gradp) int psi_q d_i phi_a   (src/feVectorSysElm.cpp:449-503, :528-578)
          {
            double  v[4], old[4];
            double *q[4];
            const double pJ    = a.cpr * J;
            const double gj[3] = {pJ * gi[0], pJ * gi[1], pJ * gi[2]};
#pragma unroll
            for(int qq = 0; qq < 4; ++qq) {
              const double *Br = ct + CT::O_B + qq * 3;
              v[qq]            = gj[0] * Br[0] + gj[1] * Br[1] + gj[2] * Br[2];
              const uint32_t o = rec_off(rec, 10 + qq);
              q[qq]            = Sl + (size_t)((act && o != 0xFFFFu) ? (int)o : Lmax) * 32;
            }
#pragma unroll
            for(int qq = 0; qq < 4; ++qq) old[qq] = *q[qq];
#pragma unroll
            for(int qq = 0; qq < 4; ++qq) *q[qq] = old[qq] + v[qq];
          }
        }
        iA = iB;
        iB = iC;
        iC = iD;
        rA = rB;
        rB = rC;
        dA = dB;
        fs = __shfl_sync(0xffffffffu, fs + 1, 0); // opaque to the induction-variable optimiser: `it` must stay a pure loop counter
      }
    }
    const bool valid = m.ia >= 0;
    if(RES && valid) a.rhs[m.row] = res;
    __syncwarp();
    // write-out: TW entries of every image per round through a swizzled [32][TW] tile, then runs of TW * 8 bytes per row with 16-byte
    // stores.  A row whose first CSR slot is odd shifts its rounds by one entry (p = 1) so that every 16-byte store is aligned; its
    // entry 0 is stored on its own.  (TW = 8 where the 4 KB tile would cost a resident warp: the long rows of the vertex nodes.)
    {
      constexpr int PR = TW / 2, RPR = 32 / PR; // 16-byte pairs per row of the tile, rows per store round
      const int mylen = valid ? m.len : 0;
      const int p     = valid ? (int)(m.ia & 1) : 0;
      if(valid && p) a.val[m.ia] = Sl[0];
      int64_t   ia_t[PR];
      int       len_t[PR];
      const int c = lane % PR;
#pragma unroll
      for(int tt = 0; tt < PR; ++tt) {
        const int rr = tt * RPR + lane / PR;
        ia_t[tt]     = __shfl_sync(0xffffffffu, m.ia, rr);
        len_t[tt]    = __shfl_sync(0xffffffffu, mylen, rr);
      }
      double2 *T2 = reinterpret_cast<double2 *>(T);
      // 16-byte bank groups: a quarter warp (8 lanes) must hit 8 different ones, in the fill (lane = row) and in the drain (lane = pair)
      auto swz = [](int r) { return TW == 16 ? (r & 7) : ((r >> 1) & 3); };
      for(int k0 = 0; k0 < Lmax; k0 += TW) {
        double v[TW];
#pragma unroll
        for(int k = 0; k < TW; ++k) v[k] = Sl[(size_t)min(k0 + p + k, Lmax + 2) * 32];
#pragma unroll
        for(int k = 0; k < PR; ++k) T2[lane * PR + (k ^ swz(lane))] = make_double2(v[2 * k], v[2 * k + 1]);
        __syncwarp();
#pragma unroll
        for(int tt = 0; tt < PR; ++tt) {
          const int     rr = tt * RPR + lane / PR;
          const double2 x  = T2[rr * PR + (c ^ swz(rr))];
          const int     kg = k0 + (int)(ia_t[tt] & 1) + 2 * c;
          if(kg + 1 < len_t[tt])
            *reinterpret_cast<double2 *>(a.val + ia_t[tt] + kg) = x;
          else if(kg < len_t[tt])
            a.val[ia_t[tt] + kg] = x.x;
        }
        __syncwarp();
      }
    }
  }
}

} // namespace b200
