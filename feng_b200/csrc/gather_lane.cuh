// Element-state pre-pass + lane-group row-owner gather (included by gather.cu after the plan structures).
//
// Same plan, same numbers and the same write-once row images as gather_u_kernel / gather_p_kernel, reorganised in two
// ways that the r01b/r01c profiles asked for (profiles/README.md):
//
//  1. Everything that is a function of the ELEMENT only is computed once per element by element_state_kernel (one
//     thread per element) instead of once per (row node, element) pair:
//         Dv[v][j][i] = c_conv J d_j u_i (vertex v)                     (src/feSpace.cpp:1352-1405)
//         C1[a][b]    = c_conv int phi_a (u . grad phi_b)               (src/feVectorSysElm.cpp:1171-1204)
//         R[a][i], R[q] = the complete element residual -Be of the fused forms
//                                                                       (computeBe of src/feVectorSysElm.cpp:1206-1242,
//                                                                        :1496-1532, :685-749, :449-503, :528-578,
//                                                                        :390-425, :112-128)
//     The residual uses that gradients of the velocity basis lie in the span of the pressure basis (checked at plan
//     time): int grad(phi_a) . grad(u_i) = sum_{v,m} Bp[v][a][m] d_m u_i(v) with Bp[v][a][m] = int psi_v d_m phi_a, the
//     same block that carries the pressure gradient.  The row kernels then only read 8-byte residual contributions
//     and the dependent chain pair -> DOF table -> solution of the old kernels is gone.
//
//  2. The D x D blocks (a, b) of one (row node, element) pair are spread over L lanes (L = 1: thread per node;
//     L = NS: lane per local column node).  Local column nodes of one element are distinct global columns, so the
//     lanes of a group never touch the same entry of the shared-memory row image, and the groups of a warp own
//     different rows: no atomics, deterministic order (ascending element index).  In 3-D, where one node's row image
//     is 2-5 KB, this multiplies the threads per kilobyte of shared memory by L.
#pragma once

namespace b200 {

template <int D, int NS, int NP> struct ES {
  static constexpr int O_DV = 0;                    // [NP][D][D]
  static constexpr int O_C1 = NP * D * D;           // [NS][NS]
  static constexpr int O_RU = O_C1 + NS * NS;       // [NS][D]
  static constexpr int O_RP = O_RU + NS * D;        // [NP]
  static constexpr int W    = (O_RP + NP + 1) / 2 * 2; // doubles per element (16-byte aligned records)
  __host__ __device__ static constexpr int ru(int la, int i) { return O_RU + la * D + i; }
  __host__ __device__ static constexpr int rp(int q) { return O_RP + q; }
};

// per-element state of the canonical kernels (gather_canon.cuh; P2/P1 tetrahedra; records are 128-byte aligned: 208 * 8 = 13 * 128)
struct ESC {
  static constexpr int O_DVT = 0;   // [3][4][3]   DvT[i][v][j] = c_conv J d_j u_i (vertex v): 96 contiguous bytes per component i; 36 .. 47 unused
  static constexpr int O_ROW = 48;  // [10][16]    per local row node la: C1[la][0..9], RU[0..2] at 10..12,
                                    //             slot 13 of rows 0..3: the pressure-row residual RP[q]
  static constexpr int W = 208;
  static constexpr int O_DV = 0, O_C1 = 0; // not used through this layout (names needed by the shared row kernel)
  __host__ __device__ static constexpr int ru(int la, int i) { return O_ROW + la * 16 + 10 + i; }
  __host__ __device__ static constexpr int rp(int q) { return O_ROW + q * 16 + 13; }
};
template <int D, int NS, int NP, bool CAN> struct ESelect {
  using type = ES<D, NS, NP>;
};
template <int D, int NS, int NP> struct ESelect<D, NS, NP, true> {
  using type = ESC;
};

struct ElementStateArgs {
  int64_t        nElm;
  const int32_t *adrU, *adrP;
  const double  *sol, *soldot, *source, *geo, *tab;
  double        *es;
  int            nq, ntab;
  THCoeffs       c;
  double         E[120];
};

template <int D, int NS, int NP> __global__ void __launch_bounds__(128) element_state_kernel(const ElementStateArgs a)
{
  using T = GT<D, NS, NP>;
  using X = ES<D, NS, NP>;
  constexpr int NU = NS * D, GW = T::GW, TOFF = T::O_T3; // the K table is not needed here
  extern __shared__ double s_tab[];
  for(int i = threadIdx.x; i < a.ntab - TOFF; i += blockDim.x) s_tab[i] = a.tab[TOFF + i];
  __syncthreads();
  const double *s_t3 = s_tab + (T::O_T3 - TOFF), *s_m = s_tab + (T::O_M - TOFF), *s_b = s_tab + (T::O_B - TOFF), *s_w = s_tab + (T::O_W - TOFF);
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
  int32_t ad[NU];
  double  U[NS][D];
  {
    const int32_t *au = a.adrU + e * NU;
    if(NU % 4 == 0) {
      const int4 *a4 = reinterpret_cast<const int4 *>(au);
#pragma unroll
      for(int k = 0; k < NU / 4; ++k) {
        const int4 v  = a4[k];
        ad[4 * k + 0] = v.x;
        ad[4 * k + 1] = v.y;
        ad[4 * k + 2] = v.z;
        ad[4 * k + 3] = v.w;
      }
    } else {
      const int2 *a2 = reinterpret_cast<const int2 *>(au); // NU is even for P2 in 2-D and 3-D (12, 30)
#pragma unroll
      for(int k = 0; k < NU / 2; ++k) {
        const int2 v  = a2[k];
        ad[2 * k + 0] = v.x;
        ad[2 * k + 1] = v.y;
      }
    }
#pragma unroll
    for(int c = 0; c < NS; ++c)
#pragma unroll
      for(int m = 0; m < D; ++m) U[c][m] = a.sol[ad[c * D + m]];
  }
  double P[NP];
#pragma unroll
  for(int q = 0; q < NP; ++q) P[q] = a.sol[a.adrP[e * NP + q]];
  const THCoeffs c   = a.c;
  double        *out = a.es + e * X::W;
  // velocity gradient at the vertices: gu[v][j][i] = d_j u_i (v)
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
  {
    const double sJ = c.c_conv * J;
    double2     *o2 = reinterpret_cast<double2 *>(out + X::O_DV);
#pragma unroll
    for(int k = 0; k < NP * D * D / 2; ++k) o2[k] = make_double2(sJ * gu[2 * k], sJ * gu[2 * k + 1]);
  }
  // contravariant velocity DOFs (scaled by c_conv J)
  double Ut[NS][D];
#pragma unroll
  for(int cc = 0; cc < NS; ++cc)
#pragma unroll
    for(int al = 0; al < D; ++al) {
      double s = 0.;
#pragma unroll
      for(int m = 0; m < D; ++m) s += U[cc][m] * G[al * D + m];
      Ut[cc][al] = c.c_conv * J * s;
    }
  // pressure / divergence part of the residual, P rows
  // Bp[q][b][j] = int psi_q d_j phi_b = J sum_al G[al][j] Bref[q][b][al]
  {
    double rp[NP];
#pragma unroll
    for(int q = 0; q < NP; ++q) rp[q] = 0.;
#pragma unroll
    for(int q = 0; q < NP; ++q)
#pragma unroll
      for(int b = 0; b < NS; ++b) {
        const double *Br = s_b + (q * NS + b) * D;
#pragma unroll
        for(int j = 0; j < D; ++j) {
          double s = 0.;
#pragma unroll
          for(int al = 0; al < D; ++al) s += G[al * D + j] * Br[al];
          rp[q] -= c.c_div * J * s * U[b][j];
        }
      }
#pragma unroll
    for(int q = 0; q < NP; ++q) out[X::O_RP + q] = rp[q];
    if((X::O_RP + NP) % 2) out[X::O_RP + NP] = 0.;
  }
  const bool   domass = (c.c_mass != 0.) && (a.soldot != nullptr);
  const double cvis1 = c.sig_mu - c.diff_k, cvis2 = c.sig_mu, cpre = c.c_gradp - c.c_sig;
#pragma unroll 1
  for(int aa = 0; aa < NS; ++aa) {
    // C1[aa][b] = sum_{al,v} E[b][al][v] Z[al][v],  Z[al][v] = sum_c Ut[c][al] T3[aa][c][v]
    double Z[D * NP];
#pragma unroll
    for(int i = 0; i < D * NP; ++i) Z[i] = 0.;
    const double *t3 = s_t3 + aa * NS * NP;
#pragma unroll
    for(int cc = 0; cc < NS; ++cc)
#pragma unroll
      for(int v = 0; v < NP; ++v) {
        const double t = t3[cc * NP + v];
#pragma unroll
        for(int al = 0; al < D; ++al) Z[al * NP + v] += Ut[cc][al] * t;
      }
    double r[D];
#pragma unroll
    for(int i = 0; i < D; ++i) r[i] = 0.;
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
    if(NS % 2 == 0) {
      double2 *o2 = reinterpret_cast<double2 *>(out + X::O_C1 + aa * NS);
#pragma unroll
      for(int k = 0; k < NS / 2; ++k) o2[k] = make_double2(c1[2 * k], c1[2 * k + 1]);
    } else {
#pragma unroll
      for(int b = 0; b < NS; ++b) out[X::O_C1 + aa * NS + b] = c1[b];
    }
    // viscous and pressure parts through Bp[v][aa][m]
#pragma unroll
    for(int v = 0; v < NP; ++v) {
      const double *Br = s_b + (v * NS + aa) * D;
#pragma unroll
      for(int m = 0; m < D; ++m) {
        double s = 0.;
#pragma unroll
        for(int al = 0; al < D; ++al) s += G[al * D + m] * Br[al];
        const double bp = J * s;
        r[m] += cpre * bp * P[v];
#pragma unroll
        for(int i = 0; i < D; ++i) r[i] += bp * (cvis1 * gu[(v * D + m) * D + i] + cvis2 * gu[(v * D + i) * D + m]);
      }
    }
    if(domass) {
#pragma unroll
      for(int b = 0; b < NS; ++b) {
        const double mab = c.c_mass * J * s_m[aa * NS + b];
#pragma unroll
        for(int i = 0; i < D; ++i) r[i] -= mab * a.soldot[ad[b * D + i]];
      }
    }
    if(a.source != nullptr) {
      const double *src = a.source + e * a.nq * D;
      for(int k = 0; k < a.nq; ++k) {
        const double wj = J * s_w[k * NS + aa];
#pragma unroll
        for(int i = 0; i < D; ++i) r[i] -= wj * src[k * D + i];
      }
    }
#pragma unroll
    for(int i = 0; i < D; ++i) out[X::O_RU + aa * D + i] = r[i];
  }
}

// PROW = false: velocity nodes (D rows per node); PROW = true: pressure nodes (1 row per node).
// L lanes per node (a divisor of NS), GPW = 32 / L node groups per warp, NW warps per CTA, NPB = NW * GPW nodes per CTA.
// CAN: the per-element records have the layout of the canonical kernels (pressure rows and residual-only passes)
template <int D, int NS, int NP, int NW, int L, int MINB, bool MAT, bool RES, bool PROW, bool CAN = false>
__global__ void __launch_bounds__(NW * 32, MINB) gather_lane_kernel(const GatherArgs a)
{
  using T = GT<D, NS, NP>;
  using X = typename ESelect<D, NS, NP, CAN>::type;
  static_assert(!CAN || PROW || !MAT, "canonical records: velocity rows are assembled by gather_canon_kernel");
  static_assert(NS % L == 0, "lanes per node must divide the number of velocity nodes");
  constexpr int NU = NS * D, GPW = 32 / L, NPB = NW * GPW, NT = NW * 32, GW = T::GW, NBL = NS / L;
  constexpr int NR   = PROW ? 1 : D;
  constexpr int OFFW = PROW ? T::OFFW_P : T::OFFW_U;
  constexpr int NLOC = PROW ? NP : NS; // local nodes per element of the row space (pair = e * NLOC + local index)
  extern __shared__ double sm[];
  double *s_tab = sm;
  double *s_buf = sm + a.ntab;
  __shared__ int32_t  s_row[NPB * NR];
  __shared__ uint32_t s_base[NPB];
  __shared__ int32_t  s_len[NPB];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane / L, l = lane - g * L;
  if(MAT)
    for(int i = tid; i < a.ntab; i += NT) s_tab[i] = a.tab[i];
  const int32_t  cta  = a.cta_perm ? a.cta_perm[a.cta0 + blockIdx.x] : a.cta0 + (int32_t)blockIdx.x;
  const int      ln   = wid * GPW + g;
  const int32_t  n    = cta * NPB + ln;
  const bool     live = (g < GPW) && n < a.nNodes;
  const uint32_t tot  = MAT ? a.cta_size[cta] : 0u;
  int32_t        row[NR];
  int            len = 0, cnt = 0, p0 = 0;
  uint32_t       base = 0;
#pragma unroll
  for(int c = 0; c < NR; ++c) row[c] = 0x7fffffff;
  if(live) {
#pragma unroll
    for(int c = 0; c < NR; ++c) row[c] = a.row[n * NR + c];
    if(MAT) {
#pragma unroll
      for(int c = NR - 1; c >= 0; --c)
        if(row[c] < a.nInc) len = (int)(a.ia[row[c] + 1] - a.ia[row[c]]);
      base = a.smoff[n];
    }
    const int2 rg = a.range[n];
    p0            = rg.x;
    cnt           = rg.y;
  }
  if(MAT) {
    if(g < GPW && l == 0) {
#pragma unroll
      for(int c = 0; c < NR; ++c) s_row[ln * NR + c] = row[c];
      s_base[ln] = base;
      s_len[ln]  = len;
    }
    for(int i = tid; i < (int)tot; i += NT) s_buf[i] = 0.;
    __syncthreads();
  }

  // shared-memory index of the row image of component c (unknown rows only, packed); entries that are not assembled
  // (essential row or column) go to a per-thread trash slot behind the CTA's images: branch-free updates
  const int trash = (int)tot + tid;
  int       bi[NR];
  {
    uint32_t o = base;
#pragma unroll
    for(int c = 0; c < NR; ++c) {
      bi[c] = row[c] < a.nInc ? (int)o : -1;
      if(row[c] < a.nInc) o += len;
    }
  }
  bool any_row = false;
#pragma unroll
  for(int c = 0; c < NR; ++c) any_row |= row[c] < a.nInc;
  if(!any_row) cnt = 0;
  double res[NR];
#pragma unroll
  for(int c = 0; c < NR; ++c) res[c] = 0.;
  const THCoeffs c     = a.c;
  const double   mass0 = c.c_mass * a.c0;

  // warp-uniform trip count when lanes cooperate (the update of element k+1 must see the update of element k)
  const int maxcnt = (L > 1 && MAT) ? __reduce_max_sync(0xffffffffu, cnt) : cnt;
  for(int it = 0; it < maxcnt; ++it) {
    if(it < cnt) {
      const int     p  = p0 + it;
      const int     ea = a.pair[p];
      const int     e  = ea / NLOC;
      const int     la = ea - e * NLOC;
      const double *es = a.es + (int64_t)e * X::W;
      if(RES && l == 0) {
        if(PROW) {
          res[0] += es[X::rp(la)];
        } else {
#pragma unroll
          for(int i = 0; i < D; ++i) res[i] += es[X::ru(la, i)];
        }
      }
      if(MAT) {
        double G[D * D], J;
        {
          const double2 *ge = reinterpret_cast<const double2 *>(a.geo + (int64_t)e * GW);
          double         gg[GW];
#pragma unroll
          for(int i = 0; i < GW / 2; ++i) {
            const double2 v = ge[i];
            gg[2 * i]     = v.x;
            gg[2 * i + 1] = v.y;
          }
#pragma unroll
          for(int i = 0; i < D * D; ++i) G[i] = gg[i];
          J = gg[D * D];
        }
        const uint16_t *po = a.off + (int64_t)p * OFFW;
        // thread-per-node: the whole offset row of the pair in registers (vector loads); lane groups: 16-bit loads
        uint4 ow4[L == 1 ? OFFW / 8 : 1];
        if(L == 1) {
          const uint4 *src = reinterpret_cast<const uint4 *>(po);
#pragma unroll
          for(int w = 0; w < OFFW / 8; ++w) ow4[w] = src[w];
        }
#define B200_OFF(j) (L == 1 ? off16(ow4, (j)) : (uint32_t)po[(j)])
        if(PROW) {
          // A[q][b, j] = c_div int psi_q d_j phi_b   (feSysElm_MixedDivergence, src/feVectorSysElm.cpp:685-749)
          const double cj = c.c_div * J;
#pragma unroll
          for(int bb = 0; bb < NBL; ++bb) {
            const int     b  = l + bb * L;
            const double *Br = s_tab + T::O_B + (la * NS + b) * D;
            double        v[D];
            int           idx[D];
#pragma unroll
            for(int j = 0; j < D; ++j) {
              double s = 0.;
#pragma unroll
              for(int al = 0; al < D; ++al) s += G[al * D + j] * Br[al];
              v[j]             = cj * s;
              const uint32_t o = B200_OFF(b * D + j);
              idx[j]           = o != 0xFFFFu ? bi[0] + (int)o : trash;
            }
            double old[D];
#pragma unroll
            for(int j = 0; j < D; ++j) old[j] = s_buf[idx[j]];
#pragma unroll
            for(int j = 0; j < D; ++j) s_buf[idx[j]] = old[j] + v[j];
          }
        } else {
          double Dv[NP * D * D];
          {
            const double2 *dv2 = reinterpret_cast<const double2 *>(es + X::O_DV);
#pragma unroll
            for(int k = 0; k < NP * D * D / 2; ++k) {
              const double2 v = dv2[k];
              Dv[2 * k]     = v.x;
              Dv[2 * k + 1] = v.y;
            }
          }
#pragma unroll
          for(int bb = 0; bb < NBL; ++bb) {
            const int     b  = l + bb * L;
            const double  C1 = es[X::O_C1 + la * NS + b];
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
            const double *T3ab = s_tab + T::O_T3 + (la * NS + b) * NP;
            double        t3[NP];
#pragma unroll
            for(int v = 0; v < NP; ++v) t3[v] = T3ab[v];
            const double Mab = J * s_tab[T::O_M + la * NS + b];
            const double s   = C1 + (c.diff_k - c.sig_mu) * trK + mass0 * Mab;
            uint32_t     ow[D];
#pragma unroll
            for(int j = 0; j < D; ++j) ow[j] = B200_OFF(b * D + j);
            double A[D][D];
            int    idx[D][D];
#pragma unroll
            for(int i = 0; i < D; ++i)
#pragma unroll
              for(int j = 0; j < D; ++j) {
                double C2 = 0.; // c_conv int phi_a phi_b d_j u_i
#pragma unroll
                for(int v = 0; v < NP; ++v) C2 += Dv[(v * D + j) * D + i] * t3[v];
                A[i][j]   = (i == j ? s : 0.) - c.sig_mu * K[j][i] + C2;
                idx[i][j] = (bi[i] >= 0 && ow[j] != 0xFFFFu) ? bi[i] + (int)ow[j] : trash;
              }
            double old[D][D];
#pragma unroll
            for(int i = 0; i < D; ++i)
#pragma unroll
              for(int j = 0; j < D; ++j) old[i][j] = s_buf[idx[i][j]];
#pragma unroll
            for(int i = 0; i < D; ++i)
#pragma unroll
              for(int j = 0; j < D; ++j) s_buf[idx[i][j]] = old[i][j] + A[i][j];
          }
          // pressure columns, spread over the lanes of the group
#pragma unroll
          for(int qq = 0; qq < (NP + L - 1) / L; ++qq) {
            const int q = l + qq * L;
            if(q < NP) {
              const double  *Br = s_tab + T::O_B + (q * NS + la) * D;
              const uint32_t o  = B200_OFF(NU + q);
              double         v[D];
              int            idx[D];
#pragma unroll
              for(int i = 0; i < D; ++i) {
                double sg = 0.;
#pragma unroll
                for(int al = 0; al < D; ++al) sg += G[al * D + i] * Br[al];
                v[i]   = (c.c_sig - c.c_gradp) * J * sg;
                idx[i] = (bi[i] >= 0 && o != 0xFFFFu) ? bi[i] + (int)o : trash;
              }
              double old[D];
#pragma unroll
              for(int i = 0; i < D; ++i) old[i] = s_buf[idx[i]];
#pragma unroll
              for(int i = 0; i < D; ++i) s_buf[idx[i]] = old[i] + v[i];
            }
          }
        }
      }
    }
#undef B200_OFF
    if(L > 1 && MAT) __syncwarp(); // orders the row-image updates of consecutive elements
  }
  if(RES) {
    if(live && l == 0) {
#pragma unroll
      for(int i = 0; i < NR; ++i)
        if(row[i] < a.nInc) a.rhs[row[i]] = res[i];
    }
  }
  if(MAT) {
    __syncthreads();
    const int64_t g0 = a.cta_g0[cta];
    if(g0 >= 0) {
      // the CTA's rows are consecutive in the CSR arrays: the row images are a contiguous image of val[g0 ...]
      double *dst = a.val + g0;
      for(int i = tid; i < (int)tot; i += NT) dst[i] = s_buf[i];
    } else {
      // general numbering: one warp per row segment
      for(int t = wid; t < NPB; t += NW) {
        const int ll = s_len[t];
        uint32_t  o  = s_base[t];
#pragma unroll
        for(int cc = 0; cc < NR; ++cc) {
          const int32_t r = s_row[t * NR + cc];
          if(r < a.nInc) {
            double *dst = a.val + a.ia[r];
            for(int k = lane; k < ll; k += 32) dst[k] = s_buf[o + k];
            o += ll;
          }
        }
      }
    }
  }
}

} // namespace b200
