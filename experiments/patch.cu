// Patch assembly: block-slot owners over spatially compact element patches.
//
// Same numbers as the row-owner gather (gather.cu) and therefore as the reference's per-form quadrature loops
// (feSysElm_*::computeAe/computeBe, src/feVectorSysElm.cpp:1171-1242, :1454-1532, :685-749, :449-503, :528-578, :390-425,
// :112-128, restated in oracle/fe_oracle.py) and its colour-ordered scatter (src/feLinearSystemMklPardiso.cpp:501-749), on
// the same pre-contracted reference tensors.  What changes is who computes what and where the operands live; the r01d
// profile (profiles/README.md) showed the gather kernels limited by L1 wavefronts of per-thread scattered loads:
//
//  * Elements are ordered along a Morton curve of their centroids and cut into patches of PE consecutive elements.  A
//    node (velocity node = D rows, pressure node = 1 row) belongs to the lowest patch among its adjacent elements; a
//    patch's CTA works on its own elements plus the halo elements adjacent to its nodes.
//  * Phase 1 (cooperative, coalesced): the CTA stages the local solution and the inverse affine map of its elements in
//    shared memory and computes everything that depends on the ELEMENT only -- velocity gradients at the vertices, the
//    convective block C1[a][b] and the complete element residual -- into one shared-memory record per element.
//  * Phase 2: one thread per BLOCK SLOT of the CSR matrix ((row node, column node): D x D, D x 1 or 1 x D entries).  The
//    thread sums the contributions (element, a, b) of its slot in registers, reading element records and the per-(a,b)
//    reference tensors from shared memory, and stores the finished entries: consecutive threads own consecutive slots of a
//    row, so the stores are coalesced; every CSR value is written exactly once -- no memset, no atomics, no shared-memory
//    row images, summation in ascending element order (deterministic).
//  * Phase 3: one thread per node sums the element residuals of its adjacent elements and stores the rhs rows.
#include <thrust/binary_search.h>
#include <thrust/copy.h>
#include <thrust/count.h>
#include <thrust/device_vector.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>

#include "device_common.cuh"
#include "system.h"

namespace b200 {

struct PatchDesc {
  int32_t e0, ne; // elements (own + halo) in PatchPlan::elems
  int32_t n0, nn; // owned nodes
  int32_t p0;     // first (node, element) pair; velocity nodes (and their pairs) come first
  int32_t np, npu; // pairs of the patch, of which velocity pairs
  int32_t b0, nb; // block slots, sorted by (kind, number of contributions descending)
  int32_t nuu;    // the first nuu slots are U-U blocks, the others U-P / P-U
  int32_t c0;     // first entry of the patch in the contribution list (slots with more than one contribution)
};

template <int D> struct NodeRec {
  int64_t  vbase[D]; // ia[row] of each unknown row, -1 = essential row (not assembled)
  int32_t  rows[D];
  uint16_t pair0;    // patch-relative
  uint8_t  npairs, kind; // kind 0 = velocity node, 1 = pressure node (row 0 only)
};

struct BlockRec {
  uint16_t node;   // patch-relative owner node
  uint16_t off0;   // row-local offset of the first present column
  uint16_t cstart; // cnt == 1: the contribution word itself; else patch-relative first entry in the contribution list
  uint8_t  cnt;    // number of contributions (adjacent elements holding both nodes)
  uint8_t  flags;  // bits 0-2: present column components, bits 3-4: 0 = U-U, 1 = U-P, 2 = P-U
};
static_assert(sizeof(BlockRec) == 8, "block records are loaded as one 8-byte word");

// shared-memory record of one element of the patch (own or halo), in doubles; W/2 is odd so that 16-byte loads of
// different elements spread over the banks.  What depends on a (row node, element) pair -- the row of the convective block
// C1 and the element residual of that node -- lives in per-pair records, which exist for the pairs of OWNED nodes only.
template <int D, int NS, int NP> struct PS {
  static constexpr int NU   = NS * D;
  static constexpr int O_G  = 0;                                    // G[al*D+m] = d xi_al / d x_m
  static constexpr int O_J  = (D * D + 1) / 2 * 2;                  // detJ, c_conv * detJ
  static constexpr int O_GU = O_J + 2;                              // gu[(v*D+j)*D+i] = d_j u_i at vertex v
  static constexpr int O_U  = O_GU + (NP * D * D + 1) / 2 * 2;      // local velocity DOFs, then pressure DOFs
  static constexpr int O_P  = O_U + NU;
  static constexpr int W0   = (O_P + NP + 1) / 2 * 2;
  static constexpr int W    = ((W0 / 2) % 2 == 1) ? W0 : W0 + 2;
  static constexpr int PRW  = NS + D;                               // velocity pair record: C1[a][0..NS), R[0..D)
};

// shared-memory tables: per-(a,b) records {Kref[a][b][al][be], T3[a][b][v], Mref[a][b]} with an odd 16-byte stride, then
// E[c][al][v], Bref[q][a][al], W[k][a]
template <int D, int NS, int NP> struct PT {
  static constexpr int ABW  = D * D + NP + 1;
  static constexpr int ABS  = ((ABW + 1) / 2 % 2 == 1) ? (ABW + 1) / 2 * 2 : (ABW + 1) / 2 * 2 + 2;
  static constexpr int O_AB = 0;
  static constexpr int O_E  = NS * NS * ABS;
  static constexpr int O_B  = O_E + NS * D * NP;
  static constexpr int O_W  = O_B + NP * NS * D;
};

struct PatchPlan {
  int32_t    nPatch = 0, nNodes = 0;
  int        PE = 0, maxE = 0, NT = 256, max_smem_doubles = 0, max_nb = 0, max_np = 0, max_nn = 0;
  bool       slice = false;        // B200_PATCH_KERNEL=slice: row-slice kernel (slice.cuh, 2-D only)
  const double *d_gt = nullptr;    // GT-layout reference tensors (owned by the gather plan)
  int64_t    nBlocks = 0, nCtr = 0, nPairs = 0, nElemsTot = 0;
  PatchDesc *desc = nullptr;
  int32_t   *elems = nullptr;
  void      *nodes = nullptr;
  uint16_t  *pairs = nullptr;
  BlockRec  *blocks = nullptr;
  uint16_t  *ctr = nullptr;
  double    *d_tab = nullptr;
  int        tab_len = 0, tab_len_src = 0;
  const double *d_geo = nullptr; // owned by the gather plan
  // CSR entries of unknown rows that no element couples (e.g. the diagonal the pattern keeps in pressure rows): they
  // are written as explicit zeros by every matrix pass, like the zero-initialised row images of the row-owner kernels
  int64_t   *zero_idx = nullptr;
  int64_t    nZero = 0;
};

struct PatchArgs {
  const PatchDesc *desc;
  const int32_t   *elems;
  const void      *nodes;
  const uint16_t  *pairs;
  const BlockRec  *blocks;
  const uint16_t  *ctr;
  const int32_t   *adrU, *adrP;
  const double    *sol, *soldot, *source, *geo, *tab;
  double          *val, *rhs;
  int64_t          nInc;
  int              nq, ntab;
  THCoeffs         c;
  double           c0;
};

template <int D, int NS, int NP, int NT, int MINB, bool MAT, bool RES>
__global__ void __launch_bounds__(NT, MINB) patch_kernel(const PatchArgs a)
{
  using S_ = PS<D, NS, NP>;
  using T_ = PT<D, NS, NP>;
  constexpr int NU = NS * D, NL = NU + NP, GW = GT<D, NS, NP>::GW, SW = S_::W, PRW = S_::PRW, ABS = T_::ABS;
  extern __shared__ double sm[];
  const int       tid = threadIdx.x;
  const PatchDesc pd  = a.desc[blockIdx.x];
  const int       nE  = pd.ne;
  double         *s_tab   = sm;
  double         *s_el    = s_tab + ((a.ntab + 1) & ~1);          // [nE][SW]
  double         *s_pr    = s_el + nE * SW;                       // [npu][PRW]
  double         *s_rp    = s_pr + pd.npu * PRW;                  // [np - npu]
  uint16_t       *s_pairs = reinterpret_cast<uint16_t *>(s_rp + (pd.np - pd.npu)); // [np] element << 4 | local node
  uint16_t       *s_np0   = s_pairs + ((pd.np + 3) & ~3);         // [nn] first pair of every owned node
  const int32_t  *el_list = a.elems + pd.e0;
  const THCoeffs  c       = a.c;
  const NodeRec<D> *nodes = static_cast<const NodeRec<D> *>(a.nodes) + pd.n0;

  // ---- phase 1a: tables, pair list, local solution and geometry of every element (coalesced over the DOF tables) -----
  for(int i = tid; i < a.ntab; i += NT) s_tab[i] = a.tab[i];
  for(int i = tid; i < pd.np; i += NT) s_pairs[i] = a.pairs[pd.p0 + i];
  for(int i = tid; i < pd.nn; i += NT) s_np0[i] = nodes[i].pair0;
  for(int idx = tid; idx < nE * NL; idx += NT) {
    const int     el  = idx / NL, k = idx - el * NL;
    const int64_t e   = el_list[el];
    const int32_t dof = k < NU ? a.adrU[e * NU + k] : a.adrP[e * NP + (k - NU)];
    s_el[el * SW + S_::O_U + k] = a.sol[dof];
  }
  for(int idx = tid; idx < nE * (D * D + 1); idx += NT) {
    const int     el = idx / (D * D + 1), k = idx - el * (D * D + 1);
    const int64_t e  = el_list[el];
    const double  v  = a.geo[e * GW + k];
    if(k < D * D)
      s_el[el * SW + S_::O_G + k] = v;
    else {
      s_el[el * SW + S_::O_J]     = v;
      s_el[el * SW + S_::O_J + 1] = c.c_conv * v;
    }
  }
  __syncthreads();

  // ---- phase 1b: velocity gradient at the vertices, gu[v][j][i] = d_j u_i (v)   (src/feSpace.cpp:1352-1405) ----------
  for(int it = tid; it < nE * D * NP; it += NT) {
    const int el = it / (D * NP), r = it - el * (D * NP), i = r / NP, v = r - i * NP;
    double   *st = s_el + el * SW;
    double    X[D];
#pragma unroll
    for(int al = 0; al < D; ++al) X[al] = 0.;
#pragma unroll
    for(int cc = 0; cc < NS; ++cc) {
      const double u = st[S_::O_U + cc * D + i];
#pragma unroll
      for(int al = 0; al < D; ++al) X[al] += u * s_tab[T_::O_E + (cc * D + al) * NP + v];
    }
#pragma unroll
    for(int j = 0; j < D; ++j) {
      double s = 0.;
#pragma unroll
      for(int al = 0; al < D; ++al) s += st[S_::O_G + al * D + j] * X[al];
      st[S_::O_GU + (v * D + j) * D + i] = s;
    }
  }
  __syncthreads();

  // ---- phase 1c: one thread per (owned node, adjacent element) pair: row aa of the convective block C1 and the
  // element residual of that node --------------------------------------------------------------------------------------
  {
    const bool   domass = (c.c_mass != 0.) && (a.soldot != nullptr);
    const double cvis1 = c.sig_mu - c.diff_k, cvis2 = c.sig_mu, cpre = c.c_gradp - c.c_sig;
    for(int it = tid; it < pd.npu; it += NT) {
      const uint32_t pw = s_pairs[it];
      const int      el = pw >> 4, aa = pw & 15;
      const double  *st = s_el + el * SW;
      double        *pr = s_pr + it * PRW;
      double         G[D * D];
#pragma unroll
      for(int i = 0; i < D * D; ++i) G[i] = st[S_::O_G + i];
      const double J = st[S_::O_J], cJ = st[S_::O_J + 1];
      // Z[al][v] = sum_c Ut[c][al] T3[aa][c][v],  Ut[c][al] = c_conv J sum_m U[c][m] G[al][m]
      double Z[D][NP];
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
          for(int m = 0; m < D; ++m) s += st[S_::O_U + cc * D + m] * G[al * D + m];
          ut[al] = cJ * s;
        }
        const double *t3 = s_tab + T_::O_AB + (aa * NS + cc) * ABS + D * D;
#pragma unroll
        for(int v = 0; v < NP; ++v) {
          const double t = t3[v];
#pragma unroll
          for(int al = 0; al < D; ++al) Z[al][v] += ut[al] * t;
        }
      }
      double r[D];
#pragma unroll
      for(int i = 0; i < D; ++i) r[i] = 0.;
#pragma unroll
      for(int b = 0; b < NS; ++b) {
        double s = 0.;
#pragma unroll
        for(int al = 0; al < D; ++al)
#pragma unroll
          for(int v = 0; v < NP; ++v) s += s_tab[T_::O_E + (b * D + al) * NP + v] * Z[al][v];
        if(MAT) pr[b] = s;
        if(RES) {
#pragma unroll
          for(int i = 0; i < D; ++i) r[i] -= s * st[S_::O_U + b * D + i];
        }
      }
      if(RES) {
        // viscous and pressure parts through Bp[v][aa][m] = int psi_v d_m phi_aa (gradients of the velocity basis lie
        // in the span of the pressure basis, checked at plan time)
#pragma unroll
        for(int v = 0; v < NP; ++v) {
          const double *Br = s_tab + T_::O_B + (v * NS + aa) * D;
          const double  pv = st[S_::O_P + v];
#pragma unroll
          for(int m = 0; m < D; ++m) {
            double s = 0.;
#pragma unroll
            for(int al = 0; al < D; ++al) s += G[al * D + m] * Br[al];
            const double bp = J * s;
            r[m] += cpre * bp * pv;
#pragma unroll
            for(int i = 0; i < D; ++i)
              r[i] += bp * (cvis1 * st[S_::O_GU + (v * D + m) * D + i] + cvis2 * st[S_::O_GU + (v * D + i) * D + m]);
          }
        }
        if(domass || a.source != nullptr) {
          const int64_t e = el_list[el];
          if(domass) {
            for(int b = 0; b < NS; ++b) {
              const double mab = c.c_mass * J * s_tab[T_::O_AB + (aa * NS + b) * ABS + D * D + NP];
#pragma unroll
              for(int i = 0; i < D; ++i) r[i] -= mab * a.soldot[a.adrU[e * NU + b * D + i]];
            }
          }
          if(a.source != nullptr) {
            const double *src = a.source + e * a.nq * D;
            for(int k = 0; k < a.nq; ++k) {
              const double wj = J * s_tab[T_::O_W + k * NS + aa];
#pragma unroll
              for(int i = 0; i < D; ++i) r[i] -= wj * src[k * D + i];
            }
          }
        }
#pragma unroll
        for(int i = 0; i < D; ++i) pr[NS + i] = r[i];
      }
    }
    if(RES) {
      // pressure row q: -c_div int psi_q div u   (feSysElm_MixedDivergence, src/feVectorSysElm.cpp:685-749)
      for(int it = pd.npu + tid; it < pd.np; it += NT) {
        const uint32_t pw = s_pairs[it];
        const int      el = pw >> 4, q = (int)(pw & 15) - NS;
        const double  *st = s_el + el * SW;
        double         G[D * D];
#pragma unroll
        for(int i = 0; i < D * D; ++i) G[i] = st[S_::O_G + i];
        const double J  = st[S_::O_J];
        double       rp = 0.;
#pragma unroll
        for(int b = 0; b < NS; ++b) {
          const double *Br = s_tab + T_::O_B + (q * NS + b) * D;
#pragma unroll
          for(int j = 0; j < D; ++j) {
            double s = 0.;
#pragma unroll
            for(int al = 0; al < D; ++al) s += G[al * D + j] * Br[al];
            rp -= c.c_div * J * s * st[S_::O_U + b * D + j];
          }
        }
        s_rp[it - pd.npu] = rp;
      }
    }
  }
  __syncthreads();

  // ---- phase 2: one thread per block slot; the slots of a patch are sorted by (kind, number of contributions), so
  // the lanes of a warp run the same code path and the same trip count --------------------------------------------------
  if(MAT) {
    const uint2    *blk   = reinterpret_cast<const uint2 *>(a.blocks + pd.b0);
    const uint16_t *ctr   = a.ctr + pd.c0;
    const double    mass0 = c.c_mass * a.c0;
    const double    cvd = c.diff_k - c.sig_mu, cup = c.c_sig - c.c_gradp;
    // U-U slots:  A[i][j] = delta_ij (C1 + (k - mu) tr K + c0 m M) - mu K[j][i] + c_conv J int phi_a phi_b d_j u_i
    // (feSysElm_VectorConvectiveAcceleration, _DivergenceNewtonianStress, _VectorDiffusion, _TransientVectorMass)
    for(int k = tid; k < pd.nuu; k += NT) {
      const uint2 bw     = blk[k];
      const int   bnode  = bw.x & 0xffffu, off0 = bw.x >> 16;
      const int   cstart = bw.y & 0xffffu, cnt = (bw.y >> 16) & 0xffu, dmask = (bw.y >> 24) & 7;
      const int   pair0  = s_np0[bnode];
      double      A[D][D];
#pragma unroll
      for(int i = 0; i < D; ++i)
#pragma unroll
        for(int j = 0; j < D; ++j) A[i][j] = 0.;
      uint32_t cw = cnt == 1 ? (uint32_t)cstart : (uint32_t)ctr[cstart];
      for(int t = 0; t < cnt; ++t) {
        const uint32_t nx   = (t + 1 < cnt) ? (uint32_t)ctr[cstart + t + 1] : 0u;
        const int      pair = pair0 + (int)(cw >> 4), lb = cw & 15;
        const uint32_t pw   = s_pairs[pair];
        const int      el = pw >> 4, la = pw & 15;
        const double  *st = s_el + el * SW;
        double         G[D * D];
        if(D == 2) {
          const double2 g0 = *reinterpret_cast<const double2 *>(st + S_::O_G), g1 = *reinterpret_cast<const double2 *>(st + S_::O_G + 2);
          G[0] = g0.x, G[1] = g0.y, G[2] = g1.x, G[3] = g1.y;
        } else {
#pragma unroll
          for(int i = 0; i < D * D; ++i) G[i] = st[S_::O_G + i];
        }
        const double2 jj = *reinterpret_cast<const double2 *>(st + S_::O_J);
        const double  J = jj.x, cJ = jj.y;
        const double *ab = s_tab + T_::O_AB + (la * NS + lb) * ABS;
        double        K[D][D];
        {
          double H[D][D];
#pragma unroll
          for(int al = 0; al < D; ++al)
#pragma unroll
            for(int nn = 0; nn < D; ++nn) {
              double s = 0.;
#pragma unroll
              for(int be = 0; be < D; ++be) s += ab[al * D + be] * G[be * D + nn];
              H[al][nn] = J * s;
            }
#pragma unroll
          for(int m = 0; m < D; ++m)
#pragma unroll
            for(int nn = 0; nn < D; ++nn) {
              double s = 0.;
#pragma unroll
              for(int al = 0; al < D; ++al) s += G[al * D + m] * H[al][nn];
              K[m][nn] = s;
            }
        }
        double trK = 0.;
#pragma unroll
        for(int m = 0; m < D; ++m) trK += K[m][m];
        double s = s_pr[pair * PRW + lb] + cvd * trK;
        if(mass0 != 0.) s += mass0 * J * ab[D * D + NP];
        double t3[NP];
#pragma unroll
        for(int v = 0; v < NP; ++v) t3[v] = cJ * ab[D * D + v];
#pragma unroll
        for(int i = 0; i < D; ++i)
#pragma unroll
          for(int j = 0; j < D; ++j) {
            double C2 = (i == j ? s : 0.) - c.sig_mu * K[j][i];
#pragma unroll
            for(int v = 0; v < NP; ++v) C2 += st[S_::O_GU + (v * D + j) * D + i] * t3[v];
            A[i][j] += C2;
          }
        cw = nx;
      }
      const NodeRec<D> &nd = nodes[bnode];
#pragma unroll
      for(int i = 0; i < D; ++i) {
        const int64_t vb = nd.vbase[i];
        if(vb >= 0) {
          double *dst = a.val + vb + off0;
          if(D == 2 && dmask == 3 && ((vb + off0) & 1) == 0) {
            *reinterpret_cast<double2 *>(dst) = make_double2(A[i][0], A[i][1]);
          } else {
            int n = 0;
#pragma unroll
            for(int j = 0; j < D; ++j)
              if((dmask >> j) & 1) dst[n++] = A[i][j];
          }
        }
      }
    }
    // U-P slots: rows of velocity node la, column of pressure node lb - NS (feSysElm_DivergenceNewtonianStress p-block,
    // feSysElm_MixedGradient);  P-U slots: row of pressure node la - NS, columns of velocity node lb (feSysElm_MixedDivergence)
    for(int k = pd.nuu + tid; k < pd.nb; k += NT) {
      const uint2  bw     = blk[k];
      const int    bnode  = bw.x & 0xffffu, off0 = bw.x >> 16;
      const int    cstart = bw.y & 0xffffu, cnt = (bw.y >> 16) & 0xffu, flags = bw.y >> 24;
      const bool   up     = (flags >> 3) == 1;
      const int    dmask  = flags & 7;
      const int    pair0  = s_np0[bnode];
      const double cf     = up ? cup : c.c_div;
      double       A[D];
#pragma unroll
      for(int i = 0; i < D; ++i) A[i] = 0.;
      uint32_t cw = cnt == 1 ? (uint32_t)cstart : (uint32_t)ctr[cstart];
      for(int t = 0; t < cnt; ++t) {
        const uint32_t nx = (t + 1 < cnt) ? (uint32_t)ctr[cstart + t + 1] : 0u;
        const int      lb = cw & 15;
        const uint32_t pw = s_pairs[pair0 + (int)(cw >> 4)];
        const int      el = pw >> 4, la = pw & 15;
        const double  *st = s_el + el * SW;
        const double  *Br = s_tab + T_::O_B + (up ? ((lb - NS) * NS + la) : ((la - NS) * NS + lb)) * D;
        const double   J  = cf * st[S_::O_J];
#pragma unroll
        for(int i = 0; i < D; ++i) {
          double sg = 0.;
#pragma unroll
          for(int al = 0; al < D; ++al) sg += st[S_::O_G + al * D + i] * Br[al];
          A[i] += J * sg;
        }
        cw = nx;
      }
      const NodeRec<D> &nd = nodes[bnode];
      if(up) {
#pragma unroll
        for(int i = 0; i < D; ++i) {
          const int64_t vb = nd.vbase[i];
          if(vb >= 0) a.val[vb + off0] = A[i];
        }
      } else {
        double *dst = a.val + nd.vbase[0] + off0;
        int     n   = 0;
#pragma unroll
        for(int j = 0; j < D; ++j)
          if((dmask >> j) & 1) dst[n++] = A[j];
      }
    }
  }

  // ---- phase 3: rhs rows, one thread per node: sum of the element residuals of its pairs -----------------------------
  if(RES) {
    for(int n = tid; n < pd.nn; n += NT) {
      const NodeRec<D> &nd = nodes[n];
      double            res[D];
#pragma unroll
      for(int i = 0; i < D; ++i) res[i] = 0.;
      if(nd.kind == 0) {
        for(int t = 0; t < nd.npairs; ++t) {
#pragma unroll
          for(int i = 0; i < D; ++i) res[i] += s_pr[(nd.pair0 + t) * PRW + NS + i];
        }
#pragma unroll
        for(int i = 0; i < D; ++i)
          if(nd.rows[i] < a.nInc) a.rhs[nd.rows[i]] = res[i];
      } else {
        for(int t = 0; t < nd.npairs; ++t) res[0] += s_rp[nd.pair0 + t - pd.npu];
        if(nd.rows[0] < a.nInc) a.rhs[nd.rows[0]] = res[0];
      }
    }
  }
}

} // namespace b200
#include "slice.cuh"
namespace b200 {

__global__ void patch_zero_kernel(int64_t n, const int64_t *idx, double *val)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) val[idx[i]] = 0.;
}

// ----------------------------------------------------------------------------------------------------------
// plan construction (set-up; device sorts, a few small host decisions)
// ----------------------------------------------------------------------------------------------------------
#define GRID_STRIDE(i, n) for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

__global__ void morton_kernel(int64_t nElm, int D, const double *xyz, const int32_t *conn, double lo0, double lo1, double lo2, double inv0,
                              double inv1, double inv2, uint64_t *key, int32_t *eidx)
{
  GRID_STRIDE(e, nElm)
  {
    double cx[3] = {0., 0., 0.};
    for(int v = 0; v <= D; ++v) {
      const int32_t p = conn[e * (D + 1) + v];
      for(int m = 0; m < D; ++m) cx[m] += xyz[(int64_t)p * D + m];
    }
    const double lo[3] = {lo0, lo1, lo2}, inv[3] = {inv0, inv1, inv2};
    uint32_t     q[3]  = {0, 0, 0};
    for(int m = 0; m < D; ++m) {
      double t = (cx[m] / (D + 1) - lo[m]) * inv[m];
      t        = t < 0. ? 0. : (t > 1. ? 1. : t);
      q[m]     = (uint32_t)(t * 1048575.);
    }
    uint64_t k = 0;
    for(int bit = 19; bit >= 0; --bit)
      for(int m = D - 1; m >= 0; --m) k = (k << 1) | ((q[m] >> bit) & 1u);
    key[e]  = k;
    eidx[e] = (int32_t)e;
  }
}

__global__ void patch_of_kernel(int64_t nElm, int PE, const int32_t *eidx, int32_t *patch_of)
{
  GRID_STRIDE(r, nElm) patch_of[eidx[r]] = (int32_t)(r / PE);
}

__device__ __forceinline__ int32_t node_dof0(const int32_t *adrU, const int32_t *adrP, int64_t e, int l, int D, int NS, int NP)
{
  return l < NS ? adrU[e * NS * D + l * D] : adrP[e * NP + (l - NS)];
}

__global__ void owner_kernel(int64_t nElm, int D, int NS, int NP, const int32_t *adrU, const int32_t *adrP, const int32_t *patch_of, int32_t *owner)
{
  const int NL = NS + NP;
  GRID_STRIDE(idx, nElm * NL)
  {
    const int64_t e = idx / NL;
    const int     l = (int)(idx - e * NL);
    atomicMin(owner + node_dof0(adrU, adrP, e, l, D, NS, NP), patch_of[e]);
  }
}

__global__ void pair_key_kernel(int64_t nElm, int D, int NS, int NP, const int32_t *adrU, const int32_t *adrP, const int32_t *owner, int64_t nInc,
                                uint64_t *key, int32_t *payload)
{
  const int NL = NS + NP;
  GRID_STRIDE(idx, nElm * NL)
  {
    const int64_t e = idx / NL;
    const int     l = (int)(idx - e * NL);
    bool          unknown = false;
    if(l < NS) {
      for(int c = 0; c < D; ++c) unknown |= adrU[e * NS * D + l * D + c] < nInc;
    } else
      unknown = adrP[e * NP + (l - NS)] < nInc;
    const int32_t d0 = node_dof0(adrU, adrP, e, l, D, NS, NP);
    // patch, then velocity nodes before pressure nodes, then the DOF of component 0 (unique per node)
    key[idx]     = unknown ? (((uint64_t)(uint32_t)owner[d0] << 32) | ((uint64_t)(l < NS ? 0u : 1u) << 31) | (uint32_t)d0) : ~0ull;
    payload[idx] = (int32_t)(e * 16 + l);
  }
}

__global__ void head_flag_kernel(int64_t n, const uint64_t *key, int32_t *flag)
{
  GRID_STRIDE(i, n) flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

__global__ void elem_key_kernel(int64_t n, const uint64_t *key, const int32_t *payload, uint64_t *ekey)
{
  GRID_STRIDE(i, n) ekey[i] = (key[i] & 0xffffffff00000000ull) | (uint32_t)(payload[i] >> 4);
}

__global__ void low32_kernel(int64_t n, const uint64_t *key, int32_t *out)
{
  GRID_STRIDE(i, n) out[i] = (int32_t)(key[i] & 0xffffffffu);
}

__device__ __forceinline__ int local_element(const int32_t *elems, int32_t lo, int32_t hi, int32_t e)
{
  const int32_t b = lo;
  --hi;
  while(lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if(elems[mid] < e)
      lo = mid + 1;
    else
      hi = mid;
  }
  return elems[lo] == e ? lo - b : -1;
}

// one thread per sorted (node, element) pair: packed pair word; the head of every node also writes the node record
template <int D>
__global__ void node_record_kernel(int64_t n, int NS, int NP, const uint64_t *key, const int32_t *payload, const int32_t *flag,
                                   const int32_t *nodeidx, const int32_t *pair_ptr, const int32_t *elem_ptr, const int32_t *elems,
                                   const int32_t *adrU, const int32_t *adrP, const int64_t *ia, int64_t nInc, uint16_t *pairs, uint16_t *pi,
                                   NodeRec<D> *nodes, int32_t *node_patch, int *err)
{
  GRID_STRIDE(i, n)
  {
    const int32_t p  = (int32_t)(key[i] >> 32);
    const int64_t e  = payload[i] >> 4;
    const int     l  = payload[i] & 15;
    const int     el = local_element(elems, elem_ptr[p], elem_ptr[p + 1], (int32_t)e);
    if(el < 0 || el > 4095) {
      atomicExch(err, 10);
      continue;
    }
    pairs[i] = (uint16_t)((el << 4) | l);
    if(flag[i]) {
      int64_t j = i + 1;
      pi[i]     = 0;
      while(j < n && key[j] == key[i]) {
        pi[j] = (uint16_t)(j - i); // position of the pair among the pairs of its node
        ++j;
      }
      NodeRec<D> nd;
      const int64_t rel = i - pair_ptr[p];
      if(j - i > 255 || rel > 65535) atomicExch(err, 11);
      nd.pair0  = (uint16_t)rel;
      nd.npairs = (uint8_t)(j - i);
      nd.kind   = l < NS ? 0 : 1;
      for(int c = 0; c < D; ++c) {
        int32_t r = 0x7fffffff;
        if(l < NS)
          r = adrU[e * NS * D + l * D + c];
        else if(c == 0)
          r = adrP[e * NP + (l - NS)];
        nd.rows[c]  = r;
        nd.vbase[c] = r < nInc ? ia[r] : -1;
      }
      const int32_t ni = nodeidx[i] - 1;
      nodes[ni]        = nd;
      node_patch[ni]   = p;
    }
  }
}

// candidates: one per (pair, local column node); key = (node index, DOF of component 0 of the column node)
__global__ void cand_kernel(int64_t nPairs, int D, int NS, int NP, const int32_t *payload, const uint16_t *pairs, const uint16_t *pi, const int32_t *nodeidx,
                            const int32_t *adrU, const int32_t *adrP, int64_t nInc, int blockmask, uint64_t *ckey, int32_t *cpay)
{
  const int NL = NS + NP;
  GRID_STRIDE(idx, nPairs * NL)
  {
    const int64_t i  = idx / NL;
    const int     lb = (int)(idx - i * NL);
    const int64_t e  = payload[i] >> 4;
    const int     la = payload[i] & 15;
    const int     rk = la < NS ? 0 : 1, ck = lb < NS ? 0 : 1;
    bool          ok = (blockmask >> (rk * 2 + ck)) & 1;
    if(ok) {
      bool any = false;
      if(lb < NS) {
        for(int d = 0; d < D; ++d) any |= adrU[e * NS * D + lb * D + d] < nInc;
      } else
        any = adrP[e * NP + (lb - NS)] < nInc;
      ok = any;
    }
    if(ok) {
      const int32_t d0 = node_dof0(adrU, adrP, e, lb, D, NS, NP);
      ckey[idx] = ((uint64_t)(uint32_t)(nodeidx[i] - 1) << 32) | (uint32_t)d0;
      // element (patch-local) in the upper half; the lower half is the contribution word: pair-within-node << 4 | column node
      cpay[idx] = (int32_t)(((uint32_t)(pairs[i] >> 4) << 16) | ((uint32_t)pi[i] << 4) | (uint32_t)lb);
    } else {
      ckey[idx] = ~0ull;
      cpay[idx] = 0;
    }
  }
}

template <int D>
__global__ void block_record_kernel(int64_t nCand, int NS, int NP, const uint64_t *ckey, const int32_t *cpay, const int32_t *flag, const int32_t *blkidx,
                                    const int32_t *node_patch, const int32_t *node_ptr, const int32_t *cand_ptr, const int32_t *elem_ptr,
                                    const int32_t *elems, const NodeRec<D> *nodes, const int32_t *adrU, const int32_t *adrP, const int64_t *ia,
                                    const int32_t *ja, int64_t nInc, BlockRec *blocks, uint64_t *bclass, uint8_t *covered, int *err)
{
  GRID_STRIDE(i, nCand)
  {
    if(!flag[i]) continue;
    int64_t j = i + 1;
    while(j < nCand && ckey[j] == ckey[i]) ++j;
    const int32_t n = (int32_t)(ckey[i] >> 32);
    const int32_t p = node_patch[n];
    const int64_t rel = i - cand_ptr[p];
    if(j - i > 255 || rel > 65535 || n - node_ptr[p] > 65535) {
      atomicExch(err, 12);
      continue;
    }
    const int     el = ((uint32_t)cpay[i]) >> 16, lb = cpay[i] & 15;
    const int64_t e  = elems[elem_ptr[p] + el];
    const NodeRec<D> &nd = nodes[n];
    if(((cpay[i] >> 4) & 0xfff) + (j - i) > 4095) atomicExch(err, 18);
    int32_t r0 = -1;
    for(int c = D - 1; c >= 0; --c)
      if(nd.vbase[c] >= 0) r0 = nd.rows[c];
    const int64_t beg = ia[r0], end = ia[r0 + 1];
    if(end - beg >= 65535) {
      atomicExch(err, 13);
      continue;
    }
    int     dmask = 0, ncols = 0;
    int64_t off0 = -1;
    const int nd_ = lb < NS ? D : 1;
    for(int d = 0; d < nd_; ++d) {
      const int32_t col = lb < NS ? adrU[e * NS * D + lb * D + d] : adrP[e * NP + (lb - NS)];
      if(col >= nInc) continue;
      int64_t lo = beg, hi = end - 1;
      while(lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if(ja[mid] < col)
          lo = mid + 1;
        else
          hi = mid;
      }
      if(lo >= end || ja[lo] != col) {
        atomicExch(err, 14); // a local (row, column) pair is missing from the pattern
        continue;
      }
      const int64_t o = lo - beg;
      if(off0 < 0)
        off0 = o;
      else if(o != off0 + ncols)
        atomicExch(err, 15); // components of one column node are not consecutive in the row
      // every unknown row of the node must hold this column at the same offset
      for(int c = 0; c < D; ++c)
        if(nd.vbase[c] >= 0) {
          if(nd.rows[c] != r0 && ja[nd.vbase[c] + o] != col) atomicExch(err, 16);
          covered[nd.vbase[c] + o] = 1;
        }
      dmask |= 1 << d;
      ++ncols;
    }
    BlockRec br;
    br.node   = (uint16_t)(n - node_ptr[p]);
    br.off0   = (uint16_t)(off0 < 0 ? 0 : off0);
    br.cstart = (uint16_t)rel;
    br.cnt    = (uint8_t)(j - i);
    const int kind = nd.kind == 0 ? (lb < NS ? 0 : 1) : 2;
    br.flags  = (uint8_t)(dmask | (kind << 3));
    blocks[blkidx[i] - 1] = br;
    // class key: patch, then U-U before the D x 1 / 1 x D kinds, then decreasing number of contributions
    bclass[blkidx[i] - 1] = ((uint64_t)(uint32_t)p << 16) | (uint64_t)(kind << 8) | (uint64_t)(255 - (j - i));
  }
}

// number of entries a slot keeps in the contribution list (a single contribution lives in the record itself)
__global__ void block_list_len_kernel(int64_t nBlocks, const BlockRec *blocks, int32_t *len)
{
  GRID_STRIDE(b, nBlocks) len[b] = blocks[b].cnt >= 2 ? blocks[b].cnt : 0;
}

// after the class sort: move the contributions of every slot to their final place and rewrite cstart
__global__ void block_relayout_kernel(int64_t nBlocks, BlockRec *blocks, const uint64_t *bclass, const int32_t *pos, const int32_t *blk_ptr,
                                      const int32_t *cand_ptr, const int32_t *cpay, uint16_t *ctr, int *err)
{
  GRID_STRIDE(b, nBlocks)
  {
    BlockRec      br  = blocks[b];
    const int32_t p   = (int32_t)(bclass[b] >> 16);
    const int64_t src = (int64_t)cand_ptr[p] + br.cstart;
    if(br.cnt == 1) {
      br.cstart = (uint16_t)(cpay[src] & 0xffff);
    } else {
      const int64_t rel = (int64_t)pos[b] - pos[blk_ptr[p]];
      if(rel + br.cnt > 65535) atomicExch(err, 17);
      for(int t = 0; t < br.cnt; ++t) ctr[pos[b] + t] = (uint16_t)(cpay[src + t] & 0xffff);
      br.cstart = (uint16_t)rel;
    }
    blocks[b] = br;
  }
}

__global__ void desc_kernel(int32_t nPatch, const int32_t *elem_ptr, const int32_t *node_ptr, const int32_t *pair_ptr, const int32_t *blk_ptr,
                            const int32_t *uu_end, const int32_t *pu_end, const int32_t *pos, int64_t nBlocks, int32_t nList, int SW, int PRW,
                            PatchDesc *desc, int32_t *smem_doubles)
{
  GRID_STRIDE(p, nPatch)
  {
    PatchDesc d;
    d.e0 = elem_ptr[p];
    d.ne = elem_ptr[p + 1] - elem_ptr[p];
    d.n0 = node_ptr[p];
    d.nn = node_ptr[p + 1] - node_ptr[p];
    d.p0  = pair_ptr[p];
    d.np  = pair_ptr[p + 1] - pair_ptr[p];
    d.npu = pu_end[p] - pair_ptr[p];
    d.b0 = blk_ptr[p];
    d.nb  = blk_ptr[p + 1] - blk_ptr[p];
    d.nuu = uu_end[p] - blk_ptr[p];
    d.c0  = blk_ptr[p] < nBlocks ? pos[blk_ptr[p]] : nList;
    desc[p] = d;
    // shared memory behind the tables: element records, pair records, pair words, first pair of every node
    smem_doubles[p] = d.ne * SW + d.npu * PRW + (d.np - d.npu) + (((d.np + 3) & ~3) + d.nn + 3) / 4;
  }
}

struct PatchKeyOp {
  __host__ __device__ uint64_t operator()(int32_t p) const { return (uint64_t)(uint32_t)p << 32; }
};
struct PtrLookupOp {
  const int32_t *idx; // inclusive scan of head flags
  const int32_t *pos; // positions to look up
  int64_t        n;
  int32_t        total;
  __host__ __device__ int32_t operator()(int32_t p) const { return pos[p] < n ? idx[pos[p]] - 1 : total; }
};
struct PEndKeyOp {
  __host__ __device__ uint64_t operator()(int32_t p) const { return ((uint64_t)(uint32_t)p << 32) | (1ull << 31); }
};
struct UUEndKeyOp {
  __host__ __device__ uint64_t operator()(int32_t p) const { return ((uint64_t)(uint32_t)p << 16) | (1ull << 8); }
};
struct IsZeroOp {
  __host__ __device__ bool operator()(uint8_t v) const { return v == 0; }
};
struct DescNe {
  __host__ __device__ int32_t operator()(const PatchDesc &d) const { return d.ne; }
};
struct DescNb {
  __host__ __device__ int32_t operator()(const PatchDesc &d) const { return d.nb; }
};
struct DescNp {
  __host__ __device__ int32_t operator()(const PatchDesc &d) const { return d.np; }
};
struct DescNn {
  __host__ __device__ int32_t operator()(const PatchDesc &d) const { return d.nn; }
};

void patch_free(System *S)
{
  PatchPlan *P = static_cast<PatchPlan *>(S->patch);
  if(!P) return;
  cudaFree(P->desc);
  cudaFree(P->elems);
  cudaFree(P->nodes);
  cudaFree(P->pairs);
  cudaFree(P->blocks);
  cudaFree(P->ctr);
  cudaFree(P->d_tab);
  cudaFree(P->zero_idx);
  delete P;
  S->patch = nullptr;
}

template <int D, int NS, int NP> static void repack_tables(const std::vector<double> &g, int nq, bool with_src, std::vector<double> &t)
{
  using G_ = GT<D, NS, NP>;
  using T_ = PT<D, NS, NP>;
  t.assign((size_t)T_::O_W + (with_src ? (size_t)nq * NS : 0), 0.);
  for(int a = 0; a < NS; ++a)
    for(int b = 0; b < NS; ++b) {
      double *r = t.data() + T_::O_AB + (size_t)(a * NS + b) * T_::ABS;
      for(int k = 0; k < D * D; ++k) r[k] = g[G_::O_K + (size_t)(a * NS + b) * D * D + k];
      for(int v = 0; v < NP; ++v) r[D * D + v] = g[G_::O_T3 + (size_t)(a * NS + b) * NP + v];
      r[D * D + NP] = g[G_::O_M + (size_t)a * NS + b];
    }
  for(int k = 0; k < NS * D * NP; ++k) t[T_::O_E + k] = g[G_::O_E + k];
  for(int k = 0; k < NP * NS * D; ++k) t[T_::O_B + k] = g[G_::O_B + k];
  if(with_src)
    for(int k = 0; k < nq * NS; ++k) t[T_::O_W + k] = g[G_::O_W + k];
}

template <int D, int NS, int NP> static int build_patch_plan_t(System *S, const GatherTables &gt)
{
  auto          pol  = thrust::cuda::par.on(S->stream);
  const int64_t nElm = S->nElm;
  const int     NL   = NS + NP;
  const int32_t *adrU = S->spaces[S->su].d_adr, *adrP = S->spaces[S->sp].d_adr;
  PatchPlan    *P    = new PatchPlan;
  S->patch           = P;
  P->d_geo           = gt.d_geo;
  P->d_gt            = gt.d_tab;
  {
    const char *k = getenv("B200_PATCH_KERNEL");
    P->slice      = k && std::string(k) == "slice";
  }
  {
    const char *k = getenv("B200_PATCH_ELEMS");
    P->PE         = k ? atoi(k) : (D == 2 ? 64 : 32);
    if(P->PE < 1) P->PE = 1;
    const char *t = getenv("B200_PATCH_THREADS");
    P->NT         = t ? atoi(t) : 256;
  }
  if(nElm * (int64_t)NL * NL >= (int64_t)2147483647) {
    set_error("patch plan: more than 2^31 candidate contributions");
    return B200_ERR_UNSUPP;
  }
  // tables
  {
    std::vector<double> g(gt.tab_len_src), t, tsrc;
    B200_CUDA(cudaMemcpy(g.data(), gt.d_tab, g.size() * sizeof(double), cudaMemcpyDeviceToHost));
    repack_tables<D, NS, NP>(g, S->nq, true, tsrc);
    P->tab_len     = PT<D, NS, NP>::O_W;
    P->tab_len_src = (int)tsrc.size();
    B200_CUDA(cudaMalloc(&P->d_tab, tsrc.size() * sizeof(double)));
    B200_CUDA(cudaMemcpy(P->d_tab, tsrc.data(), tsrc.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  // Morton order of the element centroids
  double lo[3] = {0., 0., 0.}, inv[3] = {0., 0., 0.};
  {
    std::vector<double> xyz((size_t)S->nVert * D);
    B200_CUDA(cudaMemcpy(xyz.data(), S->d_xyz, xyz.size() * sizeof(double), cudaMemcpyDeviceToHost));
    double hi[3] = {0., 0., 0.};
    for(int m = 0; m < D; ++m) lo[m] = hi[m] = xyz[m];
    for(int64_t v = 0; v < S->nVert; ++v)
      for(int m = 0; m < D; ++m) {
        lo[m] = std::min(lo[m], xyz[(size_t)v * D + m]);
        hi[m] = std::max(hi[m], xyz[(size_t)v * D + m]);
      }
    double ext = 0.;
    for(int m = 0; m < D; ++m) ext = std::max(ext, hi[m] - lo[m]);
    for(int m = 0; m < D; ++m) inv[m] = ext > 0. ? 1. / ext : 0.; // same scale in every direction: compact patches
  }
  const int32_t nPatch = (int32_t)((nElm + P->PE - 1) / P->PE);
  P->nPatch            = nPatch;
  thrust::device_vector<int32_t> patch_of(nElm);
  {
    thrust::device_vector<uint64_t> mkey(nElm);
    thrust::device_vector<int32_t>  eidx(nElm);
    morton_kernel<<<148 * 8, 256, 0, S->stream>>>(nElm, D, S->d_xyz, S->d_conn, lo[0], lo[1], lo[2], inv[0], inv[1], inv[2],
                                                 thrust::raw_pointer_cast(mkey.data()), thrust::raw_pointer_cast(eidx.data()));
    thrust::stable_sort_by_key(pol, mkey.begin(), mkey.end(), eidx.begin());
    patch_of_kernel<<<148 * 8, 256, 0, S->stream>>>(nElm, P->PE, thrust::raw_pointer_cast(eidx.data()), thrust::raw_pointer_cast(patch_of.data()));
  }
  // node owners and the sorted (patch, node, element) pairs
  const int64_t nAllPairs = nElm * NL;
  thrust::device_vector<uint64_t> key(nAllPairs);
  thrust::device_vector<int32_t>  payload(nAllPairs);
  {
    thrust::device_vector<int32_t> owner(S->nDOF, 0x7fffffff);
    owner_kernel<<<148 * 8, 256, 0, S->stream>>>(nElm, D, NS, NP, adrU, adrP, thrust::raw_pointer_cast(patch_of.data()),
                                                thrust::raw_pointer_cast(owner.data()));
    pair_key_kernel<<<148 * 8, 256, 0, S->stream>>>(nElm, D, NS, NP, adrU, adrP, thrust::raw_pointer_cast(owner.data()), S->nInc,
                                                   thrust::raw_pointer_cast(key.data()), thrust::raw_pointer_cast(payload.data()));
    thrust::stable_sort_by_key(pol, key.begin(), key.end(), payload.begin());
  }
  const int64_t nPairs = thrust::lower_bound(pol, key.begin(), key.end(), ~0ull) - key.begin();
  P->nPairs            = nPairs;
  if(nPairs == 0) {
    set_error("patch plan: no unknown rows");
    return B200_ERR_UNSUPP;
  }
  thrust::device_vector<int32_t> flag(nPairs), nodeidx(nPairs);
  head_flag_kernel<<<148 * 8, 256, 0, S->stream>>>(nPairs, thrust::raw_pointer_cast(key.data()), thrust::raw_pointer_cast(flag.data()));
  thrust::inclusive_scan(pol, flag.begin(), flag.end(), nodeidx.begin());
  const int32_t nNodes = nodeidx[nPairs - 1];
  P->nNodes            = nNodes;
  auto patch_keys      = thrust::make_transform_iterator(thrust::counting_iterator<int32_t>(0), PatchKeyOp());
  thrust::device_vector<int32_t> pair_ptr(nPatch + 1), node_ptr(nPatch + 1), elem_ptr(nPatch + 1), cand_ptr(nPatch + 1), blk_ptr(nPatch + 1);
  thrust::lower_bound(pol, key.begin(), key.begin() + nPairs, patch_keys, patch_keys + (nPatch + 1), pair_ptr.begin());
  thrust::device_vector<int32_t> pu_end(nPatch);
  {
    thrust::device_vector<uint64_t> pk(nPatch);
    thrust::transform(pol, thrust::counting_iterator<int32_t>(0), thrust::counting_iterator<int32_t>(nPatch), pk.begin(), PEndKeyOp());
    thrust::lower_bound(pol, key.begin(), key.begin() + nPairs, pk.begin(), pk.end(), pu_end.begin());
  }
  {
    PtrLookupOp op{thrust::raw_pointer_cast(nodeidx.data()), thrust::raw_pointer_cast(pair_ptr.data()), nPairs, nNodes};
    thrust::transform(pol, thrust::counting_iterator<int32_t>(0), thrust::counting_iterator<int32_t>(nPatch + 1), node_ptr.begin(), op);
  }
  // elements of every patch (own + halo), sorted by element index
  int64_t nPE = 0;
  {
    thrust::device_vector<uint64_t> ekey(nPairs);
    elem_key_kernel<<<148 * 8, 256, 0, S->stream>>>(nPairs, thrust::raw_pointer_cast(key.data()), thrust::raw_pointer_cast(payload.data()),
                                                   thrust::raw_pointer_cast(ekey.data()));
    thrust::sort(pol, ekey.begin(), ekey.end());
    nPE = thrust::unique(pol, ekey.begin(), ekey.end()) - ekey.begin();
    thrust::lower_bound(pol, ekey.begin(), ekey.begin() + nPE, patch_keys, patch_keys + (nPatch + 1), elem_ptr.begin());
    B200_CUDA(cudaMalloc(&P->elems, (size_t)nPE * sizeof(int32_t)));
    low32_kernel<<<148 * 8, 256, 0, S->stream>>>(nPE, thrust::raw_pointer_cast(ekey.data()), P->elems);
  }
  P->nElemsTot = nPE;
  int *d_err;
  B200_CUDA(cudaMalloc(&d_err, sizeof(int)));
  B200_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), S->stream));
  thrust::device_vector<int32_t> node_patch(nNodes);
  thrust::device_vector<uint16_t> pi(nPairs);
  B200_CUDA(cudaMalloc(&P->pairs, (size_t)nPairs * sizeof(uint16_t)));
  B200_CUDA(cudaMalloc(&P->nodes, (size_t)nNodes * sizeof(NodeRec<D>)));
  NodeRec<D> *nodes = static_cast<NodeRec<D> *>(P->nodes);
  node_record_kernel<D><<<148 * 8, 256, 0, S->stream>>>(nPairs, NS, NP, thrust::raw_pointer_cast(key.data()), thrust::raw_pointer_cast(payload.data()),
                                                       thrust::raw_pointer_cast(flag.data()), thrust::raw_pointer_cast(nodeidx.data()),
                                                       thrust::raw_pointer_cast(pair_ptr.data()), thrust::raw_pointer_cast(elem_ptr.data()), P->elems,
                                                       adrU, adrP, S->d_ia, S->nInc, P->pairs, thrust::raw_pointer_cast(pi.data()), nodes,
                                                       thrust::raw_pointer_cast(node_patch.data()), d_err);
  // candidates -> block slots
  int blockmask = 0;
  for(int bi = 0; bi < 2; ++bi)
    for(int bj = 0; bj < 2; ++bj)
      if(S->has_matrix_block[bi][bj]) blockmask |= 1 << (bi * 2 + bj);
  blockmask &= ~8; // no P-P block in the fused Taylor-Hood system
  const int64_t nAllCand = nPairs * NL;
  int64_t       nCand = 0, nBlocks = 0;
  {
    thrust::device_vector<uint64_t> ckey(nAllCand);
    thrust::device_vector<int32_t>  cpay(nAllCand);
    cand_kernel<<<148 * 8, 256, 0, S->stream>>>(nPairs, D, NS, NP, thrust::raw_pointer_cast(payload.data()), P->pairs,
                                               thrust::raw_pointer_cast(pi.data()), thrust::raw_pointer_cast(nodeidx.data()), adrU, adrP, S->nInc, blockmask,
                                               thrust::raw_pointer_cast(ckey.data()), thrust::raw_pointer_cast(cpay.data()));
    // the pair arrays are no longer needed: release them before the big sort
    thrust::device_vector<uint64_t>().swap(key);
    thrust::device_vector<int32_t>().swap(payload);
    thrust::device_vector<int32_t>().swap(flag);
    thrust::stable_sort_by_key(pol, ckey.begin(), ckey.end(), cpay.begin());
    nCand = thrust::lower_bound(pol, ckey.begin(), ckey.end(), ~0ull) - ckey.begin();
    if(nCand == 0) {
      cudaFree(d_err);
      set_error("patch plan: no matrix blocks");
      return B200_ERR_UNSUPP;
    }
    thrust::device_vector<int32_t> cflag(nCand), blkidx(nCand);
    head_flag_kernel<<<148 * 8, 256, 0, S->stream>>>(nCand, thrust::raw_pointer_cast(ckey.data()), thrust::raw_pointer_cast(cflag.data()));
    thrust::inclusive_scan(pol, cflag.begin(), cflag.end(), blkidx.begin());
    nBlocks = blkidx[nCand - 1];
    // first candidate of every patch: keys are (node index << 32 | column), nodes are grouped by patch
    {
      thrust::device_vector<uint64_t> nk(nPatch + 1);
      thrust::transform(pol, node_ptr.begin(), node_ptr.end(), nk.begin(), PatchKeyOp());
      thrust::lower_bound(pol, ckey.begin(), ckey.begin() + nCand, nk.begin(), nk.end(), cand_ptr.begin());
      PtrLookupOp op{thrust::raw_pointer_cast(blkidx.data()), thrust::raw_pointer_cast(cand_ptr.data()), nCand, (int32_t)nBlocks};
      thrust::transform(pol, thrust::counting_iterator<int32_t>(0), thrust::counting_iterator<int32_t>(nPatch + 1), blk_ptr.begin(), op);
    }
    B200_CUDA(cudaMalloc(&P->blocks, (size_t)nBlocks * sizeof(BlockRec)));
    thrust::device_vector<uint64_t> bclass(nBlocks);
    thrust::device_vector<uint8_t>  covered(S->nnz, 0);
    block_record_kernel<D><<<148 * 8, 256, 0, S->stream>>>(
      nCand, NS, NP, thrust::raw_pointer_cast(ckey.data()), thrust::raw_pointer_cast(cpay.data()), thrust::raw_pointer_cast(cflag.data()),
      thrust::raw_pointer_cast(blkidx.data()), thrust::raw_pointer_cast(node_patch.data()), thrust::raw_pointer_cast(node_ptr.data()),
      thrust::raw_pointer_cast(cand_ptr.data()), thrust::raw_pointer_cast(elem_ptr.data()), P->elems, nodes, adrU, adrP, S->d_ia, S->d_ja, S->nInc,
      P->blocks, thrust::raw_pointer_cast(bclass.data()), thrust::raw_pointer_cast(covered.data()), d_err);
    B200_CUDA(cudaStreamSynchronize(S->stream));
    thrust::device_vector<uint64_t>().swap(ckey);
    thrust::device_vector<int32_t>().swap(cflag);
    thrust::device_vector<int32_t>().swap(blkidx);
    // class sort inside every patch: warps of the block phase then run one kind and one trip count
    {
      thrust::device_ptr<uint64_t> bp(reinterpret_cast<uint64_t *>(P->blocks));
      thrust::stable_sort_by_key(pol, bclass.begin(), bclass.end(), bp);
    }
    thrust::device_vector<int32_t> len(nBlocks), pos(nBlocks), uu_end(nPatch);
    block_list_len_kernel<<<148 * 8, 256, 0, S->stream>>>(nBlocks, P->blocks, thrust::raw_pointer_cast(len.data()));
    thrust::exclusive_scan(pol, len.begin(), len.end(), pos.begin());
    const int32_t nList = (int32_t)pos[nBlocks - 1] + (int32_t)len[nBlocks - 1];
    P->nCtr             = nList;
    B200_CUDA(cudaMalloc(&P->ctr, (size_t)std::max(nList, 1) * sizeof(uint16_t)));
    block_relayout_kernel<<<148 * 8, 256, 0, S->stream>>>(nBlocks, P->blocks, thrust::raw_pointer_cast(bclass.data()),
                                                         thrust::raw_pointer_cast(pos.data()), thrust::raw_pointer_cast(blk_ptr.data()),
                                                         thrust::raw_pointer_cast(cand_ptr.data()), thrust::raw_pointer_cast(cpay.data()), P->ctr,
                                                         d_err);
    {
      // end of the U-U slots of every patch: first class key with the kind bit set
      thrust::device_vector<uint64_t> uk(nPatch);
      thrust::transform(pol, thrust::counting_iterator<int32_t>(0), thrust::counting_iterator<int32_t>(nPatch), uk.begin(), UUEndKeyOp());
      thrust::lower_bound(pol, bclass.begin(), bclass.end(), uk.begin(), uk.end(), uu_end.begin());
    }
    B200_CUDA(cudaMalloc(&P->desc, (size_t)nPatch * sizeof(PatchDesc)));
    thrust::device_vector<int32_t> smem_d(nPatch);
    desc_kernel<<<(nPatch + 255) / 256, 256, 0, S->stream>>>(nPatch, thrust::raw_pointer_cast(elem_ptr.data()), thrust::raw_pointer_cast(node_ptr.data()),
                                                            thrust::raw_pointer_cast(pair_ptr.data()), thrust::raw_pointer_cast(blk_ptr.data()),
                                                            thrust::raw_pointer_cast(uu_end.data()), thrust::raw_pointer_cast(pu_end.data()),
                                                            thrust::raw_pointer_cast(pos.data()), nBlocks, nList, PS<D, NS, NP>::W,
                                                            PS<D, NS, NP>::PRW, P->desc, thrust::raw_pointer_cast(smem_d.data()));
    B200_CUDA(cudaStreamSynchronize(S->stream));
    thrust::device_vector<int32_t>().swap(cpay);
    P->max_smem_doubles = thrust::reduce(pol, smem_d.begin(), smem_d.end(), 0, thrust::maximum<int32_t>());
    P->nZero = thrust::count(pol, covered.begin(), covered.end(), (uint8_t)0);
    if(P->nZero > 0) {
      B200_CUDA(cudaMalloc(&P->zero_idx, (size_t)P->nZero * sizeof(int64_t)));
      thrust::copy_if(pol, thrust::counting_iterator<int64_t>(0), thrust::counting_iterator<int64_t>(S->nnz), covered.begin(),
                      thrust::device_pointer_cast(P->zero_idx), IsZeroOp());
    }
  }
  P->nBlocks = nBlocks;
  count_launch(12);
  int h_err = 0;
  B200_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, S->stream));
  B200_CUDA(cudaStreamSynchronize(S->stream));
  cudaFree(d_err);
  if(h_err) {
    set_error("patch plan: the CSR pattern / numbering does not have the regular node-block structure (code " + std::to_string(h_err) + ")");
    return B200_ERR_UNSUPP;
  }
  {
    thrust::device_ptr<PatchDesc> dp(P->desc);
    P->maxE = thrust::transform_reduce(pol, dp, dp + nPatch, DescNe(), 0, thrust::maximum<int32_t>());
    P->max_nb = thrust::transform_reduce(pol, dp, dp + nPatch, DescNb(), 0, thrust::maximum<int32_t>());
    P->max_np = thrust::transform_reduce(pol, dp, dp + nPatch, DescNp(), 0, thrust::maximum<int32_t>());
    P->max_nn = thrust::transform_reduce(pol, dp, dp + nPatch, DescNn(), 0, thrust::maximum<int32_t>());
  }
  const size_t smem = ((size_t)((P->tab_len_src + 1) & ~1) + (size_t)P->max_smem_doubles) * sizeof(double);
  if(smem > 224 * 1024) {
    set_error("patch plan: " + std::to_string(P->maxE) + " elements in one patch (" + std::to_string(smem) + " bytes) exceed the shared-memory budget");
    return B200_ERR_UNSUPP;
  }
  return B200_OK;
}

int build_patch_plan(System *S)
{
  patch_free(S);
  GatherTables gt;
  if(!gather_tables(S, &gt)) return B200_ERR_UNSUPP;
  const int D = S->dim, NS = S->spaces[S->su].nS, NP = S->spaces[S->sp].nS;
  int       rc = B200_ERR_UNSUPP;
  try {
    if(D == 2 && NS == 6 && NP == 3)
      rc = build_patch_plan_t<2, 6, 3>(S, gt);
    else if(D == 3 && NS == 10 && NP == 4)
      rc = build_patch_plan_t<3, 10, 4>(S, gt);
    else
      set_error("patch plan: only P2/P1 simplices");
  } catch(const std::exception &ex) {
    set_error(std::string("patch plan: ") + ex.what());
    rc = B200_ERR_CUDA;
  }
  if(rc != B200_OK) patch_free(S);
  return rc;
}

template <int D, int NS, int NP, int NT, int MINB> static int launch_patch_t(System *S, int what, const THCoeffs &c)
{
  PatchPlan *P = static_cast<PatchPlan *>(S->patch);
  PatchArgs  a;
  a.desc   = P->desc;
  a.elems  = P->elems;
  a.nodes  = P->nodes;
  a.pairs  = P->pairs;
  a.blocks = P->blocks;
  a.ctr    = P->ctr;
  a.adrU   = S->spaces[S->su].d_adr;
  a.adrP   = S->spaces[S->sp].d_adr;
  a.sol    = S->d_sol;
  a.soldot = S->have_soldot ? S->d_soldot : nullptr;
  a.source = (c.c_src != 0.) ? S->d_source : nullptr;
  a.geo    = P->d_geo;
  a.tab    = P->d_tab;
  a.val    = S->d_val;
  a.rhs    = S->d_rhs;
  a.nInc   = S->nInc;
  a.nq     = S->nq;
  a.ntab   = a.source ? P->tab_len_src : P->tab_len;
  a.c      = c;
  a.c0     = S->c0;
  const size_t smem = ((size_t)((a.ntab + 1) & ~1) + (size_t)P->max_smem_doubles) * sizeof(double);
#define B200_LAUNCH_P(MAT, RES)                                                                                                          \
  do {                                                                                                                                   \
    auto kern = patch_kernel<D, NS, NP, NT, MINB, MAT, RES>;                                                                             \
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                                       \
    kern<<<P->nPatch, NT, smem, S->stream>>>(a);                                                                                         \
  } while(0)
  if(what == 3)
    B200_LAUNCH_P(true, true);
  else if(what == 2)
    B200_LAUNCH_P(true, false);
  else
    B200_LAUNCH_P(false, true);
#undef B200_LAUNCH_P
  count_launch();
  if((what & 2) && P->nZero > 0) {
    patch_zero_kernel<<<(unsigned)std::min<int64_t>((P->nZero + 255) / 256, 148 * 8), 256, 0, S->stream>>>(P->nZero, P->zero_idx, S->d_val);
    count_launch();
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

// row-slice kernel: B200_ERR_UNSUPP when the pass needs something it does not do (transient residual, sources, oversized
// patches); the caller then runs the patch kernel
static int launch_slice_2d(System *S, int what, const THCoeffs &c)
{
  constexpr int NT = 256, MAXS = 12;
  PatchPlan *P = static_cast<PatchPlan *>(S->patch);
  if((c.c_mass != 0. && S->have_soldot && (what & 1)) || c.c_src != 0. || P->max_nb > NT * MAXS || P->maxE > NT) return B200_ERR_UNSUPP;
  PatchArgs a;
  a.desc   = P->desc;
  a.elems  = P->elems;
  a.nodes  = P->nodes;
  a.pairs  = P->pairs;
  a.blocks = P->blocks;
  a.ctr    = P->ctr;
  a.adrU   = S->spaces[S->su].d_adr;
  a.adrP   = S->spaces[S->sp].d_adr;
  a.sol    = S->d_sol;
  a.soldot = nullptr;
  a.source = nullptr;
  a.geo    = P->d_geo;
  a.tab    = P->d_tab;
  a.val    = S->d_val;
  a.rhs    = S->d_rhs;
  a.nInc   = S->nInc;
  a.nq     = S->nq;
  a.ntab   = P->tab_len;
  a.c      = c;
  a.c0     = S->c0;
  B200_CUDA(cudaMemcpyToSymbolAsync(c_gt, P->d_gt, sizeof(double) * GT<2, 6, 3>::O_W, 0, cudaMemcpyDeviceToDevice, S->stream));
  const size_t smem = (size_t)P->maxE * (PS<2, 6, 3>::W + 32 + 16) * sizeof(double) + (size_t)(((P->max_np + 3) & ~3) + P->max_nn + 8) * sizeof(uint16_t);
  if(smem > 224 * 1024) return B200_ERR_UNSUPP;
  if(what & 1) {
    B200_CUDA(cudaFuncSetAttribute(slice_kernel_2d<NT, MAXS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    slice_kernel_2d<NT, MAXS, true><<<P->nPatch, NT, smem, S->stream>>>(a);
  } else {
    B200_CUDA(cudaFuncSetAttribute(slice_kernel_2d<NT, MAXS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    slice_kernel_2d<NT, MAXS, false><<<P->nPatch, NT, smem, S->stream>>>(a);
  }
  count_launch();
  if(P->nZero > 0) {
    patch_zero_kernel<<<(unsigned)std::min<int64_t>((P->nZero + 255) / 256, 148 * 8), 256, 0, S->stream>>>(P->nZero, P->zero_idx, S->d_val);
    count_launch();
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

// what: bit 0 residual, bit 1 matrix; OVERWRITES val / rhs (every unknown row is written exactly once)
int launch_patch(System *S, int what, const THCoeffs &c)
{
  const PatchPlan *P = static_cast<const PatchPlan *>(S->patch);
  if(S->dim == 2 && P->slice && (what & 2)) {
    const int rc = launch_slice_2d(S, what, c);
    if(rc != B200_ERR_UNSUPP) return rc;
  }
  if(S->dim == 2) {
    if(P->NT == 128) return launch_patch_t<2, 6, 3, 128, 4>(S, what, c);
    if(P->NT == 512) return launch_patch_t<2, 6, 3, 512, 1>(S, what, c);
    return launch_patch_t<2, 6, 3, 256, 3>(S, what, c);
  }
  if(P->NT == 128) return launch_patch_t<3, 10, 4, 128, 2>(S, what, c);
  if(P->NT == 512) return launch_patch_t<3, 10, 4, 512, 1>(S, what, c);
  return launch_patch_t<3, 10, 4, 256, 1>(S, what, c);
}

} // namespace b200
