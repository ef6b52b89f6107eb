// Fused element kernels: gather -> geometry -> quadrature loop -> local Jacobian + residual -> scatter.
//
// B200-first restatement of the reference's per-form, per-element host loops
//   feBilinearForm::initialize            src/feBilinearForm.cpp:284-367   (gather)
//   feSysElm_*::computeAe / computeBe     src/feSysElm.cpp, src/feVectorSysElm.cpp (quadrature loops)
//   feLinearSystemMklPardiso::assemble*   src/feLinearSystemMklPardiso.cpp:501-749 (scatter)
// Instead of one mesh traversal per weak form with dense nF x nF vector-basis contractions (half of them
// multiplying the structural zeros of the vector Lagrange layout, src/feSpace_2D.cpp:41-55), ALL registered forms
// are fused into one pass that works on the scalar basis and the block structure phi_{a*dim+c} = phi_a e_c.
//
// Thread mapping: one thread per LOCAL ROW of the fused element system (2-D Taylor-Hood: 15 rows, 3-D: 34 rows),
// EPB elements per CTA.  Per chunk of CH quadrature points:
//   phase 1  one thread per (element, quadrature point): physical gradients, u, grad u, p and the row-independent
//            combinations, written to shared memory in SoA layout (conflict-free);
//   phase 2  one thread per row accumulates its M Jacobian entries + residual entry in registers from shared
//            memory (all lanes of an element read the same words -> broadcast).
// Scatter: precomputed CSR slot per local entry, red.global.add.f64 (or plain adds inside one element colour).
#include <cstdio>

#include "system.h"
#include "device_common.cuh"

namespace b200 {

// ----------------------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------------------
template <bool ATOMIC> __device__ __forceinline__ void add_to(double *p, double v)
{
  if(ATOMIC)
    atomicAdd(p, v); // result unused -> RED.E.ADD.F64
  else
    *p += v;
}

// ----------------------------------------------------------------------------------------------------------
// Taylor-Hood (vector P_k velocity + scalar pressure) fused kernel
// ----------------------------------------------------------------------------------------------------------
struct THArgs {
  const double  *xyz;
  const int32_t *conn, *adrU, *adrP;
  const double  *sol, *soldot, *source, *tab;
  const int32_t *slot;
  double        *val, *rhs;
  const int32_t *elem_list; // colour-sorted element list (coloured scatter) or nullptr
  int64_t        elem_begin, elem_end, nInc;
  int            nq;
  THCoeffs       c;
  double         c0;
};

template <int DIM, int NS, int NP> struct THShape {
  static constexpr int NU = NS * DIM, M = NU + NP;
  static constexpr int F_G = 0;                   // g[b][m]      NS*DIM
  static constexpr int F_UGP = F_G + NS * DIM;    // c_conv u.g_b NS
  static constexpr int F_CG = F_UGP + NS;         // c_conv d_j u_i  [j][i]
  static constexpr int F_R0 = F_CG + DIM * DIM;   // DIM
  static constexpr int F_Q = F_R0 + DIM;          // Q[m][i]
  static constexpr int F_DIVU = F_Q + DIM * DIM;  // 1
  static constexpr int F_JW = F_DIVU + 1;         // 1
  static constexpr int NF = F_JW + 1;
  static constexpr int EL = DIM * DIM + 1 + NU + NP + NU; // per-element smem doubles: G, detJ, U, P, Udot
};

template <int DIM, int NS, int NP, int EPB, int CH>
size_t th_smem_bytes(int nq)
{
  using S = THShape<DIM, NS, NP>;
  size_t d = (size_t)nq * (1 + NS + NS * DIM + NP) + (size_t)S::NF * EPB * CH + (size_t)EPB * S::EL;
  return d * sizeof(double) + (size_t)EPB * S::M * sizeof(int32_t);
}

template <int DIM, int NS, int NP, int EPB, int CH, bool MAT, bool RES, bool ATOMIC>
__global__ void __launch_bounds__(EPB *(NS *DIM + NP)) th_kernel(const THArgs a)
{
  using S = THShape<DIM, NS, NP>;
  constexpr int NU = S::NU, M = S::M, NT = EPB * M, NPAIR = EPB * CH, EL = S::EL;
  extern __shared__ double sm[];
  const int nq   = a.nq;
  double   *s_w  = sm;
  double   *s_LU = s_w + nq;
  double   *s_dLU = s_LU + nq * NS;
  double   *s_LP = s_dLU + nq * NS * DIM;
  double   *s_qp = s_LP + nq * NP;
  double   *s_el = s_qp + S::NF * NPAIR;
  int32_t  *s_adr = reinterpret_cast<int32_t *>(s_el + EPB * EL);

  const int     tid   = threadIdx.x;
  const int64_t ebase = a.elem_begin + (int64_t)blockIdx.x * EPB;

  // ---- phase 0: tables, gather (feBilinearForm::initialize) and element geometry --------------------------
  const int tab_len = nq * (1 + NS + NS * DIM + NP);
  for(int i = tid; i < tab_len; i += NT) sm[i] = a.tab[i];

  for(int idx = tid; idx < EPB * M; idx += NT) {
    const int     el = idx / M, i = idx - el * M;
    const int64_t ei = ebase + el;
    if(ei < a.elem_end) {
      const int64_t e   = a.elem_list ? (int64_t)a.elem_list[ei] : ei;
      double       *E   = s_el + el * EL;
      int32_t       dof;
      if(i < NU) {
        dof                      = a.adrU[e * NU + i];
        E[DIM * DIM + 1 + i]     = a.sol[dof];
        E[DIM * DIM + 1 + NU + NP + i] = a.soldot ? a.soldot[dof] : 0.;
      } else {
        dof                  = a.adrP[e * NP + (i - NU)];
        E[DIM * DIM + 1 + i] = a.sol[dof];
      }
      s_adr[idx] = dof;
      if(i == 0) {
        int32_t vtx[DIM + 1];
#pragma unroll
        for(int v = 0; v <= DIM; ++v) vtx[v] = a.conn[e * (DIM + 1) + v];
        element_geometry<DIM>(a.xyz, vtx, E, E + DIM * DIM);
      }
    }
  }
  __syncthreads();

  const int     el_me = tid / M, i_me = tid - el_me * M;
  const int64_t ei_me = ebase + el_me;
  const bool    live  = ei_me < a.elem_end;
  const int64_t e_me  = live ? (a.elem_list ? (int64_t)a.elem_list[ei_me] : ei_me) : 0;

  double acc[M];
#pragma unroll
  for(int j = 0; j < M; ++j) acc[j] = 0.;
  double res = 0.;

  const THCoeffs c      = a.c;
  const double   massc0 = c.c_mass * a.c0;

  for(int k0 = 0; k0 < nq; k0 += CH) {
    // ---- phase 1: one thread per (element, quadrature point) -----------------------------------------
    for(int pidx = tid; pidx < NPAIR; pidx += NT) {
      const int     el = pidx / CH, kk = pidx - el * CH, k = k0 + kk;
      const int64_t ei = ebase + el;
      if(k < nq && ei < a.elem_end) {
        const double *E = s_el + el * EL;
        const double *G = E, *U = E + DIM * DIM + 1, *P = U + NU, *Ud = P + NP;
        double        g[NS][DIM];
#pragma unroll
        for(int b = 0; b < NS; ++b)
#pragma unroll
          for(int m = 0; m < DIM; ++m) {
            double v = 0.;
#pragma unroll
            for(int al = 0; al < DIM; ++al) v += s_dLU[(k * NS + b) * DIM + al] * G[al * DIM + m];
            g[b][m]                          = v;
            s_qp[(S::F_G + b * DIM + m) * NPAIR + pidx] = v;
          }
        double u[DIM], ud[DIM], gu[DIM][DIM];
#pragma unroll
        for(int n = 0; n < DIM; ++n) {
          u[n] = 0.;
          ud[n] = 0.;
#pragma unroll
          for(int m = 0; m < DIM; ++m) gu[m][n] = 0.;
        }
#pragma unroll
        for(int b = 0; b < NS; ++b) {
          const double phi = s_LU[k * NS + b];
#pragma unroll
          for(int n = 0; n < DIM; ++n) {
            const double ubn = U[b * DIM + n];
            u[n] += phi * ubn;
            ud[n] += phi * Ud[b * DIM + n];
#pragma unroll
            for(int m = 0; m < DIM; ++m) gu[m][n] += g[b][m] * ubn; // gu[m][n] = d_m u_n (src/feSpace.cpp:1391-1394)
          }
        }
        double p = 0.;
#pragma unroll
        for(int q = 0; q < NP; ++q) p += s_LP[k * NP + q] * P[q];
#pragma unroll
        for(int b = 0; b < NS; ++b) {
          double v = 0.;
#pragma unroll
          for(int m = 0; m < DIM; ++m) v += u[m] * g[b][m];
          s_qp[(S::F_UGP + b) * NPAIR + pidx] = c.c_conv * v;
        }
        double divu = 0.;
#pragma unroll
        for(int m = 0; m < DIM; ++m) divu += gu[m][m];
        const int64_t e = a.elem_list ? (int64_t)a.elem_list[ei] : ei;
#pragma unroll
        for(int i = 0; i < DIM; ++i) {
          double ugu = 0.; // (u.grad u)_i = u_n d_n u_i (src/feVectorSysElm.cpp:1228-1233)
#pragma unroll
          for(int n = 0; n < DIM; ++n) ugu += u[n] * gu[n][i];
          double r0 = -c.c_conv * ugu - c.c_mass * ud[i];
          if(a.source) r0 -= c.c_src * a.source[(e * nq + k) * DIM + i];
          s_qp[(S::F_R0 + i) * NPAIR + pidx] = r0;
#pragma unroll
          for(int m = 0; m < DIM; ++m) {
            s_qp[(S::F_CG + m * DIM + i) * NPAIR + pidx] = c.c_conv * gu[m][i];
            double qv = c.sig_mu * (gu[m][i] + gu[i][m]) - c.diff_k * gu[m][i];
            if(m == i) qv += (c.c_gradp - c.c_sig) * p;
            s_qp[(S::F_Q + m * DIM + i) * NPAIR + pidx] = qv;
          }
        }
        s_qp[S::F_DIVU * NPAIR + pidx] = divu;
        s_qp[S::F_JW * NPAIR + pidx]   = E[DIM * DIM] * s_w[k];
      }
    }
    __syncthreads();

    // ---- phase 2: one thread per local row ---------------------------------------------------------------
    if(live) {
      const int kend = (nq - k0 < CH) ? (nq - k0) : CH;
      if(i_me < NU) {
        const int ar = i_me / DIM, cr = i_me - ar * DIM;
        for(int kk = 0; kk < kend; ++kk) {
          const int     k    = k0 + kk;
          const double *qp   = s_qp + el_me * CH + kk;
          const double  jw   = qp[S::F_JW * NPAIR];
          const double  phia = s_LU[k * NS + ar];
          double        ga[DIM];
#pragma unroll
          for(int m = 0; m < DIM; ++m) ga[m] = qp[(S::F_G + ar * DIM + m) * NPAIR];
          const double gac = qp[(S::F_G + ar * DIM + cr) * NPAIR];
          if(MAT) {
            double T[DIM], V[DIM];
#pragma unroll
            for(int j = 0; j < DIM; ++j) {
              T[j] = jw * phia * qp[(S::F_CG + j * DIM + cr) * NPAIR];
              V[j] = -jw * c.sig_mu * ga[j];
            }
            const double Nn = jw * (c.diff_k - c.sig_mu), A1 = jw * phia, Mm = jw * massc0 * phia;
#pragma unroll
            for(int b = 0; b < NS; ++b) {
              const double phib = s_LU[k * NS + b];
              double       dot  = 0.;
#pragma unroll
              for(int m = 0; m < DIM; ++m) dot += ga[m] * qp[(S::F_G + b * DIM + m) * NPAIR];
              const double gbc = qp[(S::F_G + b * DIM + cr) * NPAIR];
              const double s   = A1 * qp[(S::F_UGP + b) * NPAIR] + Nn * dot + Mm * phib;
#pragma unroll
              for(int j = 0; j < DIM; ++j) acc[b * DIM + j] += T[j] * phib + V[j] * gbc + (j == cr ? s : 0.);
            }
            const double up = jw * (c.c_sig - c.c_gradp) * gac;
#pragma unroll
            for(int q = 0; q < NP; ++q) acc[NU + q] += up * s_LP[k * NP + q];
          }
          if(RES) {
            double r = qp[(S::F_R0 + cr) * NPAIR] * phia;
#pragma unroll
            for(int m = 0; m < DIM; ++m) r += ga[m] * qp[(S::F_Q + m * DIM + cr) * NPAIR];
            res += jw * r;
          }
        }
      } else {
        const int q = i_me - NU;
        for(int kk = 0; kk < kend; ++kk) {
          const int     k    = k0 + kk;
          const double *qp   = s_qp + el_me * CH + kk;
          const double  coef = qp[S::F_JW * NPAIR] * c.c_div * s_LP[k * NP + q];
          if(MAT) {
#pragma unroll
            for(int j = 0; j < NU; ++j) acc[j] += coef * qp[(S::F_G + j) * NPAIR];
          }
          if(RES) res -= coef * qp[S::F_DIVU * NPAIR];
        }
      }
    }
    __syncthreads();
  }

  // ---- scatter (src/feLinearSystemMklPardiso.cpp:565-660, :731-737) ------------------------------------------
  if(live) {
    if(MAT) {
      const int32_t *sl = a.slot + (e_me * M + i_me) * (int64_t)M;
#pragma unroll
      for(int j = 0; j < M; ++j) {
        const int32_t s = sl[j];
        if(s >= 0) add_to<ATOMIC>(a.val + s, acc[j]);
      }
    }
    if(RES) {
      const int32_t dof = s_adr[el_me * M + i_me];
      if(dof < a.nInc) add_to<ATOMIC>(a.rhs + dof, res);
    }
  }
}

// ----------------------------------------------------------------------------------------------------------
// Scalar Lagrange fused kernel: DIFFUSION + SOURCE + TRANSIENT_MASS
// ----------------------------------------------------------------------------------------------------------
struct SCArgs {
  const double  *xyz;
  const int32_t *conn, *adr;
  const double  *sol, *soldot, *source, *tab;
  const double  *kcoef; // [nElm][nq] tabulated diffusivity (a space-dependent coefficient callback), or nullptr
  const int32_t *slot;
  double        *val, *rhs;
  const int32_t *elem_list;
  int64_t        elem_begin, elem_end, nInc;
  int            nq;
  ScalarCoeffs   c;
  double         c0;
};

template <int DIM, int NS> struct SCShape {
  static constexpr int F_G = 0;               // NS*DIM
  static constexpr int F_GU = F_G + NS * DIM; // DIM  k*grad u
  static constexpr int F_R0 = F_GU + DIM;     // 1
  static constexpr int F_JW = F_R0 + 1;
  static constexpr int F_KQ = F_JW + 1;       // diffusivity at the quadrature node
  static constexpr int NF = F_KQ + 1;
  static constexpr int EL = DIM * DIM + 1 + NS + NS;
};

template <int DIM, int NS, int EPB, int CH> size_t sc_smem_bytes(int nq)
{
  using S = SCShape<DIM, NS>;
  size_t d = (size_t)nq * (1 + NS + NS * DIM) + (size_t)S::NF * EPB * CH + (size_t)EPB * S::EL;
  return d * sizeof(double) + (size_t)EPB * NS * sizeof(int32_t);
}

template <int DIM, int NS, int EPB, int CH, bool MAT, bool RES, bool ATOMIC>
__global__ void __launch_bounds__(EPB *NS) sc_kernel(const SCArgs a)
{
  using S = SCShape<DIM, NS>;
  constexpr int M = NS, NT = EPB * M, NPAIR = EPB * CH, EL = S::EL;
  extern __shared__ double sm[];
  const int nq    = a.nq;
  double   *s_w   = sm;
  double   *s_L   = s_w + nq;
  double   *s_dL  = s_L + nq * NS;
  double   *s_qp  = s_dL + nq * NS * DIM;
  double   *s_el  = s_qp + S::NF * NPAIR;
  int32_t  *s_adr = reinterpret_cast<int32_t *>(s_el + EPB * EL);

  const int     tid   = threadIdx.x;
  const int64_t ebase = a.elem_begin + (int64_t)blockIdx.x * EPB;
  const int     tab_len = nq * (1 + NS + NS * DIM);
  for(int i = tid; i < tab_len; i += NT) sm[i] = a.tab[i];

  {
    const int     el = tid / M, i = tid - el * M;
    const int64_t ei = ebase + el;
    if(ei < a.elem_end) {
      const int64_t e   = a.elem_list ? (int64_t)a.elem_list[ei] : ei;
      double       *E   = s_el + el * EL;
      const int32_t dof = a.adr[e * NS + i];
      E[DIM * DIM + 1 + i]      = a.sol[dof];
      E[DIM * DIM + 1 + NS + i] = a.soldot ? a.soldot[dof] : 0.;
      s_adr[tid]                = dof;
      if(i == 0) {
        int32_t vtx[DIM + 1];
#pragma unroll
        for(int v = 0; v <= DIM; ++v) vtx[v] = a.conn[e * (DIM + 1) + v];
        element_geometry<DIM>(a.xyz, vtx, E, E + DIM * DIM);
      }
    }
  }
  __syncthreads();

  const int     el_me = tid / M, i_me = tid - el_me * M;
  const int64_t ei_me = ebase + el_me;
  const bool    live  = ei_me < a.elem_end;
  const int64_t e_me  = live ? (a.elem_list ? (int64_t)a.elem_list[ei_me] : ei_me) : 0;
  double        acc[M];
#pragma unroll
  for(int j = 0; j < M; ++j) acc[j] = 0.;
  double             res = 0.;
  const ScalarCoeffs c   = a.c;
  const double       massc0 = c.c_mass * a.c0;

  for(int k0 = 0; k0 < nq; k0 += CH) {
    for(int pidx = tid; pidx < NPAIR; pidx += NT) {
      const int     el = pidx / CH, kk = pidx - el * CH, k = k0 + kk;
      const int64_t ei = ebase + el;
      if(k < nq && ei < a.elem_end) {
        const double *E = s_el + el * EL;
        const double *G = E, *U = E + DIM * DIM + 1, *Ud = U + NS;
        double        gu[DIM], uv = 0., ud = 0.;
#pragma unroll
        for(int m = 0; m < DIM; ++m) gu[m] = 0.;
#pragma unroll
        for(int b = 0; b < NS; ++b) {
          const double phi = s_L[k * NS + b];
          uv += phi * U[b];
          ud += phi * Ud[b];
#pragma unroll
          for(int m = 0; m < DIM; ++m) {
            double v = 0.;
#pragma unroll
            for(int al = 0; al < DIM; ++al) v += s_dL[(k * NS + b) * DIM + al] * G[al * DIM + m];
            s_qp[(S::F_G + b * DIM + m) * NPAIR + pidx] = v;
            gu[m] += v * U[b];
          }
        }
        const int64_t e  = a.elem_list ? (int64_t)a.elem_list[ei] : ei;
        // the reference evaluates the coefficient callback at every quadrature node (src/feSysElm.cpp:538, :566)
        const double  kq = a.kcoef ? c.k * a.kcoef[e * nq + k] : c.k;
#pragma unroll
        for(int m = 0; m < DIM; ++m) s_qp[(S::F_GU + m) * NPAIR + pidx] = kq * gu[m];
        s_qp[S::F_KQ * NPAIR + pidx] = kq;
        double        r0 = -c.c_mass * ud;
        if(a.source) r0 -= c.c_src * a.source[e * nq + k];
        s_qp[S::F_R0 * NPAIR + pidx] = r0;
        s_qp[S::F_JW * NPAIR + pidx] = E[DIM * DIM] * s_w[k];
      }
    }
    __syncthreads();
    if(live) {
      const int kend = (nq - k0 < CH) ? (nq - k0) : CH;
      for(int kk = 0; kk < kend; ++kk) {
        const int     k    = k0 + kk;
        const double *qp   = s_qp + el_me * CH + kk;
        const double  jw   = qp[S::F_JW * NPAIR];
        const double  phia = s_L[k * NS + i_me];
        double        ga[DIM];
#pragma unroll
        for(int m = 0; m < DIM; ++m) ga[m] = qp[(S::F_G + i_me * DIM + m) * NPAIR];
        if(MAT) {
          const double kj = jw * qp[S::F_KQ * NPAIR], mj = jw * massc0 * phia;
#pragma unroll
          for(int b = 0; b < NS; ++b) {
            double dot = 0.;
#pragma unroll
            for(int m = 0; m < DIM; ++m) dot += ga[m] * qp[(S::F_G + b * DIM + m) * NPAIR];
            acc[b] += kj * dot + mj * s_L[k * NS + b];
          }
        }
        if(RES) {
          double r = qp[S::F_R0 * NPAIR] * phia;
#pragma unroll
          for(int m = 0; m < DIM; ++m) r -= ga[m] * qp[(S::F_GU + m) * NPAIR];
          res += jw * r;
        }
      }
    }
    __syncthreads();
  }
  if(live) {
    if(MAT) {
      const int32_t *sl = a.slot + (e_me * M + i_me) * (int64_t)M;
#pragma unroll
      for(int j = 0; j < M; ++j) {
        const int32_t s = sl[j];
        if(s >= 0) add_to<ATOMIC>(a.val + s, acc[j]);
      }
    }
    if(RES) {
      const int32_t dof = s_adr[tid];
      if(dof < a.nInc) add_to<ATOMIC>(a.rhs + dof, res);
    }
  }
}

// ----------------------------------------------------------------------------------------------------------
// CSR slot map: slot[e][i][j] = position of (adr_i, adr_j) in ja, -1 if essential / block absent
// (precomputed form of the row scan at src/feLinearSystemMklPardiso.cpp:619-660)
// ----------------------------------------------------------------------------------------------------------
__global__ void slot_map_kernel(int64_t nElm, int M, int NU, const int32_t *adrU, const int32_t *adrP, int NP, const int64_t *ia,
                                const int32_t *ja, int64_t nInc, int blockmask, int32_t *slot, int *err)
{
  const int64_t tot = nElm * M * (int64_t)M;
  for(int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx / (M * M);
    const int     r = (int)(idx - e * M * M);
    const int     i = r / M, j = r - i * M;
    const int     bi = i < NU ? 0 : 1, bj = j < NU ? 0 : 1;
    int32_t       out = -1;
    if(blockmask & (1 << (bi * 2 + bj))) {
      const int32_t I = bi == 0 ? adrU[e * NU + i] : adrP[e * NP + (i - NU)];
      const int32_t J = bj == 0 ? adrU[e * NU + j] : adrP[e * NP + (j - NU)];
      if(I < nInc && J < nInc) {
        int64_t lo = ia[I], hi = ia[I + 1] - 1;
        while(lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          if(ja[mid] < J)
            lo = mid + 1;
          else
            hi = mid;
        }
        if(lo < ia[I + 1] && ja[lo] == J)
          out = (int32_t)lo;
        else
          atomicExch(err, 1);
      }
    }
    slot[idx] = out;
  }
}

// ----------------------------------------------------------------------------------------------------------
// host side: plan + launch
// ----------------------------------------------------------------------------------------------------------
static bool is_scalar_kind(int k) { return k == B200_FORM_SOURCE || k == B200_FORM_TRANSIENT_MASS || k == B200_FORM_DIFFUSION; }

// Classifies the registered forms (scalar / Taylor-Hood), sums their constant coefficients into the fused-kernel
// coefficient structs and records which blocks of the fused local matrix exist.
int analyze_forms(System *S)
{
  if(S->forms.empty() || S->nElm == 0 || S->nq == 0) {
    set_error("b200_finalize: mesh, quadrature, spaces and forms must be set first");
    return B200_ERR_ARG;
  }
  if(S->nv != S->dim + 1 || (S->dim != 2 && S->dim != 3)) {
    set_error("b200_finalize: only straight triangles / tetrahedra are supported");
    return B200_ERR_UNSUPP;
  }
  if(S->chns_active) {
    for(int bi = 0; bi < 2; ++bi)
      for(int bj = 0; bj < 2; ++bj) S->has_matrix_block[bi][bj] = false;
    return chns_analyze(S);
  }
  bool all_scalar = true, any_scalar = false;
  for(auto &f : S->forms) {
    if(is_scalar_kind(f.kind))
      any_scalar = true;
    else
      all_scalar = false;
  }
  if(any_scalar && !all_scalar) {
    set_error("b200_finalize: mixing scalar and vector weak forms in one system is not supported");
    return B200_ERR_UNSUPP;
  }
  S->th = THCoeffs();
  S->th_transient = THCoeffs();
  S->sc = ScalarCoeffs();
  S->sc_transient = ScalarCoeffs();
  S->d_source = nullptr;
  S->d_kcoef  = nullptr;
  for(int bi = 0; bi < 2; ++bi)
    for(int bj = 0; bj < 2; ++bj) S->has_matrix_block[bi][bj] = false;

  if(all_scalar) {
    S->plan = PLAN_SCALAR;
    S->su   = S->forms[0].su;
    S->sp   = -1;
    for(auto &f : S->forms) {
      if(f.su != S->su || S->spaces[f.su].nc != 1) {
        set_error("b200_finalize: scalar forms must share one scalar space");
        return B200_ERR_UNSUPP;
      }
      if(f.kind == B200_FORM_DIFFUSION) {
        if(f.d_coeff_table) {
          if(S->d_kcoef) {
            set_error("b200_finalize: at most one diffusion form with a tabulated coefficient");
            return B200_ERR_UNSUPP;
          }
          S->d_kcoef = f.d_coeff_table;
        }
        S->sc.k += f.coeff * f.param; // diffusivity = coeff x param (the reference form has one callback: pass param = 1)
        S->has_matrix_block[0][0] = true;
      } else if(f.kind == B200_FORM_TRANSIENT_MASS) {
        S->sc.c_mass += f.coeff;
        S->sc_transient.c_mass += f.coeff;
        S->has_matrix_block[0][0] = true;
      } else {
        if(S->d_source) {
          set_error("b200_finalize: at most one source form (pre-sum the tabulated sources on the host)");
          return B200_ERR_UNSUPP;
        }
        S->sc.c_src = 1.;
        S->d_source = f.d_source;
      }
    }
    S->M = S->spaces[S->su].nS;
  } else {
    S->plan = PLAN_TAYLOR_HOOD;
    S->su = S->sp = -1;
    for(auto &f : S->forms) {
      int u = f.su, p = f.sp;
      // MIXED_DIVERGENCE is declared on {p, u} (tests/withLinearSolver/navier_stokes.cpp:85)
      if(f.kind == B200_FORM_MIXED_DIVERGENCE) {
        u = f.sp;
        p = f.su;
      }
      if(u >= 0) {
        if(S->su >= 0 && S->su != u) {
          set_error("b200_finalize: vector forms must share one velocity space");
          return B200_ERR_UNSUPP;
        }
        S->su = u;
      }
      if(p >= 0) {
        if(S->sp >= 0 && S->sp != p) {
          set_error("b200_finalize: mixed forms must share one pressure space");
          return B200_ERR_UNSUPP;
        }
        S->sp = p;
      }
      switch(f.kind) {
        case B200_FORM_VECTOR_CONVECTIVE_ACCELERATION:
          S->th.c_conv += f.coeff;
          S->has_matrix_block[0][0] = true;
          break;
        case B200_FORM_DIV_NEWTONIAN_STRESS:
          S->th.c_sig += f.coeff;
          S->th.sig_mu += f.coeff * f.param;
          S->has_matrix_block[0][0] = S->has_matrix_block[0][1] = true;
          break;
        case B200_FORM_MIXED_DIVERGENCE:
          S->th.c_div += f.coeff;
          S->has_matrix_block[1][0] = true;
          break;
        case B200_FORM_VECTOR_DIFFUSION:
          S->th.diff_k += f.coeff * f.param;
          S->has_matrix_block[0][0] = true;
          break;
        case B200_FORM_MIXED_GRADIENT:
          S->th.c_gradp += f.coeff;
          S->has_matrix_block[0][1] = true;
          break;
        case B200_FORM_TRANSIENT_VECTOR_MASS:
          S->th.c_mass += f.coeff;
          S->th_transient.c_mass += f.coeff;
          S->has_matrix_block[0][0] = true;
          break;
        case B200_FORM_VECTOR_SOURCE:
          if(S->d_source) {
            set_error("b200_finalize: at most one source form (pre-sum the tabulated sources on the host)");
            return B200_ERR_UNSUPP;
          }
          S->th.c_src = 1.;
          S->d_source = f.d_source;
          break;
        default: set_error("b200_finalize: weak form kind " + std::to_string(f.kind) + " is not supported"); return B200_ERR_UNSUPP;
      }
    }
    if(S->su < 0 || S->sp < 0) {
      set_error("b200_finalize: Taylor-Hood plan needs a velocity and a pressure space");
      return B200_ERR_UNSUPP;
    }
    const Space &U = S->spaces[S->su], &P = S->spaces[S->sp];
    if(U.nc != S->dim || P.nc != 1) {
      set_error("b200_finalize: velocity space must have dim components, pressure one");
      return B200_ERR_UNSUPP;
    }
    const bool ok2 = S->dim == 2 && U.nS == 6 && P.nS == 3, ok3 = S->dim == 3 && U.nS == 10 && P.nS == 4;
    if(!ok2 && !ok3) {
      set_error("b200_finalize: fused Taylor-Hood kernel is built for P2/P1 only");
      return B200_ERR_UNSUPP;
    }
    S->M = U.nS * U.nc + P.nS;
  }
  return B200_OK;
}

// CSR slot of every local entry, for the scatter kernels (th_kernel / sc_kernel).  4 nElm M^2 bytes (21.6 GB at T3D(92)):
// built at finalize only when no write-once plan exists, otherwise on the first scatter-mode assembly.
static int build_slot_map(System *S)
{
  const Space &U  = S->spaces[S->su];
  const int    NU = U.nS * U.nc, NP = S->sp >= 0 ? S->spaces[S->sp].nS : 0;
  const int    M  = S->M;
  int          mask = 0;
  for(int bi = 0; bi < 2; ++bi)
    for(int bj = 0; bj < 2; ++bj)
      if(S->has_matrix_block[bi][bj]) mask |= 1 << (bi * 2 + bj);
  if(S->d_slot) cudaFree(S->d_slot);
  S->d_slot = nullptr;
  B200_CUDA(cudaMalloc(&S->d_slot, (size_t)S->nElm * M * M * sizeof(int32_t)));
  int *d_err;
  B200_CUDA(cudaMalloc(&d_err, sizeof(int)));
  B200_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), S->stream));
  slot_map_kernel<<<148 * 8, 256, 0, S->stream>>>(S->nElm, M, NU, U.d_adr, S->sp >= 0 ? S->spaces[S->sp].d_adr : nullptr, NP, S->d_ia,
                                                  S->d_ja, S->nInc, mask, S->d_slot, d_err);
  count_launch();
  int h_err = 0;
  B200_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, S->stream));
  B200_CUDA(cudaStreamSynchronize(S->stream));
  cudaFree(d_err);
  if(h_err) {
    cudaFree(S->d_slot);
    S->d_slot = nullptr;
    set_error("b200_finalize: a local (row, col) pair is missing from the CSR pattern");
    return B200_ERR_ARG;
  }
  log_stage("plan: slot map");
  return B200_OK;
}

int build_plan(System *S)
{
  if(S->d_ia == nullptr) {
    set_error("b200_finalize: set or build the CSR pattern first");
    return B200_ERR_ARG;
  }
  if(S->nnz >= (int64_t)2147483647) {
    set_error("b200_finalize: nnz >= 2^31 needs the 64-bit slot map (not built)");
    return B200_ERR_UNSUPP;
  }
  const int arc = analyze_forms(S);
  if(arc != B200_OK) return arc;

  if(S->plan == PLAN_CHNS) {
    gather_free(S);
    if(S->assembly_mode == B200_ASSEMBLY_GATHER) {
      set_error("b200_finalize: the gather assembly needs the fused Taylor-Hood system");
      return B200_ERR_UNSUPP;
    }
    return chns_build_plan(S);
  }
  // packed tables: w | LU | dLU | LP
  {
    const Space        &U = S->spaces[S->su];
    std::vector<double> tab(S->w);
    tab.insert(tab.end(), U.L.begin(), U.L.end());
    tab.insert(tab.end(), U.dL.begin(), U.dL.end());
    if(S->sp >= 0) tab.insert(tab.end(), S->spaces[S->sp].L.begin(), S->spaces[S->sp].L.end());
    if(S->d_tab) cudaFree(S->d_tab);
    S->tab_len = (int)tab.size();
    B200_CUDA(cudaMalloc(&S->d_tab, tab.size() * sizeof(double)));
    B200_CUDA(cudaMemcpyAsync(S->d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, S->stream));
    B200_CUDA(cudaStreamSynchronize(S->stream));
  }

  if(S->d_slot) cudaFree(S->d_slot);
  S->d_slot = nullptr;
  // row-owner gather plan where the problem qualifies
  gather_free(S);
  if(S->plan == PLAN_TAYLOR_HOOD && S->assembly_mode != B200_ASSEMBLY_SCATTER) {
    const int grc = build_gather_plan(S);
    if(grc != B200_OK && (grc != B200_ERR_UNSUPP || S->assembly_mode == B200_ASSEMBLY_GATHER)) return grc;
  } else if(S->assembly_mode == B200_ASSEMBLY_GATHER) {
    set_error("b200_finalize: the gather assembly needs the fused Taylor-Hood system");
    return B200_ERR_UNSUPP;
  }
  // the write-once plans validate the pattern themselves; without one the scatter kernels run and need their slot map now
  if(S->gather == nullptr) return build_slot_map(S);
  return B200_OK;
}

template <typename K, typename A> static int launch_colored_or_atomic(System *S, K kern_atomic, K kern_plain, A args, int EPB, int threads, size_t smem)
{
  if(smem > 48 * 1024) {
    B200_CUDA(cudaFuncSetAttribute(kern_atomic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_CUDA(cudaFuncSetAttribute(kern_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if(S->scatter_mode == B200_SCATTER_COLORED) {
    if(S->n_colors <= 0) {
      set_error("coloured scatter requested but b200_set_colors was not called");
      return B200_ERR_ARG;
    }
    for(int col = 0; col < S->n_colors; ++col) {
      args.elem_list  = S->d_color_elems;
      args.elem_begin = S->color_ptr[col];
      args.elem_end   = S->color_ptr[col + 1];
      const int64_t n = args.elem_end - args.elem_begin;
      if(n <= 0) continue;
      kern_plain<<<(unsigned)((n + EPB - 1) / EPB), threads, smem, S->stream>>>(args);
      count_launch();
    }
  } else {
    args.elem_list  = nullptr;
    args.elem_begin = 0;
    args.elem_end   = S->nElm;
    kern_atomic<<<(unsigned)((S->nElm + EPB - 1) / EPB), threads, smem, S->stream>>>(args);
    count_launch();
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

template <int DIM, int NS, int NP, int EPB, int CH> static int launch_th(System *S, int what, const THCoeffs &c)
{
  THArgs a;
  a.xyz    = S->d_xyz;
  a.conn   = S->d_conn;
  a.adrU   = S->spaces[S->su].d_adr;
  a.adrP   = S->spaces[S->sp].d_adr;
  a.sol    = S->d_sol;
  a.soldot = S->have_soldot ? S->d_soldot : nullptr;
  a.source = (c.c_src != 0.) ? S->d_source : nullptr;
  a.tab    = S->d_tab;
  a.slot   = S->d_slot;
  a.val    = S->d_val;
  a.rhs    = S->d_rhs;
  a.nInc   = S->nInc;
  a.nq     = S->nq;
  a.c      = c;
  a.c0     = S->c0;
  a.elem_list = nullptr;
  a.elem_begin = 0;
  a.elem_end = S->nElm;
  const size_t smem = th_smem_bytes<DIM, NS, NP, EPB, CH>(S->nq);
  const int    thr  = EPB * (NS * DIM + NP);
  if(what == 3) return launch_colored_or_atomic(S, th_kernel<DIM, NS, NP, EPB, CH, true, true, true>, th_kernel<DIM, NS, NP, EPB, CH, true, true, false>, a, EPB, thr, smem);
  if(what == 2) return launch_colored_or_atomic(S, th_kernel<DIM, NS, NP, EPB, CH, true, false, true>, th_kernel<DIM, NS, NP, EPB, CH, true, false, false>, a, EPB, thr, smem);
  return launch_colored_or_atomic(S, th_kernel<DIM, NS, NP, EPB, CH, false, true, true>, th_kernel<DIM, NS, NP, EPB, CH, false, true, false>, a, EPB, thr, smem);
}

template <int DIM, int NS, int EPB, int CH> static int launch_sc(System *S, int what, const ScalarCoeffs &c)
{
  SCArgs a;
  a.xyz    = S->d_xyz;
  a.conn   = S->d_conn;
  a.adr    = S->spaces[S->su].d_adr;
  a.sol    = S->d_sol;
  a.soldot = S->have_soldot ? S->d_soldot : nullptr;
  a.source = (c.c_src != 0.) ? S->d_source : nullptr;
  a.kcoef  = S->d_kcoef;
  a.tab    = S->d_tab;
  a.slot   = S->d_slot;
  a.val    = S->d_val;
  a.rhs    = S->d_rhs;
  a.nInc   = S->nInc;
  a.nq     = S->nq;
  a.c      = c;
  a.c0     = S->c0;
  a.elem_list = nullptr;
  a.elem_begin = 0;
  a.elem_end = S->nElm;
  const size_t smem = sc_smem_bytes<DIM, NS, EPB, CH>(S->nq);
  const int    thr  = EPB * NS;
  if(what == 3) return launch_colored_or_atomic(S, sc_kernel<DIM, NS, EPB, CH, true, true, true>, sc_kernel<DIM, NS, EPB, CH, true, true, false>, a, EPB, thr, smem);
  if(what == 2) return launch_colored_or_atomic(S, sc_kernel<DIM, NS, EPB, CH, true, false, true>, sc_kernel<DIM, NS, EPB, CH, true, false, false>, a, EPB, thr, smem);
  return launch_colored_or_atomic(S, sc_kernel<DIM, NS, EPB, CH, false, true, true>, sc_kernel<DIM, NS, EPB, CH, false, true, false>, a, EPB, thr, smem);
}

// what: bit 0 residual, bit 1 matrix.  only_transient: matrix restricted to the transient forms
// (assembleOnlyTransientMatrices, src/feLinearSystemMklPardiso.cpp:528); the residual always uses every form.
int launch_assemble(System *S, int what, int only_transient)
{
  if(S->plan == PLAN_NONE) {
    set_error("b200_assemble: call b200_finalize first");
    return B200_ERR_ARG;
  }
  if(what < 1 || what > 3) {
    set_error("b200_assemble: what must be 1, 2 or 3");
    return B200_ERR_ARG;
  }
  if(S->gather != nullptr && S->assembly_mode != B200_ASSEMBLY_SCATTER) {
    // every row is overwritten: a pending setToZero of the parts written here is dropped, the rest materialised
    if(only_transient && (what & 2)) {
      int rc = B200_OK;
      if(what & 1) {
        S->pending_zero &= ~1;
        rc = launch_gather(S, 1, S->th);
      }
      if(rc != B200_OK) return rc;
      S->pending_zero &= ~2;
      return launch_gather(S, 2, S->th_transient);
    }
    S->pending_zero &= ~what;
    return launch_gather(S, what, S->th);
  }
  {
    const int zrc = flush_zero(S, 3);
    if(zrc != B200_OK) return zrc;
  }
  if(S->plan != PLAN_CHNS && S->d_slot == nullptr) {
    const int src = build_slot_map(S);
    if(src != B200_OK) return src;
  }
  if(S->plan == PLAN_CHNS) {
    // the monolithic form is not a "transient matrix" form (feBilinearForm::isTransientMatrix is false for it)
    if(only_transient) what &= ~2;
    return what ? chns_launch(S, what) : B200_OK;
  }
  if(only_transient && (what & 2)) {
    // transient-only matrix and full residual cannot share coefficients: two passes
    int rc = B200_OK;
    if(what & 1) rc = launch_assemble(S, 1, 0);
    if(rc != B200_OK) return rc;
    if(S->plan == PLAN_TAYLOR_HOOD) {
      if(S->dim == 2) return launch_th<2, 6, 3, 16, 8>(S, 2, S->th_transient);
      return launch_th<3, 10, 4, 8, 8>(S, 2, S->th_transient);
    }
    const int nS = S->spaces[S->su].nS;
    if(S->dim == 2 && nS == 6) return launch_sc<2, 6, 32, 8>(S, 2, S->sc_transient);
    if(S->dim == 2 && nS == 3) return launch_sc<2, 3, 64, 8>(S, 2, S->sc_transient);
    if(S->dim == 3 && nS == 10) return launch_sc<3, 10, 16, 8>(S, 2, S->sc_transient);
    if(S->dim == 3 && nS == 4) return launch_sc<3, 4, 48, 8>(S, 2, S->sc_transient);
    set_error("b200_assemble: unsupported scalar space");
    return B200_ERR_UNSUPP;
  }
  if(S->plan == PLAN_TAYLOR_HOOD) {
    if(S->dim == 2) return launch_th<2, 6, 3, 16, 8>(S, what, S->th);
    return launch_th<3, 10, 4, 8, 8>(S, what, S->th);
  }
  const int nS = S->spaces[S->su].nS;
  if(S->dim == 2 && nS == 6) return launch_sc<2, 6, 32, 8>(S, what, S->sc);
  if(S->dim == 2 && nS == 3) return launch_sc<2, 3, 64, 8>(S, what, S->sc);
  if(S->dim == 3 && nS == 10) return launch_sc<3, 10, 16, 8>(S, what, S->sc);
  if(S->dim == 3 && nS == 4) return launch_sc<3, 4, 48, 8>(S, what, S->sc);
  set_error("b200_assemble: unsupported scalar space (P1/P2 Lagrange on triangles/tetrahedra)");
  return B200_ERR_UNSUPP;
}

} // namespace b200
