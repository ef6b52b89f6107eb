// Monolithic Cahn-Hilliard Navier-Stokes weak form with finite-difference Jacobian.
//
// Reference: CHNS_Abels<2>::computeBe (src/feSysElmCHNS.cpp:66-273, MODEL 0) and CHNS_MassAveraged<2>::computeBe
// (src/feSysElmCHNS.cpp:347-602, MODEL 1: mass-averaged velocity, pressure-dependent diffusive flux, time-averaged double
// well with phi at the previous time step from b200_set_solution_n) and CHNS_Khanwale<2>::computeBe
// (src/feSysElmCHNS.cpp:678-938, MODEL 2: non-dimensional, every field averaged with the previous time step) evaluated N+1
// times per element by
// feBilinearForm::computeMatrixFiniteDifference (src/feBilinearForm.cpp:388-428), one weak form on the fields
// [U (vector P2), P (P1), Phi, Mu (P1 or P2)] (layout src/feSysElmCHNS.cpp:13-14), property laws of CHNS_Solver
// (src/CHNS_Solver.cpp:124-235) as enums instead of host callbacks.
//
// GPU organisation: ONE WARP PER ELEMENT.  Lane j < N carries the residual of the state perturbed in local column j,
// lane N the unperturbed residual R0, so the N+1 residual evaluations of the reference run side by side and the
// Jacobian column is -(Rh - R0)/delta after one shuffle broadcast of R0.  The interpolated fields at a quadrature
// node are linear in the local DOFs: they are computed once per element (lane k <-> node k) and every lane adds
// delta x (its own basis function) to them.  The quadrature tables live in shared memory.  Local entries are added
// to the CSR arrays with red.global.add.f64 after a binary search of the column in the row (the reference scans the
// row, src/feLinearSystemMklPardiso.cpp:648-658).
#include <cfloat>
#include <cmath>

#include "device_common.cuh"
#include "system.h"

namespace b200 {

struct ChnsArgs {
  int64_t         nElm;
  const double   *xyz;
  const int32_t  *conn, *adr; // adr: [nElm][M] local DOFs in field order U | P | Phi | Mu
  const uint16_t *off;        // [nElm][M][M] row-local CSR offset of local entry (i, j), 0xFFFF = not assembled
  const double   *sol, *soldot, *soln, *tab;
  const int64_t  *ia;
  const int32_t  *ja;
  double         *val, *rhs;
  int64_t         nInc;
  int             nq, ntab, what;
  double          c0, h0, dt;
  b200_chns_params prm;
};

constexpr int CHNS_NSU = 6, CHNS_NSP = 3, CHNS_WPB = 4;
// fields per quadrature node: 16 for CHNS_Abels; + grad p (2) + phi at the previous time step for CHNS_MassAveraged;
// + the 13 fields of the previous time step {u, p, phi, mu, grad u, grad phi, grad mu} for CHNS_Khanwale
__host__ __device__ constexpr int chns_nfld(int model) { return model == 0 ? 16 : (model == 1 ? 20 : 30); }

template <int NSF> struct ChnsT {
  static constexpr int NU = CHNS_NSU * 2, M = NU + CHNS_NSP + 2 * NSF;
  // table layout (doubles): w[nq] | LU[nq][6] | dLU[nq][6][2] | LP[nq][3] | LF[nq][NSF] | dLF[nq][NSF][2] | dLP[nq][3][2]
  __host__ __device__ static int o_lu(int nq) { return nq; }
  __host__ __device__ static int o_dlu(int nq) { return nq + nq * CHNS_NSU; }
  __host__ __device__ static int o_lp(int nq) { return nq + nq * CHNS_NSU * 3; }
  __host__ __device__ static int o_lf(int nq) { return nq + nq * CHNS_NSU * 3 + nq * CHNS_NSP; }
  __host__ __device__ static int o_dlf(int nq) { return nq + nq * CHNS_NSU * 3 + nq * CHNS_NSP + nq * NSF; }
  __host__ __device__ static int o_dlp(int nq) { return nq + nq * CHNS_NSU * 3 + nq * CHNS_NSP + nq * NSF * 3; }
  __host__ __device__ static int len(int nq) { return o_dlp(nq) + nq * CHNS_NSP * 2; }
};

template <int NSF, int MODEL> __global__ void __launch_bounds__(CHNS_WPB * 32, MODEL == 0 ? 4 : 3) chns_kernel(const ChnsArgs a)
{
  using T = ChnsT<NSF>;
  constexpr int M = T::M, NU = T::NU, NSU = CHNS_NSU, NSP = CHNS_NSP, CHNS_NFLD = chns_nfld(MODEL);
  static_assert(M + 1 <= 32, "one lane per local column plus the base lane");
  extern __shared__ double sm[];
  double *s_tab = sm;                                    // ntab
  double *s_loc = s_tab + a.ntab;                        // [WPB][M]
  double *s_dot = s_loc + CHNS_WPB * M;                  // [WPB][M]
  double *s_fld = s_dot + CHNS_WPB * M;                  // [WPB][nq][NFLD]
  double *s_ln  = s_fld + (size_t)CHNS_WPB * a.nq * CHNS_NFLD; // [WPB][32] local DOFs at the previous time step (MODEL 1: Phi only)
  int32_t *s_adr = reinterpret_cast<int32_t *>(s_ln + CHNS_WPB * 32); // [WPB][M]

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nq = a.nq;
  for(int i = tid; i < a.ntab; i += CHNS_WPB * 32) s_tab[i] = a.tab[i];
  __syncthreads();
  const int64_t e = blockIdx.x * (int64_t)CHNS_WPB + wid;
  if(e >= a.nElm) return;
  const double *w = s_tab, *LU = s_tab + T::o_lu(nq), *dLU = s_tab + T::o_dlu(nq), *LP = s_tab + T::o_lp(nq), *LF = s_tab + T::o_lf(nq),
               *dLF = s_tab + T::o_dlf(nq), *dLP = s_tab + T::o_dlp(nq);
  double *locn = s_ln + wid * 32;
  double  *loc = s_loc + wid * M, *dot = s_dot + wid * M, *fld = s_fld + (size_t)wid * nq * CHNS_NFLD;
  int32_t *adr = s_adr + wid * M;
  for(int i = lane; i < M; i += 32) {
    const int32_t d = a.adr[e * M + i];
    adr[i] = d;
    loc[i] = a.sol[d];
    dot[i] = a.soldot ? a.soldot[d] : 0.;
    // feBilinearForm::initialize fills _solAtTimeN from the global solAtTimeN (src/feBilinearForm.cpp:347); without
    // b200_set_solution_n it is the current solution, as in a stationary solve
    if(MODEL == 1 && i >= NU + NSP && i < NU + NSP + NSF) locn[i - NU - NSP] = a.soln ? a.soln[d] : a.sol[d];
    if(MODEL == 2) locn[i] = a.soln ? a.soln[d] : a.sol[d];
  }
  double G[4], J;
  {
    int32_t vtx[3];
#pragma unroll
    for(int v = 0; v < 3; ++v) vtx[v] = a.conn[e * 3 + v];
    element_geometry<2>(a.xyz, vtx, G, &J);
  }
  __syncwarp();
  const double *Ul = loc, *Pl = loc + NU, *Fl = loc + NU + NSP, *Ml = loc + NU + NSP + NSF;
  const double *Ud = dot, *Fd = dot + NU + NSP;
  // ---- unperturbed fields at the quadrature nodes (lane k <-> node k):
  // fld = {u0,u1,p,phi,mu,dudt0,dudt1,dphidt, gu[m][n] (4, d_m u_n, src/feSpace.cpp:1391-1394), gphi[2], gmu[2]}
  for(int k = lane; k < nq; k += 32) {
    double f[CHNS_NFLD];
#pragma unroll
    for(int i = 0; i < CHNS_NFLD; ++i) f[i] = 0.;
    for(int b = 0; b < NSU; ++b) {
      const double L = LU[k * NSU + b], dr = dLU[(k * NSU + b) * 2], ds = dLU[(k * NSU + b) * 2 + 1];
      const double gx = dr * G[0] + ds * G[2], gy = dr * G[1] + ds * G[3];
      f[0] += L * Ul[b * 2];
      f[1] += L * Ul[b * 2 + 1];
      f[5] += L * Ud[b * 2];
      f[6] += L * Ud[b * 2 + 1];
      f[8] += gx * Ul[b * 2];      // d_x u_0
      f[9] += gx * Ul[b * 2 + 1];  // d_x u_1
      f[10] += gy * Ul[b * 2];     // d_y u_0
      f[11] += gy * Ul[b * 2 + 1]; // d_y u_1
    }
    for(int q = 0; q < NSP; ++q) {
      f[2] += LP[k * NSP + q] * Pl[q];
      if(MODEL == 1) {
        const double dr = dLP[(k * NSP + q) * 2], ds = dLP[(k * NSP + q) * 2 + 1];
        f[16] += (dr * G[0] + ds * G[2]) * Pl[q];
        f[17] += (dr * G[1] + ds * G[3]) * Pl[q];
      }
    }
    if(MODEL == 1)
      for(int i = 0; i < NSF; ++i) f[18] += LF[k * NSF + i] * locn[i];
    for(int i = 0; i < NSF; ++i) {
      const double L = LF[k * NSF + i], dr = dLF[(k * NSF + i) * 2], ds = dLF[(k * NSF + i) * 2 + 1];
      const double gx = dr * G[0] + ds * G[2], gy = dr * G[1] + ds * G[3];
      f[3] += L * Fl[i];
      f[7] += L * Fd[i];
      f[12] += gx * Fl[i];
      f[13] += gy * Fl[i];
      f[4] += L * Ml[i];
      f[14] += gx * Ml[i];
      f[15] += gy * Ml[i];
    }
    if(MODEL == 2) {
      // previous time step: f[16..28] = {u_n (2), p_n, phi_n, mu_n, grad u_n (4), grad phi_n (2), grad mu_n (2)}
      const double *Un = locn, *Pn = locn + NU, *Fn = locn + NU + NSP, *Mn = locn + NU + NSP + NSF;
      for(int b = 0; b < NSU; ++b) {
        const double L = LU[k * NSU + b], dr = dLU[(k * NSU + b) * 2], ds = dLU[(k * NSU + b) * 2 + 1];
        const double gx = dr * G[0] + ds * G[2], gy = dr * G[1] + ds * G[3];
        f[16] += L * Un[b * 2];
        f[17] += L * Un[b * 2 + 1];
        f[21] += gx * Un[b * 2];
        f[22] += gx * Un[b * 2 + 1];
        f[23] += gy * Un[b * 2];
        f[24] += gy * Un[b * 2 + 1];
      }
      for(int q = 0; q < NSP; ++q) f[18] += LP[k * NSP + q] * Pn[q];
      for(int i = 0; i < NSF; ++i) {
        const double L = LF[k * NSF + i], dr = dLF[(k * NSF + i) * 2], ds = dLF[(k * NSF + i) * 2 + 1];
        const double gx = dr * G[0] + ds * G[2], gy = dr * G[1] + ds * G[3];
        f[19] += L * Fn[i];
        f[25] += gx * Fn[i];
        f[26] += gy * Fn[i];
        f[20] += L * Mn[i];
        f[27] += gx * Mn[i];
        f[28] += gy * Mn[i];
      }
    }
#pragma unroll
    for(int i = 0; i < CHNS_NFLD; ++i) fld[k * CHNS_NFLD + i] = f[i];
  }
  __syncwarp();

  // ---- my perturbation: delta = h0 max(|u_j|, 1), solDot += delta c0 (src/feBilinearForm.cpp:404-410)
  const int  j      = lane;
  const bool column = j < M;
  const bool active = (column && (a.what & 2)) || j == M;
  double     delta = 0.;
  if(column) delta = a.h0 * fmax(fabs(loc[j]), 1.);
  double du0 = 0., du1 = 0., dp = 0., dphi = 0., dmu = 0.;
  int    aU = 0, qP = 0, iF = 0, iM = 0;
  if(j < NU) {
    aU = j >> 1;
    if(j & 1)
      du1 = delta;
    else
      du0 = delta;
  } else if(j < NU + NSP) {
    qP = j - NU;
    dp = delta;
  } else if(j < NU + NSP + NSF) {
    iF   = j - NU - NSP;
    dphi = delta;
  } else if(j < M) {
    iM  = j - NU - NSP - NSF;
    dmu = delta;
  }
  double R[M];
#pragma unroll
  for(int i = 0; i < M; ++i) R[i] = 0.;
  const b200_chns_params pr  = a.prm;
  const double           lam = 3. / (2. * sqrt(2.)) * pr.surface_tension * pr.epsilon; // src/feSysElm.h:1338
  const double           dw  = lam / (pr.epsilon * pr.epsilon);
  const double           drho = (pr.rho_a - pr.rho_b) * 0.5;

  if(active) {
    for(int k = 0; k < nq; ++k) {
      const double jw = J * w[k];
      const double *f = fld + k * CHNS_NFLD;
      double u0 = f[0], u1 = f[1], p = f[2], phi = f[3], mu = f[4], dt0 = f[5], dt1 = f[6], dphidt = f[7];
      double gu00 = f[8], gu01 = f[9], gu10 = f[10], gu11 = f[11], gp0 = f[12], gp1 = f[13], gm0 = f[14], gm1 = f[15];
      double gpr0 = 0., gpr1 = 0., phin = 0.; // grad p and phi at the previous time step (MODEL 1)
      if(MODEL == 1) {
        gpr0 = f[16];
        gpr1 = f[17];
        phin = f[18];
      }
      (void)gpr0, (void)gpr1, (void)phin;
      // test functions and their physical gradients at this node
      double lu[NSU], gux[NSU], guy[NSU], lf[NSF], gfx[NSF], gfy[NSF];
#pragma unroll
      for(int b = 0; b < NSU; ++b) {
        const double dr = dLU[(k * NSU + b) * 2], ds = dLU[(k * NSU + b) * 2 + 1];
        lu[b]  = LU[k * NSU + b];
        gux[b] = dr * G[0] + ds * G[2];
        guy[b] = dr * G[1] + ds * G[3];
      }
#pragma unroll
      for(int i = 0; i < NSF; ++i) {
        const double dr = dLF[(k * NSF + i) * 2], ds = dLF[(k * NSF + i) * 2 + 1];
        lf[i]  = LF[k * NSF + i];
        gfx[i] = dr * G[0] + ds * G[2];
        gfy[i] = dr * G[1] + ds * G[3];
      }
      // perturbed fields: linear in the local DOFs
      {
        const double L = LU[k * NSU + aU], dr = dLU[(k * NSU + aU) * 2], ds = dLU[(k * NSU + aU) * 2 + 1];
        const double gx = dr * G[0] + ds * G[2], gy = dr * G[1] + ds * G[3];
        u0 += du0 * L;
        u1 += du1 * L;
        dt0 += du0 * a.c0 * L;
        dt1 += du1 * a.c0 * L;
        gu00 += du0 * gx;
        gu01 += du1 * gx;
        gu10 += du0 * gy;
        gu11 += du1 * gy;
        p += dp * LP[k * NSP + qP];
        if(MODEL == 1) {
          const double pr_ = dLP[(k * NSP + qP) * 2], ps_ = dLP[(k * NSP + qP) * 2 + 1];
          gpr0 += dp * (pr_ * G[0] + ps_ * G[2]);
          gpr1 += dp * (pr_ * G[1] + ps_ * G[3]);
        }
        const double Lf = LF[k * NSF + iF], fr = dLF[(k * NSF + iF) * 2], fs = dLF[(k * NSF + iF) * 2 + 1];
        phi += dphi * Lf;
        dphidt += dphi * a.c0 * Lf;
        gp0 += dphi * (fr * G[0] + fs * G[2]);
        gp1 += dphi * (fr * G[1] + fs * G[3]);
        const double Lm = LF[k * NSF + iM], mr = dLF[(k * NSF + iM) * 2], ms = dLF[(k * NSF + iM) * 2 + 1];
        mu += dmu * Lm;
        gm0 += dmu * (mr * G[0] + ms * G[2]);
        gm1 += dmu * (mr * G[1] + ms * G[3]);
      }
      // property laws (src/CHNS_Solver.cpp:124-235)
      const double pc  = pr.limiter ? fmax(-1., fmin(1., phi)) : phi;
      const double rho = drho * pc + (pr.rho_a + pr.rho_b) * 0.5;
      const double eta = (pr.visc_a - pr.visc_b) * 0.5 * pc + (pr.visc_a + pr.visc_b) * 0.5;
      const double Mob = pr.degenerate_mobility ? pr.mobility * fabs(1. - phi * phi) : pr.mobility;
      const double divu = gu00 + gu11;
      // Every block of the residual has the form  -jw (c phi_i + X d_x phi_i + Y d_y phi_i): the formulation only decides
      // the coefficients, which carry the quadrature weight so that each test function costs three FMAs.
      double a0, a1, X0, Y0, X1, Y1; // momentum rows (components 0, 1) of velocity test function phi_b
      double cP, PX = 0., PY = 0.;   // continuity rows
      double cF, FX, FY, cM, MX, MY; // tracer and potential rows
      if(MODEL == 0) {
        // CHNS_Abels, src/feSysElmCHNS.cpp:158-271
        const double ugu0 = u0 * gu00 + u1 * gu10, ugu1 = u0 * gu01 + u1 * gu11;
        const double gmgu0 = gm0 * gu00 + gm1 * gu10, gmgu1 = gm0 * gu01 + gm1 * gu11;
        const double S00 = gu00 + gu00, S01 = gu01 + gu10, S11 = gu11 + gu11;
        a0 = jw * (rho * (dt0 + ugu0 - pr.force[0]) - drho * Mob * gmgu0 + phi * gm0 + pr.source_u[0]);
        a1 = jw * (rho * (dt1 + ugu1 - pr.force[1]) - drho * Mob * gmgu1 + phi * gm1 + pr.source_u[1]);
        const double je = jw * eta;
        X0 = je * S00 - jw * p;
        Y0 = je * S01;
        X1 = Y0;
        Y1 = je * S11 - jw * p;
        cP = jw * (divu + pr.source_p);
        cF = jw * (dphidt + u0 * gp0 + u1 * gp1 + pr.source_phi);
        FX = jw * Mob * gm0;
        FY = jw * Mob * gm1;
        cM = jw * (mu - dw * phi * (phi * phi - 1.) + pr.source_mu);
        MX = -jw * lam * gp0;
        MY = -jw * lam * gp1;
      } else if(MODEL == 1) {
        // CHNS_MassAveraged, src/feSysElmCHNS.cpp:446-600: div(rho u), time-averaged double well (Simpson in time)
        const double alpha = pr.mass_alpha;
        const double beta  = 3. / (2. * sqrt(2.)) * pr.surface_tension / pr.epsilon; // src/feSysElm.h:1425
        const double ugu0 = u0 * gu00 + u1 * gu10, ugu1 = u0 * gu01 + u1 * gu11;
        const double S00 = gu00 + gu00, S01 = gu01 + gu10, S11 = gu11 + gu11;
        const double divRhoU = rho * divu + drho * (u0 * gp0 + u1 * gp1);
        const double pavg = 0.5 * (phi + phin);
        const double well = (phi * (phi * phi - 1.) + 4. * pavg * (pavg * pavg - 1.) + phin * (phin * phin - 1.)) * beta / 6.;
        const double mc = 0.5 * (drho * dphidt + divRhoU);
        a0 = jw * (rho * (dt0 + ugu0 - pr.force[0]) + mc * u0 + phi * gm0 + pr.source_u[0]);
        a1 = jw * (rho * (dt1 + ugu1 - pr.force[1]) + mc * u1 + phi * gm1 + pr.source_u[1]);
        const double je = jw * eta, jpe = jw * (p + eta * divu); // - p div(phi_i) - eta (2/dim) div u div(phi_i), dim = 2
        X0 = je * S00 - jpe;
        Y0 = je * S01;
        X1 = Y0;
        Y1 = je * S11 - jpe;
        const double fx = Mob * (gm0 + alpha * gpr0), fy = Mob * (gm1 + alpha * gpr1);
        cP = jw * (divu + pr.source_p);
        PX = jw * alpha * fx;
        PY = jw * alpha * fy;
        cF = jw * (dphidt + pr.source_phi);
        FX = jw * (fx - phi * u0);
        FY = jw * (fy - phi * u1);
        cM = jw * (mu - well + pr.source_mu);
        MX = -jw * lam * gp0;
        MY = -jw * lam * gp1;
      } else {
        // CHNS_Khanwale, src/feSysElmCHNS.cpp:717-936: every field averaged with the previous time step; the volume
        // force of this class is the constant (0, -1) (:623)
        const double *kw = pr.khanwale; // Re, Pe, Cn, We, Fr, rhoA, rhoB
        const double Re = kw[0], Pe = kw[1], Cn = kw[2], We = kw[3], Fr = kw[4], rA = kw[5], rB = kw[6];
        const double ua0 = 0.5 * (u0 + f[16]), ua1 = 0.5 * (u1 + f[17]), pa = 0.5 * (p + f[18]), fa = 0.5 * (phi + f[19]),
                     ma = 0.5 * (mu + f[20]);
        const double ga00 = 0.5 * (gu00 + f[21]), ga01 = 0.5 * (gu01 + f[22]), ga10 = 0.5 * (gu10 + f[23]), ga11 = 0.5 * (gu11 + f[24]);
        const double gfa0 = 0.5 * (gp0 + f[25]), gfa1 = 0.5 * (gp1 + f[26]), gma0 = 0.5 * (gm0 + f[27]), gma1 = 0.5 * (gm1 + f[28]);
        auto lin = [&](double x, double va, double vb) {
          const double cl = pr.limiter ? fmax(-1., fmin(1., x)) : x;
          return (va - vb) * 0.5 * cl + (va + vb) * 0.5;
        };
        const double rho_n = lin(f[19], pr.rho_a, pr.rho_b), rho_a = lin(fa, pr.rho_a, pr.rho_b), eta_a = lin(fa, pr.visc_a, pr.visc_b);
        const double well = fa * (fa * fa - 1.);
        const double divua = ga00 + ga11;
        const double ug0 = ua0 * ga00 + ua1 * ga10, ug1 = ua0 * ga01 + ua1 * ga11;
        const double T00 = ga00 + ga00, T01 = ga01 + ga10, T11 = ga11 + ga11;
        const double jc = (rB - rA) / (2. * rA * Cn), j0 = jc * gma0, j1 = jc * gma1;
        const double jg0 = j0 * ga00 + j1 * ga10, jg1 = j0 * ga01 + j1 * ga11;
        const double div_rau = drho * (gfa0 * ua0 + gfa1 * ua1) + rho_a * divua;
        a0 = jw * (rho_a * (dt0 + ug0) + jg0 / Pe + pr.source_u[0]);
        a1 = jw * (rho_a * (dt1 + ug1) + jg1 / Pe + rho_a / Fr + pr.source_u[1]);
        const double jk = jw * Cn / We, jp = jw * pa / We, je = jw * eta_a / Re;
        X0 = je * T00 - jp - jk * gfa0 * gfa0;
        Y0 = je * T01 - jk * gfa1 * gfa0;
        X1 = je * T01 - jk * gfa0 * gfa1;
        Y1 = je * T11 - jp - jk * gfa1 * gfa1;
        cP = jw * (divu + (rho - rho_n) / a.dt + div_rau + pr.source_p);
        PX = -jw * j0 / Pe;
        PY = -jw * j1 / Pe;
        const double dm = 1. / (Pe * Cn);
        cF = jw * (dphidt + pr.source_phi);
        FX = jw * (dm * gma0 - fa * ua0);
        FY = jw * (dm * gma1 - fa * ua1);
        cM = jw * (ma - well + pr.source_mu);
        MX = -jw * Cn * Cn * gfa0;
        MY = -jw * Cn * Cn * gfa1;
        (void)eta, (void)Mob;
      }
#pragma unroll
      for(int b = 0; b < NSU; ++b) {
        R[2 * b]     = fma(-a0, lu[b], fma(-X0, gux[b], fma(-Y0, guy[b], R[2 * b])));
        R[2 * b + 1] = fma(-a1, lu[b], fma(-X1, gux[b], fma(-Y1, guy[b], R[2 * b + 1])));
      }
#pragma unroll
      for(int q = 0; q < NSP; ++q) {
        double r = fma(-cP, LP[k * NSP + q], R[NU + q]);
        if(MODEL != 0) {
          const double dr = dLP[(k * NSP + q) * 2], ds = dLP[(k * NSP + q) * 2 + 1];
          r = fma(-PX, dr * G[0] + ds * G[2], fma(-PY, dr * G[1] + ds * G[3], r));
        }
        R[NU + q] = r;
      }
#pragma unroll
      for(int i = 0; i < NSF; ++i) {
        R[NU + NSP + i]       = fma(-cF, lf[i], fma(-FX, gfx[i], fma(-FY, gfy[i], R[NU + NSP + i])));
        R[NU + NSP + NSF + i] = fma(-cM, lf[i], fma(-MX, gfx[i], fma(-MY, gfy[i], R[NU + NSP + NSF + i])));
      }
    }
  }
  // ---- Jacobian column: Ae[i][j] = -(Rh[i] - R0[i]) / delta (src/feBilinearForm.cpp:419-422); residual = R0
  const double inv = column ? 1. / delta : 0.;
#pragma unroll
  for(int i = 0; i < M; ++i) {
    const double  r0 = __shfl_sync(0xffffffffu, R[i], M);
    const int32_t I  = adr[i];
    if(I >= a.nInc) continue;
    if(j == M && (a.what & 1)) atomicAdd(a.rhs + I, r0);
    if(column && (a.what & 2)) {
      // the reference scans the row for the column (src/feLinearSystemMklPardiso.cpp:648-658); here the row-local offset of
      // every local entry was found once at plan time (chns_offsets_kernel)
      const uint32_t o = a.off[((int64_t)e * M + i) * M + j];
      if(o != 0xFFFFu) atomicAdd(a.val + a.ia[I] + o, -(R[i] - r0) * inv);
    }
  }
}

// row-local CSR offsets of the M x M local entries of every element (binary search of the column in the row, once)
__global__ void chns_offsets_kernel(int64_t nElm, int M, const int32_t *adr, const int64_t *ia, const int32_t *ja, int64_t nInc, uint16_t *off,
                                    int *err)
{
  const int64_t tot = nElm * M * (int64_t)M;
  for(int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx / (M * M);
    const int     r = (int)(idx - e * M * M), i = r / M, j = r - i * M;
    const int32_t I = adr[e * M + i], Jc = adr[e * M + j];
    uint16_t      o = 0xFFFF;
    if(I < nInc && Jc < nInc) {
      int64_t lo = ia[I], hi = ia[I + 1] - 1;
      const int64_t beg = lo, end = hi + 1;
      while(lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if(ja[mid] < Jc)
          lo = mid + 1;
        else
          hi = mid;
      }
      if(lo < end && ja[lo] == Jc && lo - beg < 0xFFFF)
        o = (uint16_t)(lo - beg);
      else
        atomicExch(err, 1);
    }
    off[idx] = o;
  }
}

// ----------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------
__global__ void chns_concat_adr_kernel(int64_t nElm, int M, int n0, int n1, int n2, int n3, const int32_t *a0, const int32_t *a1, const int32_t *a2,
                                       const int32_t *a3, int32_t *out)
{
  const int64_t tot = nElm * M;
  for(int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx / M;
    int           i = (int)(idx - e * M);
    int32_t       v;
    if(i < n0)
      v = a0[e * n0 + i];
    else if((i -= n0) < n1)
      v = a1[e * n1 + i];
    else if((i -= n1) < n2)
      v = a2[e * n2 + i];
    else
      v = a3[e * n3 + (i - n2)];
    out[idx] = v;
  }
}

void chns_free(System *S)
{
  cudaFree(S->chns_adr);
  cudaFree(S->chns_tab);
  cudaFree(S->chns_off);
  S->chns_adr = nullptr;
  S->chns_tab = nullptr;
  S->chns_off = nullptr;
}

// Validates the CHNS registration and builds the concatenated element->DOF table (also used by the pattern builder).
int chns_analyze(System *S)
{
  if(S->dim != 2) {
    set_error("CHNS weak forms exist for dim = 2 only (src/feSysElmCHNS.cpp:275)");
    return B200_ERR_UNSUPP;
  }
  if(S->forms.size() != 1) {
    set_error("the CHNS weak form is monolithic: it must be the only form of the system");
    return B200_ERR_UNSUPP;
  }
  const Space &U = S->spaces[S->chns_space[0]], &P = S->spaces[S->chns_space[1]], &F = S->spaces[S->chns_space[2]],
              &Mu = S->spaces[S->chns_space[3]];
  if(U.nS != CHNS_NSU || U.nc != 2 || P.nS != CHNS_NSP || P.nc != 1 || F.nc != 1 || Mu.nc != 1 || F.nS != Mu.nS || (F.nS != 3 && F.nS != 6)) {
    set_error("CHNS kernel: U must be vector P2, P scalar P1, Phi and Mu the same scalar P1 or P2 space");
    return B200_ERR_UNSUPP;
  }
  S->plan = PLAN_CHNS;
  S->su   = S->chns_space[0];
  S->sp   = S->chns_space[1];
  S->M    = U.nS * 2 + P.nS + 2 * F.nS;
  S->has_matrix_block[0][0] = true;
  if(S->chns_adr == nullptr) {
    B200_CUDA(cudaMalloc(&S->chns_adr, (size_t)S->nElm * S->M * sizeof(int32_t)));
    chns_concat_adr_kernel<<<148 * 8, 256, 0, S->stream>>>(S->nElm, S->M, U.nS * 2, P.nS, F.nS, Mu.nS, U.d_adr, P.d_adr, F.d_adr, Mu.d_adr,
                                                          S->chns_adr);
    count_launch();
    B200_CUDA(cudaGetLastError());
  }
  return B200_OK;
}

int chns_build_plan(System *S)
{
  const Space &U = S->spaces[S->chns_space[0]], &P = S->spaces[S->chns_space[1]], &F = S->spaces[S->chns_space[2]];
  std::vector<double> tab(S->w);
  tab.insert(tab.end(), U.L.begin(), U.L.end());
  tab.insert(tab.end(), U.dL.begin(), U.dL.end());
  tab.insert(tab.end(), P.L.begin(), P.L.end());
  tab.insert(tab.end(), F.L.begin(), F.L.end());
  tab.insert(tab.end(), F.dL.begin(), F.dL.end());
  tab.insert(tab.end(), P.dL.begin(), P.dL.end());
  const int expect = F.nS == 3 ? ChnsT<3>::len(S->nq) : ChnsT<6>::len(S->nq);
  if((int)tab.size() != expect) {
    set_error("CHNS kernel: basis tables have unexpected sizes");
    return B200_ERR_ARG;
  }
  cudaFree(S->chns_tab);
  S->chns_tab     = nullptr;
  S->chns_tab_len = (int)tab.size();
  B200_CUDA(cudaMalloc(&S->chns_tab, tab.size() * sizeof(double)));
  B200_CUDA(cudaMemcpyAsync(S->chns_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, S->stream));
  // row-local offsets of the local entries
  cudaFree(S->chns_off);
  S->chns_off = nullptr;
  B200_CUDA(cudaMalloc(&S->chns_off, (size_t)S->nElm * S->M * S->M * sizeof(uint16_t)));
  int *d_err;
  B200_CUDA(cudaMalloc(&d_err, sizeof(int)));
  B200_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), S->stream));
  chns_offsets_kernel<<<148 * 8, 256, 0, S->stream>>>(S->nElm, S->M, S->chns_adr, S->d_ia, S->d_ja, S->nInc, S->chns_off, d_err);
  count_launch();
  int h_err = 0;
  B200_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, S->stream));
  B200_CUDA(cudaStreamSynchronize(S->stream));
  cudaFree(d_err);
  if(h_err) {
    set_error("b200_finalize: a local (row, col) pair of the CHNS form is missing from the CSR pattern");
    return B200_ERR_ARG;
  }
  return B200_OK;
}

// what: bit 0 residual, bit 1 matrix; ADDS into val / rhs (the caller has zeroed them)
int chns_launch(System *S, int what)
{
  ChnsArgs a;
  a.nElm   = S->nElm;
  a.xyz    = S->d_xyz;
  a.conn   = S->d_conn;
  a.adr    = S->chns_adr;
  a.off    = S->chns_off;
  a.sol    = S->d_sol;
  a.soldot = S->have_soldot ? S->d_soldot : nullptr;
  a.tab    = S->chns_tab;
  a.ia     = S->d_ia;
  a.ja     = S->d_ja;
  a.val    = S->d_val;
  a.rhs    = S->d_rhs;
  a.nInc   = S->nInc;
  a.nq     = S->nq;
  a.ntab   = S->chns_tab_len;
  a.what   = what;
  a.c0     = S->c0;
  a.h0     = sqrt(DBL_EPSILON); // src/feBilinearForm.cpp:170
  a.prm    = S->chns_prm;
  const int    nsf  = S->spaces[S->chns_space[2]].nS;
  const int    M    = S->M;
  const int    model = S->chns_model;
  a.soln = (model != 0 && S->have_soln) ? S->d_soln : nullptr;
  a.dt   = S->dt;
  if(model == 2 && !(S->dt > 0.)) {
    set_error("CHNS_Khanwale needs the time step: call b200_set_solution_n(s, sol_n, dt) with dt > 0");
    return B200_ERR_ARG;
  }
  const size_t smem = ((size_t)a.ntab + 2 * CHNS_WPB * M + (size_t)CHNS_WPB * a.nq * chns_nfld(model) + CHNS_WPB * 32) * sizeof(double) +
                      (size_t)CHNS_WPB * M * sizeof(int32_t);
  const unsigned grid = (unsigned)((S->nElm + CHNS_WPB - 1) / CHNS_WPB);
#define B200_LAUNCH_CHNS(NSF, MODEL)                                                                                                     \
  do {                                                                                                                                   \
    B200_CUDA(cudaFuncSetAttribute(chns_kernel<NSF, MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
    chns_kernel<NSF, MODEL><<<grid, CHNS_WPB * 32, smem, S->stream>>>(a);                                                                \
  } while(0)
  if(nsf == 3 && model == 0)
    B200_LAUNCH_CHNS(3, 0);
  else if(nsf == 6 && model == 0)
    B200_LAUNCH_CHNS(6, 0);
  else if(nsf == 3 && model == 1)
    B200_LAUNCH_CHNS(3, 1);
  else if(nsf == 6 && model == 1)
    B200_LAUNCH_CHNS(6, 1);
  else if(nsf == 3)
    B200_LAUNCH_CHNS(3, 2);
  else
    B200_LAUNCH_CHNS(6, 2);
#undef B200_LAUNCH_CHNS
  count_launch();
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

} // namespace b200
