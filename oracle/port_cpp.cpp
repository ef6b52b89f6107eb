// =============================================================================
// TEST / BASELINE INFRASTRUCTURE ONLY -- C++/OpenMP restatement of the reference's CPU assembly path, generalised to
// dim = 3 (the reference instantiates its vector weak forms for dim = 2 only, src/feVectorSysElm.cpp:1246,1536, so the
// "~20 M-DOF tetrahedral Navier-Stokes" workload of BASELINE.json has no reference implementation to time).
//
// It follows the reference's organisation, not the product's:
//   * one mesh traversal per weak form (feLinearSystemMklPardiso::assembleMatrices loops the forms,
//     src/feLinearSystemMklPardiso.cpp:524-663), each visit gathering coordinates, addressing vectors and local
//     solution like feBilinearForm::initialize (src/feBilinearForm.cpp:284-367);
//   * quadrature-point-major loops of computeAe / computeBe with physical gradients recomputed per point
//     (src/feVectorSysElm.cpp:1171-1242 convective acceleration, :1454-1532 divergence of the Newtonian stress, :685-749
//     mixed divergence, :449-503 vector diffusion, :528-578 mixed gradient);
//   * colour loop with an OpenMP parallel-for over the elements of one colour and a sorted scatter into the CSR row
//     (src/feLinearSystemMklPardiso.cpp:548-660), essential DOFs (>= nInc) filtered;
//   * the pattern of feEZCompressedRowStorage (src/feCompressedRowStorage.cpp:15-133): per-row vectors, sort, unique,
//     forced diagonal.
// It is block-structured (the structural zeros of the vector-Lagrange layout, src/feSpace_2D.cpp:41-55, are skipped), so it
// does LESS arithmetic per element than the reference's dense loops: as a CPU baseline it errs on the fast side.
// Pinned on oracle/fe_oracle.py (numpy) in 2-D and 3-D and on the compiled reference in 2-D (tests/test_port_cpp.py).
// Nothing in feng_b200/ links or loads this file.
// =============================================================================
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#if defined(_OPENMP)
#include <omp.h>
#endif

namespace {

enum { VECTOR_DIFFUSION = 18, CONVECTIVE = 22, DIV_STRESS = 25, MIXED_GRADIENT = 26, MIXED_DIVERGENCE = 31 };

struct Mesh {
  int            dim, nv;
  int64_t        nE;
  const double  *xyz;   // [nVert][3]
  const int32_t *cells; // [nE][nv]
};

// inverse affine map G[alpha][m] = d xi_alpha / d x_m and detJ (src/feCncGeo.cpp:332,340,385; :651-692)
inline double geometry(const Mesh &M, int64_t e, double G[3][3])
{
  const int d = M.dim;
  double    F[3][3];
  const int32_t *c = M.cells + e * M.nv;
  for(int a = 0; a < d; ++a)
    for(int m = 0; m < d; ++m) F[m][a] = M.xyz[3 * (int64_t)c[a + 1] + m] - M.xyz[3 * (int64_t)c[0] + m];
  if(d == 2) {
    const double J = F[0][0] * F[1][1] - F[1][0] * F[0][1];
    G[0][0] = F[1][1] / J;
    G[1][0] = -F[1][0] / J;
    G[0][1] = -F[0][1] / J;
    G[1][1] = F[0][0] / J;
    return J;
  }
  const double J = F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) - F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
                   F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
  G[0][0] = (F[1][1] * F[2][2] - F[1][2] * F[2][1]) / J;
  G[0][1] = (F[0][2] * F[2][1] - F[0][1] * F[2][2]) / J;
  G[0][2] = (F[0][1] * F[1][2] - F[0][2] * F[1][1]) / J;
  G[1][0] = (F[1][2] * F[2][0] - F[1][0] * F[2][2]) / J;
  G[1][1] = (F[0][0] * F[2][2] - F[0][2] * F[2][0]) / J;
  G[1][2] = (F[0][2] * F[1][0] - F[0][0] * F[1][2]) / J;
  G[2][0] = (F[1][0] * F[2][1] - F[1][1] * F[2][0]) / J;
  G[2][1] = (F[0][1] * F[2][0] - F[0][0] * F[2][1]) / J;
  G[2][2] = (F[0][0] * F[1][1] - F[0][1] * F[1][0]) / J;
  return J;
}

constexpr int MAXS = 10, MAXP = 4, MAXD = 3, MAXU = MAXS * MAXD;

struct Tables {
  int           nS, nP, nq;
  const double *w, *LU, *dLU, *LP; // LU[k][a], dLU[k][a][alpha], LP[k][q]
};

// one weak form on one element: Ae (rows M x cols N, row-major, leading dimension MAXU + MAXP) and Be
struct Local {
  double Ae[MAXU][MAXU + MAXP];
  double Be[MAXU];
  int    M, N;
};

void element_form(int kind, double c, double prm, const Mesh &mesh, const Tables &T, int64_t e, const int64_t *adrU, const int64_t *adrP,
                  const double *sol, bool matrix, Local &L)
{
  const int d = mesh.dim, nS = T.nS, nP = T.nP, nU = nS * d;
  double    G[3][3];
  const double J = geometry(mesh, e, G);
  double    ul[MAXS][MAXD], pl[MAXP];
  for(int a = 0; a < nS; ++a)
    for(int i = 0; i < d; ++i) ul[a][i] = sol[adrU[e * nU + a * d + i]];
  for(int q = 0; q < nP; ++q) pl[q] = adrP ? sol[adrP[e * nP + q]] : 0.;
  const bool rowsP = kind == MIXED_DIVERGENCE;
  L.M = rowsP ? nP : nU;
  L.N = kind == DIV_STRESS ? nU + nP : (kind == MIXED_GRADIENT ? nP : nU);
  for(int i = 0; i < L.M; ++i) {
    L.Be[i] = 0.;
    for(int j = 0; j < L.N; ++j) L.Ae[i][j] = 0.;
  }
  for(int k = 0; k < T.nq; ++k) {
    const double  jw = J * T.w[k];
    const double *phi = T.LU + (size_t)k * nS, *psi = T.LP ? T.LP + (size_t)k * nP : nullptr;
    double        g[MAXS][MAXD]; // physical gradients (src/feSpace.cpp:669-710)
    for(int a = 0; a < nS; ++a)
      for(int m = 0; m < d; ++m) {
        double s = 0.;
        for(int al = 0; al < d; ++al) s += T.dLU[((size_t)k * nS + a) * d + al] * G[al][m];
        g[a][m] = s;
      }
    double u[MAXD] = {0., 0., 0.}, gu[MAXD][MAXD] = {{0.}}, p = 0.; // gu[m][n] = d_m u_n (src/feSpace.cpp:1391-1394)
    for(int a = 0; a < nS; ++a)
      for(int n = 0; n < d; ++n) {
        u[n] += phi[a] * ul[a][n];
        for(int m = 0; m < d; ++m) gu[m][n] += g[a][m] * ul[a][n];
      }
    for(int q = 0; q < nP; ++q) p += psi[q] * pl[q];
    switch(kind) {
      case CONVECTIVE: {
        double ugu[MAXD];
        for(int i = 0; i < d; ++i) {
          double s = 0.;
          for(int n = 0; n < d; ++n) s += u[n] * gu[n][i];
          ugu[i] = s;
        }
        for(int a = 0; a < nS; ++a) {
          for(int i = 0; i < d; ++i) L.Be[a * d + i] -= c * ugu[i] * phi[a] * jw;
          if(!matrix) continue;
          for(int b = 0; b < nS; ++b) {
            double ugp = 0.;
            for(int m = 0; m < d; ++m) ugp += u[m] * g[b][m];
            const double pab = phi[a] * phi[b];
            for(int i = 0; i < d; ++i)
              for(int j = 0; j < d; ++j) L.Ae[a * d + i][b * d + j] += c * ((i == j ? ugp * phi[a] : 0.) + pab * gu[j][i]) * jw;
          }
        }
      } break;
      case DIV_STRESS: {
        for(int a = 0; a < nS; ++a) {
          for(int i = 0; i < d; ++i) {
            double s = 0.;
            for(int m = 0; m < d; ++m) s += g[a][m] * (gu[m][i] + gu[i][m]);
            L.Be[a * d + i] -= c * (p * g[a][i] - prm * s) * jw;
          }
          if(!matrix) continue;
          for(int b = 0; b < nS; ++b) {
            double gg = 0.;
            for(int m = 0; m < d; ++m) gg += g[a][m] * g[b][m];
            for(int i = 0; i < d; ++i)
              for(int j = 0; j < d; ++j) L.Ae[a * d + i][b * d + j] += -c * prm * ((i == j ? gg : 0.) + g[a][j] * g[b][i]) * jw;
          }
          for(int q = 0; q < nP; ++q)
            for(int i = 0; i < d; ++i) L.Ae[a * d + i][nU + q] += c * psi[q] * g[a][i] * jw;
        }
      } break;
      case MIXED_DIVERGENCE: {
        double divu = 0.;
        for(int m = 0; m < d; ++m) divu += gu[m][m];
        for(int q = 0; q < nP; ++q) {
          L.Be[q] -= c * divu * psi[q] * jw;
          if(!matrix) continue;
          for(int b = 0; b < nS; ++b)
            for(int j = 0; j < d; ++j) L.Ae[q][b * d + j] += c * psi[q] * g[b][j] * jw;
        }
      } break;
      case VECTOR_DIFFUSION: {
        for(int a = 0; a < nS; ++a) {
          for(int i = 0; i < d; ++i) {
            double s = 0.;
            for(int m = 0; m < d; ++m) s += g[a][m] * gu[m][i];
            L.Be[a * d + i] -= c * prm * s * jw;
          }
          if(!matrix) continue;
          for(int b = 0; b < nS; ++b) {
            double gg = 0.;
            for(int m = 0; m < d; ++m) gg += g[a][m] * g[b][m];
            for(int i = 0; i < d; ++i) L.Ae[a * d + i][b * d + i] += c * prm * gg * jw;
          }
        }
      } break;
      case MIXED_GRADIENT: {
        for(int a = 0; a < nS; ++a)
          for(int i = 0; i < d; ++i) {
            L.Be[a * d + i] += c * p * g[a][i] * jw;
            if(matrix)
              for(int q = 0; q < nP; ++q) L.Ae[a * d + i][q] += -c * psi[q] * g[a][i] * jw;
          }
      } break;
      default: break;
    }
  }
}

} // namespace

extern "C" {

int port_max_threads()
{
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void port_set_threads(int n)
{
#if defined(_OPENMP)
  if(n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// Pattern of feEZCompressedRowStorage for the couplings U x U, U x P, P x U (+ forced diagonal): first call with ja = NULL
// returns nnz and fills ia[nInc + 1]; second call fills ja (int32).
int64_t port_pattern(int64_t nE, int nU, int nP, const int64_t *adrU, const int64_t *adrP, int64_t nInc, int64_t *ia, int32_t *ja)
{
  static std::vector<std::vector<int32_t>> rows; // kept between the two calls
  if(!ja) {
    rows.assign(nInc, std::vector<int32_t>());
    for(int64_t i = 0; i < nInc; ++i) rows[i].push_back((int32_t)i);
    for(int64_t e = 0; e < nE; ++e) {
      for(int i = 0; i < nU + nP; ++i) {
        const int64_t I = i < nU ? adrU[e * nU + i] : adrP[e * nP + (i - nU)];
        if(I >= nInc) continue;
        for(int j = 0; j < nU + nP; ++j) {
          if(i >= nU && j >= nU) continue; // no P x P form
          const int64_t Jc = j < nU ? adrU[e * nU + j] : adrP[e * nP + (j - nU)];
          if(Jc < nInc) rows[I].push_back((int32_t)Jc);
        }
      }
    }
#pragma omp parallel for schedule(dynamic, 256)
    for(int64_t i = 0; i < nInc; ++i) {
      std::sort(rows[i].begin(), rows[i].end());
      rows[i].erase(std::unique(rows[i].begin(), rows[i].end()), rows[i].end());
    }
    ia[0] = 0;
    for(int64_t i = 0; i < nInc; ++i) ia[i + 1] = ia[i] + (int64_t)rows[i].size();
    return ia[nInc];
  }
  for(int64_t i = 0; i < nInc; ++i) std::copy(rows[i].begin(), rows[i].end(), ja + ia[i]);
  rows.clear();
  rows.shrink_to_fit();
  return ia[nInc];
}

// what: bit 0 residual, bit 1 matrix.  Colours: elements of colour c are color_elems[color_ptr[c] .. color_ptr[c+1]).
// Returns the wall time in seconds (assembly only).
double port_assemble(int dim, int64_t nE, const double *xyz, const int32_t *cells, const int64_t *adrU, const int64_t *adrP, int nS, int nP,
                     int nq, const double *w, const double *LU, const double *dLU, const double *LP, int64_t nInc, const int64_t *ia,
                     const int32_t *ja, const double *sol, int nForms, const int32_t *kinds, const double *coeff, const double *param,
                     int nColors, const int64_t *color_ptr, const int32_t *color_elems, int what, double *vals, double *rhs)
{
  const Mesh   mesh{dim, dim + 1, nE, xyz, cells};
  const Tables T{nS, nP, nq, w, LU, dLU, LP};
  const int    nU = nS * dim;
  const bool   matrix = (what & 2) != 0, residual = (what & 1) != 0;
  const auto   t0 = std::chrono::steady_clock::now();
  if(matrix) std::memset(vals, 0, (size_t)ia[nInc] * sizeof(double));
  if(residual) std::memset(rhs, 0, (size_t)nInc * sizeof(double));
  for(int f = 0; f < nForms; ++f) {
    const int kind = kinds[f];
    for(int col = 0; col < nColors; ++col) {
      const int64_t c0 = color_ptr[col], c1 = color_ptr[col + 1];
#pragma omp parallel
      {
        Local L;
#pragma omp for schedule(dynamic, 16)
        for(int64_t t = c0; t < c1; ++t) {
          const int64_t e = color_elems[t];
          element_form(kind, coeff[f], param[f], mesh, T, e, adrU, adrP, sol, matrix, L);
          const bool rowsP = kind == MIXED_DIVERGENCE;
          for(int i = 0; i < L.M; ++i) {
            const int64_t I = rowsP ? adrP[e * nP + i] : adrU[e * nU + i];
            if(I >= nInc) continue;
            if(residual) rhs[I] += L.Be[i];
            if(!matrix) continue;
            const int32_t *row = ja + ia[I];
            const int64_t  len = ia[I + 1] - ia[I];
            for(int j = 0; j < L.N; ++j) {
              int64_t Jc;
              if(kind == MIXED_GRADIENT)
                Jc = adrP[e * nP + j];
              else
                Jc = j < nU ? adrU[e * nU + j] : adrP[e * nP + (j - nU)];
              if(Jc >= nInc) continue;
              const int32_t *pos = std::lower_bound(row, row + len, (int32_t)Jc);
              vals[ia[I] + (pos - row)] += L.Ae[i][j];
            }
          }
        }
      }
    }
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Matrix-free application of the same operator: y += A(sol) x and rhs, element by element, without a CSR matrix -- the
// size-independent checker of the GPU assembly at the bench size (tests/test_gpu_fullsize.py), where the assembled
// matrix (15 GB of values) is not worth building on the host.  OpenMP over all elements, atomic row updates.
double port_apply(int dim, int64_t nE, const double *xyz, const int32_t *cells, const int64_t *adrU, const int64_t *adrP, int nS, int nP, int nq,
                  const double *w, const double *LU, const double *dLU, const double *LP, int64_t nInc, const double *sol, int nForms,
                  const int32_t *kinds, const double *coeff, const double *param, int nVec, const double *x, double *y, double *rhs)
{
  const Mesh   mesh{dim, dim + 1, nE, xyz, cells};
  const Tables T{nS, nP, nq, w, LU, dLU, LP};
  const int    nU = nS * dim;
  const auto   t0 = std::chrono::steady_clock::now();
  std::memset(y, 0, (size_t)nInc * nVec * sizeof(double));
  std::memset(rhs, 0, (size_t)nInc * sizeof(double));
#pragma omp parallel
  {
    Local L;
#pragma omp for schedule(dynamic, 64)
    for(int64_t e = 0; e < nE; ++e) {
      for(int f = 0; f < nForms; ++f) {
        const int kind = kinds[f];
        element_form(kind, coeff[f], param[f], mesh, T, e, adrU, adrP, sol, true, L);
        const bool rowsP = kind == MIXED_DIVERGENCE;
        for(int i = 0; i < L.M; ++i) {
          const int64_t I = rowsP ? adrP[e * nP + i] : adrU[e * nU + i];
          if(I >= nInc) continue;
#pragma omp atomic
          rhs[I] += L.Be[i];
          for(int v = 0; v < nVec; ++v) {
            const double *xv = x + (size_t)v * nInc;
            double        s = 0.;
            for(int j = 0; j < L.N; ++j) {
              const int64_t Jc = kind == MIXED_GRADIENT ? adrP[e * nP + j] : (j < nU ? adrU[e * nU + j] : adrP[e * nP + (j - nU)]);
              if(Jc < nInc) s += L.Ae[i][j] * xv[Jc];
            }
#pragma omp atomic
            y[(size_t)v * nInc + I] += s;
          }
        }
      }
    }
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

} // extern "C"
