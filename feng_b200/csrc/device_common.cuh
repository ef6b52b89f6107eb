// Device helpers shared by the assembly kernels.
#pragma once
#include <cstdint>

namespace b200 {

// Inverse affine map of a straight simplex; conventions of feCncGeo::computeElementTransformation
// (src/feCncGeo.cpp:651-692): G[alpha*DIM+m] = d(xi_alpha)/d(x_m); detJ as src/feCncGeo.cpp:332,340,385.
template <int DIM> __device__ __forceinline__ void element_geometry(const double *__restrict__ xyz, const int32_t *vtx, double *G, double *detJ)
{
  if(DIM == 2) {
    const double x0 = xyz[2 * vtx[0]], y0 = xyz[2 * vtx[0] + 1];
    const double dxdr = xyz[2 * vtx[1]] - x0, dydr = xyz[2 * vtx[1] + 1] - y0;
    const double dxds = xyz[2 * vtx[2]] - x0, dyds = xyz[2 * vtx[2] + 1] - y0;
    const double J = dxdr * dyds - dydr * dxds;
    G[0] = dyds / J;  // dr/dx
    G[1] = -dxds / J; // dr/dy
    G[2] = -dydr / J; // ds/dx
    G[3] = dxdr / J;  // ds/dy
    *detJ = J;
  } else {
    double F[3][3]; // F[m][alpha] = dx_m / dxi_alpha
    const double *p0 = xyz + 3 * vtx[0];
#pragma unroll
    for(int al = 0; al < 3; ++al) {
      const double *p = xyz + 3 * vtx[al + 1];
#pragma unroll
      for(int m = 0; m < 3; ++m) F[m][al] = p[m] - p0[m];
    }
    const double c00 = F[1][1] * F[2][2] - F[1][2] * F[2][1];
    const double c01 = F[1][2] * F[2][0] - F[1][0] * F[2][2];
    const double c02 = F[1][0] * F[2][1] - F[1][1] * F[2][0];
    const double J = F[0][0] * c00 + F[0][1] * c01 + F[0][2] * c02;
    const double iJ = 1. / J;
    // inverse of F: G[alpha][m]
    G[0] = c00 * iJ;
    G[1] = (F[0][2] * F[2][1] - F[0][1] * F[2][2]) * iJ;
    G[2] = (F[0][1] * F[1][2] - F[0][2] * F[1][1]) * iJ;
    G[3] = c01 * iJ;
    G[4] = (F[0][0] * F[2][2] - F[0][2] * F[2][0]) * iJ;
    G[5] = (F[0][2] * F[1][0] - F[0][0] * F[1][2]) * iJ;
    G[6] = c02 * iJ;
    G[7] = (F[0][1] * F[2][0] - F[0][0] * F[2][1]) * iJ;
    G[8] = (F[0][0] * F[1][1] - F[0][1] * F[1][0]) * iJ;
    *detJ = J;
  }
}


// Layout of the pre-contracted reference tensors (built by gather.cu:build_tables from the host tables)
template <int D, int NS, int NP> struct GT {
  static constexpr int NU = NS * D, M = NU + NP;
  static constexpr int O_K = 0;                       // Kref[a][b][al][be]
  static constexpr int O_T3 = O_K + NS * NS * D * D;  // T3[a][b][v]
  static constexpr int O_M = O_T3 + NS * NS * NP;     // Mref[a][b]
  static constexpr int O_E = O_M + NS * NS;           // E[c][al][v]
  static constexpr int O_B = O_E + NS * D * NP;       // Bref[q][a][al]
  static constexpr int O_W = O_B + NP * NS * D;       // W[k][a] = w_k phi_a(k)   (source forms only)
  static constexpr int OFFW_U = (M + 7) / 8 * 8;      // uint16 offsets per (U node, element) pair
  static constexpr int OFFW_P = (NU + 7) / 8 * 8;     // per (P node, element) pair
  static constexpr int GW = (D * D + 1 + 1) / 2 * 2;  // doubles per element of the geometry table (16-byte aligned records)
};

} // namespace b200
