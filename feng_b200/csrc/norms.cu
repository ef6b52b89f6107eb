// Error norms of the device-resident state as reductions on the device.
//
// Same sums as the reference's feNorm (src/feNorm.cpp):
//   computeLpNorm / computeVectorLpNorm (:323-398, :1399-1443):  ( sum_e sum_k w_k J_e sum_i |u_i(x_k) - uh_i(x_k)|^p )^(1/p)
//   computeH1SemiNorm / computeVectorH1SemiNorm (:1643-1732, :1794-1846):  sqrt( sum_e sum_k w_k J_e |grad u - grad uh|^2 )
// The exact field is a host callback in the reference (feFunction::eval at the physical quadrature node); here the caller tabulates
// it at every (element, quadrature node), as tabulated sources and coefficients are.  One thread per element, one partial sum per
// CTA, summed in a fixed order by a second kernel: the result does not depend on the launch.
#include <cmath>
#include <string>

#include "device_common.cuh"
#include "system.h"

namespace b200 {

struct NormArgs {
  int64_t        nElm;
  const double  *xyz;
  const int32_t *conn, *adr;
  const double  *sol, *exact, *tab; // tab: L[nq][nS], then dL[nq][nS][dim], then w[nq]
  double        *partial;
  int            nS, nc, nq, kind, p;
};

template <int D> __global__ void __launch_bounds__(128) norm_kernel(const NormArgs a)
{
  extern __shared__ double sm[];
  const int ntab = a.nq * a.nS * (1 + D) + a.nq;
  for(int i = threadIdx.x; i < ntab; i += blockDim.x) sm[i] = a.tab[i];
  __shared__ double red[4];
  __syncthreads();
  const double *L = sm, *dL = sm + a.nq * a.nS, *w = dL + a.nq * a.nS * D;
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double        acc = 0.;
  if(e < a.nElm) {
    int32_t vtx[D + 1];
#pragma unroll
    for(int v = 0; v <= D; ++v) vtx[v] = a.conn[e * (D + 1) + v];
    double G[D * D], J;
    element_geometry<D>(a.xyz, vtx, G, &J);
    const int nf = a.nS * a.nc;
    const int32_t *ad = a.adr + e * nf;
    for(int k = 0; k < a.nq; ++k) {
      double s = 0.;
      for(int c = 0; c < a.nc; ++c) {
        if(a.kind == 0) {
          double uh = 0.;
          for(int b = 0; b < a.nS; ++b) uh += L[k * a.nS + b] * a.sol[ad[b * a.nc + c]];
          const double d = fabs((a.exact ? a.exact[(e * a.nq + k) * a.nc + c] : 0.) - uh);
          s += a.p == 2 ? d * d : a.p == 1 ? d : pow(d, (double)a.p);
        } else {
          double gr[D]; // reference gradient of uh_c
#pragma unroll
          for(int al = 0; al < D; ++al) gr[al] = 0.;
          for(int b = 0; b < a.nS; ++b) {
            const double u = a.sol[ad[b * a.nc + c]];
#pragma unroll
            for(int al = 0; al < D; ++al) gr[al] += dL[(k * a.nS + b) * D + al] * u;
          }
#pragma unroll
          for(int m = 0; m < D; ++m) {
            double gp = 0.;
#pragma unroll
            for(int al = 0; al < D; ++al) gp += G[al * D + m] * gr[al];
            const double d = gp - (a.exact ? a.exact[((e * a.nq + k) * a.nc + c) * D + m] : 0.);
            s += d * d;
          }
        }
      }
      acc += s * J * w[k];
    }
  }
  // CTA sum in a fixed order: warp shuffles, then the four warp sums
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if(threadIdx.x == 0) a.partial[blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
}

__global__ void norm_sum_kernel(int64_t n, const double *partial, double *out)
{
  // one CTA, strided partial sums in a fixed order
  __shared__ double s[256];
  double            acc = 0.;
  for(int64_t i = threadIdx.x; i < n; i += 256) acc += partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for(int o = 128; o > 0; o >>= 1) {
    if((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if(threadIdx.x == 0) out[0] = s[0];
}

int error_norm(System *S, int space, int kind, int p, const double *exact, double *out)
{
  if(space < 0 || space >= (int)S->spaces.size() || !out || (kind != 0 && kind != 1) || p < 1) {
    set_error("b200_error_norm: bad arguments");
    return B200_ERR_ARG;
  }
  if(!S->d_sol || S->nElm == 0) {
    set_error("b200_error_norm: no mesh / no state on the device");
    return B200_ERR_ARG;
  }
  if(comm_active(S)) {
    set_error("b200_error_norm: sub-domain systems hold ghost elements; evaluate the norm on the undecomposed system");
    return B200_ERR_UNSUPP;
  }
  const Space &sp = S->spaces[space];
  const int    D = S->dim, nq = S->nq;
  const size_t ntab = (size_t)nq * sp.nS * (1 + D) + nq;
  std::vector<double> tab;
  tab.insert(tab.end(), sp.L.begin(), sp.L.end());
  tab.insert(tab.end(), sp.dL.begin(), sp.dL.end());
  tab.insert(tab.end(), S->w.begin(), S->w.end());
  const size_t nex = exact ? (size_t)S->nElm * nq * sp.nc * (kind == 1 ? D : 1) : 0;
  const int64_t nblk = (S->nElm + 127) / 128;
  double *d_tab = nullptr, *d_ex = nullptr, *d_part = nullptr;
  B200_CUDA(cudaMalloc(&d_tab, ntab * sizeof(double)));
  B200_CUDA(cudaMalloc(&d_part, (size_t)(nblk + 1) * sizeof(double)));
  int rc = B200_OK;
  do {
    if(cudaMemcpyAsync(d_tab, tab.data(), ntab * sizeof(double), cudaMemcpyHostToDevice, S->stream) != cudaSuccess) {
      rc = B200_ERR_CUDA;
      break;
    }
    if(nex) {
      if(cudaMalloc(&d_ex, nex * sizeof(double)) != cudaSuccess ||
         cudaMemcpyAsync(d_ex, exact, nex * sizeof(double), cudaMemcpyHostToDevice, S->stream) != cudaSuccess) {
        rc = B200_ERR_CUDA;
        break;
      }
    }
    NormArgs a;
    a.nElm    = S->nElm;
    a.xyz     = S->d_xyz;
    a.conn    = S->d_conn;
    a.adr     = sp.d_adr;
    a.sol     = S->d_sol;
    a.exact   = d_ex;
    a.tab     = d_tab;
    a.partial = d_part;
    a.nS      = sp.nS;
    a.nc      = sp.nc;
    a.nq      = nq;
    a.kind    = kind;
    a.p       = p;
    const size_t smem = ntab * sizeof(double);
    if(D == 2)
      norm_kernel<2><<<(unsigned)nblk, 128, smem, S->stream>>>(a);
    else
      norm_kernel<3><<<(unsigned)nblk, 128, smem, S->stream>>>(a);
    norm_sum_kernel<<<1, 256, 0, S->stream>>>(nblk, d_part, d_part + nblk);
    count_launch(2);
    double h = 0.;
    if(cudaMemcpyAsync(&h, d_part + nblk, sizeof(double), cudaMemcpyDeviceToHost, S->stream) != cudaSuccess ||
       cudaStreamSynchronize(S->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
      rc = B200_ERR_CUDA;
      break;
    }
    // negative quadrature weights can sum to a very small negative integral (src/feNorm.cpp:391-395)
    if(h < 0. && fabs(h) < 1e-14) h = fabs(h);
    *out = kind == 0 ? pow(h, 1. / (double)p) : sqrt(h);
  } while(0);
  cudaFree(d_tab);
  cudaFree(d_ex);
  cudaFree(d_part);
  if(rc != B200_OK) set_error(std::string("b200_error_norm: ") + cudaGetErrorString(cudaGetLastError()));
  return rc;
}

} // namespace b200
