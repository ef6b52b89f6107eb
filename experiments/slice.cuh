// Row-slice kernel (experimental, 2-D P2/P1, B200_PATCH_KERNEL=slice): the "next kernel" of DESIGN.md section 5.
//
// Same plan as the patch kernel (Morton patches, block slots with contribution lists), different division of labour:
//  * PRODUCER: one thread per element of the patch computes the element matrix one ROW SLICE at a time -- slice a = the
//    blocks (a, b) of local row node a -- with every operand in registers or in the constant bank: the local row index is
//    a compile-time constant of the unrolled slice sequence, so the reference tensors Kref[a][b], T3[a][b][v], E, Bref are
//    constant-bank operands of the DFMAs (no table loads at all), and the element state is loaded once per slice.
//  * The slice (30 doubles per element) goes through a shared-memory stage.
//  * CONSUMER: one thread per block slot keeps its sum in registers across the slices and adds the contributions whose
//    local row index is the current slice; finished slots are stored once, coalesced, like in the patch kernel.
// Per block contribution the shared-memory traffic is one store and one load of the block itself (8 doubles) instead of
// the ~26 doubles of operands the slot-owner patch kernel re-reads.
#pragma once

namespace b200 {

__constant__ double c_gt[512]; // GT<2,6,3> reference tensors of the system being assembled (set before every launch)

struct SliceCoef {
  double cvd, sig_mu, mass0, cup, c_div, cvis1, cvis2, cpre;
};

template <int A, bool RES>
__device__ __forceinline__ void slice_produce_u(const double *st, double *stage, double *R, const SliceCoef k)
{
  using S_ = PS<2, 6, 3>;
  using G_ = GT<2, 6, 3>;
  double G[4], U[6][2], P[3], gu[12];
#pragma unroll
  for(int i = 0; i < 4; ++i) G[i] = st[S_::O_G + i];
  const double J = st[S_::O_J], cJ = st[S_::O_J + 1];
#pragma unroll
  for(int b = 0; b < 6; ++b) {
    U[b][0] = st[S_::O_U + 2 * b];
    U[b][1] = st[S_::O_U + 2 * b + 1];
  }
#pragma unroll
  for(int q = 0; q < 3; ++q) P[q] = st[S_::O_P + q];
#pragma unroll
  for(int i = 0; i < 12; ++i) gu[i] = st[S_::O_GU + i];
  double Z[2][3];
#pragma unroll
  for(int al = 0; al < 2; ++al)
#pragma unroll
    for(int v = 0; v < 3; ++v) Z[al][v] = 0.;
#pragma unroll
  for(int c = 0; c < 6; ++c) {
    double ut[2];
#pragma unroll
    for(int al = 0; al < 2; ++al) ut[al] = cJ * (U[c][0] * G[al * 2] + U[c][1] * G[al * 2 + 1]);
#pragma unroll
    for(int v = 0; v < 3; ++v) {
      const double t = c_gt[G_::O_T3 + (A * 6 + c) * 3 + v];
#pragma unroll
      for(int al = 0; al < 2; ++al) Z[al][v] += ut[al] * t;
    }
  }
  double r[2] = {0., 0.};
#pragma unroll
  for(int b = 0; b < 6; ++b) {
    double c1 = 0.;
#pragma unroll
    for(int al = 0; al < 2; ++al)
#pragma unroll
      for(int v = 0; v < 3; ++v) c1 += c_gt[G_::O_E + (b * 2 + al) * 3 + v] * Z[al][v];
    double H[2][2], K[2][2];
#pragma unroll
    for(int al = 0; al < 2; ++al)
#pragma unroll
      for(int n = 0; n < 2; ++n)
        H[al][n] = J * (c_gt[G_::O_K + ((A * 6 + b) * 2 + al) * 2] * G[n] + c_gt[G_::O_K + ((A * 6 + b) * 2 + al) * 2 + 1] * G[2 + n]);
#pragma unroll
    for(int m = 0; m < 2; ++m)
#pragma unroll
      for(int n = 0; n < 2; ++n) K[m][n] = G[m] * H[0][n] + G[2 + m] * H[1][n];
    double s = c1 + k.cvd * (K[0][0] + K[1][1]);
    if(k.mass0 != 0.) s += k.mass0 * J * c_gt[G_::O_M + A * 6 + b];
    double t3[3];
#pragma unroll
    for(int v = 0; v < 3; ++v) t3[v] = cJ * c_gt[G_::O_T3 + (A * 6 + b) * 3 + v];
    double Ab[4];
#pragma unroll
    for(int i = 0; i < 2; ++i)
#pragma unroll
      for(int j = 0; j < 2; ++j) {
        double v = (i == j ? s : 0.) - k.sig_mu * K[j][i];
#pragma unroll
        for(int w = 0; w < 3; ++w) v += gu[(w * 2 + j) * 2 + i] * t3[w];
        Ab[i * 2 + j] = v;
      }
    *reinterpret_cast<double2 *>(stage + b * 4)     = make_double2(Ab[0], Ab[1]);
    *reinterpret_cast<double2 *>(stage + b * 4 + 2) = make_double2(Ab[2], Ab[3]);
    if(RES) {
      r[0] -= c1 * U[b][0];
      r[1] -= c1 * U[b][1];
    }
  }
#pragma unroll
  for(int q = 0; q < 3; ++q) {
    double bp[2];
#pragma unroll
    for(int i = 0; i < 2; ++i) bp[i] = J * (G[i] * c_gt[G_::O_B + (q * 6 + A) * 2] + G[2 + i] * c_gt[G_::O_B + (q * 6 + A) * 2 + 1]);
    *reinterpret_cast<double2 *>(stage + 24 + q * 2) = make_double2(k.cup * bp[0], k.cup * bp[1]);
    if(RES) {
#pragma unroll
      for(int m = 0; m < 2; ++m) {
        r[m] += k.cpre * bp[m] * P[q];
#pragma unroll
        for(int i = 0; i < 2; ++i) r[i] += bp[m] * (k.cvis1 * gu[(q * 2 + m) * 2 + i] + k.cvis2 * gu[(q * 2 + i) * 2 + m]);
      }
    }
  }
  if(RES) {
    R[A * 2]     = r[0];
    R[A * 2 + 1] = r[1];
  }
}

template <int Q, bool RES> __device__ __forceinline__ void slice_produce_p(const double *st, double *stage, double *R, const SliceCoef k)
{
  using S_ = PS<2, 6, 3>;
  using G_ = GT<2, 6, 3>;
  double G[4];
#pragma unroll
  for(int i = 0; i < 4; ++i) G[i] = st[S_::O_G + i];
  const double cj = k.c_div * st[S_::O_J];
  double       rp = 0.;
#pragma unroll
  for(int b = 0; b < 6; ++b) {
    double v[2];
#pragma unroll
    for(int j = 0; j < 2; ++j) v[j] = cj * (G[j] * c_gt[G_::O_B + (Q * 6 + b) * 2] + G[2 + j] * c_gt[G_::O_B + (Q * 6 + b) * 2 + 1]);
    *reinterpret_cast<double2 *>(stage + b * 2) = make_double2(v[0], v[1]);
    if(RES) rp -= v[0] * st[S_::O_U + 2 * b] + v[1] * st[S_::O_U + 2 * b + 1];
  }
  if(RES) R[12 + Q] = rp;
}

template <int NT, int MAXS, bool RES> __global__ void __launch_bounds__(NT, 1) slice_kernel_2d(const PatchArgs a)
{
  constexpr int D = 2, NS = 6, NP = 3, NU = 12, NL = 15, GW = 6;
  using S_ = PS<D, NS, NP>;
  using G_ = GT<D, NS, NP>;
  constexpr int SW = S_::W, STW = 32, RW = 16;
  extern __shared__ double sm[];
  const int       tid = threadIdx.x;
  const PatchDesc pd  = a.desc[blockIdx.x];
  const int       nE  = pd.ne;
  double         *s_el    = sm;                                   // [nE][SW]   element records
  double         *s_stage = s_el + nE * SW;                       // [nE][STW]  current row slice of every element matrix
  double         *s_R     = s_stage + nE * STW;                   // [nE][RW]   element residuals
  uint16_t       *s_pairs = reinterpret_cast<uint16_t *>(s_R + nE * RW); // [np]
  uint16_t       *s_np0   = s_pairs + ((pd.np + 3) & ~3);         // [nn]
  const int32_t  *el_list = a.elems + pd.e0;
  const THCoeffs  c       = a.c;
  const NodeRec<D> *nodes = static_cast<const NodeRec<D> *>(a.nodes) + pd.n0;

  // phase 1a / 1b: as in patch_kernel (staging of the local solution and geometry, vertex gradients)
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
  for(int it = tid; it < nE * D * NP; it += NT) {
    const int el = it / (D * NP), r = it - el * (D * NP), i = r / NP, v = r - i * NP;
    double   *st = s_el + el * SW;
    double    X[D] = {0., 0.};
#pragma unroll
    for(int cc = 0; cc < NS; ++cc) {
      const double u = st[S_::O_U + cc * D + i];
#pragma unroll
      for(int al = 0; al < D; ++al) X[al] += u * a.tab[PT<D, NS, NP>::O_E + (cc * D + al) * NP + v];
    }
#pragma unroll
    for(int j = 0; j < D; ++j) st[S_::O_GU + (v * D + j) * D + i] = st[S_::O_G + j] * X[0] + st[S_::O_G + D + j] * X[1];
  }
  __syncthreads();

  // my block slots and their sums
  const uint2    *blk = reinterpret_cast<const uint2 *>(a.blocks + pd.b0);
  const uint16_t *ctr = a.ctr + pd.c0;
  uint2           rec[MAXS];
  double          acc[MAXS][4];
#pragma unroll
  for(int s = 0; s < MAXS; ++s) {
    const int k = tid + s * NT;
    rec[s]      = k < pd.nb ? blk[k] : make_uint2(0u, 0u); // cnt = 0: nothing to add, nothing to store
#pragma unroll
    for(int i = 0; i < 4; ++i) acc[s][i] = 0.;
  }
  SliceCoef kf;
  kf.cvd    = c.diff_k - c.sig_mu;
  kf.sig_mu = c.sig_mu;
  kf.mass0  = c.c_mass * a.c0;
  kf.cup    = c.c_sig - c.c_gradp;
  kf.c_div  = c.c_div;
  kf.cvis1  = c.sig_mu - c.diff_k;
  kf.cvis2  = c.sig_mu;
  kf.cpre   = c.c_gradp - c.c_sig;
  const double *my_el    = s_el + tid * SW;
  double       *my_stage = s_stage + tid * STW, *my_R = s_R + tid * RW;

  // the nine slices share ONE copy of the consumer code (the slice index is a run-time value there); only the producers
  // are specialised per slice.  Fully unrolling the sequence makes ~36 k instructions and the kernel instruction-fetch
  // bound (measured: 118 Melem/s).
#pragma unroll 1
  for(int sl = 0; sl < NS + NP; ++sl) {
    if(tid < nE) {
      switch(sl) {
      case 0: slice_produce_u<0, RES>(my_el, my_stage, my_R, kf); break;
      case 1: slice_produce_u<1, RES>(my_el, my_stage, my_R, kf); break;
      case 2: slice_produce_u<2, RES>(my_el, my_stage, my_R, kf); break;
      case 3: slice_produce_u<3, RES>(my_el, my_stage, my_R, kf); break;
      case 4: slice_produce_u<4, RES>(my_el, my_stage, my_R, kf); break;
      case 5: slice_produce_u<5, RES>(my_el, my_stage, my_R, kf); break;
      case 6: slice_produce_p<0, RES>(my_el, my_stage, my_R, kf); break;
      case 7: slice_produce_p<1, RES>(my_el, my_stage, my_R, kf); break;
      default: slice_produce_p<2, RES>(my_el, my_stage, my_R, kf); break;
      }
    }
    __syncthreads();
#pragma unroll
    for(int s = 0; s < MAXS; ++s) {
      const uint2 bw     = rec[s];
      const int   cnt    = (bw.y >> 16) & 0xffu;
      const int   cstart = bw.y & 0xffffu, kind = bw.y >> 27;
      const int   pair0  = cnt ? s_np0[bw.x & 0xffffu] : 0;
      for(int t = 0; t < cnt; ++t) {
        const uint32_t cw = cnt == 1 ? (uint32_t)cstart : (uint32_t)ctr[cstart + t];
        const uint32_t pw = s_pairs[pair0 + (int)(cw >> 4)];
        if((int)(pw & 15) != sl) continue;
        const int     lb  = cw & 15;
        const double *src = s_stage + (pw >> 4) * STW + (kind == 0 ? lb * 4 : (kind == 1 ? 24 + (lb - NS) * 2 : lb * 2));
        const double2 v0  = *reinterpret_cast<const double2 *>(src);
        acc[s][0] += v0.x;
        acc[s][1] += v0.y;
        if(kind == 0) {
          const double2 v1 = *reinterpret_cast<const double2 *>(src + 2);
          acc[s][2] += v1.x;
          acc[s][3] += v1.y;
        }
      }
    }
    __syncthreads();
  }

  // finished slots: U-U blocks acc = A[i][j] (i*2+j), U-P acc = {A[0], A[1]}, P-U acc = {A[0][0], A[0][1]}
#pragma unroll
  for(int s = 0; s < MAXS; ++s) {
    const uint2 bw  = rec[s];
    const int   cnt = (bw.y >> 16) & 0xffu;
    if(cnt == 0) continue;
    const int         off0 = bw.x >> 16, flags = bw.y >> 24, kind = flags >> 3, dmask = flags & 7;
    const NodeRec<D> &nd = nodes[bw.x & 0xffffu];
    if(kind == 0) {
#pragma unroll
      for(int i = 0; i < D; ++i) {
        const int64_t vb = nd.vbase[i];
        if(vb >= 0) {
          double *dst = a.val + vb + off0;
          if(dmask == 3 && ((vb + off0) & 1) == 0)
            *reinterpret_cast<double2 *>(dst) = make_double2(acc[s][i * 2], acc[s][i * 2 + 1]);
          else {
            int n = 0;
#pragma unroll
            for(int j = 0; j < D; ++j)
              if((dmask >> j) & 1) dst[n++] = acc[s][i * 2 + j];
          }
        }
      }
    } else if(kind == 1) {
#pragma unroll
      for(int i = 0; i < D; ++i) {
        const int64_t vb = nd.vbase[i];
        if(vb >= 0) a.val[vb + off0] = acc[s][i];
      }
    } else {
      double *dst = a.val + nd.vbase[0] + off0;
      int     n   = 0;
#pragma unroll
      for(int j = 0; j < D; ++j)
        if((dmask >> j) & 1) dst[n++] = acc[s][j];
    }
  }
  if(RES) {
    for(int n = tid; n < pd.nn; n += NT) {
      const NodeRec<D> &nd = nodes[n];
      double            res[D] = {0., 0.};
      for(int t = 0; t < nd.npairs; ++t) {
        const uint32_t pw = s_pairs[nd.pair0 + t];
        const double  *Re = s_R + (pw >> 4) * RW;
        if(nd.kind == 0) {
          res[0] += Re[(pw & 15) * 2];
          res[1] += Re[(pw & 15) * 2 + 1];
        } else
          res[0] += Re[12 + (int)(pw & 15) - NS];
      }
      if(nd.kind == 0) {
#pragma unroll
        for(int i = 0; i < D; ++i)
          if(nd.rows[i] < a.nInc) a.rhs[nd.rows[i]] = res[i];
      } else if(nd.rows[0] < a.nInc)
        a.rhs[nd.rows[0]] = res[0];
    }
  }
}

} // namespace b200
