// Device-side construction of the CSR sparsity pattern with the rules of feEZCompressedRowStorage
// (src/feCompressedRowStorage.cpp:15-133): forced diagonal (:33), only unknown x unknown pairs (:80), periodic
// (slave, master) extras (:98-107), columns ascending and unique per row (:110-116).  The reference pushes the
// adrI x adrJ pairs of a dry assembly into one std::vector per row and sorts each row; here the pairs are packed as
// 64-bit keys row * nInc + col, sorted and merged chunk by chunk (bounded memory), and split back into ia / ja.
// Set-up code: thrust (shipped with the CUDA toolkit) does the sort / set-union.
#include <thrust/binary_search.h>
#include <thrust/device_vector.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/set_operations.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include "system.h"

namespace b200 {

__global__ void pattern_keys_kernel(int64_t e0, int64_t e1, int M, int NU, const int32_t *adrU, const int32_t *adrP, int NP, int64_t nInc,
                                    int blockmask, uint64_t *keys)
{
  const int64_t tot = (e1 - e0) * M * (int64_t)M;
  for(int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = e0 + idx / (M * M);
    const int     r = (int)(idx % (M * M));
    const int     i = r / M, j = r - i * M;
    const int     bi = i < NU ? 0 : 1, bj = j < NU ? 0 : 1;
    uint64_t      key = ~0ull;
    if(blockmask & (1 << (bi * 2 + bj))) {
      const int64_t I = bi == 0 ? adrU[e * NU + i] : adrP[e * NP + (i - NU)];
      const int64_t J = bj == 0 ? adrU[e * NU + j] : adrP[e * NP + (j - NU)];
      if(I < nInc && J < nInc) key = (uint64_t)(I * nInc + J);
    }
    keys[idx] = key;
  }
}

__global__ void diag_keys_kernel(int64_t nInc, uint64_t *keys)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nInc; i += (int64_t)gridDim.x * blockDim.x)
    keys[i] = (uint64_t)(i * nInc + i);
}

__global__ void split_keys_kernel(int64_t nnz, int64_t nInc, const uint64_t *keys, int32_t *ja)
{
  for(int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x)
    ja[k] = (int32_t)(keys[k] % (uint64_t)nInc);
}

__global__ void row_start_keys_kernel(int64_t nInc, uint64_t *q)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= nInc; i += (int64_t)gridDim.x * blockDim.x)
    q[i] = (uint64_t)(i * nInc);
}

int analyze_forms(System *S); // assemble.cu
int alloc_linear_system_public(System *S);

int build_pattern_device(System *S, int64_t n_inc, int64_t n_dof, const std::vector<int64_t> &per_master, const std::vector<int64_t> &per_slave)
{
  S->nInc = n_inc;
  S->nDOF = n_dof;
  int rc  = analyze_forms(S);
  if(rc != B200_OK) return rc;
  const Space &U  = S->spaces[S->su];
  const bool   chns = S->plan == PLAN_CHNS; // one monolithic M x M block on the concatenated element->DOF table
  const int    M = S->M, NU = chns ? M : U.nS * U.nc, NP = chns ? 0 : (S->sp >= 0 ? S->spaces[S->sp].nS : 0);
  const int32_t *d_adr_u = chns ? S->chns_adr : U.d_adr;
  const int32_t *d_adr_p = (!chns && S->sp >= 0) ? S->spaces[S->sp].d_adr : nullptr;
  int          mask = 0;
  for(int bi = 0; bi < 2; ++bi)
    for(int bj = 0; bj < 2; ++bj)
      if(S->has_matrix_block[bi][bj]) mask |= 1 << (bi * 2 + bj);
  try {
    auto pol = thrust::cuda::par.on(S->stream);
    thrust::device_vector<uint64_t> acc(n_inc), chunk, merged;
    diag_keys_kernel<<<148 * 8, 256, 0, S->stream>>>(n_inc, thrust::raw_pointer_cast(acc.data()));
    count_launch();
    if(!per_master.empty()) {
      std::vector<uint64_t> extra;
      for(size_t p = 0; p < per_master.size(); ++p)
        if(per_slave[p] < n_inc && per_master[p] < n_inc) extra.push_back((uint64_t)(per_slave[p] * n_inc + per_master[p]));
      thrust::device_vector<uint64_t> ex(extra.begin(), extra.end());
      thrust::sort(pol, ex.begin(), ex.end());
      merged.resize(acc.size() + ex.size());
      auto end = thrust::set_union(pol, acc.begin(), acc.end(), ex.begin(), ex.end(), merged.begin());
      merged.resize(end - merged.begin());
      acc.swap(merged);
    }
    log_stage("pattern: diagonal keys");
    const int64_t per_elem = (int64_t)M * M;
    const int64_t chunk_elems = std::max<int64_t>(1, (int64_t)(1ll << 28) / per_elem); // <= 2 GiB of keys per chunk
    for(int64_t e0 = 0; e0 < S->nElm; e0 += chunk_elems) {
      const int64_t e1 = std::min(S->nElm, e0 + chunk_elems);
      chunk.resize((e1 - e0) * per_elem);
      pattern_keys_kernel<<<148 * 16, 256, 0, S->stream>>>(e0, e1, M, NU, d_adr_u, d_adr_p, NP, n_inc, mask,
                                                          thrust::raw_pointer_cast(chunk.data()));
      count_launch();
      thrust::sort(pol, chunk.begin(), chunk.end());
      auto uend = thrust::unique(pol, chunk.begin(), chunk.end());
      int64_t nu = uend - chunk.begin();
      if(nu > 0 && chunk[nu - 1] == ~0ull) --nu; // drop the sentinel of filtered pairs
      merged.resize(acc.size() + nu);
      // thrust's set operations index with 32 bits: merge key range by key range (both inputs are sorted, so the
      // pieces concatenate), every piece well below 2^31 entries
      {
        const int64_t na = (int64_t)acc.size(), piece = (int64_t)1 << 27;
        const int64_t nseg = std::max<int64_t>(1, (na + nu + piece - 1) / piece);
        int64_t       a0 = 0, c0 = 0, o = 0;
        for(int64_t sgm = 1; sgm <= nseg; ++sgm) {
          int64_t a1 = na, c1 = nu;
          if(sgm < nseg) {
            a1 = std::min(na, sgm * (na / nseg + 1));
            if(a1 < na) {
              const uint64_t split = acc[a1]; // first key of the next piece
              c1 = thrust::lower_bound(pol, chunk.begin() + c0, chunk.begin() + nu, split) - chunk.begin();
            }
          }
          auto end = thrust::set_union(pol, acc.begin() + a0, acc.begin() + a1, chunk.begin() + c0, chunk.begin() + c1, merged.begin() + o);
          o  = end - merged.begin();
          a0 = a1;
          c0 = c1;
        }
        merged.resize(o);
      }
      acc.swap(merged);
      log_stage("pattern: chunk merged");
    }
    chunk.clear();
    chunk.shrink_to_fit();
    merged.clear();
    merged.shrink_to_fit();
    S->nnz = (int64_t)acc.size();
    cudaFree(S->d_ia);
    cudaFree(S->d_ja);
    B200_CUDA(cudaMalloc(&S->d_ia, (size_t)(n_inc + 1) * sizeof(int64_t)));
    B200_CUDA(cudaMalloc(&S->d_ja, (size_t)S->nnz * sizeof(int32_t)));
    thrust::device_vector<uint64_t> q(n_inc + 1);
    row_start_keys_kernel<<<148 * 8, 256, 0, S->stream>>>(n_inc, thrust::raw_pointer_cast(q.data()));
    thrust::lower_bound(pol, acc.begin(), acc.end(), q.begin(), q.end(), thrust::device_pointer_cast(S->d_ia));
    split_keys_kernel<<<148 * 16, 256, 0, S->stream>>>(S->nnz, n_inc, thrust::raw_pointer_cast(acc.data()), S->d_ja);
    count_launch(2);
    B200_CUDA(cudaStreamSynchronize(S->stream));
    log_stage("pattern: ia/ja");
  } catch(const std::exception &ex) {
    set_error(std::string("b200_build_pattern: ") + ex.what());
    return B200_ERR_CUDA;
  }
  S->plan = PLAN_NONE;
  krylov_free(S);
  return alloc_linear_system_public(S);
}

} // namespace b200
