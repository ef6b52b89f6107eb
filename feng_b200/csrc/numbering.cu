// Edge tags of a simplicial mesh on the device ("P2 edge numbering from connectivity").
//
// The reference gives every mesh edge a tag in order of FIRST APPEARANCE while its reader sweeps the boundary triangles and then the
// cells, local edges in the order of src/feTriangle.cpp:3-26 / src/feTetrahedron.h:31 (std::set<Edge> insertion,
// src/feMeshRead.cpp:1317-1338, :1412-1456, :1498, :1613-1614); feNumber then numbers the mid-edge DOFs of a P2 space in tag order
// (src/feNumber.cpp:370-483).  feng_b200/numbering.py restates that with numpy.unique; at T3D(92) (28 M vertex pairs) this is the
// slowest step of the set-up that is left on the host.  Here: pack (min, max) keys, stable sort by key carrying the sweep position,
// one segment per distinct edge, rank the segments by the sweep position of their first member.  Bit-identical to the host version
// (tests/test_gpu_numbering.py).
#include <thrust/device_vector.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>

#include <string>

#include "system.h"

namespace b200 {

__global__ void edge_key_kernel(int64_t n, int64_t nV, const int32_t *__restrict__ pairs, uint64_t *key, int32_t *pos)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = pairs[2 * i], b = pairs[2 * i + 1];
    key[i] = (uint64_t)(a < b ? a : b) * (uint64_t)nV + (uint64_t)(a < b ? b : a);
    pos[i] = (int32_t)i;
  }
}

__global__ void edge_flag_kernel(int64_t n, const uint64_t *__restrict__ key, int32_t *flag)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

// first[seg] = sweep position of the first member of the segment (the sort is stable: it is the member at the segment start)
__global__ void edge_first_kernel(int64_t n, const int32_t *__restrict__ flag, const int32_t *__restrict__ seg, const int32_t *__restrict__ pos,
                                  int32_t *first, int32_t *segid)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if(flag[i]) {
      first[seg[i] - 1] = pos[i];
      segid[seg[i] - 1] = seg[i] - 1;
    }
}

__global__ void edge_rank_kernel(int64_t m, const int32_t *__restrict__ segid_sorted, const int32_t *__restrict__ first_sorted,
                                 const int32_t *__restrict__ pairs, int32_t *rank, int32_t *edges)
{
  for(int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < m; r += (int64_t)gridDim.x * blockDim.x) {
    rank[segid_sorted[r]] = (int32_t)r;
    if(edges) {
      edges[2 * r]     = pairs[2 * (int64_t)first_sorted[r]];
      edges[2 * r + 1] = pairs[2 * (int64_t)first_sorted[r] + 1];
    }
  }
}

__global__ void edge_assign_kernel(int64_t n, const int32_t *__restrict__ seg, const int32_t *__restrict__ pos, const int32_t *__restrict__ rank,
                                   int32_t *edge_of_pair)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    edge_of_pair[pos[i]] = rank[seg[i] - 1];
}

int unique_edges(int device, int64_t nV, int64_t n, const int32_t *h_pairs, int32_t *h_edge_of_pair, int32_t *h_edges, int64_t *n_edges)
{
  if(n <= 0 || nV <= 0 || !h_pairs || !h_edge_of_pair || !n_edges || n >= (int64_t)2147483647) {
    set_error("b200_unique_edges: bad arguments (at most 2^31 - 1 vertex pairs)");
    return B200_ERR_ARG;
  }
  B200_CUDA(cudaSetDevice(device));
  try {
    thrust::device_vector<int32_t>  pairs(h_pairs, h_pairs + 2 * n), pos(n), flag(n), seg(n);
    thrust::device_vector<uint64_t> key(n);
    const int grid = 148 * 8;
    edge_key_kernel<<<grid, 256>>>(n, nV, thrust::raw_pointer_cast(pairs.data()), thrust::raw_pointer_cast(key.data()), thrust::raw_pointer_cast(pos.data()));
    thrust::stable_sort_by_key(key.begin(), key.end(), pos.begin());
    edge_flag_kernel<<<grid, 256>>>(n, thrust::raw_pointer_cast(key.data()), thrust::raw_pointer_cast(flag.data()));
    thrust::inclusive_scan(flag.begin(), flag.end(), seg.begin());
    const int64_t m = (int64_t)(int32_t)seg[n - 1];
    thrust::device_vector<int32_t> first(m), segid(m), rank(m), edges(h_edges ? 2 * m : 0), eop(n);
    edge_first_kernel<<<grid, 256>>>(n, thrust::raw_pointer_cast(flag.data()), thrust::raw_pointer_cast(seg.data()), thrust::raw_pointer_cast(pos.data()),
                                    thrust::raw_pointer_cast(first.data()), thrust::raw_pointer_cast(segid.data()));
    thrust::sort_by_key(first.begin(), first.end(), segid.begin()); // sweep positions are distinct
    edge_rank_kernel<<<grid, 256>>>(m, thrust::raw_pointer_cast(segid.data()), thrust::raw_pointer_cast(first.data()), thrust::raw_pointer_cast(pairs.data()),
                                   thrust::raw_pointer_cast(rank.data()), h_edges ? thrust::raw_pointer_cast(edges.data()) : nullptr);
    edge_assign_kernel<<<grid, 256>>>(n, thrust::raw_pointer_cast(seg.data()), thrust::raw_pointer_cast(pos.data()), thrust::raw_pointer_cast(rank.data()),
                                     thrust::raw_pointer_cast(eop.data()));
    count_launch(5);
    B200_CUDA(cudaMemcpy(h_edge_of_pair, thrust::raw_pointer_cast(eop.data()), (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if(h_edges) B200_CUDA(cudaMemcpy(h_edges, thrust::raw_pointer_cast(edges.data()), (size_t)2 * m * sizeof(int32_t), cudaMemcpyDeviceToHost));
    B200_CUDA(cudaGetLastError());
    *n_edges = m;
  } catch(const std::exception &ex) {
    set_error(std::string("b200_unique_edges: ") + ex.what());
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

} // namespace b200
