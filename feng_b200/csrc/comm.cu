// Multi-GPU plumbing of the Newton linear solve (north-star subsystem 5): one process per GPU,
// elements partitioned with one ghost layer (owner-computes assembly needs no exchange), rows owned by exactly one rank.
// The only exchange steps of the path are
//   * the halo update of the SpMV input vector (neighbour ncclSend / ncclRecv inside one group), and
//   * the reductions of the Krylov dot products / norms (ncclAllReduce of <= m+2 doubles),
// both enqueued on the system's own stream, so a GMRES iteration still has a single host synchronisation.
//
// The reference's own scheme (replicated mesh, row filtering, MPI_Allgatherv of the full solution every Newton
// iteration: src/feLinearSystemPETSc.cpp:626-640, :1055) is "replicas only" and is not reproduced.
//
// NCCL is resolved at run time from the library already loaded by torch.distributed (dlopen), so that libfeng_b200.so
// loads on machines without NCCL and both sides share one NCCL instance.
#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include <nccl.h>

#include "system.h"

namespace b200 {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static bool load_nccl()
{
  if(g_nccl.lib) return true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if(!h) h = dlopen("libnccl.so.2", RTLD_NOW);
  if(!h) h = dlopen("libnccl.so", RTLD_NOW);
  if(!h) {
    set_error("NCCL library not found (dlopen libnccl.so.2)");
    return false;
  }
#define B200_SYM(field, name)                                                                                 \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));                                    \
  if(!g_nccl.field) {                                                                                         \
    set_error(std::string("NCCL symbol missing: ") + name);                                                   \
    return false;                                                                                             \
  }
  B200_SYM(GetUniqueId, "ncclGetUniqueId")
  B200_SYM(CommInitRank, "ncclCommInitRank")
  B200_SYM(CommDestroy, "ncclCommDestroy")
  B200_SYM(AllReduce, "ncclAllReduce")
  B200_SYM(AllGather, "ncclAllGather")
  B200_SYM(Send, "ncclSend")
  B200_SYM(Recv, "ncclRecv")
  B200_SYM(GroupStart, "ncclGroupStart")
  B200_SYM(GroupEnd, "ncclGroupEnd")
  B200_SYM(GetErrorString, "ncclGetErrorString")
#undef B200_SYM
  g_nccl.lib = h;
  return true;
}

#define B200_NCCL(call)                                                                                       \
  do {                                                                                                        \
    ncclResult_t _r = (call);                                                                                 \
    if(_r != ncclSuccess) {                                                                                   \
      set_error(std::string(#call) + ": " + g_nccl.GetErrorString(_r));                                       \
      return B200_ERR_CUDA;                                                                                   \
    }                                                                                                         \
  } while(0)

struct Comm {
  ncclComm_t comm = nullptr;
  int        rank = 0, world = 1;
  // halo plan
  int                  n_nbr = 0;
  std::vector<int>     nbr;
  std::vector<int64_t> send_ptr, recv_ptr;
  int32_t             *d_send_idx = nullptr, *d_recv_idx = nullptr;
  double              *d_send_buf = nullptr, *d_recv_buf = nullptr;
  double              *d_mask = nullptr; // [nInc] 1 for owned rows, 0 for ghost rows
  int64_t              n_owned = 0;
  // interior / boundary split of the SpMV (rows whose columns are all owned do not wait for the halo)
  uint8_t             *d_bflag = nullptr; // [nInc] 1 = the row reads a ghost column (or is a ghost row)
  int32_t             *d_brows = nullptr; // list of those rows
  int64_t              n_brows = 0;
  cudaStream_t         cstream = nullptr; // halo traffic runs here, next to the interior product on the system's stream
  cudaEvent_t          ev_x = nullptr, ev_halo = nullptr;
};

__global__ void boundary_flag_kernel(int64_t n, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja, const double *__restrict__ mask,
                                     uint8_t *flag)
{
  const int     lane = threadIdx.x & 7;
  const int64_t g0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3, ng = (gridDim.x * (int64_t)blockDim.x) >> 3;
  for(int64_t i = g0; i < n; i += ng) {
    int f = mask[i] == 0. ? 1 : 0;
    for(int64_t k = ia[i] + lane; k < ia[i + 1] && !f; k += 8) f |= mask[ja[k]] == 0. ? 1 : 0;
    f |= __shfl_xor_sync(0xffffffffu, f, 1, 8);
    f |= __shfl_xor_sync(0xffffffffu, f, 2, 8);
    f |= __shfl_xor_sync(0xffffffffu, f, 4, 8);
    if(lane == 0) flag[i] = (uint8_t)f;
  }
}

__global__ void pack_kernel(int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ x, double *__restrict__ buf)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) buf[i] = x[idx[i]];
}

__global__ void unpack_kernel(int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ buf, double *__restrict__ x)
{
  for(int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[idx[i]] = buf[i];
}

void comm_free(System *S)
{
  Comm *C = static_cast<Comm *>(S->comm);
  if(!C) return;
  cudaFree(C->d_send_idx);
  cudaFree(C->d_recv_idx);
  cudaFree(C->d_send_buf);
  cudaFree(C->d_recv_buf);
  cudaFree(C->d_mask);
  cudaFree(C->d_bflag);
  cudaFree(C->d_brows);
  if(C->cstream) cudaStreamDestroy(C->cstream);
  if(C->ev_x) cudaEventDestroy(C->ev_x);
  if(C->ev_halo) cudaEventDestroy(C->ev_halo);
  if(C->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(C->comm);
  delete C;
  S->comm = nullptr;
}

const double *comm_mask(const System *S)
{
  const Comm *C = static_cast<const Comm *>(S->comm);
  return C ? C->d_mask : nullptr;
}

int comm_rank(const System *S)
{
  const Comm *C = static_cast<const Comm *>(S->comm);
  return C ? C->rank : 0;
}

int comm_world(const System *S)
{
  const Comm *C = static_cast<const Comm *>(S->comm);
  return C ? C->world : 1;
}

bool comm_active(const System *S)
{
  const Comm *C = static_cast<const Comm *>(S->comm);
  return C && C->world > 1;
}

static int halo_on_stream(Comm *C, double *d_x, cudaStream_t st)
{
  const int64_t ns = C->send_ptr[C->n_nbr], nr = C->recv_ptr[C->n_nbr];
  if(ns > 0) {
    pack_kernel<<<(unsigned)std::min<int64_t>((ns + 255) / 256, 148 * 8), 256, 0, st>>>(ns, C->d_send_idx, d_x, C->d_send_buf);
    count_launch();
  }
  B200_NCCL(g_nccl.GroupStart());
  for(int k = 0; k < C->n_nbr; ++k) {
    const int64_t s0 = C->send_ptr[k], s1 = C->send_ptr[k + 1], r0 = C->recv_ptr[k], r1 = C->recv_ptr[k + 1];
    if(s1 > s0) B200_NCCL(g_nccl.Send(C->d_send_buf + s0, (size_t)(s1 - s0), ncclDouble, C->nbr[k], C->comm, st));
    if(r1 > r0) B200_NCCL(g_nccl.Recv(C->d_recv_buf + r0, (size_t)(r1 - r0), ncclDouble, C->nbr[k], C->comm, st));
  }
  B200_NCCL(g_nccl.GroupEnd());
  if(nr > 0) {
    unpack_kernel<<<(unsigned)std::min<int64_t>((nr + 255) / 256, 148 * 8), 256, 0, st>>>(nr, C->d_recv_idx, C->d_recv_buf, d_x);
    count_launch();
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

// x[ghost] <- owner's value, on the system's stream
int comm_halo_exchange(System *S, double *d_x)
{
  Comm *C = static_cast<Comm *>(S->comm);
  if(!C || C->world == 1 || C->n_nbr == 0) return B200_OK;
  return halo_on_stream(C, d_x, S->stream);
}

// y = A x with the halo update of x overlapped: the exchange runs on its own stream while the system's stream multiplies the
// rows that read owned columns only; the rows touching ghost columns follow once the halo has landed
int comm_spmv_overlapped(System *S, double *d_x, double *d_y)
{
  Comm *C = static_cast<Comm *>(S->comm);
  if(!C || C->world == 1 || C->n_nbr == 0 || !C->d_bflag) {
    const int rc = comm_halo_exchange(S, d_x);
    return rc != B200_OK ? rc : spmv(S, d_x, d_y);
  }
  B200_CUDA(cudaEventRecord(C->ev_x, S->stream));
  B200_CUDA(cudaStreamWaitEvent(C->cstream, C->ev_x, 0));
  int rc = halo_on_stream(C, d_x, C->cstream);
  if(rc != B200_OK) return rc;
  B200_CUDA(cudaEventRecord(C->ev_halo, C->cstream));
  rc = spmv_rows(S, d_x, d_y, C->d_bflag, nullptr, 0); // interior rows
  if(rc != B200_OK) return rc;
  B200_CUDA(cudaStreamWaitEvent(S->stream, C->ev_halo, 0));
  return spmv_rows(S, d_x, d_y, nullptr, C->d_brows, C->n_brows); // rows that needed the halo
}

int comm_allreduce(System *S, double *d_buf, int count, bool max_op)
{
  Comm *C = static_cast<Comm *>(S->comm);
  if(!C || C->world == 1) return B200_OK;
  B200_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)count, ncclDouble, max_op ? ncclMax : ncclSum, C->comm, S->stream));
  return B200_OK;
}

// every rank contributes `count` 8-byte words; recv holds world * count words in rank order
int comm_allgather64(System *S, const void *send, void *recv, size_t count)
{
  Comm *C = static_cast<Comm *>(S->comm);
  if(!C || C->world == 1) {
    B200_CUDA(cudaMemcpyAsync(recv, send, count * 8, cudaMemcpyDeviceToDevice, S->stream));
    return B200_OK;
  }
  B200_NCCL(g_nccl.AllGather(send, recv, count, ncclUint64, C->comm, S->stream));
  return B200_OK;
}

// rows that read a ghost column (the interior / boundary split of the overlapped SpMV)
void comm_boundary_rows(const System *S, const int32_t **rows, int64_t *n)
{
  const Comm *C = static_cast<const Comm *>(S->comm);
  *rows = C ? C->d_brows : nullptr;
  *n    = C ? C->n_brows : 0;
}

} // namespace b200

using namespace b200;

extern "C" {

int b200_comm_unique_id(char *id128)
{
  if(!id128 || !load_nccl()) return B200_ERR_CUDA;
  ncclUniqueId id;
  B200_NCCL(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return B200_OK;
}

int b200_comm_init(b200_system *s, const char *id128, int rank, int world)
{
  if(!s || !id128 || world < 1 || rank < 0 || rank >= world) {
    set_error("b200_comm_init: bad arguments");
    return B200_ERR_ARG;
  }
  if(cudaSetDevice(s->device) != cudaSuccess) return B200_ERR_CUDA;
  comm_free(s);
  Comm *C  = new Comm;
  C->rank  = rank;
  C->world = world;
  s->comm  = C;
  if(world > 1) {
    if(!load_nccl()) return B200_ERR_CUDA;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    B200_NCCL(g_nccl.CommInitRank(&C->comm, world, id, rank));
  }
  return B200_OK;
}

int b200_set_halo(b200_system *s, const uint8_t *owned, int n_nbr, const int32_t *nbr_rank, const int64_t *send_ptr, const int32_t *send_idx,
                  const int64_t *recv_ptr, const int32_t *recv_idx)
{
  if(!s || !s->comm || !owned || s->nInc == 0) {
    set_error("b200_set_halo: call b200_comm_init and set the pattern first");
    return B200_ERR_ARG;
  }
  if(cudaSetDevice(s->device) != cudaSuccess) return B200_ERR_CUDA;
  Comm *C  = static_cast<Comm *>(s->comm);
  C->n_nbr = n_nbr;
  C->nbr.assign(nbr_rank, nbr_rank + n_nbr);
  C->send_ptr.assign(send_ptr, send_ptr + n_nbr + 1);
  C->recv_ptr.assign(recv_ptr, recv_ptr + n_nbr + 1);
  const int64_t ns = n_nbr ? send_ptr[n_nbr] : 0, nr = n_nbr ? recv_ptr[n_nbr] : 0;
  cudaFree(C->d_send_idx);
  cudaFree(C->d_recv_idx);
  cudaFree(C->d_send_buf);
  cudaFree(C->d_recv_buf);
  cudaFree(C->d_mask);
  C->d_send_idx = C->d_recv_idx = nullptr;
  C->d_send_buf = C->d_recv_buf = nullptr;
  if(ns > 0) {
    B200_CUDA(cudaMalloc(&C->d_send_idx, ns * sizeof(int32_t)));
    B200_CUDA(cudaMalloc(&C->d_send_buf, ns * sizeof(double)));
    B200_CUDA(cudaMemcpy(C->d_send_idx, send_idx, ns * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  if(nr > 0) {
    B200_CUDA(cudaMalloc(&C->d_recv_idx, nr * sizeof(int32_t)));
    B200_CUDA(cudaMalloc(&C->d_recv_buf, nr * sizeof(double)));
    B200_CUDA(cudaMemcpy(C->d_recv_idx, recv_idx, nr * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  std::vector<double> mask(s->nInc);
  C->n_owned = 0;
  for(int64_t i = 0; i < s->nInc; ++i) {
    mask[i] = owned[i] ? 1. : 0.;
    C->n_owned += owned[i] ? 1 : 0;
  }
  B200_CUDA(cudaMalloc(&C->d_mask, (size_t)s->nInc * sizeof(double)));
  B200_CUDA(cudaMemcpy(C->d_mask, mask.data(), (size_t)s->nInc * sizeof(double), cudaMemcpyHostToDevice));
  // interior / boundary rows for the overlapped product
  cudaFree(C->d_bflag);
  cudaFree(C->d_brows);
  C->d_bflag = nullptr;
  C->d_brows = nullptr;
  C->n_brows = 0;
  if(C->world > 1 && n_nbr > 0 && s->d_ia && !getenv("B200_NO_OVERLAP")) {
    B200_CUDA(cudaMalloc(&C->d_bflag, (size_t)s->nInc));
    boundary_flag_kernel<<<148 * 8, 256, 0, s->stream>>>(s->nInc, s->d_ia, s->d_ja, C->d_mask, C->d_bflag);
    count_launch();
    std::vector<uint8_t> flag(s->nInc);
    B200_CUDA(cudaMemcpyAsync(flag.data(), C->d_bflag, (size_t)s->nInc, cudaMemcpyDeviceToHost, s->stream));
    B200_CUDA(cudaStreamSynchronize(s->stream));
    std::vector<int32_t> rows;
    for(int64_t i = 0; i < s->nInc; ++i)
      if(flag[i]) rows.push_back((int32_t)i);
    C->n_brows = (int64_t)rows.size();
    if(C->n_brows > 0) {
      B200_CUDA(cudaMalloc(&C->d_brows, rows.size() * sizeof(int32_t)));
      B200_CUDA(cudaMemcpy(C->d_brows, rows.data(), rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    if(!C->cstream) {
      B200_CUDA(cudaStreamCreateWithFlags(&C->cstream, cudaStreamNonBlocking));
      B200_CUDA(cudaEventCreateWithFlags(&C->ev_x, cudaEventDisableTiming));
      B200_CUDA(cudaEventCreateWithFlags(&C->ev_halo, cudaEventDisableTiming));
    }
  }
  return B200_OK;
}

// test / host hook: exchange the ghost entries of a host vector of n_inc doubles through the device path
int b200_halo_exchange_host(b200_system *s, double *x)
{
  if(!s || !x) return B200_ERR_ARG;
  if(cudaSetDevice(s->device) != cudaSuccess) return B200_ERR_CUDA;
  double *d = nullptr;
  B200_CUDA(cudaMalloc(&d, (size_t)s->nInc * sizeof(double)));
  B200_CUDA(cudaMemcpyAsync(d, x, (size_t)s->nInc * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  int rc = comm_halo_exchange(s, d);
  if(rc == B200_OK) {
    cudaMemcpyAsync(x, d, (size_t)s->nInc * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    if(cudaStreamSynchronize(s->stream) != cudaSuccess) rc = B200_ERR_CUDA;
  }
  cudaFree(d);
  return rc;
}

} // extern "C"
