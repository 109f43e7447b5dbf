// Inter-GPU communication: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// Replaces the MPI calls of the reference's hot path:
//   TACSBVecDistribute::beginForward/endForward   src/bpmat/TACSBVecDistribute.cpp:543-640   (halo gather)
//   TACSBVecDistribute::beginReverse/endReverse   :651-743                                  (halo scatter-add)
//   TACSMatDistribute::beginAssembly/endAssembly  src/bpmat/TACSMatDistribute.cpp:1154-1267 (off-rank rows)
//   TACSBVec::norm/dot/mdot MPI_Allreduce         src/bpmat/TACSBVec.cpp:220, 268, 324
// NCCL is loaded with dlopen at communicator creation so that a single-GPU process has no NCCL
// dependency and a torch process shares the copy torch already mapped.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tb2_host.h"

namespace tb2 {

typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(void **, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_send)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_recv)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_group)(void);
typedef int (*fn_allgather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef const char *(*fn_errstr)(int);

static struct {
  void *lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_send send = nullptr;
  fn_recv recv = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
  fn_allgather allgather = nullptr;
  fn_errstr errstr = nullptr;
} nccl;

static const int kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;

static int load_nccl() {
  if (nccl.lib) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (nccl.lib) break;
  }
  if (!nccl.lib) {
    fprintf(stderr, "tacs_b200: cannot load libnccl.so.2 (%s)\n", dlerror());
    return 1;
  }
  nccl.get_uid = (fn_get_uid)dlsym(nccl.lib, "ncclGetUniqueId");
  nccl.init_rank = (fn_init_rank)dlsym(nccl.lib, "ncclCommInitRank");
  nccl.allreduce = (fn_allreduce)dlsym(nccl.lib, "ncclAllReduce");
  nccl.send = (fn_send)dlsym(nccl.lib, "ncclSend");
  nccl.recv = (fn_recv)dlsym(nccl.lib, "ncclRecv");
  nccl.group_start = (fn_group)dlsym(nccl.lib, "ncclGroupStart");
  nccl.group_end = (fn_group)dlsym(nccl.lib, "ncclGroupEnd");
  nccl.errstr = (fn_errstr)dlsym(nccl.lib, "ncclGetErrorString");
  nccl.allgather = (fn_allgather)dlsym(nccl.lib, "ncclAllGather");
  if (!nccl.get_uid || !nccl.init_rank || !nccl.allreduce || !nccl.send || !nccl.recv || !nccl.group_start ||
      !nccl.group_end) {
    fprintf(stderr, "tacs_b200: libnccl is missing required symbols\n");
    return 1;
  }
  return 0;
}

static bool nccl_ok(int rc, const char *what) {
  if (rc == 0) return true;
  fprintf(stderr, "tacs_b200: NCCL error in %s: %s\n", what, nccl.errstr ? nccl.errstr(rc) : "?");
  return false;
}

int comm_unique_id(unsigned char id[128]) {
  if (load_nccl()) return 1;
  nccl_uid uid;
  if (!nccl_ok(nccl.get_uid(&uid), "ncclGetUniqueId")) return 1;
  memcpy(id, uid.internal, 128);
  return 0;
}

int comm_peer_init();
int comm_allreduce_sum(double *dev_buf, int n);

int comm_init(int rank, int size, const unsigned char id[128]) {
  Context &c = ctx();
  if (c.device < 0) {
    fprintf(stderr, "tacs_b200: call tacsb200_init before tacsb200_comm_init\n");
    return 1;
  }
  c.rank = rank;
  c.size = size;
  if (size <= 1) return 0;
  if (load_nccl()) return 1;
  nccl_uid uid;
  memcpy(uid.internal, id, 128);
  if (!nccl_ok(nccl.init_rank(&c.nccl_comm, size, uid, rank), "ncclCommInitRank")) return 1;
  comm_peer_init();  // optional: failure leaves the NCCL reductions in place
  return 0;
}

// Peer-mapped exchange buffers for the fused scalar all-reduce (kernels.h PeerExchange): every rank allocates a small
// buffer, the CUDA IPC handles travel through one ncclAllGather, and each rank maps the buffers of all the others
// (NVLink / NVSwitch peer access within the node). TACSB200_NO_PEER=1 keeps the NCCL path.
int comm_peer_init() {
  Context &c = ctx();
  if (c.size <= 1 || c.size > PeerExchange::kMaxRanks || c.d_peer || !nccl.allgather || getenv("TACSB200_NO_PEER"))
    return 1;
  const size_t nval = (size_t)PeerExchange::kSlots * c.size;
  // layout of a rank's buffer: vals[kSlots * size] doubles | flags[kSlots * size] u64 | produced | consumed
  const size_t bytes = nval * 8 + nval * 8 + 16;
  // A rank whose own buffer or handle cannot be made still takes part in the two collectives below (with a zeroed
  // handle, which its peers fail to open): leaving early on one rank would leave the others waiting in ncclAllGather.
  unsigned char *mine = nullptr;
  bool ok = true;
  if (cudaMalloc(&mine, bytes) != cudaSuccess) { cudaGetLastError(); mine = nullptr; ok = false; }
  if (ok) cudaMemset(mine, 0, bytes);
  cudaIpcMemHandle_t handle;
  memset(&handle, 0, sizeof(handle));
  if (ok && cudaIpcGetMemHandle(&handle, mine) != cudaSuccess) { cudaGetLastError(); ok = false; }
  unsigned char *d_handles = nullptr;
  const size_t hb = sizeof(cudaIpcMemHandle_t);
  if (cudaMalloc(&d_handles, hb * (c.size + 1)) != cudaSuccess) {   // nothing else would work on this rank either
    cudaGetLastError();
    if (mine) cudaFree(mine);
    return 1;
  }
  cudaMemcpy(d_handles + hb * c.size, &handle, hb, cudaMemcpyHostToDevice);
  const int kNcclChar = 0;
  const bool gathered = nccl_ok(nccl.allgather(d_handles + hb * c.size, d_handles, hb, kNcclChar, c.nccl_comm, c.stream.s),
                                "ncclAllGather") &&
                        cuda_ok(cudaStreamSynchronize(c.stream.s), "peer handle exchange");
  ok = ok && gathered;
  std::vector<cudaIpcMemHandle_t> all(c.size);
  if (gathered) cudaMemcpy(all.data(), d_handles, hb * c.size, cudaMemcpyDeviceToHost);
  cudaFree(d_handles);
  PeerExchange px;
  memset(&px, 0, sizeof(px));
  px.rank = c.rank;
  px.size = c.size;
  for (int p = 0; p < c.size && ok; p++) {
    unsigned char *base = mine;
    if (p != c.rank) {
      void *mapped = nullptr;
      if (cudaIpcOpenMemHandle(&mapped, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = false;
        break;
      }
      base = static_cast<unsigned char *>(mapped);
    }
    px.vals[p] = reinterpret_cast<double *>(base);
    px.flags[p] = reinterpret_cast<unsigned long long *>(base + nval * 8);
  }
  // every rank must agree on the outcome: a rank that failed to map a peer would otherwise wait for flags for ever
  double *d_flag = nullptr;
  double h_flag = ok ? 0.0 : 1.0;
  if (cudaMalloc(&d_flag, sizeof(double)) == cudaSuccess) {
    cudaMemcpy(d_flag, &h_flag, sizeof(double), cudaMemcpyHostToDevice);
    comm_allreduce_sum(d_flag, 1);
    cudaStreamSynchronize(c.stream.s);
    cudaMemcpy(&h_flag, d_flag, sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_flag);
  } else {
    cudaGetLastError();
    h_flag = 1.0;
  }
  if (h_flag != 0.0) {
    if (c.rank == 0) fprintf(stderr, "tacs_b200: peer-mapped exchange buffers unavailable; reductions stay on NCCL\n");
    return 1;
  }
  px.produced = reinterpret_cast<unsigned long long *>(mine + 2 * nval * 8);
  px.consumed = px.produced + 1;
  if (cudaMalloc(&c.d_peer, sizeof(PeerExchange)) != cudaSuccess) { cudaGetLastError(); c.d_peer = nullptr; return 1; }
  cudaMemcpy(c.d_peer, &px, sizeof(PeerExchange), cudaMemcpyHostToDevice);
  return 0;
}

int comm_allreduce_sum(double *dev_buf, int n) {
  Context &c = ctx();
  if (c.size <= 1) return 0;
  return nccl_ok(nccl.allreduce(dev_buf, dev_buf, (size_t)n, kNcclFloat64, kNcclSum, c.nccl_comm, c.stream),
                 "ncclAllReduce") ? 0 : 1;
}

int comm_allreduce_max(double *dev_buf, int n) {
  Context &c = ctx();
  if (c.size <= 1) return 0;
  return nccl_ok(nccl.allreduce(dev_buf, dev_buf, (size_t)n, kNcclFloat64, kNcclMax, c.nccl_comm, c.stream),
                 "ncclAllReduce") ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------
// neighbour exchanges
// ---------------------------------------------------------------------------------------------
int comm_setup_exchange(DeviceExchange &dx, const ExchangePlan &x) {
  dx.send_peers = x.send_peers;
  dx.send_ptr = x.send_ptr;
  dx.recv_peers = x.recv_peers;
  dx.recv_ptr = x.recv_ptr;
  if (!x.send_idx.empty() && !dx.d_send_idx.upload(x.send_idx)) return 1;
  return 0;
}

// pack src[send_idx] (chunks of `chunk` doubles) and exchange with the neighbours; peer p's chunks
// land at recv_base(k) for the k-th receive peer. All on stream `s`.
template <class RecvBase>
static int exchange(DeviceExchange &dx, int chunk, const double *src, RecvBase recv_base, cudaStream_t s) {
  Context &c = ctx();
  const long ns = dx.sendTotal();
  if (ns > 0) {
    if (dx.sendbuf.count < (size_t)ns * chunk && !dx.sendbuf.alloc((size_t)ns * chunk)) return 1;
    if (s == c.stream.s) {
      KernelTimer kt(K_HALO);  // event timing only makes sense on the stream the events are recorded on
      if (!cuda_ok(launch_pack_blocks(chunk, ns, dx.d_send_idx.ptr, src, dx.sendbuf.ptr, c.num_sms, s), "halo pack"))
        return 1;
    } else {
      c.kernel_launches++;
      if (!cuda_ok(launch_pack_blocks(chunk, ns, dx.d_send_idx.ptr, src, dx.sendbuf.ptr, c.num_sms, s), "halo pack"))
        return 1;
    }
  }
  if (!nccl_ok(nccl.group_start(), "ncclGroupStart")) return 1;
  for (size_t k = 0; k < dx.send_peers.size(); k++) {
    const size_t off = (size_t)dx.send_ptr[k] * chunk, cnt = (size_t)(dx.send_ptr[k + 1] - dx.send_ptr[k]) * chunk;
    if (!nccl_ok(nccl.send(dx.sendbuf.ptr + off, cnt, kNcclFloat64, dx.send_peers[k], c.nccl_comm, s), "ncclSend"))
      return 1;
  }
  for (size_t k = 0; k < dx.recv_peers.size(); k++) {
    const size_t cnt = (size_t)(dx.recv_ptr[k + 1] - dx.recv_ptr[k]) * chunk;
    if (!nccl_ok(nccl.recv(recv_base((int)k), cnt, kNcclFloat64, dx.recv_peers[k], c.nccl_comm, s), "ncclRecv"))
      return 1;
  }
  return nccl_ok(nccl.group_end(), "ncclGroupEnd") ? 0 : 1;
}

// TACSBVecDistribute::beginForward/endForward for the assembler's external nodes: the ext blocks of a
// local-order vector are [ext < range | owned | ext >= range]; each owner's nodes are contiguous.
int halo_forward(TACSAssembler *a, TACSBVec *v) {
  if (a->size <= 1) return 0;
  DeviceExchange &dx = a->x_state;
  const int chunk = v->bsize, eb = a->ext_before, no = a->nowned;
  return exchange(dx, chunk, v->owned(),
           [&](int k) {
             const int first = dx.recv_ptr[k];  // position in the sorted external list
             const long local = first < eb ? first : (long)no + first;
             return v->local() + local * chunk;
           },
           ctx().stream);
}

// Off-rank rows (TACSMatDistribute::beginAssembly/endAssembly, TACSBVecDistribute reverse): rows of the
// residual staging (and the matching block rows of the matrix staging) that belong to nodes owned by a
// neighbour are sent to it and land in the tail of its staging arrays, where its gather plans find them.
int staging_exchange(TACSAssembler *a, bool with_blocks) {
  if (a->size <= 1) return 0;
  HostPlan &P = *a->plan;
  const int bs = a->bs;
  if (exchange(a->x_rows, bs, a->Re.ptr,
               [&](int k) { return a->Re.ptr + ((size_t)P.local_node_slots + a->x_rows.recv_ptr[k]) * bs; },
               ctx().stream))
    return 1;
  if (with_blocks)
    return exchange(a->x_blocks, bs * bs, a->Ke.ptr,
                    [&](int k) { return a->Ke.ptr + ((size_t)P.local_blocks + a->x_blocks.recv_ptr[k]) * bs * bs; },
                    ctx().stream);
  return 0;
}

// SpMV halo (TACSParallelMat::mult, TACSParallelMat.cpp:248-265): gather the external columns on the
// communication stream while the compute stream runs the local product.
static cudaEvent_t ev_x_ready = nullptr, ev_halo_done = nullptr;
int spmv_halo_begin(TACSParallelMat *A, TACSBVec *x) {
  Context &c = ctx();
  if (!ev_x_ready) {
    cudaEventCreateWithFlags(&ev_x_ready, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev_halo_done, cudaEventDisableTiming);
  }
  cudaEventRecord(ev_x_ready, c.stream);
  cudaStreamWaitEvent(c.comm_stream, ev_x_ready, 0);
  DeviceExchange &dx = A->x_cols;
  const int chunk = A->Aloc.bsize;
  const int rc = exchange(dx, chunk, x->owned(), [&](int k) { return A->x_ext.ptr + (size_t)dx.recv_ptr[k] * chunk; },
                          c.comm_stream);
  cudaEventRecord(ev_halo_done, c.comm_stream);
  return rc;
}
void spmv_halo_end(TACSParallelMat *A) {
  (void)A;
  cudaStreamWaitEvent(ctx().stream, ev_halo_done, 0);
}

}  // namespace tb2
