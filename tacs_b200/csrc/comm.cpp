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
#include <string.h>

#include "tb2_host.h"

namespace tb2 {

typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(void **, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_send)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_recv)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_group)(void);
typedef const char *(*fn_errstr)(int);

static struct {
  void *lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_send send = nullptr;
  fn_recv recv = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
  fn_errstr errstr = nullptr;
} nccl;

static const int kNcclFloat64 = 8, kNcclSum = 0;

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
  return nccl_ok(nccl.init_rank(&c.nccl_comm, size, uid, rank), "ncclCommInitRank") ? 0 : 1;
}

int comm_allreduce_sum(double *dev_buf, int n) {
  Context &c = ctx();
  if (c.size <= 1) return 0;
  return nccl_ok(nccl.allreduce(dev_buf, dev_buf, (size_t)n, kNcclFloat64, kNcclSum, c.nccl_comm, c.stream),
                 "ncclAllReduce") ? 0 : 1;
}

// The distributed halo / off-rank-row exchanges are installed by the distributed plan (next
// milestone); on one rank they are never reached.
void halo_forward(TACSAssembler *a, TACSBVec *v) { (void)a; (void)v; }
void residual_exchange(TACSAssembler *a, TACSBVec *res) { (void)a; (void)res; }
void matrix_exchange(TACSAssembler *a, TACSParallelMat *A) { (void)a; (void)A; }
void spmv_halo_begin(TACSParallelMat *A, TACSBVec *x) { (void)A; (void)x; }
void spmv_halo_end(TACSParallelMat *A) { (void)A; }

}  // namespace tb2
