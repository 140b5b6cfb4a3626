// C1 — the one collective of the data-parallel path: SUM all-reduce of the flat fp32 gradient buffer over the GPUs of a
// node (NCCL over NVLink / NVSwitch), behind the C ABI so that a host without torch.distributed can run data parallel
// (gcs_comm_init / gcs_allreduce_grads, §8e).  The reference itself is single-process (no tf.distribute).
//
// NCCL is bound at run time (dlopen of libnccl.so.2, only the six entry points used): a process that already carries an
// NCCL - PyTorch's bundled one - hands out that same library, a plain C host gets the system one.  Nothing here depends on
// a particular NCCL header; the types below are NCCL's stable ABI (128-byte unique id, opaque communicator, enum values
// ncclFloat32 = 7, ncclFloat64 = 8, ncclSum = 0).
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclUniqueId, int);
typedef int (*CommDestroyFn)(NcclComm);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef const char* (*GetErrorStringFn)(int);
typedef int (*GroupFn)(void);

struct Nccl {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  AllReduceFn all_reduce = nullptr;
  GetErrorStringFn error_string = nullptr;
  GroupFn group_start = nullptr, group_end = nullptr;
  bool ok = false;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      n.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (n.handle) break;
    }
    if (!n.handle) return;
    n.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(n.handle, "ncclGetUniqueId"));
    n.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(n.handle, "ncclCommInitRank"));
    n.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(n.handle, "ncclCommDestroy"));
    n.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(n.handle, "ncclAllReduce"));
    n.error_string = reinterpret_cast<GetErrorStringFn>(dlsym(n.handle, "ncclGetErrorString"));
    n.group_start = reinterpret_cast<GroupFn>(dlsym(n.handle, "ncclGroupStart"));
    n.group_end = reinterpret_cast<GroupFn>(dlsym(n.handle, "ncclGroupEnd"));
    n.ok = n.get_unique_id && n.comm_init_rank && n.comm_destroy && n.all_reduce && n.error_string;
  });
  return n;
}

int nccl_fail(const char* what, int rc) {
  return gcs::fail(GCS_ERR_NCCL, "%s: NCCL error %d (%s)", what, rc, nccl().error_string ? nccl().error_string(rc) : "?");
}

}  // namespace

struct gcs_comm {
  NcclComm comm;
  int rank, world;
};

using namespace gcs;

extern "C" int gcs_comm_unique_id(void* id_host) {
  GCS_CHECK_ARG(id_host, "gcs_comm_unique_id: null pointer");
  if (!nccl().ok) return fail(GCS_ERR_NCCL, "gcs_comm_unique_id: libnccl.so.2 could not be loaded");
  NcclUniqueId id;
  const int rc = nccl().get_unique_id(&id);
  if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id_host, &id, sizeof(id));
  return GCS_OK;
}

extern "C" int gcs_comm_init(const void* id_host, int32_t rank, int32_t world_size, gcs_comm** comm) {
  GCS_CHECK_ARG(id_host && comm && world_size >= 1 && rank >= 0 && rank < world_size, "gcs_comm_init: bad argument");
  if (!nccl().ok) return fail(GCS_ERR_NCCL, "gcs_comm_init: libnccl.so.2 could not be loaded");
  NcclUniqueId id;
  memcpy(&id, id_host, sizeof(id));
  NcclComm c = nullptr;
  const int rc = nccl().comm_init_rank(&c, world_size, id, rank);     // binds to the calling thread's current device
  if (rc != 0) return nccl_fail("ncclCommInitRank", rc);
  *comm = new gcs_comm{c, rank, world_size};
  return GCS_OK;
}

extern "C" int gcs_comm_destroy(gcs_comm* comm) {
  if (!comm) return GCS_OK;
  const int rc = nccl().comm_destroy(comm->comm);
  delete comm;
  return rc == 0 ? GCS_OK : nccl_fail("ncclCommDestroy", rc);
}

extern "C" int gcs_comm_rank(const gcs_comm* comm) { return comm ? comm->rank : -1; }
extern "C" int gcs_comm_world_size(const gcs_comm* comm) { return comm ? comm->world : -1; }

extern "C" int gcs_allreduce_grads(gcs_comm* comm, float* grads, int64_t n, gcs_stream stream) {
  GCS_CHECK_ARG(comm && n >= 0 && (grads || n == 0), "gcs_allreduce_grads: bad argument");
  if (n == 0 || comm->world == 1) return GCS_OK;
  const int rc = nccl().all_reduce(grads, grads, static_cast<size_t>(n), /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm->comm, as_stream(stream));
  return rc == 0 ? GCS_OK : nccl_fail("ncclAllReduce", rc);
}

extern "C" int gcs_allreduce_f64(gcs_comm* comm, double* buf, int64_t n, gcs_stream stream) {
  GCS_CHECK_ARG(comm && n >= 0 && (buf || n == 0), "gcs_allreduce_f64: bad argument");
  if (n == 0 || comm->world == 1) return GCS_OK;
  const int rc = nccl().all_reduce(buf, buf, static_cast<size_t>(n), /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm->comm, as_stream(stream));
  return rc == 0 ? GCS_OK : nccl_fail("ncclAllReduce", rc);
}
