// jvmc_comm_* -- the reductions of jVMC/mpi_wrapper.py on NCCL, behind the C ABI.
//
// The reference stages every reduction through the host: pmap psum -> np.array (D2H) -> MPI.Allreduce -> device_put
// (jVMC/mpi_wrapper.py:114-147), gathers with pickle (:278-292) and broadcasts checkpoints with MPI.Bcast (:246-275).
// Here one process drives one GPU and the same operations run in-stream on device buffers over NVLink 5 / NVSwitch:
//   global_sum / global_mean / global_variance / global_covariance  -> jvmc_comm_allreduce_sum_f64 (complex = 2 doubles)
//   gather                                                          -> jvmc_comm_allgather_bytes
//   bcast_unknown_size                                              -> jvmc_comm_bcast_bytes
//   reduce-scatter of the Gram tile rows (distributed solve)        -> jvmc_comm_reduce_scatter_sum_f64
// NCCL is resolved at run time (dlopen of libnccl.so.2: inside a torch process that is the copy torch already
// loaded), so the library has no link-time dependency on it and loads on hosts without NCCL; the entry points then
// return JVMC_ERR_UNSUPPORTED.  Only the host-side bootstrap (unique id exchange) is left to the caller.
#include <dlfcn.h>
#include <stddef.h>
#include <string.h>

#include "common.cuh"

namespace {

// the subset of nccl.h that is used (ABI-stable since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;

struct NcclApi {
  void* handle = nullptr;
  bool tried = false;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;

bool nccl_ready() {
  if (!g_nccl.tried) {
    g_nccl.tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      g_nccl.handle = h;
      *(void**)&g_nccl.GetUniqueId = dlsym(h, "ncclGetUniqueId");
      *(void**)&g_nccl.CommInitRank = dlsym(h, "ncclCommInitRank");
      *(void**)&g_nccl.CommDestroy = dlsym(h, "ncclCommDestroy");
      *(void**)&g_nccl.AllReduce = dlsym(h, "ncclAllReduce");
      *(void**)&g_nccl.AllGather = dlsym(h, "ncclAllGather");
      *(void**)&g_nccl.Broadcast = dlsym(h, "ncclBroadcast");
      *(void**)&g_nccl.ReduceScatter = dlsym(h, "ncclReduceScatter");
      *(void**)&g_nccl.GetVersion = dlsym(h, "ncclGetVersion");
    }
  }
  return g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce && g_nccl.AllGather &&
         g_nccl.Broadcast && g_nccl.ReduceScatter;
}

}  // namespace

/* NCCL version (major*10000 + minor*100 + patch), 0 when no NCCL library can be loaded. */
extern "C" int jvmc_comm_nccl_version(void) {
  if (!nccl_ready() || !g_nccl.GetVersion) return 0;
  int v = 0;
  return g_nccl.GetVersion(&v) == ncclSuccess ? v : 0;
}

/* Rank 0 creates the 128-byte rendezvous id; the caller ships it to the other ranks (any host channel). */
extern "C" int jvmc_comm_unique_id(void* id128) {
  if (!id128) return JVMC_ERR_ARG;
  if (!nccl_ready()) return JVMC_ERR_UNSUPPORTED;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return JVMC_ERR_CUDA;
  memcpy(id128, &id, sizeof(id));
  return JVMC_OK;
}

/* Collective over all ranks: joins the communicator on the calling thread's current CUDA device. */
extern "C" int jvmc_comm_init(const void* id128, int rank, int world, void** comm) {
  if (!id128 || !comm || world <= 0 || rank < 0 || rank >= world) return JVMC_ERR_ARG;
  if (!nccl_ready()) return JVMC_ERR_UNSUPPORTED;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  if (g_nccl.CommInitRank(&c, world, id, rank) != ncclSuccess) return JVMC_ERR_CUDA;
  *comm = (void*)c;
  return JVMC_OK;
}

extern "C" int jvmc_comm_destroy(void* comm) {
  if (!comm) return JVMC_ERR_ARG;
  if (!nccl_ready()) return JVMC_ERR_UNSUPPORTED;
  return g_nccl.CommDestroy((ncclComm_t)comm) == ncclSuccess ? JVMC_OK : JVMC_ERR_CUDA;
}

/* In-place SUM all-reduce of `count` doubles (a complex128 array counts twice its length). */
extern "C" int jvmc_comm_allreduce_sum_f64(void* comm, double* buf, long long count, void* stream) {
  if (!comm || !buf || count < 0) return JVMC_ERR_ARG;
  if (!nccl_ready()) return JVMC_ERR_UNSUPPORTED;
  if (count == 0) return JVMC_OK;
  return g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)comm, (cudaStream_t)stream) == ncclSuccess
             ? JVMC_OK : JVMC_ERR_CUDA;
}

/* recv [world][bytesPerRank] <- every rank's send [bytesPerRank] (equal sizes; ragged gathers pad on the caller side). */
extern "C" int jvmc_comm_allgather_bytes(void* comm, const void* send, void* recv, long long bytesPerRank, void* stream) {
  if (!comm || !send || !recv || bytesPerRank < 0) return JVMC_ERR_ARG;
  if (!nccl_ready()) return JVMC_ERR_UNSUPPORTED;
  if (bytesPerRank == 0) return JVMC_OK;
  return g_nccl.AllGather(send, recv, (size_t)bytesPerRank, ncclInt8, (ncclComm_t)comm, (cudaStream_t)stream) == ncclSuccess
             ? JVMC_OK : JVMC_ERR_CUDA;
}

extern "C" int jvmc_comm_bcast_bytes(void* comm, void* buf, long long bytes, int root, void* stream) {
  if (!comm || !buf || bytes < 0 || root < 0) return JVMC_ERR_ARG;
  if (!nccl_ready()) return JVMC_ERR_UNSUPPORTED;
  if (bytes == 0) return JVMC_OK;
  return g_nccl.Broadcast(buf, buf, (size_t)bytes, ncclInt8, root, (ncclComm_t)comm, (cudaStream_t)stream) == ncclSuccess
             ? JVMC_OK : JVMC_ERR_CUDA;
}

/* recv [recvCount] <- this rank's block of the element-wise SUM over ranks of send [world * recvCount]. */
extern "C" int jvmc_comm_reduce_scatter_sum_f64(void* comm, const double* send, double* recv, long long recvCount,
                                                void* stream) {
  if (!comm || !send || !recv || recvCount < 0) return JVMC_ERR_ARG;
  if (!nccl_ready()) return JVMC_ERR_UNSUPPORTED;
  if (recvCount == 0) return JVMC_OK;
  return g_nccl.ReduceScatter(send, recv, (size_t)recvCount, ncclFloat64, ncclSum, (ncclComm_t)comm,
                              (cudaStream_t)stream) == ncclSuccess ? JVMC_OK : JVMC_ERR_CUDA;
}
