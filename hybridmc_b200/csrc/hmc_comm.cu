// hmc_comm.cu — the path's one exchange step over NVLink: NCCL all-reduce (u64 sum) of the
// per-state counts and of the first-passage-time histograms, in place on the engine's
// accumulator blocks and on the engine's stream, callable from a plain C/C++ host
// (no torch).  SURVEY.md §8(e); replaces the reference's "one process per transition,
// results meet in files" (tools/init_json.py:14-36) when one transition's replicas are
// spread over several GPUs.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2", or $HMC_NCCL_LIB): a single-GPU host
// needs no NCCL at all, and inside a Python process that already carries torch's own NCCL
// the library keeps its private copy (RTLD_LOCAL), so the two never share state.
#include "hmc_internal.h"

#include <dlfcn.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

namespace {

// the slice of nccl.h this file needs (ABI-stable since NCCL 2.0)
typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclUint64 = 5, ncclFloat64 = 8, ncclUint8 = 1 };
enum { ncclSum = 0, ncclMax = 2, ncclMin = 3 };

struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int *) = nullptr;
  std::string err;
};

NcclApi &nccl() {
  static NcclApi api;
  if (api.lib || !api.err.empty()) return api;
  const char *names[] = {getenv("HMC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    if (!n || !*n) continue;
    api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (api.lib) break;
  }
  if (!api.lib) {
    api.err = std::string("NCCL not found (dlopen libnccl.so.2: ") + dlerror() +
              "); set HMC_NCCL_LIB";
    return api;
  }
#define SYM(field, name)                                                                  \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name));                \
  if (!api.field) api.err = std::string("NCCL symbol missing: ") + name;
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Broadcast, "ncclBroadcast");
  SYM(GetErrorString, "ncclGetErrorString");
  SYM(GetVersion, "ncclGetVersion");
#undef SYM
  return api;
}

int cfail(const std::string &msg, int code = -7) {
  hmc_set_global_error(msg.c_str());
  return code;
}

} // namespace

struct hmc_comm_s {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  cudaStream_t stream = nullptr; // for the host-buffer helpers
  void *d_stage = nullptr;
  size_t stage_bytes = 0;
  uint64_t n_collectives = 0;
  std::string id_file; // rank 0 of hmc_comm_create_from_file: removed again on destroy
  bool borrowed = false; // hmc_comm_wrap_nccl: the ncclComm_t belongs to the caller
};

#define NC(call)                                                                          \
  do {                                                                                    \
    const int r_ = (call);                                                                \
    if (r_ != ncclSuccess)                                                                \
      return cfail(std::string(#call) + ": " + nccl().GetErrorString(r_));                \
  } while (0)
#define CU(call)                                                                          \
  do {                                                                                    \
    const cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) return cfail(std::string(#call) + ": " + cudaGetErrorString(e_), -2); \
  } while (0)

extern "C" {

int hmc_comm_unique_id(uint8_t id[HMC_COMM_ID_BYTES]) {
  NcclApi &api = nccl();
  if (!api.err.empty()) return cfail(api.err);
  ncclUniqueId u;
  NC(api.GetUniqueId(&u));
  std::memcpy(id, u.internal, sizeof u.internal);
  return 0;
}

int hmc_comm_create(const uint8_t id[HMC_COMM_ID_BYTES], int rank, int world, int device,
                    hmc_comm *out) {
  if (!id || !out || world < 1 || rank < 0 || rank >= world) return cfail("hmc_comm_create: bad arguments", -1);
  *out = nullptr;
  NcclApi &api = nccl();
  if (!api.err.empty()) return cfail(api.err);
  CU(cudaSetDevice(device));
  hmc_comm c = new hmc_comm_s();
  c->rank = rank;
  c->world = world;
  c->device = device;
  ncclUniqueId u;
  std::memcpy(u.internal, id, sizeof u.internal);
  const int r = api.CommInitRank(&c->comm, world, u, rank);
  if (r != ncclSuccess) {
    delete c;
    return cfail(std::string("ncclCommInitRank: ") + api.GetErrorString(r));
  }
  cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  *out = c;
  return 0;
}

// Rendezvous through a file on a filesystem all ranks see (single node): rank 0 writes the
// id to `path` (temporary file + rename), the others wait for it.  For hosts that have no
// other way to pass 128 bytes around (the C++ `hybridmc` executable started N times).
int hmc_comm_create_from_file(const char *path, int rank, int world, int device, hmc_comm *out) {
  if (!path || !out) return cfail("hmc_comm_create_from_file: null", -1);
  uint8_t id[HMC_COMM_ID_BYTES];
  if (rank == 0) {
    const int rc = hmc_comm_unique_id(id);
    if (rc) return rc;
    const std::string tmp = std::string(path) + ".tmp";
    FILE *f = std::fopen(tmp.c_str(), "wb");
    if (!f) return cfail(std::string("cannot write ") + tmp);
    std::fwrite(id, 1, sizeof id, f);
    std::fclose(f);
    if (std::rename(tmp.c_str(), path) != 0) return cfail(std::string("cannot rename to ") + path);
  } else {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
      FILE *f = std::fopen(path, "rb");
      if (f) {
        const size_t n = std::fread(id, 1, sizeof id, f);
        std::fclose(f);
        if (n == sizeof id) break;
      }
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120))
        return cfail(std::string("timed out waiting for the NCCL id file ") + path);
      std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
  }
  const int rc = hmc_comm_create(id, rank, world, device, out);
  // (a stale id file would make the other ranks of the NEXT launch join a dead group)
  if (rc == 0 && rank == 0) (*out)->id_file = path;
  return rc;
}

// Adopt a communicator the host already has (ncclCommInitRank / MPI-bootstrapped):
// hmc_allreduce_counts(h, wrapped) is then SURVEY section 8(b)'s hmc_allreduce_counts(h,
// ncclComm_t).  The ncclComm_t must belong to the libnccl.so.2 this library resolves
// (dlopen by soname returns the copy the process has already loaded).  Not destroyed here.
int hmc_comm_wrap_nccl(void *nccl_comm, int rank, int world, int device, hmc_comm *out) {
  if (!nccl_comm || !out || world < 1 || rank < 0 || rank >= world)
    return cfail("hmc_comm_wrap_nccl: bad arguments", -1);
  *out = nullptr;
  NcclApi &api = nccl();
  if (!api.err.empty()) return cfail(api.err);
  CU(cudaSetDevice(device));
  hmc_comm c = new hmc_comm_s();
  c->comm = static_cast<ncclComm_t>(nccl_comm);
  c->rank = rank;
  c->world = world;
  c->device = device;
  c->borrowed = true;
  cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  *out = c;
  return 0;
}

int hmc_comm_destroy(hmc_comm c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->comm && !c->borrowed) nccl().CommDestroy(c->comm);
  if (c->d_stage) cudaFree(c->d_stage);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (!c->id_file.empty()) std::remove(c->id_file.c_str());
  delete c;
  return 0;
}

int hmc_comm_rank(hmc_comm c) { return c ? c->rank : 0; }
int hmc_comm_world(hmc_comm c) { return c ? c->world : 1; }
uint64_t hmc_comm_collectives(hmc_comm c) { return c ? c->n_collectives : 0; }

int hmc_comm_nccl_version(void) {
  NcclApi &api = nccl();
  if (!api.err.empty()) return cfail(api.err);
  int v = 0;
  api.GetVersion(&v);
  return v;
}

// u64 sum over all ranks, in place on device memory, ordered on `stream`
int hmc_comm_allreduce_device_u64(hmc_comm c, void *dptr, uint64_t n, void *stream) {
  if (!c || c->world == 1 || n == 0) return 0;
  NC(nccl().AllReduce(dptr, dptr, size_t(n), ncclUint64, ncclSum, c->comm,
                      static_cast<cudaStream_t>(stream)));
  c->n_collectives++;
  return 0;
}

namespace {
int stage(hmc_comm c, size_t bytes) {
  if (bytes <= c->stage_bytes) return 0;
  CU(cudaStreamSynchronize(c->stream));
  if (c->d_stage) cudaFree(c->d_stage);
  c->d_stage = nullptr;
  CU(cudaMalloc(&c->d_stage, bytes));
  c->stage_bytes = bytes;
  return 0;
}
} // namespace

// Host-buffer helpers for the O(states) bookkeeping of the phase loop (flags, the exact
// fixed-point sums of the Wang-Landau estimates, start structures).  Blocking.
int hmc_comm_allreduce_host_u64(hmc_comm c, uint64_t *buf, uint64_t n, int op /*0 sum, 1 max, 2 min*/) {
  if (!c || c->world == 1 || n == 0) return 0;
  CU(cudaSetDevice(c->device));
  int rc = stage(c, n * 8);
  if (rc) return rc;
  CU(cudaMemcpyAsync(c->d_stage, buf, n * 8, cudaMemcpyHostToDevice, c->stream));
  NC(nccl().AllReduce(c->d_stage, c->d_stage, size_t(n), ncclUint64,
                      op == 1 ? ncclMax : (op == 2 ? ncclMin : ncclSum), c->comm, c->stream));
  CU(cudaMemcpyAsync(buf, c->d_stage, n * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->n_collectives++;
  return 0;
}

int hmc_comm_bcast_host(hmc_comm c, void *buf, uint64_t nbytes, int root) {
  if (!c || c->world == 1 || nbytes == 0) return 0;
  CU(cudaSetDevice(c->device));
  int rc = stage(c, nbytes);
  if (rc) return rc;
  if (c->rank == root) CU(cudaMemcpyAsync(c->d_stage, buf, nbytes, cudaMemcpyHostToDevice, c->stream));
  NC(nccl().Broadcast(c->d_stage, c->d_stage, size_t(nbytes), ncclUint8, root, c->comm, c->stream));
  CU(cudaMemcpyAsync(buf, c->d_stage, nbytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->n_collectives++;
  return 0;
}

} // extern "C"
