// dist.cu -- the multi-GPU entries of the C ABI (include/slamfe.h, "one box, several GPUs").
//
// The front-end shards only where the reference's data model allows it:
//   * frame pairs of a replayed sequence are independent units: every rank replays its own block
//     (sfe_shard_range + sfe_replay_sequence) and there is NO collective on the data path; the per-feature
//     result rows can be collected with sfe_allgather_rows when the host-side map bookkeeping wants them in
//     one place;
//   * the descriptor matcher has one real exchange (BASELINE config 5): the train set is replicated
//     (ncclBroadcast from the rank that owns it), query rows are sharded, and the per-row top-2 results are
//     all-gathered.  Top-2 is per query row, so no cross-rank reduction exists.
// One context per GPU (one process or one host thread per GPU, as NCCL requires); the collectives are
// enqueued on the context's stream like every kernel, so they are ordered with the matcher without host
// synchronisation.  NCCL is resolved at run time from libnccl.so.2 (the copy the process already loaded, e.g.
// PyTorch's, else the system's): libslamfe.so itself has no link-time dependency on it and every other entry
// works without it.
#include <ctype.h>
#include <dlfcn.h>
#include <sched.h>
#include <nccl.h>  // types and enums only; the functions are looked up with dlsym
#include <stdio.h>
#include <string.h>

#include "ctx.cuh"

struct sfe_dist {
  ncclComm_t comm;
  int rank, world;
  bool own;  // created by sfe_dist_init (destroyed with the context) vs attached by the caller
};

namespace {

struct NcclApi {
  void* lib;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
};

NcclApi* nccl() {
  static NcclApi api;
  static int state = 0;  // 0 untried, 1 ready, -1 unavailable
  if (state == 0) {
    state = -1;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.lib = h;
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
      api.Broadcast = (decltype(api.Broadcast))dlsym(h, "ncclBroadcast");
      api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
      api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
      if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.AllGather && api.GroupStart &&
          api.GroupEnd && api.GetErrorString)
        state = 1;
    }
  }
  return state == 1 ? &api : nullptr;
}

int dfail(sfe_ctx* c, int code, const char* what, const char* detail) {
  if (c) snprintf(c->err, sizeof(c->err), "%s%s%s", what, detail ? ": " : "", detail ? detail : "");
  return code;
}

#define NC(call)                                                                                  \
  do {                                                                                            \
    ncclResult_t r_ = (call);                                                                     \
    if (r_ != ncclSuccess) return dfail(ctx, SFE_ERR_CUDA, #call, nccl()->GetErrorString(r_));     \
  } while (0)
#define DCU(call)                                                                                 \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) return dfail(ctx, SFE_ERR_CUDA, #call, cudaGetErrorString(e_));         \
  } while (0)

void shard(int64_t n, int rank, int world, int64_t* lo, int64_t* hi) {
  const int64_t base = n / world, rem = n % world;
  *lo = rank * base + (rank < rem ? rank : rem);
  *hi = *lo + base + (rank < rem ? 1 : 0);
}

// All-gather of row blocks that already sit at their final place inside `all` (rank r owns rows
// [lo_r, hi_r) of n_total): one ncclAllGather when the blocks are equal, else one in-place broadcast per
// rank inside a group (NCCL fuses the group into one launch).
int gather_in_place(sfe_ctx* ctx, void* all, size_t row_bytes, int64_t n_total) {
  sfe_dist* d = ctx->dist;
  NcclApi* N = nccl();
  if (d->world == 1 || n_total == 0) return SFE_SUCCESS;
  if (n_total % d->world == 0) {
    const size_t bytes = (size_t)(n_total / d->world) * row_bytes;
    NC(N->AllGather((const char*)all + (size_t)d->rank * bytes, all, bytes, ncclUint8, d->comm, ctx->stream));
    return SFE_SUCCESS;
  }
  NC(N->GroupStart());
  for (int r = 0; r < d->world; ++r) {
    int64_t lo, hi;
    shard(n_total, r, d->world, &lo, &hi);
    char* p = (char*)all + (size_t)lo * row_bytes;
    ncclResult_t rc = N->Broadcast(p, p, (size_t)(hi - lo) * row_bytes, ncclUint8, r, d->comm, ctx->stream);
    if (rc != ncclSuccess) {
      N->GroupEnd();
      return dfail(ctx, SFE_ERR_CUDA, "ncclBroadcast", N->GetErrorString(rc));
    }
  }
  NC(N->GroupEnd());
  return SFE_SUCCESS;
}

int need_dist(sfe_ctx* ctx) {
  if (!ctx) return SFE_ERR_INVALID;
  if (!ctx->dist) return dfail(ctx, SFE_ERR_INVALID, "no communicator: call sfe_dist_init or sfe_dist_attach first", nullptr);
  DCU(cudaSetDevice(ctx->device));
  return SFE_SUCCESS;
}

}  // namespace

void sfe_dist_release(sfe_ctx* ctx) {
  if (!ctx || !ctx->dist) return;
  if (ctx->dist->own && nccl()) nccl()->CommDestroy(ctx->dist->comm);
  delete ctx->dist;
  ctx->dist = nullptr;
}

namespace {
// CPUs close to a device, as NVML reports them (nvmlDeviceGetCpuAffinity; libnvidia-ml is loaded at run time) or, when
// NVML is not there, as sysfs does (/sys/bus/pci/devices/<id>/local_cpulist).  Returns the number of CPUs found.
int device_cpus(int device, cpu_set_t* out) {
  CPU_ZERO(out);
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), device) != cudaSuccess) return 0;
  void* h = dlopen("libnvidia-ml.so.1", RTLD_NOW);
  if (h) {
    typedef int (*init_t)(void);
    typedef int (*byid_t)(const char*, void**);
    typedef int (*aff_t)(void*, unsigned, unsigned long*);
    init_t init = (init_t)dlsym(h, "nvmlInit_v2");
    byid_t byid = (byid_t)dlsym(h, "nvmlDeviceGetHandleByPciBusId_v2");
    aff_t aff = (aff_t)dlsym(h, "nvmlDeviceGetCpuAffinity");
    void* dev = nullptr;
    unsigned long words[CPU_SETSIZE / (8 * sizeof(unsigned long))] = {0};
    const unsigned nwords = (unsigned)(sizeof(words) / sizeof(words[0]));
    if (init && byid && aff && init() == 0 && byid(bus, &dev) == 0 && aff(dev, nwords, words) == 0)
      for (unsigned c = 0; c < CPU_SETSIZE; ++c)
        if (words[c / (8 * sizeof(unsigned long))] >> (c % (8 * sizeof(unsigned long))) & 1ul) CPU_SET(c, out);
  }
  if (CPU_COUNT(out) == 0) {
    for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE* f = fopen(path, "r");
    if (f) {  // "0-31,64-95"
      int lo, hi;
      while (fscanf(f, "%d", &lo) == 1) {
        hi = lo;
        int ch = fgetc(f);
        if (ch == '-') {
          if (fscanf(f, "%d", &hi) != 1) break;
          ch = fgetc(f);
        }
        for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c) CPU_SET(c, out);
        if (ch != ',') break;
      }
      fclose(f);
    }
  }
  return CPU_COUNT(out);
}
}  // namespace

extern "C" {

int sfe_bind_host_to_device(int device) {
  cpu_set_t near, cur, both;
  if (device_cpus(device, &near) == 0) return 0;
  if (sched_getaffinity(0, sizeof(cur), &cur) != 0) return 0;
  CPU_AND(&both, &near, &cur);  // never widen what the caller (a cpuset, taskset) allowed
  const int n = CPU_COUNT(&both);
  if (n == 0 || sched_setaffinity(0, sizeof(both), &both) != 0) return 0;
  return n;
}

int sfe_shard_range(int64_t n, int rank, int world, int64_t* lo, int64_t* hi) {
  if (n < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) return SFE_ERR_INVALID;
  shard(n, rank, world, lo, hi);
  return SFE_SUCCESS;
}

int sfe_dist_unique_id(uint8_t* id128) {
  if (!id128 || !nccl()) return SFE_ERR_CUDA;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (nccl()->GetUniqueId(&id) != ncclSuccess) return SFE_ERR_CUDA;
  memcpy(id128, &id, 128);
  return SFE_SUCCESS;
}

int sfe_dist_init(sfe_ctx* ctx, const uint8_t* id128, int rank, int world) {
  if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return dfail(ctx, SFE_ERR_INVALID, "bad sfe_dist_init arguments", nullptr);
  if (!nccl()) return dfail(ctx, SFE_ERR_CUDA, "libnccl.so.2 could not be loaded", dlerror());
  sfe_dist_release(ctx);
  DCU(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm;
  NC(nccl()->CommInitRank(&comm, world, id, rank));
  ctx->dist = new sfe_dist{comm, rank, world, true};
  return SFE_SUCCESS;
}

int sfe_dist_attach(sfe_ctx* ctx, void* nccl_comm, int rank, int world) {
  if (!ctx || !nccl_comm || world < 1 || rank < 0 || rank >= world) return dfail(ctx, SFE_ERR_INVALID, "bad sfe_dist_attach arguments", nullptr);
  if (!nccl()) return dfail(ctx, SFE_ERR_CUDA, "libnccl.so.2 could not be loaded", dlerror());
  sfe_dist_release(ctx);
  ctx->dist = new sfe_dist{(ncclComm_t)nccl_comm, rank, world, false};
  return SFE_SUCCESS;
}

int sfe_dist_shutdown(sfe_ctx* ctx) {
  if (!ctx) return SFE_ERR_INVALID;
  if (ctx->dist) cudaStreamSynchronize(ctx->stream);
  sfe_dist_release(ctx);
  return SFE_SUCCESS;
}

int sfe_allgather_rows_dev(sfe_ctx* ctx, const void* local_dev, size_t row_bytes, int64_t n_total, void* all_dev) {
  int rc = need_dist(ctx);
  if (rc) return rc;
  if (n_total < 0 || row_bytes == 0 || !all_dev) return dfail(ctx, SFE_ERR_INVALID, "bad sfe_allgather_rows arguments", nullptr);
  int64_t lo, hi;
  shard(n_total, ctx->dist->rank, ctx->dist->world, &lo, &hi);
  char* mine = (char*)all_dev + (size_t)lo * row_bytes;
  if (hi > lo && local_dev && local_dev != mine)
    DCU(cudaMemcpyAsync(mine, local_dev, (size_t)(hi - lo) * row_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return gather_in_place(ctx, all_dev, row_bytes, n_total);
}

int sfe_match_hamming256_sharded_dev(sfe_ctx* ctx, const uint32_t* q_local, int64_t nq_total, uint32_t* t, int nt,
                                     int train_root, int ratio_num, int ratio_den, int max_dist, int32_t* idx_all,
                                     int32_t* dist_all, uint8_t* pass_all) {
  int rc = need_dist(ctx);
  if (rc) return rc;
  sfe_dist* d = ctx->dist;
  if (nq_total < 0 || nt < 0 || nt > (1 << 22) || !idx_all || !dist_all || (nt && !t) || train_root < 0 || train_root >= d->world)
    return dfail(ctx, SFE_ERR_INVALID, "bad sfe_match_hamming256_sharded arguments", nullptr);
  int64_t lo, hi;
  shard(nq_total, d->rank, d->world, &lo, &hi);
  if (hi - lo > 0x7fffffff || (hi > lo && !q_local)) return dfail(ctx, SFE_ERR_INVALID, "bad query block", nullptr);
  int local_rc = SFE_SUCCESS;
  // 1. replicate the train set (32 B per descriptor) from the rank that owns it
  if (d->world > 1 && nt > 0) NC(nccl()->Broadcast(t, t, (size_t)nt * 32, ncclUint8, train_root, d->comm, ctx->stream));
  // 2. this rank's query rows against the whole train set, written straight into their rows of the gathered arrays
  if (hi > lo) {
    if (ctx->ham_pending) DCU(cudaStreamWaitEvent(ctx->stream, ctx->ham_done, 0));
    int nl = launch_hamming256(q_local, (int)(hi - lo), t, nt, 1, ratio_num, ratio_den, max_dist, idx_all + 2 * lo, dist_all + 2 * lo,
                               pass_all ? pass_all + lo : nullptr, &ctx->ham_ws, &ctx->ham_cap, ctx->stream);
    if (nl < 0) {
      // a rank that cannot match still enters the collectives below -- its peers would otherwise wait in the all-gather
      // for ever -- with its rows marked invalid (idx = dist = -1, pass = 0), and reports the error when they are done
      local_rc = dfail(ctx, SFE_ERR_CUDA, "hamming launch", cudaGetErrorString((cudaError_t)(-nl)));
      cudaMemsetAsync(idx_all + 2 * lo, 0xff, 8 * (size_t)(hi - lo), ctx->stream);
      cudaMemsetAsync(dist_all + 2 * lo, 0xff, 8 * (size_t)(hi - lo), ctx->stream);
      if (pass_all) cudaMemsetAsync(pass_all + lo, 0, (size_t)(hi - lo), ctx->stream);
    } else {
      ctx->launches += nl;
    }
  }
  // 3. all-gather the result rows in place
  if ((rc = gather_in_place(ctx, idx_all, 8, nq_total))) return rc;
  if ((rc = gather_in_place(ctx, dist_all, 8, nq_total))) return rc;
  if (pass_all && (rc = gather_in_place(ctx, pass_all, 1, nq_total))) return rc;
  return local_rc;
}

int sfe_match_hamming256_sharded(sfe_ctx* ctx, const uint32_t* q_local, int64_t nq_total, const uint32_t* t, int nt,
                                 int train_root, int ratio_num, int ratio_den, int max_dist, int32_t* idx_all,
                                 int32_t* dist_all, uint8_t* pass_all) {
  int rc = need_dist(ctx);
  if (rc) return rc;
  sfe_dist* d = ctx->dist;
  if (nq_total < 0 || nt < 0 || !idx_all || !dist_all || (d->rank == train_root && nt && !t))
    return dfail(ctx, SFE_ERR_INVALID, "bad sfe_match_hamming256_sharded arguments", nullptr);
  int64_t lo, hi;
  shard(nq_total, d->rank, d->world, &lo, &hi);
  const size_t qb = 32 * (size_t)(hi - lo), tb = 32 * (size_t)nt, ob = 8 * (size_t)nq_total;
  auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t need = pad(qb) + pad(tb) + 2 * pad(ob) + pad((size_t)nq_total);
  if (need > ctx->scratch_cap) {
    if (ctx->scratch) DCU(cudaFree(ctx->scratch));
    ctx->scratch = nullptr;
    ctx->scratch_cap = 0;
    cudaError_t e = cudaMalloc(&ctx->scratch, need);
    if (e != cudaSuccess) return dfail(ctx, SFE_ERR_NOMEM, "cudaMalloc(scratch)", cudaGetErrorString(e));
    ctx->scratch_cap = need;
  }
  char* base = (char*)ctx->scratch;
  uint32_t* d_q = (uint32_t*)base; base += pad(qb);
  uint32_t* d_t = (uint32_t*)base; base += pad(tb);
  int32_t* d_i = (int32_t*)base; base += pad(ob);
  int32_t* d_d = (int32_t*)base; base += pad(ob);
  uint8_t* d_p = (uint8_t*)base;
  cudaStream_t s = ctx->stream;
  if (qb) DCU(cudaMemcpyAsync(d_q, q_local, qb, cudaMemcpyHostToDevice, s));
  if (d->rank == train_root && tb) DCU(cudaMemcpyAsync(d_t, t, tb, cudaMemcpyHostToDevice, s));
  rc = sfe_match_hamming256_sharded_dev(ctx, d_q, nq_total, d_t, nt, train_root, ratio_num, ratio_den, max_dist, d_i, d_d,
                                        pass_all ? d_p : nullptr);
  if (rc) return rc;
  DCU(cudaMemcpyAsync(idx_all, d_i, ob, cudaMemcpyDeviceToHost, s));
  DCU(cudaMemcpyAsync(dist_all, d_d, ob, cudaMemcpyDeviceToHost, s));
  if (pass_all) DCU(cudaMemcpyAsync(pass_all, d_p, (size_t)nq_total, cudaMemcpyDeviceToHost, s));
  DCU(cudaStreamSynchronize(s));
  return SFE_SUCCESS;
}

}  // extern "C"
