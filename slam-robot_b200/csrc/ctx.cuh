// ctx.cuh -- the objects behind the opaque handles of include/slamfe.h (shared by capi.cu and replay.cu).
#pragma once
#include "sfe_common.cuh"

struct sfe_replay;  // replay.cu: the chunk pipeline's streams, events and staging buffers
struct sfe_dist;    // dist.cu: the NCCL communicator of the multi-GPU entries

struct sfe_ctx {
  int device;
  cudaStream_t own_stream;
  cudaStream_t stream;
  float* d_mask;
  // work-queue heads of the persistent tracking kernels: a ring of 64-byte slots, one per launch, so that trackers
  // enqueued on different streams of one context (sfe_set_stream, the replay pipeline) never share a counter
  int* d_counter;
  unsigned counter_seq;
  int num_sms;
  float h_mask[SFE_PLEN];
  // grow-on-demand device scratch for the host-pointer entry points
  void* scratch;
  size_t scratch_cap;
  void* h_stage;   // pinned host mirror of the scratch head: small host-pointer calls move their arrays in two copies
  size_t h_stage_cap;
  void* ham_ws;
  size_t ham_cap;
  void* ham_io;   // device buffers of sfe_match_hamming256_async (must outlive the call, so not the shared scratch)
  size_t ham_io_cap;
  // sfe_match_hamming256_async runs on a side stream (forked from / joined into the context's stream by events), so
  // that its uploads and kernels overlap with whatever the caller enqueues next, e.g. the sfe_replay_pairs pipeline
  cudaStream_t ham_stream;
  cudaEvent_t ham_fork, ham_done;
  bool ham_pending;
  int64_t launches;
  sfe_replay* replay;  // lazily created by sfe_replay_pairs
  sfe_dist* dist;      // sfe_dist_init / sfe_dist_attach
  void* gftt_ws;   // corner-seeding workspace: response maps, maxima, candidate keys
  size_t gftt_cap;
  char err[512];
};

struct sfe_pyr {
  sfe_ctx* ctx;
  int flavor;
  int planes;
  PyrView view;
  float* storage;
  size_t storage_floats;
  int64_t bytes_per_frame;
};


void sfe_replay_release(sfe_ctx* ctx);  // replay.cu
void sfe_dist_release(sfe_ctx* ctx);    // dist.cu

constexpr int SFE_COUNTER_SLOTS = 64;
inline int* sfe_next_counter(sfe_ctx* ctx) { return ctx->d_counter + 16 * (ctx->counter_seq++ % SFE_COUNTER_SLOTS); }
