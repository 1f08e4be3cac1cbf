// replay.cu -- sfe_replay_pairs: the host-side batching/stream layer for replaying independent frame
// pairs (BASELINE config 4; the per-frame body is matcher.cpp:317 MakePyramid + :208-271 FindMatches).
//
// Host buffers in, host buffers out.  The pairs are cut into chunks; CUDA streams form a pipeline
//   copy stream     H2D of the BGR frames of chunk k+2 (double-buffered device staging)
//   pyramid stream  MakePyramid of both frames of chunk k+1, HIGH priority and one of three pyramid sets: the
//                   tracker is a persistent kernel that holds every SM until its queue is empty, so the pyramids of
//                   the next chunk can only run in the gap between two trackers -- with the priority they take the
//                   first SMs tracker k-1 frees, ahead of the CTAs of tracker k (whose pyramids were built one gap
//                   earlier), and no tracker ever waits for its pyramids
//   compute streams forward/backward tracking of chunk k; chunks alternate between two streams so that the next
//                   tracker fills the SMs the previous one's tail frees
//   output stream   D2H of the results of chunk k-1
// ordered by events only, so with pinned host buffers (sfe_host_alloc) the PCIe traffic of a step hides
// behind its kernels.  Staging buffers, pyramids and events are cached in the context between calls.
#include <new>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include "ctx.cuh"

struct sfe_replay {
  cudaStream_t copy_stream, out_stream, compute2, pyr_stream;  // tracking alternates between the context's stream and compute2
  cudaEvent_t copied[2], consumed[2], built[3], tracked[3], drained;
  int w, h, depth, chunk;  // geometry the frame buffers / pyramids were built for (calls with chunk <= this reuse them)
  uint8_t* d_frames[2];    // per buffer: the c from-frames of a chunk followed by its c to-frames (c <= chunk)
  uint8_t* d_raw[2];       // YUYV input only: the frames as uploaded, converted into d_frames on the pyramid stream
  sfe_pyr* pyr[3];         // pyramid sets of 2 * chunk slots (from-frames, then to-frames): chunk k uses set k % 3
  size_t n_cap;            // feature capacity of the arrays below
  float *d_from, *d_to, *d_back;
  int32_t *d_lv, *d_s1, *d_s2, *d_steps;
  uint8_t* d_acc;
};

namespace {

thread_local const char* g_entry = "sfe_replay_pairs";  // the entry point the error text names

int rfail(sfe_ctx* c, int code, const char* what, cudaError_t e) {
  snprintf(c->err, sizeof(c->err), "%s: %s%s%s", g_entry, what, e != cudaSuccess ? ": " : "",
           e != cudaSuccess ? cudaGetErrorString(e) : "");
  return code;
}

#define RCU(call)                                                        \
  do {                                                                   \
    cudaError_t e_ = (call);                                             \
    if (e_ != cudaSuccess) return rfail(ctx, SFE_ERR_CUDA, #call, e_);   \
  } while (0)

void free_geometry(sfe_replay* r) {
  for (int b = 0; b < 3; ++b) {
    if (b < 2 && r->d_frames[b]) cudaFree(r->d_frames[b]);
    if (b < 2 && r->d_raw[b]) cudaFree(r->d_raw[b]);
    if (r->pyr[b]) sfe_pyr_destroy(r->pyr[b]);
    if (b < 2) r->d_frames[b] = r->d_raw[b] = nullptr;
    r->pyr[b] = nullptr;
  }
  r->w = r->h = r->depth = r->chunk = 0;
}

void free_features(sfe_replay* r) {
  void* p[] = {r->d_from, r->d_to, r->d_back, r->d_lv, r->d_s1, r->d_s2, r->d_steps, r->d_acc};
  for (void* q : p)
    if (q) cudaFree(q);
  r->d_from = r->d_to = r->d_back = nullptr;
  r->d_lv = r->d_s1 = r->d_s2 = r->d_steps = nullptr;
  r->d_acc = nullptr;
  r->n_cap = 0;
}

int ensure(sfe_ctx* ctx, int w, int h, int depth, int chunk, size_t n, bool yuyv) {
  sfe_replay* r = ctx->replay;
  if (!r) {
    r = new (std::nothrow) sfe_replay();
    if (!r) return rfail(ctx, SFE_ERR_NOMEM, "out of host memory", cudaSuccess);
    memset(r, 0, sizeof(*r));
    ctx->replay = r;
    RCU(cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking));
    RCU(cudaStreamCreateWithFlags(&r->out_stream, cudaStreamNonBlocking));
    RCU(cudaStreamCreateWithFlags(&r->compute2, cudaStreamNonBlocking));
    int prio_least = 0, prio_greatest = 0;
    RCU(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    RCU(cudaStreamCreateWithPriority(&r->pyr_stream, cudaStreamNonBlocking, prio_greatest));
    for (int b = 0; b < 3; ++b) {
      if (b < 2) RCU(cudaEventCreateWithFlags(&r->copied[b], cudaEventDisableTiming));
      if (b < 2) RCU(cudaEventCreateWithFlags(&r->consumed[b], cudaEventDisableTiming));
      RCU(cudaEventCreateWithFlags(&r->built[b], cudaEventDisableTiming));
      RCU(cudaEventCreateWithFlags(&r->tracked[b], cudaEventDisableTiming));
    }
    RCU(cudaEventCreateWithFlags(&r->drained, cudaEventDisableTiming));
  }
  if (r->w != w || r->h != h || r->depth != depth || r->chunk < chunk) {   // smaller chunks reuse the buffers
    RCU(cudaStreamSynchronize(ctx->stream));
    free_geometry(r);
    for (int b = 0; b < 3; ++b) {
      cudaError_t e = b < 2 ? cudaMalloc(&r->d_frames[b], (size_t)2 * chunk * 3 * w * h) : cudaSuccess;
      if (e != cudaSuccess) return rfail(ctx, SFE_ERR_NOMEM, "cudaMalloc(frame staging)", e);
      int rc = sfe_pyr_create(ctx, w, h, depth, SFE_HESSIAN, 2 * chunk, &r->pyr[b]);
      if (rc) return rc;
    }
    r->w = w; r->h = h; r->depth = depth; r->chunk = chunk;
  }
  if (yuyv && !r->d_raw[0]) {
    for (int b = 0; b < 2; ++b) {
      cudaError_t e = cudaMalloc(&r->d_raw[b], (size_t)2 * r->chunk * 2 * w * h);
      if (e != cudaSuccess) return rfail(ctx, SFE_ERR_NOMEM, "cudaMalloc(YUYV staging)", e);
    }
  }
  if (n > r->n_cap) {
    RCU(cudaStreamSynchronize(ctx->stream));
    free_features(r);
    const size_t cap = n + n / 4 + 1024;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&r->d_from, 8 * cap);
    if (e == cudaSuccess) e = cudaMalloc(&r->d_to, 8 * cap);
    if (e == cudaSuccess) e = cudaMalloc(&r->d_back, 8 * cap);
    if (e == cudaSuccess) e = cudaMalloc(&r->d_lv, 4 * cap);
    if (e == cudaSuccess) e = cudaMalloc(&r->d_s1, 4 * cap);
    if (e == cudaSuccess) e = cudaMalloc(&r->d_s2, 4 * cap);
    if (e == cudaSuccess) e = cudaMalloc(&r->d_steps, 4 * cap);
    if (e == cudaSuccess) e = cudaMalloc(&r->d_acc, cap);
    if (e != cudaSuccess) return rfail(ctx, SFE_ERR_NOMEM, "cudaMalloc(feature arrays)", e);
    r->n_cap = cap;
  }
  return SFE_SUCCESS;
}

int upload_frames(sfe_ctx* ctx, uint8_t* dst, const uint8_t* src, int w, int h, size_t row_stride, size_t frame_stride,
                  int count, cudaStream_t s, int bpp = 3) {
  const size_t dense_row = (size_t)bpp * w, dense_frame = dense_row * h;
  if (row_stride == dense_row && (frame_stride == dense_frame || count == 1)) {
    RCU(cudaMemcpyAsync(dst, src, dense_frame * count, cudaMemcpyHostToDevice, s));
  } else {
    for (int f = 0; f < count; ++f)
      RCU(cudaMemcpy2DAsync(dst + f * dense_frame, dense_row, src + f * frame_stride, row_stride, dense_row, h,
                            cudaMemcpyHostToDevice, s));
  }
  return SFE_SUCCESS;
}

}  // namespace

void sfe_replay_release(sfe_ctx* ctx) {
  sfe_replay* r = ctx->replay;
  if (!r) return;
  free_geometry(r);
  free_features(r);
  for (int b = 0; b < 3; ++b) {
    if (b < 2 && r->copied[b]) cudaEventDestroy(r->copied[b]);
    if (b < 2 && r->consumed[b]) cudaEventDestroy(r->consumed[b]);
    if (r->built[b]) cudaEventDestroy(r->built[b]);
    if (r->tracked[b]) cudaEventDestroy(r->tracked[b]);
  }
  if (r->drained) cudaEventDestroy(r->drained);
  if (r->copy_stream) cudaStreamDestroy(r->copy_stream);
  if (r->out_stream) cudaStreamDestroy(r->out_stream);
  if (r->compute2) cudaStreamDestroy(r->compute2);
  if (r->pyr_stream) cudaStreamDestroy(r->pyr_stream);
  delete r;
  ctx->replay = nullptr;
}

namespace {
// The pipeline behind both entry points.  seq_stride == 0: independent pairs, pair i = (from_bgr[i], to_bgr[i]).
// seq_stride > 0: a replayed SEQUENCE in from_bgr (to_bgr unused), pair i = (frame i, frame i + seq_stride) -- a chunk of c
// pairs then needs the c + seq_stride consecutive frames [p0, p0 + c + seq_stride), which are uploaded and built once
// (the reference builds one pyramid per new frame, matcher.cpp:317), not once as a `from` and once as a `to` frame.
int replay_run(sfe_ctx* ctx, int w, int h, int depth, int npairs, int seq_stride, const uint8_t* from_bgr,
               const uint8_t* to_bgr, size_t row_stride, size_t frame_stride, int n_per_pair,
               const float* from_xy, float* to_xy, const int32_t* levels, int default_levels, float thr,
               int maxit, double fb_max, float* back_xy, int32_t* status_fwd, int32_t* status_bwd,
               uint8_t* accepted, int32_t* steps, int chunk_pairs, bool yuyv = false) {
  if (!ctx) return SFE_ERR_INVALID;
  if (npairs == 0) return SFE_SUCCESS;
  const int bpp = yuyv ? 2 : 3;
  if (w < 1 || h < 1 || depth < 1 || depth > SFE_MAX_LEVELS || npairs < 0 || n_per_pair < 1 || !from_bgr ||
      (seq_stride == 0 && !to_bgr) || seq_stride < 0 || !from_xy || !to_xy || default_levels < 1 || maxit < 0 ||
      row_stride < (size_t)bpp * w || (yuyv && ((w & 1) || ((size_t)w * h) % 4)))
    return rfail(ctx, SFE_ERR_INVALID, "bad arguments", cudaSuccess);
  if (levels)
    for (size_t i = 0; i < (size_t)npairs * n_per_pair; ++i)
      if (levels[i] < 1) return rfail(ctx, SFE_ERR_INVALID, "levels[] entries must be >= 1", cudaSuccess);
  RCU(cudaSetDevice(ctx->device));
  // chunk size: default = an eighth of the batch (at least 8 pairs so that the kernels still fill the GPU; measured
  // best on B200 for 128 and 512 VGA pairs) but at most 128 pairs -- three pyramid sets and two staging buffers of
  // 2 * chunk frames each are 1.4 GB then, whatever the length of the replay; the first chunk is a quarter of
  // that and the second a half, so compute starts after a short upload and each upload (54 GB/s measured) stays ahead
  // of the tracking of the chunk before it (tracking a pair takes 2.75x its upload)
  int chunk = chunk_pairs > 0 ? chunk_pairs : (npairs + 7) / 8;
  if (chunk_pairs <= 0 && chunk < 8) chunk = 8;
  if (chunk_pairs <= 0 && chunk > 128) chunk = 128;
  if (chunk > npairs) chunk = npairs;
  if (chunk < seq_stride) chunk = seq_stride;   // a chunk's c + seq_stride frames must fit the 2 * chunk staging slots
  const int first_chunk = chunk >= 16 ? chunk / 4 : chunk;
  const size_t n = (size_t)npairs * n_per_pair;
  int rc = ensure(ctx, w, h, depth, chunk, n, yuyv);
  if (rc) return rc;
  sfe_replay* r = ctx->replay;
  cudaStream_t cs0 = ctx->stream, xs = r->copy_stream, os = r->out_stream, ps = r->pyr_stream;

  // SFE_REPLAY_TRACE=1 (diagnostics): per-chunk event times of the pipeline on stderr
  static const bool trace = getenv("SFE_REPLAY_TRACE") != nullptr;
  constexpr int TRACE_MAX = 64;
  cudaEvent_t tr_ev[TRACE_MAX][5];
  cudaEvent_t tr_t0 = nullptr;
  int tr_n = 0;
  if (trace) {
    cudaEventCreate(&tr_t0);
    cudaEventRecord(tr_t0, cs0);
  }
  // the other streams must not run ahead of whatever the caller queued on the context's stream before this call
  // (which also covers the staging buffers of a previous call)
  RCU(cudaEventRecord(r->drained, cs0));
  RCU(cudaStreamWaitEvent(xs, r->drained, 0));
  RCU(cudaStreamWaitEvent(os, r->drained, 0));
  RCU(cudaStreamWaitEvent(ps, r->drained, 0));
  // the feature lists travel with their chunk on the copy stream (slices of a few hundred KB in front of the chunk's
  // frames).  Round 1 uploaded them in one piece on the context's stream: measured with SFE_REPLAY_TRACE, everything
  // queued on that stream behind those copies -- the first chunk's tracker -- then waited until the copy engine had
  // drained the first four frame uploads (1.2 ms of a 29 ms call).  The compute streams carry no copy-engine work now.
  RCU(cudaStreamWaitEvent(r->compute2, r->drained, 0));
  const size_t dense_frame = (size_t)3 * w * h;
  int p0 = 0;
  for (int k = 0; p0 < npairs; ++k) {
    const int b = k & 1, set = k % 3, want = k == 0 ? first_chunk : (k == 1 && 2 * first_chunk < chunk ? 2 * first_chunk : chunk),
              c = npairs - p0 < want ? npairs - p0 : want;
    cudaStream_t cs = b ? r->compute2 : cs0;
    // ---- copy stream: frames of chunk k into staging buffer b (free once chunk k-2's pyramids are built)
    if (k >= 2) RCU(cudaStreamWaitEvent(xs, r->consumed[b], 0));
    // independent pairs: the c from-frames, then the c to-frames; sequence: the c + seq_stride frames the chunk's pairs touch
    const int nbuild = seq_stride ? c + seq_stride : 2 * c, to_first = seq_stride ? seq_stride : c;
    uint8_t* up = yuyv ? r->d_raw[b] : r->d_frames[b];
    const size_t up_frame = (size_t)bpp * w * h;
    {
      const size_t f0 = (size_t)p0 * n_per_pair, nf = (size_t)c * n_per_pair;
      RCU(cudaMemcpyAsync(r->d_from + 2 * f0, from_xy + 2 * f0, 8 * nf, cudaMemcpyHostToDevice, xs));
      RCU(cudaMemcpyAsync(r->d_to + 2 * f0, to_xy + 2 * f0, 8 * nf, cudaMemcpyHostToDevice, xs));
      if (levels) RCU(cudaMemcpyAsync(r->d_lv + f0, levels + f0, 4 * nf, cudaMemcpyHostToDevice, xs));
    }
    rc = upload_frames(ctx, up, from_bgr + (size_t)p0 * frame_stride, w, h, row_stride, frame_stride,
                       seq_stride ? nbuild : c, xs, bpp);
    if (!rc && !seq_stride)
      rc = upload_frames(ctx, up + (size_t)c * up_frame, to_bgr + (size_t)p0 * frame_stride, w, h,
                         row_stride, frame_stride, c, xs, bpp);
    if (rc) return rc;
    RCU(cudaEventRecord(r->copied[b], xs));
    const bool tr = trace && k < TRACE_MAX;
    if (tr) {
      for (int e = 0; e < 5; ++e) cudaEventCreate(&tr_ev[k][e]);
      cudaEventRecord(tr_ev[k][0], xs);
      tr_n = k + 1;
    }
    // ---- pyramid stream: the pyramids of the chunk's frames into set k % 3 (free once chunk k-3 has been tracked)
    RCU(cudaStreamWaitEvent(ps, r->copied[b], 0));
    if (k >= 3) RCU(cudaStreamWaitEvent(ps, r->tracked[set], 0));
    if (yuyv) {  // video.cpp:187-223 on the device: the staging buffer then holds the BGR frames MakePyramid takes
      int nc = launch_yuyv_to_bgr(r->d_raw[b], (size_t)nbuild * w * h, r->d_frames[b], ps);
      if (nc < 0) return rfail(ctx, SFE_ERR_CUDA, "yuyv_to_bgr launch", (cudaError_t)(-nc));
      ctx->launches += nc;
    }
    // one build for the frames of the chunk: slots [0, c) the from-frames, [to_first, to_first + c) the to-frames
    int nl = launch_pyr_build(r->pyr[set]->view, SFE_HESSIAN, r->d_frames[b], (size_t)3 * w, dense_frame, 0, nbuild, ps);
    if (nl < 0) return rfail(ctx, SFE_ERR_CUDA, "pyramid launch", (cudaError_t)(-nl));
    ctx->launches += nl;
    RCU(cudaEventRecord(r->consumed[b], ps));
    RCU(cudaEventRecord(r->built[set], ps));
    if (tr) cudaEventRecord(tr_ev[k][1], ps);
    // ---- compute stream b: forward/backward tracking of the chunk's features (the work-queue counter is per stream)
    RCU(cudaStreamWaitEvent(cs, r->built[set], 0));
    if (tr) cudaEventRecord(tr_ev[k][2], cs);
    const size_t f0 = (size_t)p0 * n_per_pair;
    const int nf = c * n_per_pair;
    TrackArgs ta{nf, n_per_pair, 0, to_first, r->d_from + 2 * f0, r->d_to + 2 * f0, levels ? r->d_lv + f0 : nullptr, default_levels,
                 thr, maxit, fb_max, r->d_back + 2 * f0, r->d_s1 + f0, r->d_s2 + f0, r->d_acc + f0, r->d_steps + f0, 2};
    nl = launch_track_hessian(r->pyr[set]->view, r->pyr[set]->view, ta, ctx->d_mask, sfe_next_counter(ctx), ctx->num_sms, cs);
    if (nl < 0) return rfail(ctx, SFE_ERR_CUDA, "track launch", (cudaError_t)(-nl));
    ctx->launches += nl;
    RCU(cudaEventRecord(r->tracked[set], cs));
    if (tr) cudaEventRecord(tr_ev[k][3], cs);
    // ---- output stream: results of chunk k back to the caller's buffers
    RCU(cudaStreamWaitEvent(os, r->tracked[set], 0));
    RCU(cudaMemcpyAsync(to_xy + 2 * f0, r->d_to + 2 * f0, 8 * (size_t)nf, cudaMemcpyDeviceToHost, os));
    if (back_xy) RCU(cudaMemcpyAsync(back_xy + 2 * f0, r->d_back + 2 * f0, 8 * (size_t)nf, cudaMemcpyDeviceToHost, os));
    if (status_fwd) RCU(cudaMemcpyAsync(status_fwd + f0, r->d_s1 + f0, 4 * (size_t)nf, cudaMemcpyDeviceToHost, os));
    if (status_bwd) RCU(cudaMemcpyAsync(status_bwd + f0, r->d_s2 + f0, 4 * (size_t)nf, cudaMemcpyDeviceToHost, os));
    if (accepted) RCU(cudaMemcpyAsync(accepted + f0, r->d_acc + f0, (size_t)nf, cudaMemcpyDeviceToHost, os));
    if (steps) RCU(cudaMemcpyAsync(steps + f0, r->d_steps + f0, 4 * (size_t)nf, cudaMemcpyDeviceToHost, os));
    if (tr) cudaEventRecord(tr_ev[k][4], os);
    p0 += c;
  }
  cudaStream_t cs = cs0;
  // join: the context's stream is complete only when the last download is
  RCU(cudaEventRecord(r->drained, os));
  RCU(cudaStreamWaitEvent(cs, r->drained, 0));
  RCU(cudaStreamSynchronize(cs));
  if (trace) {
    fprintf(stderr, "replay trace (ms after the call's first stream operation): chunk: copied built track_start tracked downloaded\n");

    for (int k = 0; k < tr_n; ++k) {
      float t[5];
      for (int e = 0; e < 5; ++e) {
        cudaEventElapsedTime(&t[e], tr_t0, tr_ev[k][e]);
        cudaEventDestroy(tr_ev[k][e]);
      }
      fprintf(stderr, "  %2d: %7.3f %7.3f %7.3f %7.3f %7.3f\n", k, t[0], t[1], t[2], t[3], t[4]);
    }
    cudaEventDestroy(tr_t0);
  }
  return SFE_SUCCESS;
}
}  // namespace

extern "C" int sfe_replay_pairs(sfe_ctx* ctx, int w, int h, int depth, int npairs, const uint8_t* from_bgr,
                                const uint8_t* to_bgr, size_t row_stride, size_t frame_stride, int n_per_pair,
                                const float* from_xy, float* to_xy, const int32_t* levels, int default_levels, float thr,
                                int maxit, double fb_max, float* back_xy, int32_t* status_fwd, int32_t* status_bwd,
                                uint8_t* accepted, int32_t* steps, int chunk_pairs) {
  g_entry = "sfe_replay_pairs";
  return replay_run(ctx, w, h, depth, npairs, 0, from_bgr, to_bgr, row_stride, frame_stride, n_per_pair, from_xy, to_xy, levels,
                    default_levels, thr, maxit, fb_max, back_xy, status_fwd, status_bwd, accepted, steps, chunk_pairs);
}

extern "C" int sfe_replay_sequence(sfe_ctx* ctx, int w, int h, int depth, int nframes, int pair_stride,
                                   const uint8_t* frames_bgr, size_t row_stride, size_t frame_stride, int n_per_pair,
                                   const float* from_xy, float* to_xy, const int32_t* levels, int default_levels, float thr,
                                   int maxit, double fb_max, float* back_xy, int32_t* status_fwd, int32_t* status_bwd,
                                   uint8_t* accepted, int32_t* steps, int chunk_pairs) {
  if (!ctx) return SFE_ERR_INVALID;
  g_entry = "sfe_replay_sequence";
  if (pair_stride < 1 || nframes < 0) return rfail(ctx, SFE_ERR_INVALID, "bad sequence arguments", cudaSuccess);
  if (nframes <= pair_stride) return SFE_SUCCESS;   // no pair
  return replay_run(ctx, w, h, depth, nframes - pair_stride, pair_stride, frames_bgr, nullptr, row_stride, frame_stride, n_per_pair,
                    from_xy, to_xy, levels, default_levels, thr, maxit, fb_max, back_xy, status_fwd, status_bwd, accepted, steps,
                    chunk_pairs);
}

extern "C" int sfe_replay_sequence_yuyv(sfe_ctx* ctx, int w, int h, int depth, int nframes, int pair_stride,
                                        const uint8_t* frames_yuyv, size_t row_stride, size_t frame_stride, int n_per_pair,
                                        const float* from_xy, float* to_xy, const int32_t* levels, int default_levels,
                                        float thr, int maxit, double fb_max, float* back_xy, int32_t* status_fwd,
                                        int32_t* status_bwd, uint8_t* accepted, int32_t* steps, int chunk_pairs) {
  if (!ctx) return SFE_ERR_INVALID;
  g_entry = "sfe_replay_sequence_yuyv";
  if (pair_stride < 1 || nframes < 0) return rfail(ctx, SFE_ERR_INVALID, "bad sequence arguments", cudaSuccess);
  if (nframes <= pair_stride) return SFE_SUCCESS;   // no pair
  return replay_run(ctx, w, h, depth, nframes - pair_stride, pair_stride, frames_yuyv, nullptr, row_stride, frame_stride, n_per_pair,
                    from_xy, to_xy, levels, default_levels, thr, maxit, fb_max, back_xy, status_fwd, status_bwd, accepted, steps,
                    chunk_pairs, true);
}
