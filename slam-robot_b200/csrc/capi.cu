// capi.cu -- the C ABI of libslamfe.so (include/slamfe.h): contexts, pyramid handles, and the
// host-pointer / device-pointer entry points.  No CPU fallback anywhere: every entry point
// either runs the CUDA path or reports SFE_ERR_CUDA.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "ctx.cuh"

namespace {

int fail(sfe_ctx* c, int code, const char* fmt, const char* detail) {
  if (c) snprintf(c->err, sizeof(c->err), fmt, detail);
  return code;
}

#define CU(call)                                                                     \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) return fail(ctx, SFE_ERR_CUDA, #call ": %s", cudaGetErrorString(e_)); \
  } while (0)

int use_device(sfe_ctx* ctx) {
  CU(cudaSetDevice(ctx->device));
  return SFE_SUCCESS;
}

int launched(sfe_ctx* ctx, int r, const char* what) {
  if (r < 0) return fail(ctx, SFE_ERR_CUDA, what, cudaGetErrorString((cudaError_t)(-r)));
  ctx->launches += r;
  return SFE_SUCCESS;
}

// hessian.h:11-30 (same arithmetic as oracle.c orc_mask13)
void build_mask(float* mask) {
  const int n = SFE_PATCH;
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x) {
      double rx = 0.5 * n - x, ry = 0.5 * n - y;
      mask[y * n + x] = (float)(1. / (15. + (rx * rx + ry * ry)));
    }
  double sum = 0;
  for (int i = 0; i < SFE_PLEN; ++i) sum += mask[i];
  double scale = SFE_PLEN / sum;
  for (int i = 0; i < SFE_PLEN; ++i) mask[i] = (float)(mask[i] * scale);
}

int ensure_scratch(sfe_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->scratch_cap) return SFE_SUCCESS;
  if (ctx->scratch) CU(cudaFree(ctx->scratch));
  ctx->scratch = nullptr;
  ctx->scratch_cap = 0;
  size_t cap = bytes + bytes / 4 + 4096;
  cudaError_t e = cudaMalloc(&ctx->scratch, cap);
  if (e != cudaSuccess) return fail(ctx, SFE_ERR_NOMEM, "cudaMalloc(scratch): %s", cudaGetErrorString(e));
  ctx->scratch_cap = cap;
  return SFE_SUCCESS;
}

// bump allocator over the scratch buffer (256-byte aligned pieces)
struct Carver {
  char* base;
  size_t off;
  template <class T>
  T* take(size_t count) {
    T* p = (T*)(base + off);
    off += (count * sizeof(T) + 255) & ~(size_t)255;
    return p;
  }
};
size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

int check_pairs(sfe_ctx* ctx, const sfe_pyr* from, int from_first, const sfe_pyr* to, int to_first, int n,
                int n_per_pair) {
  if (!from || !to || n < 0 || n_per_pair <= 0) return fail(ctx, SFE_ERR_INVALID, "%s", "bad track arguments");
  if (n == 0) return SFE_SUCCESS;
  int npairs = (n + n_per_pair - 1) / n_per_pair;
  if (from_first < 0 || to_first < 0 || from_first + npairs > from->view.batch || to_first + npairs > to->view.batch)
    return fail(ctx, SFE_ERR_INVALID, "%s", "frame pair range exceeds the pyramid batch");
  if (from->view.w[0] != to->view.w[0] || from->view.h[0] != to->view.h[0])
    return fail(ctx, SFE_ERR_INVALID, "%s", "from/to pyramids differ in frame size");
  return SFE_SUCCESS;
}

}  // namespace

namespace {
// shared host-pointer wrapper of the two forward/backward trackers
template <class Launch>
int track_host(sfe_ctx* ctx, int n, const float* from_xy, float* to_xy, const int32_t* levels, float* back_xy,
               int32_t* status_fwd, int32_t* status_bwd, uint8_t* accepted, int32_t* steps, Launch launch) {
  if (levels)  // hessian.h:249: lvls = min(stack.size(), patches.size()) >= 1 for every call the reference makes
    for (int i = 0; i < n; ++i)
      if (levels[i] < 1) return fail(ctx, SFE_ERR_INVALID, "%s", "levels[] entries must be >= 1");
  size_t need = padded(8 * (size_t)n) * 3 + padded(4 * (size_t)n) * 4 + padded(n);
  int rc = ensure_scratch(ctx, need);
  if (rc) return rc;
  // inputs first (from, levels, seeds), then the outputs (to, back, statuses, steps, accepted): each group is one
  // contiguous range of the scratch
  Carver c{(char*)ctx->scratch, 0};
  float* d_from = c.take<float>(2 * (size_t)n);
  int32_t* d_lv = c.take<int32_t>(n);
  float* d_to = c.take<float>(2 * (size_t)n);
  const size_t in_bytes = c.off;
  float* d_back = c.take<float>(2 * (size_t)n);
  int32_t* d_s1 = c.take<int32_t>(n);
  int32_t* d_s2 = c.take<int32_t>(n);
  int32_t* d_steps = c.take<int32_t>(n);
  uint8_t* d_acc = c.take<uint8_t>(n);
  const size_t out_off = (char*)d_to - (char*)ctx->scratch, out_bytes = c.off - out_off;
  cudaStream_t s = ctx->stream;
  // Small calls (the live robot tracks a few hundred features per frame, matcher.cpp:208-271) are latency-bound: nine
  // copies of a few KB from pageable memory cost more than the kernel.  They go through a pinned mirror of the scratch
  // head instead -- one upload, one download -- and are scattered to the caller's arrays on the host.
  const bool staged = n <= 65536;
  if (staged) {
    if (c.off > ctx->h_stage_cap) {
      if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
      ctx->h_stage = nullptr;
      ctx->h_stage_cap = 0;
      cudaError_t e = cudaMallocHost(&ctx->h_stage, c.off);
      if (e != cudaSuccess) return fail(ctx, SFE_ERR_NOMEM, "cudaMallocHost(staging): %s", cudaGetErrorString(e));
      ctx->h_stage_cap = c.off;
    }
    char* h = (char*)ctx->h_stage;
    memcpy(h + ((char*)d_from - (char*)ctx->scratch), from_xy, 8 * (size_t)n);
    if (levels) memcpy(h + ((char*)d_lv - (char*)ctx->scratch), levels, 4 * (size_t)n);
    memcpy(h + out_off, to_xy, 8 * (size_t)n);
    CU(cudaMemcpyAsync(ctx->scratch, h, in_bytes, cudaMemcpyHostToDevice, s));
  } else {
    CU(cudaMemcpyAsync(d_from, from_xy, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d_to, to_xy, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
    if (levels) CU(cudaMemcpyAsync(d_lv, levels, 4 * (size_t)n, cudaMemcpyHostToDevice, s));
  }
  rc = launch(d_from, d_to, levels ? d_lv : nullptr, d_back, d_s1, d_s2, d_acc, d_steps);
  if (rc) return rc;
  if (staged) {
    char* h = (char*)ctx->h_stage;
    CU(cudaMemcpyAsync(h + out_off, d_to, out_bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    auto at = [&](const void* d) { return h + ((const char*)d - (const char*)ctx->scratch); };
    memcpy(to_xy, at(d_to), 8 * (size_t)n);
    if (back_xy) memcpy(back_xy, at(d_back), 8 * (size_t)n);
    if (status_fwd) memcpy(status_fwd, at(d_s1), 4 * (size_t)n);
    if (status_bwd) memcpy(status_bwd, at(d_s2), 4 * (size_t)n);
    if (accepted) memcpy(accepted, at(d_acc), (size_t)n);
    if (steps) memcpy(steps, at(d_steps), 4 * (size_t)n);
    return SFE_SUCCESS;
  }
  CU(cudaMemcpyAsync(to_xy, d_to, 8 * (size_t)n, cudaMemcpyDeviceToHost, s));
  if (back_xy) CU(cudaMemcpyAsync(back_xy, d_back, 8 * (size_t)n, cudaMemcpyDeviceToHost, s));
  if (status_fwd) CU(cudaMemcpyAsync(status_fwd, d_s1, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
  if (status_bwd) CU(cudaMemcpyAsync(status_bwd, d_s2, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
  if (accepted) CU(cudaMemcpyAsync(accepted, d_acc, (size_t)n, cudaMemcpyDeviceToHost, s));
  if (steps) CU(cudaMemcpyAsync(steps, d_steps, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SFE_SUCCESS;
}
}  // namespace

namespace {
template <class Launch>
int point_pair_host(sfe_ctx* ctx, int n, const float* txy, const float* xy, float* out, int out_per, Launch launch) {
  size_t need = 2 * padded(8 * (size_t)n) + padded(4 * (size_t)n * out_per);
  int rc = ensure_scratch(ctx, need);
  if (rc) return rc;
  Carver c{(char*)ctx->scratch, 0};
  float* d_t = c.take<float>(2 * (size_t)n);
  float* d_x = c.take<float>(2 * (size_t)n);
  float* d_o = c.take<float>((size_t)n * out_per);
  cudaStream_t s = ctx->stream;
  CU(cudaMemcpyAsync(d_t, txy, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(d_x, xy, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
  rc = launched(ctx, launch(d_t, d_x, d_o, s), "launch: %s");
  if (rc) return rc;
  CU(cudaMemcpyAsync(out, d_o, 4 * (size_t)n * out_per, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SFE_SUCCESS;
}
bool frame_ok(const sfe_pyr* p, int frame, int level) {
  return p && frame >= 0 && frame < p->view.batch && level >= 0 && level < p->view.depth;
}
}  // namespace

extern "C" {

int sfe_create(int device, sfe_ctx** out) {
  if (!out) return SFE_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return SFE_ERR_CUDA;
  sfe_ctx* ctx = new (std::nothrow) sfe_ctx();
  if (!ctx) return SFE_ERR_NOMEM;
  memset(ctx, 0, sizeof(*ctx));
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return SFE_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  build_mask(ctx->h_mask);
  cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
  if (cudaMalloc(&ctx->d_counter, 64 * SFE_COUNTER_SLOTS) != cudaSuccess || cudaMalloc(&ctx->d_mask, sizeof(ctx->h_mask)) != cudaSuccess ||
      cudaMemcpy(ctx->d_mask, ctx->h_mask, sizeof(ctx->h_mask), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return SFE_ERR_CUDA;
  }
  *out = ctx;
  return SFE_SUCCESS;
}

void sfe_destroy(sfe_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  sfe_replay_release(ctx);
  sfe_dist_release(ctx);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->ham_ws) cudaFree(ctx->ham_ws);
  if (ctx->ham_stream) {
    cudaStreamSynchronize(ctx->ham_stream);
    cudaStreamDestroy(ctx->ham_stream);
    cudaEventDestroy(ctx->ham_fork);
    cudaEventDestroy(ctx->ham_done);
  }
  if (ctx->ham_io) cudaFree(ctx->ham_io);
  if (ctx->gftt_ws) cudaFree(ctx->gftt_ws);
  cudaFree(ctx->d_mask);
  cudaFree(ctx->d_counter);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* sfe_last_error(const sfe_ctx* ctx) { return ctx ? ctx->err : "null context"; }

int sfe_set_stream(sfe_ctx* ctx, void* cuda_stream) {
  if (!ctx) return SFE_ERR_INVALID;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return SFE_SUCCESS;
}

int sfe_sync(sfe_ctx* ctx) {
  if (!ctx) return SFE_ERR_INVALID;
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->ham_pending) {
    CU(cudaStreamSynchronize(ctx->ham_stream));
    ctx->ham_pending = false;
  }
  return SFE_SUCCESS;
}

int sfe_get_mask(sfe_ctx* ctx, float* mask169) {
  if (!ctx || !mask169) return SFE_ERR_INVALID;
  CU(cudaMemcpy(mask169, ctx->d_mask, sizeof(float) * SFE_PLEN, cudaMemcpyDeviceToHost));
  return SFE_SUCCESS;
}

int sfe_host_alloc(sfe_ctx* ctx, size_t bytes, void** out) {
  if (!ctx || !out) return SFE_ERR_INVALID;
  CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return SFE_SUCCESS;
}

int sfe_host_free(sfe_ctx* ctx, void* p) {
  if (!ctx) return SFE_ERR_INVALID;
  CU(cudaFreeHost(p));
  return SFE_SUCCESS;
}

int64_t sfe_launch_count(const sfe_ctx* ctx) { return ctx ? ctx->launches : 0; }

/* ---- pyramids ---------------------------------------------------------------------------- */

int sfe_pyr_create(sfe_ctx* ctx, int w, int h, int depth, int flavor, int batch, sfe_pyr** out) {
  if (!ctx || !out) return SFE_ERR_INVALID;
  *out = nullptr;
  if (w < 1 || h < 1 || depth < 1 || depth > SFE_MAX_LEVELS || batch < 1 || flavor < SFE_HESSIAN || flavor > SFE_BRUTE)
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad pyramid geometry");
  if (use_device(ctx)) return SFE_ERR_CUDA;
  sfe_pyr* p = new (std::nothrow) sfe_pyr();
  if (!p) return SFE_ERR_NOMEM;
  memset(p, 0, sizeof(*p));
  p->ctx = ctx;
  p->flavor = flavor;
  p->planes = flavor == SFE_KLT ? 3 : 1;
  PyrView& v = p->view;
  v.depth = depth;
  v.batch = batch;
  size_t total = 0;
  int64_t px = 0;
  int lw = w, lh = h;
  for (int l = 0; l < depth; ++l) {
    v.w[l] = lw;
    v.h[l] = lh;
    v.pitch[l] = (lw + 3) & ~3;
    // frame stride rounded to 32 floats so every frame plane starts 128-byte aligned
    v.frame_stride[l] = ((long long)v.pitch[l] * lh + 31) & ~31LL;
    total += (size_t)v.frame_stride[l] * batch * p->planes;
    px += (int64_t)lw * lh;
    lw = (lw + 1) / 2;  // hessian.h:108
    lh = (lh + 1) / 2;
  }
  p->bytes_per_frame = 3LL * w * h + 4LL * px * p->planes;
  cudaError_t e = cudaMalloc(&p->storage, total * sizeof(float));
  if (e != cudaSuccess) {
    delete p;
    return fail(ctx, SFE_ERR_NOMEM, "cudaMalloc(pyramid): %s", cudaGetErrorString(e));
  }
  p->storage_floats = total;
  size_t off = 0;
  for (int pl = 0; pl < p->planes; ++pl)
    for (int l = 0; l < depth; ++l) {
      v.base[pl][l] = p->storage + off;
      off += (size_t)v.frame_stride[l] * batch;
    }
  *out = p;
  return SFE_SUCCESS;
}

void sfe_pyr_destroy(sfe_pyr* pyr) {
  if (!pyr) return;
  cudaSetDevice(pyr->ctx->device);
  cudaStreamSynchronize(pyr->ctx->stream);
  cudaFree(pyr->storage);
  delete pyr;
}

int sfe_pyr_level_size(const sfe_pyr* pyr, int level, int* w, int* h, int* pitch_floats) {
  if (!pyr || level < 0 || level >= pyr->view.depth) return SFE_ERR_INVALID;
  if (w) *w = pyr->view.w[level];
  if (h) *h = pyr->view.h[level];
  if (pitch_floats) *pitch_floats = pyr->view.pitch[level];
  return SFE_SUCCESS;
}

int64_t sfe_pyr_bytes_per_frame(const sfe_pyr* pyr) { return pyr ? pyr->bytes_per_frame : 0; }

int sfe_pyr_build_dev(sfe_ctx* ctx, sfe_pyr* pyr, const uint8_t* bgr_dev, size_t row_stride, size_t frame_stride,
                      int first, int count) {
  if (!ctx || !pyr || !bgr_dev) return SFE_ERR_INVALID;
  if (count == 0) return SFE_SUCCESS;
  if (first < 0 || count < 0 || first + count > pyr->view.batch || row_stride < (size_t)3 * pyr->view.w[0])
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad frame range or stride");
  if (use_device(ctx)) return SFE_ERR_CUDA;
  return launched(ctx, launch_pyr_build(pyr->view, pyr->flavor, bgr_dev, row_stride, frame_stride, first, count, ctx->stream),
                  "pyramid launch: %s");
}

int sfe_pyr_build(sfe_ctx* ctx, sfe_pyr* pyr, const uint8_t* bgr_host, size_t row_stride, size_t frame_stride,
                  int first, int count) {
  if (!ctx || !pyr || !bgr_host) return SFE_ERR_INVALID;
  if (count == 0) return SFE_SUCCESS;
  if (first < 0 || count < 0 || first + count > pyr->view.batch || row_stride < (size_t)3 * pyr->view.w[0])
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad frame range or stride");
  if (use_device(ctx)) return SFE_ERR_CUDA;
  const int w = pyr->view.w[0], h = pyr->view.h[0];
  const size_t dense_row = (size_t)3 * w, dense_frame = dense_row * h;
  int rc = ensure_scratch(ctx, dense_frame * count);
  if (rc) return rc;
  uint8_t* d = (uint8_t*)ctx->scratch;
  if (row_stride == dense_row && (frame_stride == dense_frame || count == 1)) {
    CU(cudaMemcpyAsync(d, bgr_host, dense_frame * count, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    for (int f = 0; f < count; ++f)
      CU(cudaMemcpy2DAsync(d + f * dense_frame, dense_row, bgr_host + f * frame_stride, row_stride, dense_row, h,
                           cudaMemcpyHostToDevice, ctx->stream));
  }
  rc = launched(ctx, launch_pyr_build(pyr->view, pyr->flavor, d, dense_row, dense_frame, first, count, ctx->stream),
                "pyramid launch: %s");
  if (rc) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return SFE_SUCCESS;
}

int sfe_pyr_download(sfe_ctx* ctx, const sfe_pyr* pyr, int frame, int level, int plane, float* host_dst) {
  if (!ctx || !pyr || !host_dst || frame < 0 || frame >= pyr->view.batch || level < 0 || level >= pyr->view.depth ||
      plane < 0 || plane >= pyr->planes)
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad download arguments");
  if (use_device(ctx)) return SFE_ERR_CUDA;
  ImgView im = img_of(pyr->view, plane, level, frame);
  CU(cudaMemcpy2DAsync(host_dst, sizeof(float) * im.w, im.p, sizeof(float) * im.pitch, sizeof(float) * im.w, im.h,
                       cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SFE_SUCCESS;
}

/* ---- P1 ---------------------------------------------------------------------------------- */

int sfe_track_fb_dev(sfe_ctx* ctx, const sfe_pyr* from, int from_first, const sfe_pyr* to, int to_first, int n,
                     int n_per_pair, const float* from_xy, float* to_xy, const int32_t* levels, int default_levels,
                     float thr, int maxit, double fb_max, float* back_xy, int32_t* status_fwd, int32_t* status_bwd,
                     uint8_t* accepted, int32_t* steps) {
  if (!ctx) return SFE_ERR_INVALID;
  int rc = check_pairs(ctx, from, from_first, to, to_first, n, n_per_pair);
  if (rc || n == 0) return rc;
  if (!from_xy || !to_xy || default_levels < 1 || maxit < 0) return fail(ctx, SFE_ERR_INVALID, "%s", "bad track arguments");
  if (from->flavor != SFE_HESSIAN || to->flavor != SFE_HESSIAN)
    return fail(ctx, SFE_ERR_INVALID, "%s", "sfe_track_fb needs SFE_HESSIAN pyramids");
  if (use_device(ctx)) return SFE_ERR_CUDA;
  TrackArgs a{n, n_per_pair, from_first, to_first, from_xy, to_xy, levels, default_levels, thr, maxit, fb_max,
              back_xy, status_fwd, status_bwd, accepted, steps, 2};
  return launched(ctx, launch_track_hessian(from->view, to->view, a, ctx->d_mask, sfe_next_counter(ctx), ctx->num_sms, ctx->stream), "track launch: %s");
}


int sfe_track_fb(sfe_ctx* ctx, const sfe_pyr* from, int from_first, const sfe_pyr* to, int to_first, int n,
                 int n_per_pair, const float* from_xy, float* to_xy, const int32_t* levels, int default_levels,
                 float thr, int maxit, double fb_max, float* back_xy, int32_t* status_fwd, int32_t* status_bwd,
                 uint8_t* accepted, int32_t* steps) {
  if (!ctx) return SFE_ERR_INVALID;
  if (n == 0) return SFE_SUCCESS;
  if (n < 0 || !from_xy || !to_xy) return fail(ctx, SFE_ERR_INVALID, "%s", "bad track arguments");
  if (use_device(ctx)) return SFE_ERR_CUDA;
  return track_host(ctx, n, from_xy, to_xy, levels, back_xy, status_fwd, status_bwd, accepted, steps,
                    [&](float* df, float* dt, int32_t* dl, float* db, int32_t* s1, int32_t* s2, uint8_t* acc, int32_t* st) {
                      return sfe_track_fb_dev(ctx, from, from_first, to, to_first, n, n_per_pair, df, dt, dl,
                                              default_levels, thr, maxit, fb_max, db, s1, s2, acc, st);
                    });
}

int sfe_track_dev(sfe_ctx* ctx, const sfe_pyr* tmpl, int tmpl_first, const sfe_pyr* search, int search_first, int n,
                  int n_per_pair, const float* tmpl_xy, float* xy, const int32_t* levels, int default_levels, float thr,
                  int maxit, int32_t* status, int32_t* steps) {
  if (!ctx) return SFE_ERR_INVALID;
  int rc = check_pairs(ctx, tmpl, tmpl_first, search, search_first, n, n_per_pair);
  if (rc || n == 0) return rc;
  if (!tmpl_xy || !xy || default_levels < 1 || maxit < 0) return fail(ctx, SFE_ERR_INVALID, "%s", "bad track arguments");
  if (tmpl->flavor != search->flavor || tmpl->flavor == SFE_BRUTE)
    return fail(ctx, SFE_ERR_INVALID, "%s", "sfe_track needs two SFE_HESSIAN or two SFE_KLT pyramids");
  if (use_device(ctx)) return SFE_ERR_CUDA;
  TrackArgs a{n, n_per_pair, tmpl_first, search_first, tmpl_xy, xy, levels, default_levels, thr, maxit, 0.,
              nullptr, status, nullptr, nullptr, steps, 1};
  int r = tmpl->flavor == SFE_HESSIAN
              ? launch_track_hessian(tmpl->view, search->view, a, ctx->d_mask, sfe_next_counter(ctx), ctx->num_sms, ctx->stream)
              : launch_track_klt(tmpl->view, search->view, a, ctx->d_mask, sfe_next_counter(ctx), ctx->num_sms, ctx->stream);
  return launched(ctx, r, "track launch: %s");
}

int sfe_track(sfe_ctx* ctx, const sfe_pyr* tmpl, int tmpl_first, const sfe_pyr* search, int search_first, int n,
              int n_per_pair, const float* tmpl_xy, float* xy, const int32_t* levels, int default_levels, float thr,
              int maxit, int32_t* status, int32_t* steps) {
  if (!ctx) return SFE_ERR_INVALID;
  if (n == 0) return SFE_SUCCESS;
  if (n < 0 || !tmpl_xy || !xy) return fail(ctx, SFE_ERR_INVALID, "%s", "bad track arguments");
  if (use_device(ctx)) return SFE_ERR_CUDA;
  return track_host(ctx, n, tmpl_xy, xy, levels, nullptr, status, nullptr, nullptr, steps,
                    [&](float* df, float* dt, int32_t* dl, float*, int32_t* s1, int32_t*, uint8_t*, int32_t* st) {
                      return sfe_track_dev(ctx, tmpl, tmpl_first, search, search_first, n, n_per_pair, df, dt, dl,
                                           default_levels, thr, maxit, s1, st);
                    });
}

int sfe_get_patches(sfe_ctx* ctx, const sfe_pyr* pyr, int frame, int level, int n, const float* xy, float* patches,
                    float* mean, float* sumsq) {
  if (!ctx || !pyr || n < 0 || !xy || !patches || !mean || !sumsq || frame < 0 || frame >= pyr->view.batch || level < 0 ||
      level >= pyr->view.depth)
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad get_patches arguments");
  if (n == 0) return SFE_SUCCESS;
  if (use_device(ctx)) return SFE_ERR_CUDA;
  size_t need = padded(8 * (size_t)n) + padded(4 * (size_t)n * SFE_PLEN) + 2 * padded(4 * (size_t)n);
  int rc = ensure_scratch(ctx, need);
  if (rc) return rc;
  Carver c{(char*)ctx->scratch, 0};
  float* d_xy = c.take<float>(2 * (size_t)n);
  float* d_p = c.take<float>((size_t)n * SFE_PLEN);
  float* d_m = c.take<float>(n);
  float* d_q = c.take<float>(n);
  cudaStream_t s = ctx->stream;
  CU(cudaMemcpyAsync(d_xy, xy, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
  rc = launched(ctx, launch_get_patches(pyr->view, frame, level, n, d_xy, d_p, d_m, d_q, s), "get_patches launch: %s");
  if (rc) return rc;
  CU(cudaMemcpyAsync(patches, d_p, 4 * (size_t)n * SFE_PLEN, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(mean, d_m, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(sumsq, d_q, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SFE_SUCCESS;
}


int sfe_brute_hessian(sfe_ctx* ctx, const sfe_pyr* tmpl, int tmpl_frame, const sfe_pyr* search, int search_frame,
                      int level, int n, const float* tmpl_xy, const float* xy, float* out7) {
  if (!ctx || n < 0 || !tmpl_xy || !xy || !out7 || !frame_ok(tmpl, tmpl_frame, level) || !frame_ok(search, search_frame, level))
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad brute_hessian arguments");
  if (n == 0) return SFE_SUCCESS;
  if (use_device(ctx)) return SFE_ERR_CUDA;
  return point_pair_host(ctx, n, tmpl_xy, xy, out7, 7, [&](float* dt, float* dx, float* dout, cudaStream_t s) {
    return launch_brute_hessian(tmpl->view, tmpl_frame, search->view, search_frame, level, n, dt, dx, dout, ctx->d_mask, s);
  });
}

/* ---- P2 ---------------------------------------------------------------------------------- */

int sfe_klt_track_fb_dev(sfe_ctx* ctx, const sfe_pyr* from, int from_first, const sfe_pyr* to, int to_first, int n,
                         int n_per_pair, const float* from_xy, float* to_xy, float thr, int maxit, double fb_max,
                         float* back_xy, int32_t* status_fwd, int32_t* status_bwd, uint8_t* accepted, int32_t* steps) {
  if (!ctx) return SFE_ERR_INVALID;
  int rc = check_pairs(ctx, from, from_first, to, to_first, n, n_per_pair);
  if (rc || n == 0) return rc;
  if (!from_xy || !to_xy || maxit < 0) return fail(ctx, SFE_ERR_INVALID, "%s", "bad track arguments");
  if (from->flavor != SFE_KLT || to->flavor != SFE_KLT)
    return fail(ctx, SFE_ERR_INVALID, "%s", "sfe_klt_track_fb needs SFE_KLT pyramids");
  if (use_device(ctx)) return SFE_ERR_CUDA;
  TrackArgs a{n, n_per_pair, from_first, to_first, from_xy, to_xy, nullptr, from->view.depth, thr, maxit, fb_max,
              back_xy, status_fwd, status_bwd, accepted, steps, 2};
  return launched(ctx, launch_track_klt(from->view, to->view, a, ctx->d_mask, sfe_next_counter(ctx), ctx->num_sms, ctx->stream), "klt launch: %s");
}

int sfe_klt_track_fb(sfe_ctx* ctx, const sfe_pyr* from, int from_first, const sfe_pyr* to, int to_first, int n,
                     int n_per_pair, const float* from_xy, float* to_xy, float thr, int maxit, double fb_max,
                     float* back_xy, int32_t* status_fwd, int32_t* status_bwd, uint8_t* accepted, int32_t* steps) {
  if (!ctx) return SFE_ERR_INVALID;
  if (n == 0) return SFE_SUCCESS;
  if (n < 0 || !from_xy || !to_xy) return fail(ctx, SFE_ERR_INVALID, "%s", "bad track arguments");
  if (use_device(ctx)) return SFE_ERR_CUDA;
  return track_host(ctx, n, from_xy, to_xy, nullptr, back_xy, status_fwd, status_bwd, accepted, steps,
                    [&](float* df, float* dt, int32_t*, float* db, int32_t* s1, int32_t* s2, uint8_t* acc, int32_t* st) {
                      return sfe_klt_track_fb_dev(ctx, from, from_first, to, to_first, n, n_per_pair, df, dt, thr, maxit,
                                                  fb_max, db, s1, s2, acc, st);
                    });
}

int sfe_klt_system(sfe_ctx* ctx, const sfe_pyr* tmpl, int tmpl_frame, const sfe_pyr* search, int search_frame, int level,
                   int n, const float* tmpl_xy, const float* xy, float* out24) {
  if (!ctx || n < 0 || !tmpl_xy || !xy || !out24 || !frame_ok(tmpl, tmpl_frame, level) || !frame_ok(search, search_frame, level))
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad klt_system arguments");
  if (tmpl->flavor != SFE_KLT || search->flavor != SFE_KLT)
    return fail(ctx, SFE_ERR_INVALID, "%s", "sfe_klt_system needs SFE_KLT pyramids");
  if (n == 0) return SFE_SUCCESS;
  if (use_device(ctx)) return SFE_ERR_CUDA;
  return point_pair_host(ctx, n, tmpl_xy, xy, out24, 24, [&](float* dt, float* dx, float* dout, cudaStream_t s) {
    return launch_klt_system(tmpl->view, tmpl_frame, search->view, search_frame, level, n, dt, dx, dout, ctx->d_mask, s);
  });
}

/* ---- P3 ---------------------------------------------------------------------------------- */

int sfe_brute_track_dev(sfe_ctx* ctx, const sfe_pyr* from, int from_first, const sfe_pyr* to, int to_first, int n,
                        int n_per_pair, const float* from_xy, float* to_xy, const float* coarse_sched, int n_coarse,
                        const float* fine_sched, int n_fine, int32_t* status, float* best_sad) {
  if (!ctx) return SFE_ERR_INVALID;
  int rc = check_pairs(ctx, from, from_first, to, to_first, n, n_per_pair);
  if (rc || n == 0) return rc;
  if (!from_xy || !to_xy || n_coarse < 0 || n_fine < 0 || n_coarse > 8 || n_fine > 8 || (n_coarse && !coarse_sched) ||
      (n_fine && !fine_sched))
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad brute_track arguments");
  if (use_device(ctx)) return SFE_ERR_CUDA;
  // schedules are host arrays even in the _dev variant (they are tiny call parameters)
  return launched(ctx,
                  launch_brute_track(from->view, to->view, from_first, to_first, n, n_per_pair, from_xy, to_xy, coarse_sched,
                                     n_coarse, fine_sched, n_fine, status, best_sad, nullptr, ctx->stream),
                  "brute_track launch: %s");
}

int sfe_brute_track(sfe_ctx* ctx, const sfe_pyr* from, int from_first, const sfe_pyr* to, int to_first, int n,
                    int n_per_pair, const float* from_xy, float* to_xy, const float* coarse_sched, int n_coarse,
                    const float* fine_sched, int n_fine, int32_t* status, float* best_sad, int64_t* positions) {
  if (!ctx) return SFE_ERR_INVALID;
  if (n == 0) return SFE_SUCCESS;
  int rc = check_pairs(ctx, from, from_first, to, to_first, n, n_per_pair);
  if (rc) return rc;
  if (!from_xy || !to_xy || n_coarse < 0 || n_fine < 0 || n_coarse > 8 || n_fine > 8 || (n_coarse && !coarse_sched) ||
      (n_fine && !fine_sched))
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad brute_track arguments");
  if (use_device(ctx)) return SFE_ERR_CUDA;
  size_t need = 2 * padded(8 * (size_t)n) + 2 * padded(4 * (size_t)n) + 256;
  rc = ensure_scratch(ctx, need);
  if (rc) return rc;
  Carver c{(char*)ctx->scratch, 0};
  float* d_from = c.take<float>(2 * (size_t)n);
  float* d_to = c.take<float>(2 * (size_t)n);
  int32_t* d_st = c.take<int32_t>(n);
  float* d_sad = c.take<float>(n);
  unsigned long long* d_pos = c.take<unsigned long long>(1);
  cudaStream_t s = ctx->stream;
  CU(cudaMemcpyAsync(d_from, from_xy, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(d_to, to_xy, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(d_pos, 0, sizeof(unsigned long long), s));
  rc = launched(ctx,
                launch_brute_track(from->view, to->view, from_first, to_first, n, n_per_pair, d_from, d_to, coarse_sched,
                                   n_coarse, fine_sched, n_fine, d_st, d_sad, d_pos, s),
                "brute_track launch: %s");
  if (rc) return rc;
  CU(cudaMemcpyAsync(to_xy, d_to, 8 * (size_t)n, cudaMemcpyDeviceToHost, s));
  if (status) CU(cudaMemcpyAsync(status, d_st, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
  if (best_sad) CU(cudaMemcpyAsync(best_sad, d_sad, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
  unsigned long long pos = 0;
  CU(cudaMemcpyAsync(&pos, d_pos, sizeof(pos), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (positions) *positions = (int64_t)pos;
  return SFE_SUCCESS;
}

/* ---- P4 ---------------------------------------------------------------------------------- */

int sfe_match_hamming256_dev(sfe_ctx* ctx, const uint32_t* q, int nq, const uint32_t* t, int nt, int batch,
                             int ratio_num, int ratio_den, int max_dist, int32_t* idx, int32_t* dist, uint8_t* pass) {
  if (!ctx || nq < 0 || nt < 0 || batch < 1 || !idx || !dist || (nq && !q) || (nt && !t))
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad hamming arguments");
  if (nt > (1 << 22)) return fail(ctx, SFE_ERR_INVALID, "%s", "nt exceeds 4M train descriptors per call");
  if (nq == 0) return SFE_SUCCESS;
  if (use_device(ctx)) return SFE_ERR_CUDA;
  // the workspace is shared with an asynchronous match that may still be running on the side stream
  if (ctx->ham_pending) CU(cudaStreamWaitEvent(ctx->stream, ctx->ham_done, 0));
  return launched(ctx,
                  launch_hamming256(q, nq, t, nt, batch, ratio_num, ratio_den, max_dist, idx, dist, pass, &ctx->ham_ws,
                                    &ctx->ham_cap, ctx->stream),
                  "hamming launch: %s");
}

int sfe_match_hamming256(sfe_ctx* ctx, const uint32_t* q, int nq, const uint32_t* t, int nt, int batch, int ratio_num,
                         int ratio_den, int max_dist, int32_t* idx, int32_t* dist, uint8_t* pass) {
  if (!ctx || nq < 0 || nt < 0 || batch < 1 || !idx || !dist || (nq && !q) || (nt && !t))
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad hamming arguments");
  if (nq == 0) return SFE_SUCCESS;
  if (use_device(ctx)) return SFE_ERR_CUDA;
  const size_t qb = 32 * (size_t)nq * batch, tb = 32 * (size_t)nt * batch, ob = 8 * (size_t)nq * batch;
  size_t need = padded(qb) + padded(tb) + 2 * padded(ob) + padded((size_t)nq * batch);
  int rc = ensure_scratch(ctx, need);
  if (rc) return rc;
  Carver c{(char*)ctx->scratch, 0};
  uint32_t* d_q = c.take<uint32_t>(8 * (size_t)nq * batch);
  uint32_t* d_t = c.take<uint32_t>(8 * (size_t)nt * batch);
  int32_t* d_i = c.take<int32_t>(2 * (size_t)nq * batch);
  int32_t* d_d = c.take<int32_t>(2 * (size_t)nq * batch);
  uint8_t* d_p = c.take<uint8_t>((size_t)nq * batch);
  cudaStream_t s = ctx->stream;
  CU(cudaMemcpyAsync(d_q, q, qb, cudaMemcpyHostToDevice, s));
  if (nt) CU(cudaMemcpyAsync(d_t, t, tb, cudaMemcpyHostToDevice, s));
  rc = sfe_match_hamming256_dev(ctx, d_q, nq, d_t, nt, batch, ratio_num, ratio_den, max_dist, d_i, d_d, pass ? d_p : nullptr);
  if (rc) return rc;
  CU(cudaMemcpyAsync(idx, d_i, ob, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(dist, d_d, ob, cudaMemcpyDeviceToHost, s));
  if (pass) CU(cudaMemcpyAsync(pass, d_p, (size_t)nq * batch, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SFE_SUCCESS;
}

int sfe_match_hamming256_async(sfe_ctx* ctx, const uint32_t* q, int nq, const uint32_t* t, int nt, int batch,
                               int ratio_num, int ratio_den, int max_dist, int32_t* idx, int32_t* dist, uint8_t* pass) {
  if (!ctx || nq < 0 || nt < 0 || batch < 1 || !idx || !dist || (nq && !q) || (nt && !t))
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad hamming arguments");
  if (nt > (1 << 22)) return fail(ctx, SFE_ERR_INVALID, "%s", "nt exceeds 4M train descriptors per call");
  if (nq == 0) return SFE_SUCCESS;
  if (use_device(ctx)) return SFE_ERR_CUDA;
  const size_t qb = 32 * (size_t)nq * batch, tb = 32 * (size_t)nt * batch, ob = 8 * (size_t)nq * batch;
  const size_t need = padded(qb) + padded(tb) + 2 * padded(ob) + padded((size_t)nq * batch);
  if (need > ctx->ham_io_cap) {
    if (ctx->ham_stream) CU(cudaStreamSynchronize(ctx->ham_stream));  // a previous asynchronous call may still be using the old buffers
    if (ctx->ham_io) CU(cudaFree(ctx->ham_io));
    ctx->ham_io = nullptr;
    ctx->ham_io_cap = 0;
    cudaError_t e = cudaMalloc(&ctx->ham_io, need + need / 4);
    if (e != cudaSuccess) return fail(ctx, SFE_ERR_NOMEM, "cudaMalloc(hamming io): %s", cudaGetErrorString(e));
    ctx->ham_io_cap = need + need / 4;
  }
  Carver c{(char*)ctx->ham_io, 0};
  uint32_t* d_q = c.take<uint32_t>(8 * (size_t)nq * batch);
  uint32_t* d_t = c.take<uint32_t>(8 * (size_t)nt * batch);
  int32_t* d_i = c.take<int32_t>(2 * (size_t)nq * batch);
  int32_t* d_d = c.take<int32_t>(2 * (size_t)nq * batch);
  uint8_t* d_p = c.take<uint8_t>((size_t)nq * batch);
  if (!ctx->ham_stream) {
    CU(cudaStreamCreateWithFlags(&ctx->ham_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&ctx->ham_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ham_done, cudaEventDisableTiming));
  }
  // fork: the match is ordered after everything enqueued on the context's stream so far, but nothing enqueued
  // later waits for it; sfe_sync() joins
  cudaStream_t s = ctx->ham_stream;
  CU(cudaEventRecord(ctx->ham_fork, ctx->stream));
  CU(cudaStreamWaitEvent(s, ctx->ham_fork, 0));
  CU(cudaMemcpyAsync(d_q, q, qb, cudaMemcpyHostToDevice, s));
  if (nt) CU(cudaMemcpyAsync(d_t, t, tb, cudaMemcpyHostToDevice, s));
  int rc = launched(ctx,
                    launch_hamming256(d_q, nq, d_t, nt, batch, ratio_num, ratio_den, max_dist, d_i, d_d, pass ? d_p : nullptr,
                                      &ctx->ham_ws, &ctx->ham_cap, s),
                    "hamming launch: %s");
  if (rc) return rc;
  CU(cudaMemcpyAsync(idx, d_i, ob, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(dist, d_d, ob, cudaMemcpyDeviceToHost, s));
  if (pass) CU(cudaMemcpyAsync(pass, d_p, (size_t)nq * batch, cudaMemcpyDeviceToHost, s));
  CU(cudaEventRecord(ctx->ham_done, s));
  ctx->ham_pending = true;
  return SFE_SUCCESS;
}

int sfe_hamming_impl(int impl) { return hamming_set_impl(impl); }

/* ---- corner seeding -------------------------------------------------- */

namespace {
int gftt_cap(int w, int h) {  // candidate capacity per frame: a power of two (the bitonic sort pads to one) that holds the
  // worst case of distinct responses -- 3x3 maxima cannot be 8-neighbours of each other unless they are exactly equal,
  // so at most one per 2x2 block: a densely textured frame proceeds as it does in the reference instead of failing with
  // an overflow (only plateaus of bit-identical non-zero responses can still exceed it; that is reported as an error)
  int cap = 1024;
  while (cap < ((w + 1) / 2) * ((h + 1) / 2)) cap <<= 1;
  return cap;
}
// Up to this many frames per call take the two-pass response kernels (gftt.cu): the live robot's keyframe, not a batch.
// SFE_GFTT_TWO_PASS=0 / =N (experiments, tests) moves the limit.
int gftt_two_pass_limit() {
  static const int v = getenv("SFE_GFTT_TWO_PASS") ? atoi(getenv("SFE_GFTT_TWO_PASS")) : 4;
  return v;
}
size_t gftt_ws_bytes(int w, int h, int count) {
  return padded(sizeof(float) * (size_t)w * h * count) + 2 * padded(sizeof(int) * (size_t)count) +
         padded(sizeof(unsigned long long) * (size_t)gftt_cap(w, h) * count) +
         (count <= gftt_two_pass_limit() ? padded((sizeof(double) + sizeof(float)) * 3 * (size_t)w * h * count) : 0);
}
int gftt_run(sfe_ctx* ctx, const uint8_t* bgr_dev, int w, int h, size_t row_stride, size_t frame_stride, int count,
             int max_corners, double quality, double min_distance, float* corners_dev, int32_t* ncorners_dev, float** eig_dev) {
  const size_t need = gftt_ws_bytes(w, h, count);
  if (need > ctx->gftt_cap) {
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->gftt_ws) CU(cudaFree(ctx->gftt_ws));
    ctx->gftt_ws = nullptr;
    ctx->gftt_cap = 0;
    cudaError_t e = cudaMalloc(&ctx->gftt_ws, need);
    if (e != cudaSuccess) return fail(ctx, SFE_ERR_NOMEM, "cudaMalloc(corner workspace): %s", cudaGetErrorString(e));
    ctx->gftt_cap = need;
  }
  Carver c{(char*)ctx->gftt_ws, 0};
  float* eig = c.take<float>((size_t)w * h * count);
  unsigned* mx = c.take<unsigned>(count);
  int* nc = c.take<int>(count);
  unsigned long long* keys = c.take<unsigned long long>((size_t)gftt_cap(w, h) * count);
  // row sums (doubles) and, right behind them, box sums (floats) of the few-frame path
  double* rowsums = count <= gftt_two_pass_limit() ? (double*)c.take<char>((sizeof(double) + sizeof(float)) * 3 * (size_t)w * h * count) : nullptr;
  if (eig_dev) *eig_dev = eig;
  return launched(ctx,
                  launch_good_features(bgr_dev, row_stride, frame_stride, w, h, count, max_corners, quality, min_distance, eig, mx,
                                       nc, keys, gftt_cap(w, h), corners_dev, ncorners_dev, rowsums, ctx->stream),
                  "good_features launch: %s");
}
bool gftt_args_ok(int w, int h, size_t row_stride, int count, int max_corners, double quality, double min_distance) {
  return w >= 16 && h >= 16 && row_stride >= (size_t)3 * w && count >= 0 && max_corners >= 1 && max_corners <= 8192 &&
         quality > 0 && min_distance >= 0;
}
}  // namespace

int sfe_good_features_dev(sfe_ctx* ctx, const uint8_t* bgr_dev, int w, int h, size_t row_stride, size_t frame_stride, int count,
                          int max_corners, double quality, double min_distance, float* corners, int32_t* ncorners) {
  if (!ctx || !bgr_dev || !corners || !ncorners || !gftt_args_ok(w, h, row_stride, count, max_corners, quality, min_distance))
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad good_features arguments");
  if (count == 0) return SFE_SUCCESS;
  if (use_device(ctx)) return SFE_ERR_CUDA;
  return gftt_run(ctx, bgr_dev, w, h, row_stride, frame_stride, count, max_corners, quality, min_distance, corners, ncorners, nullptr);
}

int sfe_good_features(sfe_ctx* ctx, const uint8_t* bgr_host, int w, int h, size_t row_stride, size_t frame_stride, int count,
                      int max_corners, double quality, double min_distance, float* corners, int32_t* ncorners, float* eig_out) {
  if (!ctx || !bgr_host || !corners || !ncorners || !gftt_args_ok(w, h, row_stride, count, max_corners, quality, min_distance))
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad good_features arguments");
  if (count == 0) return SFE_SUCCESS;
  if (use_device(ctx)) return SFE_ERR_CUDA;
  const size_t dense_row = (size_t)3 * w, dense_frame = dense_row * h;
  const size_t need = padded(dense_frame * count) + padded(8 * (size_t)max_corners * count) + padded(4 * (size_t)count);
  int rc = ensure_scratch(ctx, need);
  if (rc) return rc;
  Carver c{(char*)ctx->scratch, 0};
  uint8_t* d_bgr = c.take<uint8_t>(dense_frame * count);
  float* d_xy = c.take<float>(2 * (size_t)max_corners * count);
  int32_t* d_n = c.take<int32_t>(count);
  cudaStream_t s = ctx->stream;
  if (row_stride == dense_row && (frame_stride == dense_frame || count == 1)) {
    CU(cudaMemcpyAsync(d_bgr, bgr_host, dense_frame * count, cudaMemcpyHostToDevice, s));
  } else {
    for (int f = 0; f < count; ++f)
      CU(cudaMemcpy2DAsync(d_bgr + f * dense_frame, dense_row, bgr_host + f * frame_stride, row_stride, dense_row, h,
                           cudaMemcpyHostToDevice, s));
  }
  CU(cudaMemsetAsync(d_xy, 0, 8 * (size_t)max_corners * count, s));
  float* d_eig = nullptr;
  rc = gftt_run(ctx, d_bgr, w, h, dense_row, dense_frame, count, max_corners, quality, min_distance, d_xy, d_n, &d_eig);
  if (rc) return rc;
  CU(cudaMemcpyAsync(corners, d_xy, 8 * (size_t)max_corners * count, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(ncorners, d_n, 4 * (size_t)count, cudaMemcpyDeviceToHost, s));
  if (eig_out) CU(cudaMemcpyAsync(eig_out, d_eig, sizeof(float) * (size_t)w * h * count, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  for (int f = 0; f < count; ++f)
    if (ncorners[f] < 0) return fail(ctx, SFE_ERR_INVALID, "%s", "corner candidate list overflowed (raise quality)");
  return SFE_SUCCESS;
}

/* ---- seeding the search and the live capture format ---------------- */

int sfe_seed_features_dev(sfe_ctx* ctx, int n, const double* points4, const double* uncertainty, const double* rot4,
                          const double* trans3, const double* k7, const float* from_xy, int cols, int rows, float* seed_xy,
                          int32_t* levels, uint8_t* go) {
  if (!ctx || n < 0 || !rot4 || !trans3 || !k7 || (n && (!points4 || !uncertainty || !from_xy || !seed_xy || !levels || !go)))
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad seed_features arguments");
  if (n == 0) return SFE_SUCCESS;
  if (use_device(ctx)) return SFE_ERR_CUDA;
  return launched(ctx, launch_seed_features(n, points4, uncertainty, rot4, trans3, k7, from_xy, cols, rows, seed_xy, levels, go,
                                            ctx->stream), "seed_features launch: %s");
}

int sfe_seed_features(sfe_ctx* ctx, int n, const double* points4, const double* uncertainty, const double* rot4,
                      const double* trans3, const double* k7, const float* from_xy, int cols, int rows, float* seed_xy,
                      int32_t* levels, uint8_t* go) {
  if (!ctx || n < 0 || !rot4 || !trans3 || !k7 || (n && (!points4 || !uncertainty || !from_xy || !seed_xy || !levels || !go)))
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad seed_features arguments");
  if (n == 0) return SFE_SUCCESS;
  if (use_device(ctx)) return SFE_ERR_CUDA;
  const size_t need = padded(32 * (size_t)n) + padded(8 * (size_t)n) * 3 + padded(4 * (size_t)n) + padded(n);
  int rc = ensure_scratch(ctx, need);
  if (rc) return rc;
  Carver c{(char*)ctx->scratch, 0};
  double* d_pt = c.take<double>(4 * (size_t)n);
  double* d_unc = c.take<double>(n);
  float* d_from = c.take<float>(2 * (size_t)n);
  float* d_seed = c.take<float>(2 * (size_t)n);
  int32_t* d_lv = c.take<int32_t>(n);
  uint8_t* d_go = c.take<uint8_t>(n);
  cudaStream_t s = ctx->stream;
  CU(cudaMemcpyAsync(d_pt, points4, 32 * (size_t)n, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(d_unc, uncertainty, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(d_from, from_xy, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
  rc = sfe_seed_features_dev(ctx, n, d_pt, d_unc, rot4, trans3, k7, d_from, cols, rows, d_seed, d_lv, d_go);
  if (rc) return rc;
  CU(cudaMemcpyAsync(seed_xy, d_seed, 8 * (size_t)n, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(levels, d_lv, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(go, d_go, (size_t)n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SFE_SUCCESS;
}

int sfe_yuyv_to_bgr_dev(sfe_ctx* ctx, const uint8_t* yuyv, size_t npixels, uint8_t* bgr) {
  if (!ctx || (npixels && (!yuyv || !bgr)) || npixels % 4 || ((uintptr_t)yuyv & 7) || ((uintptr_t)bgr & 3))
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad yuyv_to_bgr arguments (pixel count must be a multiple of 4, buffers 8/4-byte aligned)");
  if (npixels == 0) return SFE_SUCCESS;
  if (use_device(ctx)) return SFE_ERR_CUDA;
  return launched(ctx, launch_yuyv_to_bgr(yuyv, npixels, bgr, ctx->stream), "yuyv_to_bgr launch: %s");
}

int sfe_yuyv_to_bgr(sfe_ctx* ctx, const uint8_t* yuyv, size_t npixels, uint8_t* bgr) {
  if (!ctx || (npixels && (!yuyv || !bgr)) || npixels % 4)
    return fail(ctx, SFE_ERR_INVALID, "%s", "bad yuyv_to_bgr arguments (pixel count must be a multiple of 4)");
  if (npixels == 0) return SFE_SUCCESS;
  if (use_device(ctx)) return SFE_ERR_CUDA;
  int rc = ensure_scratch(ctx, padded(2 * npixels) + padded(3 * npixels));
  if (rc) return rc;
  Carver c{(char*)ctx->scratch, 0};
  uint8_t* d_in = c.take<uint8_t>(2 * npixels);
  uint8_t* d_out = c.take<uint8_t>(3 * npixels);
  CU(cudaMemcpyAsync(d_in, yuyv, 2 * npixels, cudaMemcpyHostToDevice, ctx->stream));
  rc = sfe_yuyv_to_bgr_dev(ctx, d_in, npixels, d_out);
  if (rc) return rc;
  CU(cudaMemcpyAsync(bgr, d_out, 3 * npixels, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SFE_SUCCESS;
}

}  // extern "C"
