// gpu_tracker.hpp -- header-only C++ mirror of the reference's tracker "duck type" on top of the C ABI
// (include/slamfe.h).  matcher.cpp selects its tracker with a typedef
//     typedef HessianTracker FeatureTracker;                      (matcher.cpp:21)
// and uses exactly these members:
//     FeatureTracker(cv::Size)                                    matcher.cpp:304
//     Pyramid  MakePyramid(const cv::Mat&, int depth)             matcher.cpp:317   hessian.h:95-126
//     vector<Patch> GetPatches(const Pyramid&, Point2f, int)      matcher.cpp:175   hessian.h:175-183
//     Status   TrackFeature(const Pyramid&, const vector<Patch>&,
//                           float thr, int maxit, Point2f*)       matcher.cpp:176   hessian.h:243-264
//     Patch::size, Patch::data                                    matcher.cpp:94,108
//     Status truthiness == failure                                matcher.cpp:192
// sfe::GpuTracker provides the same names with the same argument meaning and error behaviour, so
// `typedef sfe::GpuTracker FeatureTracker;` compiles against the rest of matcher.cpp, plus the batched
// entry point (TrackFeaturesFB) that a GPU needs to be worth using.  No OpenCV headers are required:
// images are passed as anything with .data/.cols/.rows/.step (cv::Mat qualifies).
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/slamfe.h"

namespace sfe {

struct Size {
  int width, height;
  Size(int w = 0, int h = 0) : width(w), height(h) {}
};

struct Point2f {
  float x, y;
  Point2f(float x_ = 0, float y_ = 0) : x(x_), y(y_) {}
};

// Borrowed view of an 8-bit 3-channel interleaved frame (what cv::Mat CV_8UC3 is).
struct ImageView {
  const uint8_t* data;
  int cols, rows;
  size_t step;
};

class Error : public std::runtime_error {
 public:
  explicit Error(const std::string& what) : std::runtime_error(what) {}
};

// One CUDA context shared by trackers and pyramids (RAII over sfe_ctx).
class Context {
 public:
  explicit Context(int device = 0) {
    if (sfe_create(device, &ctx_) != SFE_SUCCESS)
      throw Error("sfe_create failed: no usable CUDA device (the front-end has no CPU fallback)");
  }
  ~Context() { sfe_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  sfe_ctx* get() const { return ctx_; }
  void check(int rc) const {
    if (rc != SFE_SUCCESS) throw Error(std::string("libslamfe: ") + sfe_last_error(ctx_));
  }

 private:
  sfe_ctx* ctx_ = nullptr;
};

class GpuTracker {
 public:
  enum Status { OK = SFE_OK, SMALL_DET = SFE_SMALL_DET, OUT_OF_BOUNDS = SFE_OUT_OF_BOUNDS };  // hessian.h:48-52

  // Pyramid = vector<GradImage> in the reference; here a shared handle to device storage.  Copies are
  // cheap and share the device planes, like cv::Mat headers do.
  class Pyramid {
   public:
    Pyramid() {}
    size_t size() const { return h_ ? (size_t)depth_ : 0; }  // number of levels, like vector::size()
    bool empty() const { return !h_; }
    sfe_pyr* handle() const { return h_.get(); }
    int frame() const { return frame_; }

   private:
    friend class GpuTracker;
    std::shared_ptr<sfe_pyr> h_;
    int depth_ = 0;
    int frame_ = 0;  // slot inside a batched pyramid object
  };

  // hessian.h:32-40, plus where the patch came from so TrackFeature can name its template on the device
  struct Patch {
    Size size;
    std::vector<float> data;
    float mean = 0, sumsq = 0;
    Pyramid source;
    Point2f source_pt;  // level-0 coordinates of the GetPatches call
    int level = 0;
  };

  // matcher.cpp:304.  The reference rebuilds the 13x13 mask on every Track() call; here the context
  // built it once (sfe_create).
  explicit GpuTracker(Size size, std::shared_ptr<Context> ctx = nullptr) : size_(size), ctx_(ctx) {
    if (size.width != SFE_PATCH || size.height != SFE_PATCH) throw Error("GpuTracker supports the reference's 13x13 window only");
    if (!ctx_) ctx_ = std::make_shared<Context>(0);
  }

  const std::shared_ptr<Context>& context() const { return ctx_; }

  // MakePyramid (hessian.h:95-126).  Mat: anything with data/cols/rows/step (cv::Mat, ImageView).
  template <class Mat>
  Pyramid MakePyramid(const Mat& img, int depth) const {
    if (!img.data || img.cols <= 0 || img.rows <= 0) throw Error("MakePyramid: empty image");  // CHECK_NE, matcher.cpp:306-307
    sfe_pyr* raw = nullptr;
    ctx_->check(sfe_pyr_create(ctx_->get(), img.cols, img.rows, depth, SFE_HESSIAN, 1, &raw));
    Pyramid p;
    p.h_ = std::shared_ptr<sfe_pyr>(raw, [c = ctx_](sfe_pyr* q) { sfe_pyr_destroy(q); });
    p.depth_ = depth;
    ctx_->check(sfe_pyr_build(ctx_->get(), raw, (const uint8_t*)img.data, (size_t)img.step, (size_t)img.step * img.rows, 0, 1));
    return p;
  }

  // GetPatches (hessian.h:175-183): levels = min(stack.size(), levels); patch i is taken at pt * 0.5^i.
  // The pixel data is materialised on the host (matcher.cpp only looks at it for debug drawing).
  std::vector<Patch> GetPatches(const Pyramid& stack, Point2f pt, int levels) const {
    if (stack.empty()) throw Error("GetPatches: empty pyramid");
    levels = levels < (int)stack.size() ? levels : (int)stack.size();
    std::vector<Patch> out;
    Point2f p = pt;
    for (int i = 0; i < levels; ++i) {
      Patch patch;
      patch.size = size_;
      patch.data.resize(SFE_PATCH * SFE_PATCH);
      patch.source = stack;
      patch.source_pt = pt;
      patch.level = i;
      const float xy[2] = {p.x, p.y};
      ctx_->check(sfe_get_patches(ctx_->get(), stack.handle(), stack.frame(), i, 1, xy, patch.data.data(), &patch.mean, &patch.sumsq));
      out.push_back(std::move(patch));
      p.x *= 0.5f;
      p.y *= 0.5f;
    }
    return out;
  }

  // TrackFeature (hessian.h:243-264): lvls = min(stack.size(), patches.size()); *pt is written only on
  // success.  The template patches are re-derived on the device from their recorded source.
  Status TrackFeature(const Pyramid& stack, const std::vector<Patch>& patches, float threshold, int max_iterations, Point2f* pt) const {
    if (patches.empty() || stack.empty()) throw Error("TrackFeature: empty patch stack or pyramid");
    const Patch& p0 = patches[0];
    const int levels = (int)(stack.size() < patches.size() ? stack.size() : patches.size());
    const float txy[2] = {p0.source_pt.x, p0.source_pt.y};
    float xy[2] = {pt->x, pt->y};
    int32_t st = OK;
    ctx_->check(sfe_track(ctx_->get(), p0.source.handle(), p0.source.frame(), stack.handle(), stack.frame(), 1, 1, txy, xy, nullptr,
                          levels, threshold, max_iterations, &st, nullptr));
    if (st == OK) { pt->x = xy[0]; pt->y = xy[1]; }
    return (Status)st;
  }

  // ---- batched entry point: the free function TrackFeature of matcher.cpp:173-206 for n features
  struct FBResult {
    std::vector<Point2f> to_pt, back_pt;
    std::vector<int32_t> status_fwd, status_bwd, steps;
    std::vector<uint8_t> accepted;
  };
  FBResult TrackFeaturesFB(const Pyramid& from, const Pyramid& to, const std::vector<Point2f>& from_pt,
                           const std::vector<Point2f>& seed, const std::vector<int32_t>& levels, float threshold = 0.001f,
                           int max_iterations = 10, double fb_max = 0.3) const {
    const int n = (int)from_pt.size();
    if ((int)seed.size() != n || (!levels.empty() && (int)levels.size() != n)) throw Error("TrackFeaturesFB: size mismatch");
    FBResult r;
    r.to_pt = seed;
    r.back_pt.resize(n);
    r.status_fwd.resize(n);
    r.status_bwd.resize(n);
    r.steps.resize(n);
    r.accepted.resize(n);
    if (n == 0) return r;
    static_assert(sizeof(Point2f) == 2 * sizeof(float), "Point2f must be two packed floats");
    ctx_->check(sfe_track_fb(ctx_->get(), from.handle(), from.frame(), to.handle(), to.frame(), n, n, &from_pt[0].x, &r.to_pt[0].x,
                             levels.empty() ? nullptr : levels.data(), 3, threshold, max_iterations, fb_max, &r.back_pt[0].x,
                             r.status_fwd.data(), r.status_bwd.data(), r.accepted.data(), r.steps.data()));
    return r;
  }

 private:
  Size size_;
  std::shared_ptr<Context> ctx_;
};

}  // namespace sfe
