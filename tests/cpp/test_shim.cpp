// Exercises the C++ host mirror (slam-robot_b200/host) the way matcher.cpp uses its tracker, with light
// stand-ins for the reference's map classes.  Driven by tests/test_gpu_shim.py, which writes the input
// frames / corners, runs this program on the GPU and compares its dump with the CPU oracle.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "../../slam-robot_b200/host/matcher_gpu.hpp"

using sfe::GpuTracker;
using sfe::Point2f;

struct Vec2 {
  double v[2];
  Vec2(double a = 0, double b = 0) { v[0] = a; v[1] = b; }
  double operator()(int i) const { return v[i]; }
};
struct Vec3 { double x, y, z; };
struct TrackedPoint {
  int id_; Vec3 loc; double unc = 1e8;  // localmap.h:179: new points start with uncertainty 1e8
  int id() const { return id_; }
  double uncertainty() const { return unc; }
  bool feature_usable() const { return true; }
  Vec3 location() const { return loc; }
};
struct Camera { Vec2 PixelToPlane(const Vec2& p) const { return p; } };
struct Obs { int frame, point; double x, y; };
static std::vector<Obs> g_obs;
struct Frame {
  int id_; bool is_keyframe_ = false; Camera cam;
  int id() const { return id_; }
  Camera* camera() { return &cam; }
  bool Project(const Vec3&, Vec2*) const { return false; }
  Vec3 Unproject(const Vec2& p, double d) const { return Vec3{p(0), p(1), d}; }
  void AddObservation(const Vec2& p, TrackedPoint* pt) { g_obs.push_back(Obs{id_, pt->id(), p(0), p(1)}); }
};
struct LocalMap {
  std::vector<TrackedPoint*> pts;
  TrackedPoint* AddPoint(int id, const Vec3& loc) { pts.push_back(new TrackedPoint{id, loc}); return pts.back(); }
};

static std::vector<unsigned char> read_file(const char* path) {
  std::ifstream f(path, std::ios::binary);
  return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
  if (argc < 7) { std::fprintf(stderr, "usage: test_shim A.bin B.bin corners.bin W H out.txt\n"); return 2; }
  const int W = std::atoi(argv[4]), H = std::atoi(argv[5]);
  std::vector<unsigned char> A = read_file(argv[1]), B = read_file(argv[2]), cb = read_file(argv[3]);
  if ((int)A.size() != W * H * 3 || (int)B.size() != W * H * 3) { std::fprintf(stderr, "bad frame size\n"); return 2; }
  std::vector<Point2f> corners(cb.size() / 8);
  for (size_t i = 0; i < corners.size(); ++i) { const float* p = (const float*)cb.data() + 2 * i; corners[i] = Point2f(p[0], p[1]); }
  sfe::ImageView imA{A.data(), W, H, (size_t)W * 3}, imB{B.data(), W, H, (size_t)W * 3};
  FILE* out = std::fopen(argv[6], "w");

  // ---- 1. the duck type, used exactly like matcher.cpp:173-206 does
  auto ctx = std::make_shared<sfe::Context>(0);
  GpuTracker tracker(sfe::Size(13, 13), ctx);
  GpuTracker::Pyramid pa = tracker.MakePyramid(imA, 6), pb = tracker.MakePyramid(imB, 6);
  std::vector<Point2f> seeds = corners;
  std::vector<int32_t> lv(corners.size(), 3);
  GpuTracker::FBResult fb = tracker.TrackFeaturesFB(pa, pb, corners, seeds, lv);
  int mismatches = 0;
  for (size_t i = 0; i < corners.size() && i < 24; ++i) {
    Point2f to_pt = corners[i];
    auto p1 = tracker.GetPatches(pa, corners[i], 3);
    auto s1 = tracker.TrackFeature(pb, p1, 0.001f, 10, &to_pt);
    auto p2 = tracker.GetPatches(pb, to_pt, 3);
    Point2f back_pt = corners[i];
    auto s2 = tracker.TrackFeature(pa, p2, 0.001f, 10, &back_pt);
    bool ok = !(s1 || s2);
    if (ok) {
      float dx = corners[i].x - back_pt.x, dy = corners[i].y - back_pt.y;
      if (std::sqrt((double)dx * dx + (double)dy * dy) > 0.3) ok = false;  // matcher.cpp:201
    }
    if (ok != (bool)fb.accepted[i] || to_pt.x != fb.to_pt[i].x || to_pt.y != fb.to_pt[i].y || s1 != fb.status_fwd[i]) ++mismatches;
    if (p1.size() != 3 || p1[0].data.size() != 169 || p1[0].size.width != 13) ++mismatches;
  }
  std::fprintf(out, "duck_mismatches %d\n", mismatches);
  for (size_t i = 0; i < corners.size(); ++i)
    std::fprintf(out, "fb %zu %a %a %d %d %d\n", i, fb.to_pt[i].x, fb.to_pt[i].y, fb.status_fwd[i], fb.status_bwd[i], (int)fb.accepted[i]);

  // ---- 2. the Matcher mirror: frame A becomes a keyframe and seeds the features, frame B tracks them
  typedef sfe::MatcherT<Frame, LocalMap, TrackedPoint, Vec2> Matcher;
  Matcher matcher(ctx, [&](const sfe::ImageView&) { return corners; });
  LocalMap map;
  Frame f0{0}, f1{1};
  matcher.Track(imA, &f0, 0, &map, nullptr);
  std::fprintf(out, "after_A features %zu views %zu keyframe %d obs %zu\n", matcher.live_features(), matcher.views(), (int)f0.is_keyframe_, g_obs.size());
  size_t nobs_a = g_obs.size();
  matcher.Track(imB, &f1, 0, &map, [] { return false; });
  std::fprintf(out, "after_B features %zu views %zu keyframe %d obs %zu\n", matcher.live_features(), matcher.views(), (int)f1.is_keyframe_, g_obs.size() - nobs_a);
  for (size_t i = nobs_a; i < g_obs.size(); ++i) std::fprintf(out, "obs %d %d %a %a\n", g_obs[i].frame, g_obs[i].point, g_obs[i].x, g_obs[i].y);
  std::fclose(out);
  return mismatches == 0 ? 0 : 1;
}
