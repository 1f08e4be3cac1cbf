// matcher_gpu.hpp -- host orchestration of the front-end (the reference's Matcher, matcher.h:13-30 /
// matcher.cpp:301-405) re-cut for a batched device tracker.  Same call surface:
//     bool Track(const cv::Mat& img, Frame* frame, int camera, LocalMap* map, std::function<bool()> update_frames)
// and the same observable behaviour (which observations are added to which frame, which features
// are created, when a keyframe is taken), but FindMatches (matcher.cpp:208-271) -- a per-feature loop
// of forward/backward tracks -- is executed as a few batched sfe_track_fb calls:
//
//   reference, per feature f (ordered by point id, matcher.cpp:45-51), per previous view v of f:
//       levels = uncertainty > 100 ? 6 : 3; seed = projection if uncertainty < 100 else from_pt;
//       OOB test; TrackFeature(levels); on failure and levels != 6 retry with 6 levels starting from
//       the failed attempt's to_pt (matcher.cpp:247-248); first success wins (matcher.cpp:268).
//   here: attempt k of every still-unmatched feature is gathered per source view into one batch
//       (<= 4 views are alive, matcher.cpp:397), tracked in one launch, and the accept/retry/next-view
//       decision is applied on the host in point-id order.  A feature's attempts stay in the
//       reference's order, and features never influence each other inside FindMatches, so the result
//       is the same set of (feature, to_pt) matches.  Views are iterated in insertion order (the
//       reference iterates a std::map keyed by View* pointer value, i.e. in allocation-address order,
//       which is not reproducible).
//
// The map-side types are template parameters so that the header works with the reference's own
// localmap.h (Frame, LocalMap, TrackedPoint; Eigen vectors) and with light test doubles alike.
// Requirements on the types (exactly what matcher.cpp uses, SURVEY.md 8b "side effects"):
//     TrackedPoint:  double uncertainty(); int id(); bool feature_usable(); Vec3 location()
//     Frame:         bool Project(Vec3, Vec2*); void AddObservation(Vec2, TrackedPoint*);
//                    Vec3 Unproject(Vec2, double); camera()->PixelToPlane(Vec2); int id(); bool is_keyframe_
//     LocalMap:      TrackedPoint* AddPoint(int id, Vec3)
//     Traits::Vec2 constructible from (double,double) and indexable with (int).
#pragma once

#include <algorithm>
#include <cstdio>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <vector>

#include "gpu_tracker.hpp"

namespace sfe {

// Corner seeding (matcher.cpp:313 + :125-130).  By default the GPU implementation of
// cvtColor(RGB2GRAY) + goodFeaturesToTrack(grey, 120, 0.01, 20) is used (sfe_good_features, bit-identical
// to OpenCV's corner lists); a caller may inject its own detector, which receives the BGR frame.
typedef std::function<std::vector<Point2f>(const ImageView&)> CornerDetector;

inline std::vector<Point2f> GpuGoodFeatures(const Context& ctx, const ImageView& bgr, int max_corners = 120,
                                            double quality = 0.01, double min_distance = 20.0) {
  std::vector<Point2f> out((size_t)max_corners);
  int32_t n = 0;
  ctx.check(sfe_good_features(ctx.get(), bgr.data, bgr.cols, bgr.rows, bgr.step, bgr.step * (size_t)bgr.rows, 1, max_corners,
                              quality, min_distance, reinterpret_cast<float*>(out.data()), &n, nullptr));
  out.resize((size_t)n);
  return out;
}

template <class Frame, class LocalMap, class TrackedPoint, class Vec2>
class MatcherT {
 public:
  static constexpr int kWindowSize = 13;   // matcher.cpp:27
  static constexpr int kPyramidDepth = 6;  // matcher.cpp:317
  static constexpr int kMinMatches = 40;   // matcher.cpp:338,353
  static constexpr int kMaxViews = 4;      // matcher.cpp:397
  static constexpr int kGrid = 30;         // matcher.cpp:132

  explicit MatcherT(std::shared_ptr<Context> ctx = nullptr, CornerDetector detector = nullptr)
      : tracker_(Size(kWindowSize, kWindowSize), ctx), detector_(detector), next_fid_(0) {}

  struct View {
    Frame* frame;
    GpuTracker::Pyramid pyramid;
    int cols, rows;
  };
  struct Feature {
    TrackedPoint* point;
    std::vector<std::pair<View*, Point2f>> matches;  // insertion order (see header comment)
  };

  size_t live_features() const { return features_.size(); }
  size_t views() const { return views_.size(); }
  const std::map<Feature*, Point2f>& last_matches() const { return last_matches_; }

  // matcher.cpp:301-405
  template <class Mat>
  bool Track(const Mat& img, Frame* frame, int camera, LocalMap* map, std::function<bool()> update_frames) {
    if (img.cols == 0 || img.rows == 0 || !map || !frame) throw Error("Matcher::Track: bad arguments");  // :306-309
    std::unique_ptr<View> view(new View);
    view->frame = frame;
    view->pyramid = tracker_.MakePyramid(img, kPyramidDepth);  // :317
    view->cols = img.cols;
    view->rows = img.rows;

    // :327-330 remove features whose point became unusable
    for (auto it = features_.begin(); it != features_.end();) {
      if (!(*it)->point->feature_usable()) it = features_.erase(it);
      else ++it;
    }

    std::map<Feature*, Point2f> matches;
    FindMatches(*view, &matches);  // :335
    const int before = (int)matches.size();
    if (matches.size() < (size_t)kMinMatches && update_frames && update_frames()) FindMatches(*view, &matches);  // :338-345
    std::printf("Started with %d, grew to %d after additional matching\n", before, (int)matches.size());  // :346
    last_matches_ = matches;
    if (matches.size() >= (size_t)kMinMatches) return true;  // :353 (the reference leaks the View here)

    // new keyframe (:357-364)
    frame->is_keyframe_ = true;
    View* v = view.get();
    for (auto& m : matches) m.first->matches.emplace_back(v, m.second);
    views_.push_back(std::move(view));

    // :367-394 possibly add new features
    std::vector<Point2f> added;
    AddNewFeatures(img, matches, &added);
    std::printf("Adding new keyframe for camera %d on frame %d (added %d)\n", camera, frame->id(), (int)added.size());
    for (const Point2f& pt : added) {
      Vec2 frame_point((double)pt.x, (double)pt.y);
      std::unique_ptr<Feature> f(new Feature);
      auto plane_pt = v->frame->camera()->PixelToPlane(frame_point);
      auto location = v->frame->Unproject(plane_pt, 2000);  // :380 initial depth guess
      f->point = map->AddPoint(next_fid_++, location);
      v->frame->AddObservation(frame_point, f->point);
      f->matches.emplace_back(v, pt);
      features_.insert(std::move(f));
    }

    // :397-402 drop the oldest view and the matches that refer to it
    if (views_.size() > (size_t)kMaxViews) {
      View* old = views_.front().get();
      for (auto& f : features_) {
        auto& ms = f->matches;
        ms.erase(std::remove_if(ms.begin(), ms.end(), [old](const std::pair<View*, Point2f>& m) { return m.first == old; }), ms.end());
      }
      views_.pop_front();
    }
    return true;
  }

 private:
  struct FeatureCmp {  // matcher.cpp:45-49
    bool operator()(const std::unique_ptr<Feature>& a, const std::unique_ptr<Feature>& b) const { return a->point->id() < b->point->id(); }
  };

  // One pending attempt of the reference's FindMatches loop for one feature.
  struct Cursor {
    Feature* f;
    size_t view_idx;  // index into f->matches
    int levels;       // levels of the current attempt
    bool retry;       // this attempt is the 6-level retry of matcher.cpp:248
    Point2f to_pt;    // seed of the current attempt
    bool done;
  };

  // Positions the cursor on the next attempt that passes the reference's pre-checks (:225-245).
  bool Arm(Cursor* c, const View& view) const {
    while (c->view_idx < c->f->matches.size()) {
      const Point2f from_pt = c->f->matches[c->view_idx].second;
      Point2f to_pt = from_pt;
      int levels = c->f->point->uncertainty() > 100 ? 6 : 3;  // :227-229
      if (c->f->point->uncertainty() < 100) {                 // :234-239
        Vec2 p(0.0, 0.0);
        if (view.frame->Project(c->f->point->location(), &p)) { to_pt.x = (float)p(0); to_pt.y = (float)p(1); }
      }
      if (to_pt.x < 0 || to_pt.y < 0 || to_pt.x >= view.cols || to_pt.y > view.rows) {  // :243 (sic: > on y)
        ++c->view_idx;
        continue;
      }
      c->levels = levels;
      c->retry = false;
      c->to_pt = to_pt;
      return true;
    }
    c->done = true;
    return false;
  }

  // matcher.cpp:208-271, batched
  void FindMatches(const View& view, std::map<Feature*, Point2f>* matches) {
    std::vector<Cursor> cur;
    std::set<Feature*> fresh;  // matched in this call
    for (auto& f : features_) {  // point-id order
      if (matches->count(f.get())) continue;  // :219
      Cursor c{f.get(), 0, 3, false, Point2f(), false};
      if (Arm(&c, view)) cur.push_back(c);
    }
    while (true) {
      // gather the pending attempt of every unfinished feature, grouped by source view
      std::map<View*, std::vector<size_t>> by_view;
      for (size_t i = 0; i < cur.size(); ++i)
        if (!cur[i].done) by_view[cur[i].f->matches[cur[i].view_idx].first].push_back(i);
      if (by_view.empty()) break;
      for (auto& g : by_view) {
        std::vector<Point2f> from_pt, seed;
        std::vector<int32_t> levels;
        for (size_t i : g.second) {
          from_pt.push_back(cur[i].f->matches[cur[i].view_idx].second);
          seed.push_back(cur[i].to_pt);
          levels.push_back(cur[i].levels);
        }
        GpuTracker::FBResult r = tracker_.TrackFeaturesFB(g.first->pyramid, view.pyramid, from_pt, seed, levels);  // :175-201
        for (size_t k = 0; k < g.second.size(); ++k) {
          Cursor& c = cur[g.second[k]];
          c.to_pt = r.to_pt[k];  // to_pt carries the forward result even when the attempt is rejected (:247-248)
          if (r.accepted[k]) {
            (*matches)[c.f] = c.to_pt;  // :253
            fresh.insert(c.f);
            c.done = true;  // :268 first success wins
          } else if (c.levels != 6 && !c.retry) {
            c.levels = 6;  // :248 retry with 6 levels from the failed attempt's to_pt
            c.retry = true;
          } else {
            ++c.view_idx;  // :249 continue with the next previous view
            Arm(&c, view);
          }
        }
      }
    }
    // the observations reach the map in point-id order, as in the reference's sequential loop (:256-257)
    for (auto& f : features_)
      if (fresh.count(f.get())) {
        const Point2f& p = (*matches)[f.get()];
        view.frame->AddObservation(Vec2((double)p.x, (double)p.y), f->point);
      }
  }

  // matcher.cpp:123-169: keep detected corners that fall into grid cells not occupied (3x3-dilated) by
  // an existing match
  template <class Mat>
  void AddNewFeatures(const Mat& img, const std::map<Feature*, Point2f>& matches, std::vector<Point2f>* result) const {
    ImageView iv{(const uint8_t*)img.data, img.cols, img.rows, (size_t)img.step};
    std::vector<Point2f> corners = detector_ ? detector_(iv) : GpuGoodFeatures(*tracker_.context(), iv);  // :125-130
    int grid[kGrid + 2][kGrid + 2] = {};
    for (const auto& m : matches) {
      const Point2f& pt = m.second;
      int gx = (int)((pt.x / img.cols) * kGrid + 1), gy = (int)((pt.y / img.rows) * kGrid + 1);
      if (gx <= 0 || gy <= 0 || gx >= kGrid + 2 || gy >= kGrid + 2) throw Error("match outside the occupancy grid");  // CHECKs :138-141
      for (int dx = -1; dx <= 1; ++dx)  // (the reference writes one past the array when gx == 31; guarded here)
        for (int dy = -1; dy <= 1; ++dy)
          if (gx + dx <= kGrid + 1 && gy + dy <= kGrid + 1) grid[gx + dx][gy + dy] = 1;
    }
    int added = 0;
    for (const Point2f& c : corners) {
      int gx = (int)((c.x / img.cols) * kGrid + 1), gy = (int)((c.y / img.rows) * kGrid + 1);
      if (gx <= 0 || gy <= 0 || gx >= kGrid + 2 || gy >= kGrid + 2) throw Error("corner outside the occupancy grid");
      if (grid[gx][gy]) continue;
      result->push_back(c);
      ++added;
    }
    std::printf("Added %d new features\n", added);  // :168
  }

  GpuTracker tracker_;
  CornerDetector detector_;
  std::set<std::unique_ptr<Feature>, FeatureCmp> features_;
  std::deque<std::unique_ptr<View>> views_;
  std::map<Feature*, Point2f> last_matches_;
  int next_fid_;
};

}  // namespace sfe
