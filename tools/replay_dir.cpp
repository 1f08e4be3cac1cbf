// replay_dir -- replays a recorded directory ("<dir>/%08d.png", the reference's --save/--load format, main.cpp:373-398,
// :446-448) through the GPU front-end: same-camera frame pairs (id, id+2), corners seeded on the first frame of each
// pair (matcher.cpp:123-130 parameters), forward/backward tracking with the reference's constants (matcher.cpp:176,
// :182,:201, 3 levels of a 6-level pyramid).  Prints one line per pair: "pair <id> corners <n> accepted <m>" and,
// with a third argument, writes the tracked positions as text.
// build: g++ -std=c++14 -O2 -o replay_dir tools/replay_dir.cpp -Lslam-robot_b200/csrc -lslamfe -Wl,-rpath,$PWD/slam-robot_b200/csrc
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../include/slamfe.h"
#include "../slam-robot_b200/host/replay_source.hpp"

#define CHECK(call)                                                                  \
  do {                                                                               \
    if ((call) != SFE_SUCCESS) {                                                     \
      fprintf(stderr, "%s failed: %s\n", #call, ctx ? sfe_last_error(ctx) : "?");    \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: replay_dir <dir> <max_pairs> [tracks.txt]\n");
    return 2;
  }
  sfe_ctx* ctx = nullptr;
  CHECK(sfe_create(0, &ctx));
  sfe::ImageSourceFiles src(argv[1]);
  // the whole sequence in one buffer: pair p = (frame p, frame p + 2), every frame uploaded and its pyramid built once
  std::vector<uint8_t> frames;
  int w = 0, h = 0;
  const int nframes = src.LoadSequence(0, atoi(argv[2]) + 2, &frames, &w, &h);
  const int npairs = nframes - 2;
  if (npairs <= 0) {
    fprintf(stderr, "no frame pairs under %s\n", argv[1]);
    return 1;
  }
  const int max_corners = 120;
  std::vector<float> corners((size_t)npairs * max_corners * 2, 0.f);
  std::vector<int32_t> ncorners(npairs, 0);
  CHECK(sfe_good_features(ctx, frames.data(), w, h, (size_t)3 * w, (size_t)3 * w * h, npairs, max_corners, 0.01, 20.0, corners.data(),
                          ncorners.data(), nullptr));
  // the replay wants the same feature count per pair: pad short lists with the last corner (tracked twice, ignored below)
  const int n = npairs * max_corners;
  std::vector<float> from_xy(corners), to_xy, back(2 * (size_t)n);
  for (int p = 0; p < npairs; ++p)
    for (int i = ncorners[p]; i < max_corners && ncorners[p] > 0; ++i) {
      from_xy[2 * ((size_t)p * max_corners + i)] = corners[2 * ((size_t)p * max_corners + ncorners[p] - 1)];
      from_xy[2 * ((size_t)p * max_corners + i) + 1] = corners[2 * ((size_t)p * max_corners + ncorners[p] - 1) + 1];
    }
  to_xy = from_xy;  // the seed is from_pt (uncertain map points, matcher.cpp:225)
  std::vector<int32_t> s1(n), s2(n);
  std::vector<uint8_t> acc(n);
  CHECK(sfe_replay_sequence(ctx, w, h, 6, nframes, 2, frames.data(), (size_t)3 * w, (size_t)3 * w * h, max_corners, from_xy.data(),
                            to_xy.data(), nullptr, 3, 0.001f, 10, 0.3, back.data(), s1.data(), s2.data(), acc.data(), nullptr, 0));
  FILE* out = argc > 3 ? fopen(argv[3], "w") : nullptr;
  for (int p = 0; p < npairs; ++p) {
    int m = 0;
    for (int i = 0; i < ncorners[p]; ++i) {
      const size_t k = (size_t)p * max_corners + i;
      m += acc[k];
      if (out) fprintf(out, "%d %d %a %a %a %a %d\n", p, i, from_xy[2 * k], from_xy[2 * k + 1], to_xy[2 * k], to_xy[2 * k + 1], (int)acc[k]);
    }
    printf("pair %d corners %d accepted %d\n", p, ncorners[p], m);
  }
  if (out) fclose(out);
  sfe_destroy(ctx);
  return 0;
}
