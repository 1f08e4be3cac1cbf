// Reads "<dir>/%08d.png" frames through the ImageSourceFiles mirror (slam-robot_b200/host/replay_source.hpp) and
// writes them as raw BGR ("<w> <h>\n" header + bytes) so that the Python test can compare them with cv2.imread.
// usage: test_replay_source <dir> <first_id> <count> <out_prefix>
#include <stdio.h>
#include <stdlib.h>

#include "../../slam-robot_b200/host/replay_source.hpp"

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  sfe::ImageSourceFiles src(argv[1]);
  const int first = atoi(argv[2]), count = atoi(argv[3]);
  for (int i = 0; i < count; ++i) {
    sfe::BgrImage img;
    const bool ok = src.GetObservation(i & 1, first + i, &img);
    char name[512];
    snprintf(name, sizeof(name), "%s%d.bin", argv[4], first + i);
    FILE* f = fopen(name, "wb");
    if (!f) return 3;
    if (ok) {
      fprintf(f, "%d %d\n", img.cols, img.rows);
      fwrite(img.data.data(), 1, img.data.size(), f);
    } else {
      fprintf(f, "0 0\n");
    }
    fclose(f);
  }
  // pairs (id, id+2): the alternating-camera layout of main.cpp:503-519
  std::vector<uint8_t> a, b;
  int w = 0, h = 0;
  const int n = src.LoadPairs(first, count, &a, &b, &w, &h);
  printf("pairs %d %d %d %zu %zu\n", n, w, h, a.size(), b.size());
  // the same frames as one sequence buffer (sfe_replay_sequence): frame i of it is the from-frame of pair i
  std::vector<uint8_t> seq;
  int sw = 0, sh = 0;
  const int nf = src.LoadSequence(first, count + 2, &seq, &sw, &sh);
  const size_t fb = (size_t)3 * sw * sh;
  bool same = nf == n + 2 && sw == w && sh == h && seq.size() == fb * nf;
  for (int i = 0; same && i < n; ++i)
    same = memcmp(&seq[i * fb], &a[i * fb], fb) == 0 && memcmp(&seq[(i + 2) * fb], &b[i * fb], fb) == 0;
  printf("sequence %d %d\n", nf, same ? 1 : 0);
  return 0;
}
