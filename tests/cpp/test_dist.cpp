// The multi-GPU entries of the C ABI from plain C++, the way the reference's single-process C++ caller (main.cpp:490)
// would use them: one host thread and one sfe_ctx per GPU, an NCCL communicator created inside the library
// (sfe_dist_unique_id + sfe_dist_init), then BASELINE config 5's exchange -- train set broadcast, query rows sharded,
// top-2 rows all-gathered (sfe_match_hamming256_sharded) -- and a sharded replay block gathered with
// sfe_allgather_rows_dev is left to the Python tests.  Every rank's gathered result must equal a single-GPU
// sfe_match_hamming256 of the whole problem.  Usage: test_dist <ranks> <nq> <nt>; driven by tests/test_gpu_dist.py.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/slamfe.h"

static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }

int main(int argc, char** argv) {
  const int world = argc > 1 ? atoi(argv[1]) : 2;
  const int64_t nq = argc > 2 ? atoll(argv[2]) : 3001;   // not divisible by 2: the unequal-block gather is exercised
  const int nt = argc > 3 ? atoi(argv[3]) : 4097;
  std::vector<uint32_t> q(8 * (size_t)nq), t(8 * (size_t)nt);
  uint32_t s = 12345u;
  for (auto& v : t) v = lcg(s);
  for (int64_t i = 0; i < nq; ++i)
    for (int k = 0; k < 8; ++k) {
      // queries are noisy copies of train rows (so that best distances are small and ties occur), every 5th an exact copy
      uint32_t base = t[8 * (size_t)(lcg(s) % (uint32_t)nt) + k];
      q[8 * (size_t)i + k] = (i % 5 == 0) ? base : base ^ (lcg(s) & lcg(s) & lcg(s));
    }
  // reference: the whole problem on GPU 0
  std::vector<int32_t> ridx(2 * (size_t)nq), rdist(2 * (size_t)nq);
  std::vector<uint8_t> rpass((size_t)nq);
  {
    sfe_ctx* c = nullptr;
    if (sfe_create(0, &c)) { printf("sfe_create failed\n"); return 2; }
    if (sfe_match_hamming256(c, q.data(), (int)nq, t.data(), nt, 1, 4, 5, 80, ridx.data(), rdist.data(), rpass.data())) {
      printf("single-GPU match failed: %s\n", sfe_last_error(c));
      return 2;
    }
    sfe_destroy(c);
  }
  uint8_t id[128];
  if (sfe_dist_unique_id(id)) { printf("sfe_dist_unique_id failed (NCCL not loadable)\n"); return 3; }
  std::vector<int> bad(world, -1);
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r)
    th.emplace_back([&, r]() {
      sfe_ctx* c = nullptr;
      if (sfe_create(r, &c)) return;
      if (sfe_dist_init(c, id, r, world)) { printf("rank %d: %s\n", r, sfe_last_error(c)); return; }
      int64_t lo = 0, hi = 0;
      sfe_shard_range(nq, r, world, &lo, &hi);
      std::vector<int32_t> idx(2 * (size_t)nq, -7), dist(2 * (size_t)nq, -7);
      std::vector<uint8_t> pass((size_t)nq, 9);
      // only the root passes the train set; the others receive it through the broadcast
      int rc = sfe_match_hamming256_sharded(c, q.data() + 8 * (size_t)lo, nq, r == 1 % world ? t.data() : nullptr, nt, 1 % world,
                                            4, 5, 80, idx.data(), dist.data(), pass.data());
      if (rc) { printf("rank %d: %s\n", r, sfe_last_error(c)); return; }
      int b = 0;
      for (size_t i = 0; i < 2 * (size_t)nq; ++i) b += idx[i] != ridx[i] || dist[i] != rdist[i];
      for (size_t i = 0; i < (size_t)nq; ++i) b += pass[i] != rpass[i];
      bad[r] = b;
      sfe_dist_shutdown(c);
      sfe_destroy(c);
    });
  for (auto& x : th) x.join();
  int total = 0;
  for (int r = 0; r < world; ++r) {
    printf("rank %d mismatches %d\n", r, bad[r]);
    total += bad[r] != 0;
  }
  printf("sharded_mismatching_ranks %d\n", total);
  return total ? 1 : 0;
}
