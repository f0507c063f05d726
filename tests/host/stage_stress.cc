// CPU stress test of csrc/host_stage.cu (pageable inputs -> pinned bounce buffer through worker
// threads) against a shim of the CUDA runtime calls it uses (tests/host/cuda_shim): random job
// sizes, two copies per call like points + colours, contents compared after every copy.  Built with
// -fsanitize=thread by tests/test_host_stage.py.
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <random>
#include <thread>
#include <vector>

#include "host_stage.cuh"

int main(int argc, char** argv) {
  const int iterations = argc > 1 ? atoi(argv[1]) : 400;
  cg::HostStager* st = nullptr;
  std::mt19937_64 rng(1);
  std::vector<char> src(48 << 20), dst(48 << 20);
  for (size_t i = 0; i < src.size(); ++i) src[i] = static_cast<char>((i * 2654435761u) >> 13);
  for (int it = 0; it < iterations; ++it) {
    if (cg::stage_begin(st, nullptr) != 0) return 3;
    for (int part = 0; part < 2; ++part) {
      // copies of 16 MB and more go through the pool, smaller ones take the direct path
      const size_t n = (it % 3 != 2) ? (16u << 20) + rng() % (24u << 20) : (rng() % (6u << 20)) + 1;
      const size_t off = rng() % (src.size() - n);
      memset(dst.data() + off, 0, n);
      if (cg::stage_to_device(&st, dst.data() + off, src.data() + off, n, nullptr, 3) != 0) return 1;
      if (memcmp(dst.data() + off, src.data() + off, n) != 0) {
        printf("MISMATCH at job %d (%zu bytes)\n", it, n);
        return 2;
      }
    }
    // now and then let the workers fall asleep so that a late wake-up meets the next job
    if (it % 100 == 0) std::this_thread::sleep_for(std::chrono::milliseconds(3));
  }
  cg::destroy_stager(st);
  puts("stage_stress ok");
  return 0;
}
