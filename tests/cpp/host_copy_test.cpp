// host_copy_test.cpp -- the frame-ingest copier of the streaming path
// (rpg_open_remode_b200/csrc/host_copy.h) is plain host code: exercised here
// without a GPU.  Every copy must be exact whatever the size (below / above the
// parallel threshold, not a multiple of the chunk), the helper count, the
// calibration phase it happens in, and the pause between frames (helpers
// spinning vs asleep).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../rpg_open_remode_b200/csrc/host_copy.h"

static int failures = 0;
#define CHECK(cond) do { if(!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } } while(0)

static void fill(std::vector<unsigned char> &v, unsigned seed)
{
  unsigned x = seed * 2654435761u + 12345u;
  for(size_t i = 0; i < v.size(); ++i)
  {
    x = x * 1664525u + 1013904223u;
    v[i] = (unsigned char)(x >> 24);
  }
}

int main()
{
  const size_t sizes[] = {0, 1, 4095, 4096, 65536, (256u << 10) - 1, 256u << 10, (256u << 10) + 1,
                          640 * 480, 640 * 480 * 4, 752 * 480 * 4 + 3, 1920 * 1080 * 4};
  for(int helpers : {0, 1, 3, 7})
  {
    rmdb::ParallelCopier pc(helpers);
    unsigned seed = 1;
    for(int round = 0; round < 3; ++round)   // rounds 0..: calibration calls first, the chosen mode afterwards
      for(size_t n : sizes)
      {
        std::vector<unsigned char> src(n + 64), dst(n + 128, 0xA5);
        fill(src, seed++);
        pc.copy(dst.data() + 32, src.data() + 16, n);   // unaligned on purpose
        CHECK(std::memcmp(dst.data() + 32, src.data() + 16, n) == 0);
        bool guard_ok = true;
        for(size_t i = 0; i < 32; ++i) guard_ok &= (dst[i] == 0xA5) && (dst[32 + n + i] == 0xA5);
        CHECK(guard_ok);                                 // nothing written outside [dst, dst + n)
      }
    // a stream with pauses: helpers fall asleep (2 ms spin window) and must wake up for the next frame
    std::vector<unsigned char> src(640 * 480 * 4), dst(640 * 480 * 4);
    for(int k = 0; k < 6; ++k)
    {
      fill(src, 1000 + k);
      if(k & 1) std::this_thread::sleep_for(std::chrono::milliseconds(8));
      pc.copy(dst.data(), src.data(), src.size());
      CHECK(src == dst);
    }
    std::printf("helpers %d: parallel chosen %d\n", helpers, (int)pc.parallel_chosen());
  }
  // destruction with sleeping and with spinning helpers, and without any copy at all
  { rmdb::ParallelCopier idle(5); }
  { rmdb::ParallelCopier busy(5); std::vector<unsigned char> a(1 << 20), b(1 << 20); busy.copy(b.data(), a.data(), a.size()); }
  if(failures == 0) std::printf("ALL HOST COPY TESTS PASSED\n");
  return failures ? 1 : 0;
}
