// copy_probe.cpp -- host cost of staging a 640x480 float frame into pinned memory (GPU box).
// Sources are 200 separately allocated pageable frames (as bench.py's e2e leg holds them).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "../rpg_open_remode_b200/csrc/host_copy.h"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void copy_nt(void *dst, const void *src, size_t bytes)
{
  const __m256i *s = (const __m256i*)src; __m256i *d = (__m256i*)dst;
  const size_t n = bytes / 32;
  for(size_t i = 0; i < n; i += 4)
  {
    const __m256i a = _mm256_loadu_si256(s + i), b = _mm256_loadu_si256(s + i + 1);
    const __m256i c = _mm256_loadu_si256(s + i + 2), e = _mm256_loadu_si256(s + i + 3);
    _mm256_stream_si256(d + i, a); _mm256_stream_si256(d + i + 1, b);
    _mm256_stream_si256(d + i + 2, c); _mm256_stream_si256(d + i + 3, e);
  }
  _mm_sfence();
}

int main()
{
  const size_t frame = 640 * 480 * 4;
  const int N = 200;
  printf("hardware threads: %u\n", std::thread::hardware_concurrency());
  std::vector<float*> src(N);
  for(int i = 0; i < N; ++i) { src[i] = (float*)malloc(frame); for(size_t k = 0; k < frame / 4; ++k) src[i][k] = (float)(k + i); }
  void *pin[3];
  for(int i = 0; i < 3; ++i) if(cudaHostAlloc(&pin[i], frame, cudaHostAllocDefault) != cudaSuccess) { printf("cudaHostAlloc failed\n"); return 1; }
  for(int rep = 0; rep < 2; ++rep)
  {
    double t0 = now();
    for(int i = 0; i < N; ++i) memcpy(pin[i % 3], src[i], frame);
    if(rep) printf("memcpy, 1 thread            : %.1f us per frame\n", (now() - t0) / N * 1e6);
    t0 = now();
    for(int i = 0; i < N; ++i) copy_nt(pin[i % 3], src[i], frame);
    if(rep) printf("AVX2 non-temporal, 1 thread : %.1f us per frame\n", (now() - t0) / N * 1e6);
  }
  for(int helpers : {1, 3, 7, 15})
  {
    rmdb::ParallelCopier pc(helpers);
    for(int rep = 0; rep < 2; ++rep)
    {
      const double t0 = now();
      for(int i = 0; i < N; ++i) pc.copy(pin[i % 3], src[i], frame);
      if(rep) printf("ParallelCopier helpers %2d   : %.1f us per frame (parallel chosen: %d)\n", helpers, (now() - t0) / N * 1e6, (int)pc.parallel_chosen());
    }
  }
  // the same with a 40 us pause between frames (a consumer that is not back to back)
  for(int helpers : {3, 7})
  {
    rmdb::ParallelCopier pc(helpers);
    double acc = 0;
    for(int i = 0; i < 2 * N; ++i)
    {
      const double t0 = now();
      pc.copy(pin[i % 3], src[i % N], frame);
      const double t1 = now();
      if(i >= N) acc += t1 - t0;
      while(now() - t1 < 40e-6) {}
    }
    printf("ParallelCopier helpers %2d, 40 us gaps: %.1f us per frame (parallel chosen: %d)\n", helpers, acc / N * 1e6, (int)pc.parallel_chosen());
  }
  return 0;
}
