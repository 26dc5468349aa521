// h2d_probe.cu -- how fast can a 640x480 float frame reach the GPU from pinned host memory?
// (a) cudaMemcpyAsync back to back, (b) over two streams, (c) one large copy,
// (d) a kernel pulling from mapped pinned memory (zero-copy).  GPU box only.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if(e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while(0)

__global__ void pull_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, size_t n)
{
  for(size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

int main()
{
  const size_t frame = 640 * 480 * sizeof(float);
  const int N = 200, SLOTS = 4;
  float *h[SLOTS], *d[SLOTS], *hd[SLOTS];
  for(int i = 0; i < SLOTS; ++i)
  {
    CK(cudaHostAlloc(&h[i], frame, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(&hd[i], h[i], 0));
    CK(cudaMalloc(&d[i], frame));
    for(size_t k = 0; k < frame / 4; ++k) h[i][k] = (float)k;
  }
  cudaStream_t s0, s1;
  CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms;
  for(int rep = 0; rep < 2; ++rep)
  {
    CK(cudaEventRecord(e0, s0));
    for(int k = 0; k < N; ++k) CK(cudaMemcpyAsync(d[k % SLOTS], h[k % SLOTS], frame, cudaMemcpyHostToDevice, s0));
    CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    if(rep) printf("memcpyAsync 1.2 MB x%d, one stream : %.1f us per frame (%.1f GB/s)\n", N, ms * 1e3 / N, frame * N / ms / 1e6);
  }
  for(size_t bytes : {(size_t)307200, (size_t)614400})
  {
    CK(cudaEventRecord(e0, s0));
    for(int k = 0; k < N; ++k) CK(cudaMemcpyAsync(d[k % SLOTS], h[k % SLOTS], bytes, cudaMemcpyHostToDevice, s0));
    CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("memcpyAsync %zu B x%d, one stream : %.1f us per copy (%.1f GB/s)\n", bytes, N, ms * 1e3 / N, bytes * N / ms / 1e6);
  }
  {
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, s0));
    CK(cudaStreamWaitEvent(s1, e0, 0));
    for(int k = 0; k < N; ++k) CK(cudaMemcpyAsync(d[k % SLOTS], h[k % SLOTS], frame, cudaMemcpyHostToDevice, (k & 1) ? s1 : s0));
    cudaEvent_t j; CK(cudaEventCreate(&j)); CK(cudaEventRecord(j, s1)); CK(cudaStreamWaitEvent(s0, j, 0));
    CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("memcpyAsync 1.2 MB x%d, two streams: %.1f us per frame (%.1f GB/s)\n", N, ms * 1e3 / N, frame * N / ms / 1e6);
  }
  {
    const size_t big = 256u << 20;
    float *hb, *db;
    CK(cudaHostAlloc(&hb, big, cudaHostAllocDefault)); CK(cudaMalloc(&db, big));
    for(int rep = 0; rep < 2; ++rep)
    {
      CK(cudaEventRecord(e0, s0));
      CK(cudaMemcpyAsync(db, hb, big, cudaMemcpyHostToDevice, s0));
      CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("memcpyAsync 256 MB                  : %.2f ms (%.1f GB/s)\n", ms, big / ms / 1e6);
    CK(cudaEventRecord(e0, s0));
    CK(cudaMemcpyAsync(hb, db, big, cudaMemcpyDeviceToHost, s0));
    CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("memcpyAsync 256 MB D2H              : %.2f ms (%.1f GB/s)\n", ms, big / ms / 1e6);
  }
  for(int blocks : {37, 74, 148, 296, 592})
    for(int threads : {256, 1024})
    {
      for(int rep = 0; rep < 2; ++rep)
      {
        CK(cudaEventRecord(e0, s0));
        for(int k = 0; k < N; ++k)
          pull_kernel<<<blocks, threads, 0, s0>>>((const float4*)hd[k % SLOTS], (float4*)d[k % SLOTS], frame / 16);
        CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      printf("zero-copy pull kernel %4d x %4d     : %.1f us per frame (%.1f GB/s)\n", blocks, threads, ms * 1e3 / N, frame * N / ms / 1e6);
    }
  CK(cudaGetLastError());
  return 0;
}
