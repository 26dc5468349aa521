// denoiser.cu -- weighted TV-L1 (Chambolle-Pock) smoothing of the depth map.
//
// Mathematics: src/depthmap_denoiser.cu:46-59 (weights), :62-118 (one
// primal-dual step), :124-141 (L = sqrt(8), tau = 0.02, sigma = 1/(L^2 tau),
// theta = 0.5), :226-229 (large sigma^2 = range^2 / 72).
//
// Organisation (ours).  The reference launches one kernel per iteration and
// updates p / u / u_head in place with only an intra-block barrier, so it is
// racy across its 16x16 tile seams; inside a tile it is
// a Jacobi sweep: all duals from iterate n, then all primals from the new
// duals.  Here that Jacobi sweep IS the definition (deterministic), and it is
// temporally blocked: one launch advances DENOISE_T = 8 iterations.
//
//   * state = six planar float images (u, u_head, p.x, p.y ping-ponged; g, mu
//     constant); a CTA owns a 48 x 24 output tile and stages the 64 x 40 block
//     around it (halo 8 = one pixel of dependency radius per iteration) with
//     SIX TMA tensor loads (cp.async.bulk.tensor.2d, zero fill outside the
//     image) completing on one mbarrier;
//   * the 8 iterations run in shared memory, two phases per iteration (dual,
//     primal) separated by __syncthreads; rows outside the still-valid region
//     are skipped (the valid block shrinks by one pixel per side per
//     iteration; values in the invalid ring are computed from garbage and
//     never reach a valid pixel);
//   * one warp-pass covers a whole 64-pixel row: a thread owns TWO horizontally
//     adjacent pixels, loaded with 64-bit LDS and advanced with Blackwell's
//     packed fma.rn.f32x2 / add.f32x2 / mul.f32x2 (the two lanes of a packed
//     op are independent IEEE operations, so the result does not depend on
//     the packing);
//   * launched with programmatic stream serialisation: the next launch's CTAs
//     set up their barrier while this one drains.
//
// Per pixel and iteration the solver moves 40 bytes algorithmically (SURVEY.md
// 8d); here a launch reads 24 B and writes 16 B per pixel for 8 iterations
// (x 1.8 for the halo), all of it from L2.
#include <cuda.h>

#include "denoiser.cuh"
#include "packed_f32x2.cuh"

namespace rmdb
{

namespace
{
constexpr int T_IT = DENOISE_T;
constexpr int TW = DENOISE_TILE_W, TH = DENOISE_TILE_H;
constexpr int EW = DENOISE_EXT_W, EH = DENOISE_EXT_H;
constexpr int PLANE = EW * EH;                 // floats per staged plane
constexpr int PAD = 64;                        // floats before the first / after the last plane
constexpr int NPLANES = 6;                     // u, u_head, p.x, p.y, g, mu
constexpr int NTHREADS = 256;
static_assert(EW == 64, "one warp pass of 2-pixel threads covers a row");
static_assert(TW + 2 * T_IT == EW && TH + 2 * T_IT == EH, "halo = iterations per launch");

__device__ __forceinline__ unsigned int smem_addr(const void *p)
{
  return (unsigned int)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ float2 lds2(const float *p)
{
  return *reinterpret_cast<const float2*>(p);
}
__device__ __forceinline__ void sts2(float *p, const float2 v)
{
  *reinterpret_cast<float2*>(p) = v;
}

} // namespace

// computeWeightsKernel (:46-59) + "u_ = mu; u_head_ = u_; p_.zero()" (:215-217)
template<bool FROM_SEEDS>
__global__ void __launch_bounds__(256) denoise_setup_kernel(const DenoiseSetupParams P)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if(x >= P.width || y >= P.height)
    return;
  float mu, sigma_sq, a, b;
  if(FROM_SEEDS)
  {
    const float4 s = P.seed[(size_t)y * P.seed_stride + x];
    mu = s.x; sigma_sq = s.y; a = s.z; b = s.w;
  }
  else
  {
    mu = P.mu[(size_t)y * P.mu_stride + x];
    sigma_sq = P.sigma_sq[(size_t)y * P.sigma_sq_stride + x];
    a = P.a[(size_t)y * P.a_stride + x];
    b = P.b[(size_t)y * P.b_stride + x];
  }
  const float E_pi = a / (a + b);
  const float g = fmaxf((E_pi * sigma_sq + (1.0f - E_pi) * P.large_sigma_sq) / P.large_sigma_sq, 1.0f);
  const size_t i = (size_t)y * P.stride + x;
  P.g[i] = g;
  P.noisy[i] = mu;
  P.u[i] = mu;
  P.u_head[i] = mu;
  P.p_x[i] = 0.0f;
  P.p_y[i] = 0.0f;
}

// updateTVL1PrimalDualKernel (:62-118), Jacobi form, P.n_it <= DENOISE_T iterations per launch.
__global__ void __launch_bounds__(NTHREADS, 3) denoise_block_kernel(const __grid_constant__ DenoiseBlockParams P)
{
  extern __shared__ unsigned char smem_raw[];
  const unsigned int smem_pad = (128u - (smem_addr(smem_raw) & 127u)) & 127u;
  float *const base = reinterpret_cast<float*>(smem_raw + smem_pad);
  // [pad][u][u_head][p.x][p.y][g][mu][pad]: neighbour reads one element / one row beyond a plane stay inside
  // the block (they only ever feed pixels of the invalid ring)
  float *const s_u = base + PAD;
  float *const s_uh = s_u + PLANE;
  float *const s_px = s_uh + PLANE;
  float *const s_py = s_px + PLANE;
  float *const s_g = s_py + PLANE;
  float *const s_mu = s_g + PLANE;
  unsigned long long *const mbar = reinterpret_cast<unsigned long long*>(s_mu + PLANE + PAD);

  const int lane = threadIdx.x, wid = threadIdx.y;
  const int tid = wid * 32 + lane;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int ex0 = x0 - T_IT, ey0 = y0 - T_IT;     // image coordinates of the staged block's origin

  if(tid == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // everything above overlaps the previous launch's tail; its results are read from here on
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if(tid == 0)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_addr(mbar)), "r"((unsigned int)(NPLANES * PLANE * sizeof(float))) : "memory");
    const CUtensorMap *maps[NPLANES] = {&P.in_u, &P.in_uh, &P.in_px, &P.in_py, &P.g, &P.mu};
#pragma unroll
    for(int k = 0; k < NPLANES; ++k)
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
          :: "r"(smem_addr(s_u + k * PLANE)), "l"(reinterpret_cast<unsigned long long>(maps[k])), "r"(ex0), "r"(ey0),
             "r"(smem_addr(mbar))
          : "memory");
  }
  __syncthreads();   // the barrier is initialised before anybody polls it
  {
    unsigned int ok = 0;
    for(unsigned int it = 0; !ok; ++it)
    {
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
          "selp.u32 %0, 1, 0, p;\n"
          "}\n" : "=r"(ok) : "r"(smem_addr(mbar)) : "memory");
      if(it > (1u << 22)) __trap();   // a TMA that never completes must fail the launch, not hang the GPU
    }
  }

  const int lx = 2 * lane;                  // this thread's pixel pair: columns lx, lx + 1 of the block
  const int gx = ex0 + lx;                  // image column of the pair's first pixel
  const f2 sigma2 = pack(P.sigma, P.sigma);
  const f2 theta2 = pack(P.theta, P.theta);
  const float tl = P.tau * P.lambda;
  // image-edge rules: the east neighbour of the last column and the south neighbour of the last row are the
  // pixel itself (:79-80 min(width-1, x+1)); the divergence drops the terms that would cross the edge (:88-99)
  const bool e0_own = (gx >= P.width - 1), e1_own = (gx + 1 >= P.width - 1);
  const bool w0_zero = (gx == 0), w1_zero = (gx + 1 == 0);
  const bool c0_zero = (gx != 0) && (gx >= P.width - 1), c1_zero = (gx + 1 != 0) && (gx + 1 >= P.width - 1);

  const int n_it = P.n_it;
  for(int t = 0; t < n_it; ++t)
  {
    // ---- dual ascent + projection (:70-83) on rows t .. EH-1-t
    for(int ly = t + wid; ly < EH - t; ly += NTHREADS / 32)
    {
      const int o = ly * EW + lx;
      const int gy = ey0 + ly;
      const float2 u = lds2(s_u + o), uh = lds2(s_uh + o);
      const float uh_next = s_uh[o + 2];
      float2 uh_s = lds2(s_uh + o + EW);
      if(gy >= P.height - 1) uh_s = uh;
      const float2 uh_e = make_float2(e0_own ? uh.x : uh.y, e1_own ? uh.y : uh_next);
      const f2 g2 = pack(lds2(s_g + o));
      const f2 u2 = pack(u);
      const f2 grad_x = f2_sub(pack(uh_e), u2), grad_y = f2_sub(pack(uh_s), u2);   // centre term is u, not u_head
      const f2 tx2 = f2_fma(f2_mul(g2, grad_x), sigma2, pack(lds2(s_px + o)));
      const f2 ty2 = f2_fma(f2_mul(g2, grad_y), sigma2, pack(lds2(s_py + o)));
      const float2 len_sq = unpack(f2_fma(tx2, tx2, f2_mul(ty2, ty2)));
      // p = p~ / max(1, |p~|)
      const f2 inv = pack(len_sq.x > 1.0f ? rsqrtf(len_sq.x) : 1.0f, len_sq.y > 1.0f ? rsqrtf(len_sq.y) : 1.0f);
      sts2(s_px + o, unpack(f2_mul(tx2, inv)));
      sts2(s_py + o, unpack(f2_mul(ty2, inv)));
    }
    __syncthreads();
    // ---- divergence, primal shrink towards the noisy depth, over-relaxation (:86-115) on rows t+1 .. EH-2-t
    for(int ly = t + 1 + wid; ly < EH - 1 - t; ly += NTHREADS / 32)
    {
      const int o = ly * EW + lx;
      const int gy = ey0 + ly;
      const float2 px = lds2(s_px + o), py = lds2(s_py + o);
      const float px_w = s_px[o - 1];
      float2 py_n = lds2(s_py + o - EW);
      float2 cy = py;
      if(gy == 0) py_n = make_float2(0.0f, 0.0f);
      else if(gy >= P.height - 1) cy = make_float2(0.0f, 0.0f);
      const float2 cx = make_float2(c0_zero ? 0.0f : px.x, c1_zero ? 0.0f : px.y);
      const float2 wx = make_float2(w0_zero ? 0.0f : px_w, w1_zero ? 0.0f : px.x);
      const f2 div = f2_sub(f2_add(f2_sub(pack(cx), pack(wx)), pack(cy)), pack(py_n));   // cp.x - wp.x + cp.y - np.y
      const float2 u = lds2(s_u + o), g = lds2(s_g + o), noisy = lds2(s_mu + o);
      const f2 u2 = pack(u);
      const f2 tg = pack(P.tau * g.x, P.tau * g.y);
      const float2 temp = unpack(f2_fma(tg, div, u2));
      float2 nu;
      {
        const float dx = temp.x - noisy.x, dy = temp.y - noisy.y;
        nu.x = dx > tl ? temp.x - tl : (dx < -tl ? temp.x + tl : noisy.x);
        nu.y = dy > tl ? temp.y - tl : (dy < -tl ? temp.y + tl : noisy.y);
      }
      const f2 nu2 = pack(nu);
      sts2(s_u + o, nu);
      sts2(s_uh + o, unpack(f2_fma(theta2, f2_sub(nu2, u2), nu2)));
    }
    __syncthreads();
  }

  // ---- the tile itself: rows T .. T+TH-1, columns T .. T+TW-1 of the block
  for(int k = tid; k < (TW / 2) * TH; k += NTHREADS)
  {
    const int ty = k / (TW / 2), tx = 2 * (k - ty * (TW / 2));
    const int x = x0 + tx, y = y0 + ty;
    if(y >= P.height || x >= P.width)
      continue;
    const int o = (T_IT + ty) * EW + T_IT + tx;
    const size_t i = (size_t)y * P.stride + x;
    if(x + 1 < P.width)
    {
      *reinterpret_cast<float2*>(P.out_u + i) = lds2(s_u + o);
      *reinterpret_cast<float2*>(P.out_uh + i) = lds2(s_uh + o);
      *reinterpret_cast<float2*>(P.out_px + i) = lds2(s_px + o);
      *reinterpret_cast<float2*>(P.out_py + i) = lds2(s_py + o);
    }
    else
    {
      P.out_u[i] = s_u[o]; P.out_uh[i] = s_uh[o]; P.out_px[i] = s_px[o]; P.out_py[i] = s_py[o];
    }
  }
}

static inline dim3 grid_for(int width, int height, dim3 block)
{
  return dim3((width + block.x - 1) / block.x, (height + block.y - 1) / block.y);
}

cudaError_t launch_denoise_setup(const DenoiseSetupParams &P, bool from_seeds, cudaStream_t stream)
{
  const dim3 block(32, 8);
  const dim3 grid = grid_for(P.width, P.height, block);
  if(from_seeds)
    denoise_setup_kernel<true><<<grid, block, 0, stream>>>(P);
  else
    denoise_setup_kernel<false><<<grid, block, 0, stream>>>(P);
  return cudaGetLastError();
}

size_t denoise_block_smem_bytes()
{
  return (size_t)(NPLANES * PLANE + 2 * PAD) * sizeof(float) + 16 + 128;   // planes + pads + mbarrier + alignment slack
}

cudaError_t launch_denoise_block(const DenoiseBlockParams &P, cudaStream_t stream)
{
  static bool configured[64] = {false};
  int device = 0;
  cudaGetDevice(&device);
  const size_t smem = denoise_block_smem_bytes();
  if(device < 0 || device >= 64 || !configured[device])
  {
    const cudaError_t err = cudaFuncSetAttribute(denoise_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem);
    if(err != cudaSuccess) return err;
    if(device >= 0 && device < 64) configured[device] = true;
  }
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t();
  cfg.gridDim = dim3((P.width + TW - 1) / TW, (P.height + TH - 1) / TH);
  cfg.blockDim = dim3(32, NTHREADS / 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, denoise_block_kernel, P);
}

} // namespace rmdb
