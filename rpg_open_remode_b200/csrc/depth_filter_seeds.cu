// depth_filter_seeds.cu -- fused depth filter, SEED-MAJOR variant for the steady state of a keyframe.
//
// Once most seeds of a keyframe have converged (or diverged) only a few per cent of the pixels are still
// updated every frame, a handful per 32x8 tile.  The tile-organised kernel (depth_filter_staged.cu) then spends
// its time on per-tile chains -- fetch an entry, classify 256 pixels to find 3 live ones, list the tile again --
// and a frame lasts as long as its slowest tile.  But seeds never interact (src/seed_check.cu,
// src/epipolar_match.cu, src/seed_update.cu read no neighbour's state), so nothing forces the tile structure, or
// even a frame barrier, on them:
//
//   * the live seeds (state UPDATE / NO_MATCH) are kept as a COMPACT LIST of (x, y);
//   * one warp owns a group of SEED_GROUP = 8 listed seeds: their (mu, sigma^2, a, b) records, states and 5x5 / 7x7
//     reference patches stay in registers / shared memory;
//   * the warp walks its group through ALL FRAMES OF THE LAUNCH (up to SEED_FRAMES_MAX) on its own: classify,
//     epipolar segment + candidate range (one lane per seed), NCC search (lanes = the group's (seed, candidate)
//     pairs, taps of the current frame straight from L2), triangulation + Bayesian update (one lane per seed).
//     No inter-warp, inter-CTA or inter-frame synchronisation exists; a launch ends when the slowest GROUP has
//     done its frames, not frame by frame;
//   * seeds that reach an absorbing state drop out; the survivors are appended to the next launch's list, so the
//     list stays compact without a separate pass.
//
// The per-seed steps are the shared ones of depth_filter_seed_steps.cuh / depth_filter_math.cuh (every rounding
// pinned), so results are bit-identical to the tile-organised and direct kernels (tests/test_gpu_seed_mode.py).
#include <cudaTypedefs.h>

#include <algorithm>
#include <mutex>

#include "depth_filter.cuh"
#include "depth_filter_seed_steps.cuh"

namespace rmdb
{

namespace
{

constexpr int SEED_GROUP = WARP_TILE_MAX_SEEDS;   // seeds per warp
constexpr int SEED_WARPS = 8;                     // warps per CTA

template<int PS>
struct alignas(16) SeedGroupSmem   // per warp
{
  SearchRec rec[SEED_GROUP];
  unsigned long long best[SEED_GROUP];
  float l_checkpoint[SEED_GROUP][L_CHECKPOINTS];
  float templ[SEED_GROUP][PS * PS];     // reference patches of the group's seeds (constant over the frames)
  int cum[SEED_GROUP];
};

} // namespace

// conv map -> compact list of the seeds that are still updated; also counts the CONVERGED ones.
__global__ void __launch_bounds__(256) seed_list_build_kernel(const int *conv, int conv_stride, int width, int height,
                                                              unsigned int *list, unsigned int *ctl)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  int state = RMD_BORDER;
  if(x < width && y < height)
    state = conv[(size_t)y * conv_stride + x];
  const bool live = (state == RMD_UPDATE || state == RMD_NO_MATCH || state == RMD_NOT_VISIBLE);
  const unsigned int live_mask = __ballot_sync(0xffffffffu, live);
  const unsigned int conv_mask = __ballot_sync(0xffffffffu, state == RMD_CONVERGED);
  unsigned int base = 0u;
  if(threadIdx.x == 0)
  {
    if(live_mask) base = atomicAdd(ctl + 0, (unsigned int)__popc(live_mask));
    if(conv_mask) atomicAdd(ctl + 3, (unsigned int)__popc(conv_mask));
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  if(live)
    list[base + __popc(live_mask & ((1u << threadIdx.x) - 1u))] = (unsigned int)x | ((unsigned int)y << 16);
}

template<int PS>
__global__ void __launch_bounds__(32 * SEED_WARPS, (PS <= 5 ? 3 : 2)) depth_filter_seeds_kernel(
    const __grid_constant__ SeedModeBatch B)
{
  __shared__ SeedGroupSmem<PS> s_group[SEED_WARPS];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x, wid = threadIdx.y;
  SeedGroupSmem<PS> &G = s_group[wid];
  // ctl: [cur] live seeds listed for this launch, [cur ^ 1] survivors appended for the next one, [2] group cursor,
  // [3] CONVERGED seeds of the keyframe (running total), [4] CTAs done
  const unsigned int n_listed = B.ctl[B.cur];
  const unsigned int n_groups = (n_listed + SEED_GROUP - 1) / SEED_GROUP;
  const int slot = lane;                          // lanes 0..7 own the group's seeds
  for(;;)
  {
    unsigned int gi = 0u;
    if(lane == 0)
      gi = atomicAdd(B.ctl + 2, 1u);
    gi = __shfl_sync(0xffffffffu, gi, 0);
    if(gi >= n_groups)
      break;
    const bool mine = (slot < SEED_GROUP) && (gi * SEED_GROUP + (unsigned int)slot < n_listed);
    unsigned int pos = 0u;
    int sx = 0, sy = 0, state = RMD_BORDER;
    float4 seed = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 *seed_ptr = nullptr;
    int *conv_ptr = nullptr;
    float2 stats = make_float2(0.f, 0.f);
    const FilterParams &P0 = B.p[0];
    if(mine)
    {
      pos = B.list_cur[gi * SEED_GROUP + slot];
      sx = (int)(pos & 0xffffu); sy = (int)(pos >> 16);
      seed_ptr = P0.seed + (size_t)sy * P0.seed_stride + sx;
      conv_ptr = P0.conv + (size_t)sy * P0.conv_stride + sx;
      seed = *seed_ptr;
      state = *conv_ptr;
      stats = __ldg(P0.templ + (size_t)sy * P0.templ_stride + sx);
#pragma unroll
      for(int j = 0; j < PS; ++j)
#pragma unroll
        for(int i = 0; i < PS; ++i)
          G.templ[slot][j * PS + i] = __ldg(P0.ref + (size_t)(sy - PS / 2 + j) * P0.ref_stride + (sx - PS / 2 + i));
    }
    bool alive = mine;
    __syncwarp();
#pragma unroll 1
    for(int f = 0; f < B.n_frames; ++f)
    {
      const FilterParams &P = B.p[f];
      // ---- 0. convergence check of the seed's own lane (src/seed_check.cu:53-66; never BORDER: listed seeds are interior)
      bool active = false;
      if(alive)
      {
        const int st = classify_pixel<PS>(P, sx, sy, state, seed, active);
        if(!active)
        {
          if(st != state)
            *conv_ptr = st;
          if(st == RMD_CONVERGED)
            atomicAdd(B.ctl + 3, 1u);
          state = st;
          alive = false;          // absorbing: the seed leaves the list
        }
      }
      if(!__any_sync(0xffffffffu, active))
      {
        if(!__any_sync(0xffffffffu, alive))
          break;                  // the whole group is done for good
        continue;
      }
      // ---- 1. search set-up
      EpiSegment seg;
      seg.mean = make_float2(0.f, 0.f); seg.dir = make_float2(0.f, 0.f); seg.half_len = 0.f;
      int n_cand = 0, k_lo = INT_MAX, k_hi = -1, cnt = 0;
      if(active)
      {
        seg = epipolar_segment(P, sx, sy, seed.x, seed.y);
        n_cand = count_candidates(seg.half_len, G.l_checkpoint[slot]);
        accepted_range<PS>(P, seg, n_cand, G.l_checkpoint[slot], k_lo, k_hi);
        cnt = (k_hi >= 0) ? (k_hi - k_lo + 1) : 0;
        SearchRec rec;
        rec.mean_x = seg.mean.x; rec.mean_y = seg.mean.y; rec.dir_x = seg.dir.x; rec.dir_y = seg.dir.y;
        rec.half_len = seg.half_len; rec.sum_templ = stats.x; rec.denom = stats.y;
        rec.n = (k_hi >= 0) ? k_lo : 0;
        G.rec[slot] = rec;
        G.best[slot] = K_NO_MATCH;
      }
      {
        int inc = active ? cnt : 0;
#pragma unroll
        for(int off = 1; off < SEED_GROUP; off <<= 1)
        {
          const int up = __shfl_up_sync(0xffffffffu, inc, off);
          if(lane >= off) inc += up;
        }
        if(lane < SEED_GROUP)
          G.cum[lane] = inc;
      }
      __syncwarp();
      const int total = G.cum[SEED_GROUP - 1];
      // ---- 3. NCC search: lane = (seed, candidate) pair of the group
#pragma unroll 1
      for(int q = lane; q < total; q += 32)
      {
        int j = 0;
#pragma unroll
        for(int i = 0; i < SEED_GROUP - 1; ++i) j += (G.cum[i] <= q) ? 1 : 0;
        const SearchRec R = G.rec[j];
        const int k = R.n + q - (j ? G.cum[j - 1] : 0);
        const float l = candidate_l(G.l_checkpoint[j], k);
        const float2 px = candidate_px(R.mean_x, R.mean_y, R.dir_x, R.dir_y, l);
        if(candidate_rejected<PS>(px, P.width, P.height))
          continue;
        float templ[PS * PS];
#pragma unroll
        for(int t = 0; t < PS * PS; ++t) templ[t] = G.templ[j][t];
        const TapFrame frame = tap_frame<PS>(px, P.tex_quant);
        const GlobalTaps taps(P.curr, P.curr_stride, frame);
        const float ncc = ncc_score<PS>(taps, frame, templ, R.sum_templ, R.denom);
        if(ncc > -1.0f)
          atomicMax(&G.best[j], (((unsigned long long)orderable(ncc)) << 32) |
                                    (unsigned long long)(0xffffffffu - (unsigned int)k));
      }
      __syncwarp();
      // ---- 4. triangulation + Bayesian update by the seed's lane; the record stays in registers
      if(active)
        state = apply_match<PS>(P, sx, sy, seg, n_cand, G.best[slot], G.l_checkpoint[slot], seed, seed_ptr, state, conv_ptr);
      __syncwarp();
    }
    // survivors go to the next launch's list
    {
      const unsigned int keep = __ballot_sync(0xffffffffu, alive);
      unsigned int base = 0u;
      if(lane == 0 && keep)
        base = atomicAdd(B.ctl + (B.cur ^ 1), (unsigned int)__popc(keep));
      base = __shfl_sync(0xffffffffu, base, 0);
      if(alive)
        B.list_next[base + __popc(keep & ((1u << lane) - 1u))] = pos;
    }
    __syncwarp();
  }
  // the last CTA out rewinds the cursor and empties the list just consumed (the next launch appends to it)
  __syncthreads();
  if(lane == 0 && wid == 0)
  {
    __threadfence();
    const unsigned int done = atomicAdd(B.ctl + 4, 1u);
    if(done == gridDim.x - 1u)
    {
      B.ctl[2] = 0u;
      B.ctl[4] = 0u;
      B.ctl[B.cur] = 0u;
    }
  }
}

namespace
{

template<int PS>
struct SeedsLaunch
{
  static int slots(int device)
  {
    static std::mutex mutex;
    static int cached[64] = {0};
    std::lock_guard<std::mutex> lock(mutex);
    if(device < 0 || device >= 64)
      return 0;
    if(cached[device] == 0)
    {
      int per_sm = 0, sms = 0;
      if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, depth_filter_seeds_kernel<PS>, 32 * SEED_WARPS, 0) !=
             cudaSuccess ||
         cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
        return 0;
      cached[device] = per_sm * sms;
    }
    return cached[device];
  }

  static cudaError_t launch(const SeedModeBatch &B, int max_groups, cudaStream_t stream)
  {
    int device = 0;
    cudaError_t err = cudaGetDevice(&device);
    if(err != cudaSuccess) return err;
    const int n_slots = slots(device);
    if(n_slots <= 0)
    {
      err = cudaGetLastError();
      return err != cudaSuccess ? err : cudaErrorInvalidDevice;
    }
    const int ctas = std::max(1, std::min(n_slots, (max_groups + SEED_WARPS - 1) / SEED_WARPS));
    cudaLaunchConfig_t cfg = cudaLaunchConfig_t();
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(32, SEED_WARPS); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, depth_filter_seeds_kernel<PS>, B);
  }
};

} // namespace

cudaError_t launch_seed_list_build(const int *conv, int conv_stride, int width, int height, unsigned int *list,
                                   unsigned int *ctl, cudaStream_t stream)
{
  const dim3 block(32, 8);
  const dim3 grid((width + 31) / 32, (height + 7) / 8);
  seed_list_build_kernel<<<grid, block, 0, stream>>>(conv, conv_stride, width, height, list, ctl);
  return cudaGetLastError();
}

cudaError_t launch_depth_filter_seeds(const SeedModeBatch &B, int max_listed, int patch_side, cudaStream_t stream)
{
  if(B.n_frames < 1 || B.n_frames > SEED_FRAMES_MAX)
    return cudaErrorInvalidValue;
  const int max_groups = (max_listed + SEED_GROUP - 1) / SEED_GROUP;
  if(patch_side == 5) return SeedsLaunch<5>::launch(B, max_groups, stream);
  if(patch_side == 7) return SeedsLaunch<7>::launch(B, max_groups, stream);
  return cudaErrorInvalidValue;
}

} // namespace rmdb
