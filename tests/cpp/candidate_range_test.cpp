// Host check of csrc/candidate_range.cuh (no GPU): the functions every kernel uses to decide WHICH candidates of an
// epipolar segment exist and are searched -- count_candidates (the reference's `l += 0.7f` accumulation in blocks of
// 16 with checkpoints), candidate_l (restart from a checkpoint) and accepted_range (closed-form estimate of the
// index range inside the image, fixed exactly against the real candidates) -- against the naive transcription of
// the reference's loop, src/epipolar_match.cu:85-97.  The header is the SAME file nvcc compiles for the device
// (fma_rn = fmaf here, __fmaf_rn there: both one IEEE rounding; the device divides approximately under
// -use_fast_math where the host divides exactly, which only moves the ESTIMATE the exact fix-up starts from).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#include "../../rpg_open_remode_b200/csrc/candidate_range.cuh"

using namespace rmdb;

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static double urand()
{
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (double)(rng_state >> 11) / 9007199254740992.0;
}
static bool same_bits(float a, float b) { return memcmp(&a, &b, sizeof(float)) == 0; }

static long failures = 0, cases = 0, with_range = 0, exact_scans = 0;

template<int PS>
static void check(const int width, const int height, const EpiSegment &seg)
{
  ++cases;
  // the reference's loop, literally (epipolar_match.cu:85-97); RMD_MAX_EXTENT 100 px bounds it at 143 candidates
  std::vector<float> ls;
  std::vector<char> ok;
  for(float l = -seg.half_len; l <= seg.half_len && (int)ls.size() < staged::L_CHECKPOINT_STEP * staged::L_CHECKPOINTS; l += 0.7f)
  {
    const float2 px = make_float2(fmaf(l, seg.dir.x, seg.mean.x), fmaf(l, seg.dir.y, seg.mean.y));
    const bool rejected = (px.x >= (float)(width - PS)) || (px.y >= (float)(height - PS)) || (px.x < (float)PS) || (px.y < (float)PS);
    ls.push_back(l);
    ok.push_back(!rejected);
  }
  float ckpt[staged::L_CHECKPOINTS];
  for(int i = 0; i < staged::L_CHECKPOINTS; ++i) ckpt[i] = -12345.0f;
  const int n = count_candidates(seg.half_len, ckpt);
  if(n != (int)ls.size())
  {
    if(failures++ < 10) printf("count_candidates: %d, reference loop %zu (half_len %a)\n", n, ls.size(), seg.half_len);
    return;
  }
  for(int k = 0; k < n; ++k)
  {
    if((k % staged::L_CHECKPOINT_STEP) == 0 && !same_bits(ckpt[k / staged::L_CHECKPOINT_STEP], ls[k]))
      if(failures++ < 10) printf("checkpoint %d: %a vs %a\n", k / staged::L_CHECKPOINT_STEP, ckpt[k / staged::L_CHECKPOINT_STEP], ls[k]);
    if(!same_bits(candidate_l(ckpt, k), ls[k]))
      if(failures++ < 10) printf("candidate_l(%d): %a vs %a\n", k, candidate_l(ckpt, k), ls[k]);
  }
  FilterParams P;
  memset(&P, 0, sizeof(P));
  P.width = width; P.height = height;
  int k_lo = 0, k_hi = 0;
  accepted_range<PS>(P, seg, n, ckpt, k_lo, k_hi);
  int first = -1, last = -1;
  for(int k = 0; k < n; ++k)
    if(ok[k]) { if(first < 0) first = k; last = k; }
  if(last < 0)
  {
    if(k_hi >= 0)
      if(failures++ < 10) printf("accepted_range: [%d, %d] but the loop accepts nothing\n", k_lo, k_hi);
    return;
  }
  ++with_range;
  if(k_lo != first || k_hi != last)
    if(failures++ < 10)
      printf("accepted_range: [%d, %d], reference loop [%d, %d] of %d (mean %a %a dir %a %a half %a, %dx%d P%d)\n", k_lo, k_hi, first, last, n,
             seg.mean.x, seg.mean.y, seg.dir.x, seg.dir.y, seg.half_len, width, height, PS);
  // what the kernels rely on when they skip everything outside the range: it is contiguous
  for(int k = first; k <= last; ++k)
    if(!ok[k])
    {
      if(failures++ < 10) printf("accepted candidates are not contiguous: %d rejected inside [%d, %d]\n", k, first, last);
      break;
    }
}

template<int PS>
static void sweep(const int width, const int height, const int n_cases)
{
  for(int c = 0; c < n_cases; ++c)
  {
    EpiSegment seg;
    const int kind = c % 16;
    const double ang = urand() * 6.283185307179586;
    seg.dir = make_float2((float)cos(ang), (float)sin(ang));
    if(kind == 1) seg.dir = make_float2(1.0f, 0.0f);
    if(kind == 2) seg.dir = make_float2(0.0f, -1.0f);
    if(kind == 3) seg.dir = make_float2((float)(1e-7 * (urand() - 0.5)), 1.0f);      // "does not move along x"
    if(kind == 4) seg.dir = make_float2(0.0f, 0.0f);                                  // zero-length segment (defined deviation 1)
    // centres everywhere: deep inside, straddling each border, well outside
    seg.mean = make_float2((float)(urand() * (width + 160) - 80), (float)(urand() * (height + 160) - 80));
    if(kind == 5) seg.mean.x = (float)PS + (float)(urand() - 0.5);                    // on the left bound, "too close to call"
    if(kind == 6) seg.mean.y = (float)(height - PS) + (float)(urand() - 0.5);
    if(kind == 7) seg.mean = make_float2((float)(PS + urand() * 3), (float)(PS + urand() * 3));   // a corner
    seg.half_len = (float)(urand() * 50.0);
    if(kind == 8) seg.half_len = 50.0f;                                               // the 100 px cap: 143 candidates
    if(kind == 9) seg.half_len = (float)(urand() * 0.7);                              // one or two candidates
    if(kind == 10) seg.half_len = 0.0f;
    if(kind == 11) seg.half_len = NAN;                                                // sigma^2 <= 0 (defined deviation 2): no candidate
    if(kind == 12) seg.mean.x = NAN;
    if(kind == 13) seg.mean.y = INFINITY;
    if(kind == 11 || kind == 12 || kind == 13) ++exact_scans;
    check<PS>(width, height, seg);
  }
}

int main()
{
  sweep<5>(640, 480, 120000);
  sweep<7>(1920, 1080, 120000);
  sweep<7>(203, 131, 60000);
  sweep<5>(11, 11, 20000);        // the smallest image rmd_seeds_create accepts for a 5x5 patch: one searchable pixel
  printf("%ld segments, %ld with an accepted range, %ld NaN / inf cases, %ld failures\n", cases, with_range, exact_scans, failures);
  if(failures == 0 && with_range > cases / 4)
  {
    printf("ALL CANDIDATE RANGE TESTS PASSED\n");
    return 0;
  }
  return 1;
}
