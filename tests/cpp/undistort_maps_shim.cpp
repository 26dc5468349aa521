// undistort_maps_shim.cpp -- exposes the PRODUCT's host-side map computation
// (rpg_open_remode_b200/csrc/undistort_maps.h, the code rmd_seeds_init_undistortion_map
// runs) to the CPU test suite, which compares it with OpenCV bit for bit.
#include "../../rpg_open_remode_b200/csrc/undistort_maps.h"

extern "C" void product_undistort_maps(int width, int height, float fx, float fy, float cx, float cy,
                                       float k1, float k2, float p1, float p2, int16_t *xy, uint16_t *frac)
{
  rmdb::compute_undistort_maps(width, height, fx, fy, cx, cy, k1, k2, p1, p2, xy, frac);
}
