// RoIAlign sample-grid geometry shared by the forward and backward kernels (roi_align.cu,
// roi_align_bwd.cu): bit-identical in/out decisions and cells in every kernel.
#pragma once

#include "common.cuh"

namespace nafae {

struct RoiGeom {
  float start_w, start_h, bin_w, bin_h;
  int batch;
};

// roi_align_kernel.cu:33-43 as compiled: end-start is fma(x2, s, -RN(x1*s)); "+ 1." in double
// then fmaxf's float conversion is an exact float add; bin is a double division rounded to float.
__device__ __forceinline__ RoiGeom roi_geom(const float* __restrict__ roi, float scale, int sh,
                                            int sw) {
  RoiGeom g;
  g.batch = (int)roi[0];
  g.start_w = __fmul_rn(roi[1], scale);
  g.start_h = __fmul_rn(roi[2], scale);
  const float rw = fmaxf(__fadd_rn(__fmaf_rn(roi[3], scale, -g.start_w), 1.f), 0.f);
  const float rh = fmaxf(__fadd_rn(__fmaf_rn(roi[4], scale, -g.start_h), 1.f), 0.f);
  g.bin_h = __double2float_rn(__ddiv_rn((double)rh, __dsub_rn((double)sh, 1.)));
  g.bin_w = __double2float_rn(__ddiv_rn((double)rw, __dsub_rn((double)sw, 1.)));
  return g;
}

// one axis of a sample point (roi_align_kernel.cu:45-49,54,58-59): returns false if outside
__device__ __forceinline__ bool axis_sample(float start, float bin, int p, int extent, int* cell,
                                            float* ratio) {
  const float x = __fmaf_rn((float)p, bin, start);
  if (x < 0.f || x >= (float)extent || x != x) {
    *cell = 0;
    *ratio = 0.f;
    // NaN: the reference's comparisons are all false -> it would read out of bounds; we
    // define the sample as outside instead.
    return false;
  }
  const int c = (int)fminf(floorf(x), (float)(extent - 2));
  *cell = c;
  *ratio = __fsub_rn(x, (float)c);
  return true;
}


}  // namespace nafae
