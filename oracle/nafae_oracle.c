/*
 * nafae_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the NAFAE detector-side hot path (reference jshi31/NAFAE):
 * greedy NMS, the proposal-layer tail, corner-grid RoIAlign (+2x2 avg/max post
 * pool) and max RoIPool, forward and backward.  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load it.
 *
 * The reference has NO CPU implementation of NMS or RoIAlign (nms_wrapper.py:11-18
 * always calls the GPU path; functions/roi_align.py:28-29 raises on CPU), so the
 * arithmetic below follows the reference CUDA sources *as nvcc 12.9 compiles them
 * with its default -fmad=true* (verified with cuobjdump -sass on the unmodified
 * reference files built for sm_100a, see oracle/Makefile target `_ref`):
 *   - devIoU: the column box area is FMA-fused into the union,
 *       u = fma(bw, bh, RN(aw*ah)) - inter            (nms_cuda_kernel.cu:31-39)
 *   - RoIAlign geometry: roi_width = max(fma(x2, s, -RN(x1*s)) + 1, 0) and
 *       h = fma(ph, bin_h, start_h)                   (roi_align_kernel.cu:33-46)
 *   - RoIAlign interpolation: mixed float/double exactly as C promotes it, with
 *       the two DFMA contractions ptxas/nvvm emit     (roi_align_kernel.cu:64-67)
 * Every such place uses fmaf()/fma() explicitly; build with -ffp-contract=off.
 *
 * Parity pin: tests/golden/ref_gpu_*.npz hold outputs of the unmodified reference
 * kernels (oracle/_ref) executed on a B200; tests/test_oracle_golden.py checks this
 * file against them bit-for-bit.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))
#define NMS_TPB 64 /* threadsPerBlock = sizeof(unsigned long long)*8, nms_cuda_kernel.cu:29 */

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ORACLE_API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------ NMS -- */

/* devIoU, nms_cuda_kernel.cu:31-39.  a = row box (earlier / higher score),
 * b = column box.  Compiled form: Sa is a rounded product, Sb is fused. */
static inline float oracle_iou(const float* a, const float* b) {
  float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  float width = fmaxf(right - left + 1.f, 0.f);
  float height = fmaxf(bottom - top + 1.f, 0.f);
  float interS = width * height;
  float Sa = (a[2] - a[0] + 1.f) * (a[3] - a[1] + 1.f);
  float u = fmaf(b[2] - b[0] + 1.f, b[3] - b[1] + 1.f, Sa);
  return interS / (u - interS);
}

ORACLE_API float oracle_iou_pair(const float* a, const float* b) { return oracle_iou(a, b); }

/* nms_kernel (nms_cuda_kernel.cu:41-85) + host sweep (:117-144).
 * dets: (n, dim>=4) rows [x1,y1,x2,y2,(score)], already sorted by score desc.
 * keep_out: n ints, num_out: 1 int.  mask_out (optional): n*ceil(n/64) words,
 * only the words the sweep reads (column block >= row block) are filled. */
ORACLE_API void oracle_nms(const float* dets, int n, int dim, float thresh, int* keep_out,
                           int* num_out, uint64_t* mask_out) {
  if (n <= 0) {
    *num_out = 0;
    return;
  }
  const int col_blocks = (n + NMS_TPB - 1) / NMS_TPB;
  uint64_t* mask = mask_out ? mask_out : (uint64_t*)calloc((size_t)n * col_blocks, sizeof(uint64_t));
  if (mask_out) memset(mask, 0, (size_t)n * col_blocks * sizeof(uint64_t));
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < n; ++i) {
    const int row_blk = i / NMS_TPB;
    const float* cur = dets + (size_t)i * dim;
    for (int cb = row_blk; cb < col_blocks; ++cb) {
      const int col_size = (n - cb * NMS_TPB) < NMS_TPB ? (n - cb * NMS_TPB) : NMS_TPB;
      int start = (cb == row_blk) ? (i % NMS_TPB) + 1 : 0; /* :74-76 */
      uint64_t t = 0;
      for (int j = start; j < col_size; ++j)
        if (oracle_iou(cur, dets + (size_t)(cb * NMS_TPB + j) * dim) > thresh) t |= 1ULL << j;
      mask[(size_t)i * col_blocks + cb] = t;
    }
  }
  uint64_t* remv = (uint64_t*)calloc(col_blocks, sizeof(uint64_t));
  int num_to_keep = 0;
  for (int i = 0; i < n; ++i) { /* :132-144 */
    int nblock = i / NMS_TPB, inblock = i % NMS_TPB;
    if (!(remv[nblock] & (1ULL << inblock))) {
      keep_out[num_to_keep++] = i;
      const uint64_t* p = mask + (size_t)i * col_blocks;
      for (int j = nblock; j < col_blocks; ++j) remv[j] |= p[j];
    }
  }
  *num_out = num_to_keep;
  free(remv);
  if (!mask_out) free(mask);
}

/* Proposal-layer tail, lib/model/rpn/proposal_layer.py:127-163.
 * proposals (F,n,4) and scores (F,n) are already in score-desc order (the
 * torch.sort of :125 is the caller's job), pre-NMS top-N (:139-140) is applied
 * here.  rois (F,post,5) / roi_scores (F,post) are zero-initialised, column 0 of
 * every row (padding included) is the frame index (:160). */
ORACLE_API void oracle_proposal_tail(const float* proposals, const float* scores, int F, int n,
                                     int pre_nms_topn, int post_nms_topn, float thresh,
                                     float* rois, float* roi_scores, int* num_kept) {
  /* :139 compares against scores_keep.numel() == F*n, the whole batch */
  int m = n;
  if (pre_nms_topn > 0 && (long)pre_nms_topn < (long)F * n && pre_nms_topn < n) m = pre_nms_topn;
  memset(rois, 0, (size_t)F * post_nms_topn * 5 * sizeof(float));
  memset(roi_scores, 0, (size_t)F * post_nms_topn * sizeof(float));
#pragma omp parallel for schedule(dynamic, 1)
  for (int f = 0; f < F; ++f) {
    float* dets = (float*)malloc((size_t)(m > 0 ? m : 1) * 5 * sizeof(float));
    int* keep = (int*)malloc((size_t)(m > 0 ? m : 1) * sizeof(int));
    int nk = 0;
    for (int i = 0; i < m; ++i) {
      memcpy(dets + (size_t)i * 5, proposals + ((size_t)f * n + i) * 4, 4 * sizeof(float));
      dets[(size_t)i * 5 + 4] = scores[(size_t)f * n + i];
    }
    oracle_nms(dets, m, 5, thresh, keep, &nk, NULL);
    if (post_nms_topn > 0 && nk > post_nms_topn) nk = post_nms_topn; /* :154-155 */
    for (int k = 0; k < post_nms_topn; ++k) rois[((size_t)f * post_nms_topn + k) * 5] = (float)f;
    for (int k = 0; k < nk; ++k) {
      memcpy(rois + ((size_t)f * post_nms_topn + k) * 5 + 1, dets + (size_t)keep[k] * 5,
             4 * sizeof(float));
      roi_scores[(size_t)f * post_nms_topn + k] = dets[(size_t)keep[k] * 5 + 4];
    }
    if (num_kept) num_kept[f] = nk;
    free(dets);
    free(keep);
  }
}

/* ------------------------------------------------------------- RoIAlign -- */

typedef struct {
  int valid;      /* sample inside [0,H) x [0,W) */
  int hstart, wstart;
  float hr, wr;   /* h_ratio, w_ratio */
} align_pt;

/* roi_align_kernel.cu:33-60 (same lines :106-128 in the backward kernel). */
static inline void align_geometry(const float* roi, float scale, int height, int width, int ah,
                                  int aw, int ph, int pw, align_pt* g) {
  float roi_start_w = roi[1] * scale;
  float roi_start_h = roi[2] * scale;
  /* roi_end - roi_start is contracted: fma(x2, scale, -RN(x1*scale)); the "+ 1." in
   * double followed by fmaxf's float conversion is an exact float add. */
  float roi_width = fmaxf(fmaf(roi[3], scale, -roi_start_w) + 1.f, 0.f);
  float roi_height = fmaxf(fmaf(roi[4], scale, -roi_start_h) + 1.f, 0.f);
  float bin_size_h = (float)((double)roi_height / ((double)ah - 1.));
  float bin_size_w = (float)((double)roi_width / ((double)aw - 1.));
  float h = fmaf((float)ph, bin_size_h, roi_start_h);
  float w = fmaf((float)pw, bin_size_w, roi_start_w);
  g->valid = !(h < 0 || h >= height || w < 0 || w >= width);
  if (!g->valid) return;
  g->hstart = (int)fminf(floorf(h), (float)(height - 2));
  g->wstart = (int)fminf(floorf(w), (float)(width - 2));
  g->hr = h - (float)g->hstart;
  g->wr = w - (float)g->wstart;
}

/* int img_start = roi_batch_ind * channels * height * width, evaluated in float
 * left to right and truncated (roi_align_kernel.cu:51). */
static inline long align_img_start(float roi_batch_ind, int channels, int height, int width) {
  float v = roi_batch_ind * (float)channels;
  v = v * (float)height;
  v = v * (float)width;
  return (long)(int)v;
}

/* ROIAlignForward, roi_align_kernel.cu:15-70.  top (R,C,ah,aw). */
ORACLE_API void oracle_roi_align_forward(const float* bottom, float scale, int num_rois,
                                         int height, int width, int channels, int ah, int aw,
                                         const float* rois, float* top) {
#pragma omp parallel for schedule(static)
  for (int n = 0; n < num_rois; ++n) {
    const float* roi = rois + (size_t)n * 5;
    const long img_start = align_img_start(roi[0], channels, height, width);
    for (int ph = 0; ph < ah; ++ph)
      for (int pw = 0; pw < aw; ++pw) {
        align_pt g;
        align_geometry(roi, scale, height, width, ah, aw, ph, pw, &g);
        for (int c = 0; c < channels; ++c) {
          float* out = top + (((size_t)n * channels + c) * ah + ph) * aw + pw;
          if (!g.valid) {
            *out = 0.f;
            continue;
          }
          const float* p = bottom + img_start + ((long)c * height + g.hstart) * width + g.wstart;
          float ul = p[0], ur = p[1], dl = p[width], dr = p[width + 1];
          /* :64-67 as compiled: ul,ur terms in double; dl*h_ratio and (dr*h_ratio)*w_ratio in
           * float; sum ((t1+t2)+t3)+t4 in double with t1,t3's last product fused. */
          double omh = 1. - (double)g.hr;
          double omw = 1. - (double)g.wr;
          double t2 = ((double)ur * omh) * (double)g.wr;
          double s = fma((double)ul * omh, omw, t2);
          s = fma(omw, (double)(dl * g.hr), s);
          s = s + (double)((dr * g.hr) * g.wr);
          *out = (float)s;
        }
      }
  }
}

/* ROIAlignBackward, roi_align_kernel.cu:94-143.  bottom_diff (B,C,H,W) must be
 * zero-filled by the caller (functions/roi_align.py:38-39).  The reference's
 * atomicAdd order is unspecified; this restatement accumulates in index order
 * (n, c, ph, pw; upleft, upright, downleft, downright) -- compare with tolerance. */
ORACLE_API void oracle_roi_align_backward(const float* top_diff, float scale, int batch_size,
                                          int num_rois, int height, int width, int channels,
                                          int ah, int aw, const float* rois, float* bottom_diff) {
  (void)batch_size;
  /* parallel over channels: each (image, channel) plane is private to one thread */
#pragma omp parallel for schedule(static)
  for (int c = 0; c < channels; ++c) {
    for (int n = 0; n < num_rois; ++n) {
      const float* roi = rois + (size_t)n * 5;
      const long img_start = align_img_start(roi[0], channels, height, width);
      for (int ph = 0; ph < ah; ++ph)
        for (int pw = 0; pw < aw; ++pw) {
          align_pt g;
          align_geometry(roi, scale, height, width, ah, aw, ph, pw, &g);
          if (!g.valid) continue;
          float td = top_diff[(((size_t)n * channels + c) * ah + ph) * aw + pw];
          float* p = bottom_diff + img_start + ((long)c * height + g.hstart) * width + g.wstart;
          double omh = 1. - (double)g.hr;
          float omw_f = 1.f - g.wr; /* "(1 - w_ratio)" is int - float = float, :137,139 */
          p[0] += (float)(((double)td * omh) * (double)omw_f);
          p[1] += (float)(((double)td * omh) * (double)g.wr);
          p[width] += (td * g.hr) * omw_f;
          p[width + 1] += (td * g.hr) * g.wr;
        }
    }
  }
}

/* avg_pool2d / max_pool2d(kernel_size=2, stride=1) as called by RoIAlignAvg /
 * RoIAlignMax (modules/roi_align.py:27-29, 40-42).  ATen accumulates the window
 * row-major starting from 0 and divides by the window size. */
ORACLE_API void oracle_pool2x2_forward(const float* x, int planes, int ih, int iw, int is_max,
                                       float* y) {
  const int oh = ih - 1, ow = iw - 1;
#pragma omp parallel for schedule(static)
  for (int p = 0; p < planes; ++p) {
    const float* xp = x + (size_t)p * ih * iw;
    float* yp = y + (size_t)p * oh * ow;
    for (int i = 0; i < oh; ++i)
      for (int j = 0; j < ow; ++j) {
        float a = xp[i * iw + j], b = xp[i * iw + j + 1];
        float c = xp[(i + 1) * iw + j], d = xp[(i + 1) * iw + j + 1];
        if (is_max) {
          /* ATen max_pool2d: val > maxval || isnan(val), scanning row-major from -inf */
          float m = -INFINITY;
          float v[4] = {a, b, c, d};
          for (int k = 0; k < 4; ++k)
            if (v[k] > m || isnan(v[k])) m = v[k];
          yp[i * ow + j] = m;
        } else {
          yp[i * ow + j] = (((a + b) + c) + d) / 4.f;
        }
      }
  }
}

/* Backward of the 2x2/stride-1 post pool: gy (planes,ih-1,iw-1) -> gx (planes,ih,iw).
 * avg: every window element receives gy/4, accumulated in output order.
 * max: the first maximal element (row-major, NaN wins) receives gy. */
ORACLE_API void oracle_pool2x2_backward(const float* x, const float* gy, int planes, int ih,
                                        int iw, int is_max, float* gx) {
  const int oh = ih - 1, ow = iw - 1;
#pragma omp parallel for schedule(static)
  for (int p = 0; p < planes; ++p) {
    const float* xp = x ? x + (size_t)p * ih * iw : NULL;
    const float* gp = gy + (size_t)p * oh * ow;
    float* o = gx + (size_t)p * ih * iw;
    memset(o, 0, (size_t)ih * iw * sizeof(float));
    for (int i = 0; i < oh; ++i)
      for (int j = 0; j < ow; ++j) {
        float g = gp[i * ow + j];
        if (is_max) {
          int idx[4] = {i * iw + j, i * iw + j + 1, (i + 1) * iw + j, (i + 1) * iw + j + 1};
          int best = idx[0]; /* ATen: maxindex starts at the window's first element */
          float m = -INFINITY;
          for (int k = 0; k < 4; ++k) {
            float v = xp[idx[k]];
            if (v > m || isnan(v)) {
              m = v;
              best = idx[k];
            }
          }
          o[best] += g;
        } else {
          float d = g / 4.f;
          o[i * iw + j] += d;
          o[i * iw + j + 1] += d;
          o[(i + 1) * iw + j] += d;
          o[(i + 1) * iw + j + 1] += d;
        }
      }
  }
}

/* -------------------------------------------------------------- RoIPool -- */

typedef struct {
  int batch, start_w, start_h, end_w, end_h, roi_w, roi_h;
  float bin_h, bin_w;
} pool_roi;

/* roi_pooling_kernel.cu:44-55 */
static inline pool_roi pool_roi_geometry(const float* roi, float scale, int ph_n, int pw_n) {
  pool_roi r;
  r.batch = (int)roi[0];
  r.start_w = (int)roundf(roi[1] * scale);
  r.start_h = (int)roundf(roi[2] * scale);
  r.end_w = (int)roundf(roi[3] * scale);
  r.end_h = (int)roundf(roi[4] * scale);
  r.roi_w = (int)fmaxf((float)(r.end_w - r.start_w + 1), 1.f);
  r.roi_h = (int)fmaxf((float)(r.end_h - r.start_h + 1), 1.f);
  r.bin_h = (float)r.roi_h / (float)ph_n;
  r.bin_w = (float)r.roi_w / (float)pw_n;
  return r;
}

static inline int clampi(int v, int lo, int hi) {
  return (int)fminf(fmaxf((float)v, (float)lo), (float)hi);
}

/* ROIPoolForward, roi_pooling_kernel.cu:24-93.  argmax is the flat index into the
 * whole (B,C,H,W) batch, -1 for an empty bin. */
ORACLE_API void oracle_roi_pool_forward(const float* bottom, float scale, int num_rois,
                                        int height, int width, int channels, int ph_n, int pw_n,
                                        const float* rois, float* top, int* argmax) {
#pragma omp parallel for schedule(static)
  for (int n = 0; n < num_rois; ++n) {
    pool_roi r = pool_roi_geometry(rois + (size_t)n * 5, scale, ph_n, pw_n);
    for (int c = 0; c < channels; ++c)
      for (int ph = 0; ph < ph_n; ++ph)
        for (int pw = 0; pw < pw_n; ++pw) {
          int hstart = (int)floorf((float)ph * r.bin_h);
          int wstart = (int)floorf((float)pw * r.bin_w);
          int hend = (int)ceilf((float)(ph + 1) * r.bin_h);
          int wend = (int)ceilf((float)(pw + 1) * r.bin_w);
          hstart = clampi(hstart + r.start_h, 0, height);
          hend = clampi(hend + r.start_h, 0, height);
          wstart = clampi(wstart + r.start_w, 0, width);
          wend = clampi(wend + r.start_w, 0, width);
          int is_empty = (hend <= hstart) || (wend <= wstart);
          float maxval = is_empty ? 0.f : -FLT_MAX;
          int maxidx = -1;
          long off = ((long)r.batch * channels + c) * height * width;
          for (int h = hstart; h < hend; ++h)
            for (int w = wstart; w < wend; ++w) {
              float v = bottom[off + (long)h * width + w];
              if (v > maxval) {
                maxval = v;
                maxidx = (int)(off + (long)h * width + w);
              }
            }
          size_t o = (((size_t)n * channels + c) * ph_n + ph) * pw_n + pw;
          top[o] = maxval;
          if (argmax) argmax[o] = maxidx;
        }
  }
}

/* ROIPoolBackward, roi_pooling_kernel.cu:128-203 (gather form, overwrites
 * bottom_diff).  Literal: every input cell scans all RoIs of its image. */
ORACLE_API void oracle_roi_pool_backward(const float* top_diff, const int* argmax, float scale,
                                         int batch_size, int num_rois, int height, int width,
                                         int channels, int ph_n, int pw_n, const float* rois,
                                         float* bottom_diff) {
  pool_roi* rr = (pool_roi*)malloc((size_t)(num_rois > 0 ? num_rois : 1) * sizeof(pool_roi));
  for (int n = 0; n < num_rois; ++n) rr[n] = pool_roi_geometry(rois + (size_t)n * 5, scale, ph_n, pw_n);
  const long total = (long)batch_size * channels * height * width;
#pragma omp parallel for schedule(static)
  for (long index = 0; index < total; ++index) {
    long t = index;
    int w = (int)(t % width);
    t /= width;
    int h = (int)(t % height);
    t /= height;
    int c = (int)(t % channels);
    t /= channels;
    int b = (int)t;
    float gradient = 0.f;
    for (int n = 0; n < num_rois; ++n) {
      const pool_roi* r = rr + n;
      if (b != r->batch) continue;
      if (!(w >= r->start_w && w <= r->end_w && h >= r->start_h && h <= r->end_h)) continue;
      size_t offset = (size_t)n * ph_n * pw_n * channels;
      int phstart = (int)floorf((float)(h - r->start_h) / r->bin_h);
      int phend = (int)ceilf((float)(h - r->start_h + 1) / r->bin_h);
      int pwstart = (int)floorf((float)(w - r->start_w) / r->bin_w);
      int pwend = (int)ceilf((float)(w - r->start_w + 1) / r->bin_w);
      phstart = clampi(phstart, 0, ph_n);
      phend = clampi(phend, 0, ph_n);
      pwstart = clampi(pwstart, 0, pw_n);
      pwend = clampi(pwend, 0, pw_n);
      for (int ph = phstart; ph < phend; ++ph)
        for (int pw = pwstart; pw < pwend; ++pw) {
          size_t o = offset + ((size_t)c * ph_n + ph) * pw_n + pw;
          if ((long)argmax[o] == index) gradient += top_diff[o];
        }
    }
    bottom_diff[index] = gradient;
  }
  free(rr);
}
