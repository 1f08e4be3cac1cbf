/*
 * oracle.c -- CPU restatement of the slam-robot visual front-end (see oracle.h).
 * TEST INFRASTRUCTURE ONLY: the checker for tests/ and the timed CPU arm of bench.py.
 *
 * Build (oracle/Makefile): -O2 -ffp-contract=off -fno-fast-math so that every rounding is
 * the one written here; fused multiply-adds are spelled fmaf().  A second build with the
 * reference's own flags (-O3 -ffast-math -march=native, Makefile:4) is used for CPU timing only.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ helpers */

static inline int reflect101(int i, int n) {
  /* cv::BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba */
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    else i = 2 * (n - 1) - i;
  }
  return i;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Declared reduction order (oracle.h): lane partials are filled by the caller, then a
 * pairwise tree with strides 16,8,4,2,1 -- what a __shfl_xor butterfly computes in lane 0. */
static inline float tree32(float* p) {
  for (int off = 16; off; off >>= 1)
    for (int l = 0; l < off; ++l) p[l] = p[l] + p[l + off];
  return p[0];
}

/* The Hessian tracker (P1) declares the other balanced tree: adjacent lanes first, strides 1,2,4,8,16 -- the order its
 * CUDA kernel gets from transposing the lane partials through shared memory (track_hessian.cu reduce_stats). */
static inline float tree32_adj(float* p) {
  for (int off = 1; off < 32; off <<= 1)
    for (int l = 0; l < 32; l += 2 * off) p[l] = p[l] + p[l + off];
  return p[0];
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ mask */

/* hessian.h:11-30 / klt.h:11-32.  rx,ry measured from 0.5*size (6.5), computed in double,
 * stored as float, summed in double in storage order, each float scaled by (double)len/sum. */
void orc_mask13(float* mask) {
  const int n = ORC_PATCH;
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x) {
      double rx = 0.5 * n - x, ry = 0.5 * n - y;
      double rr = rx * rx + ry * ry;
      mask[y * n + x] = (float)(1. / (15. + rr));
    }
  double sum = 0;
  for (int i = 0; i < ORC_PLEN; ++i) sum += mask[i];
  double scale = ORC_PLEN / sum;
  for (int i = 0; i < ORC_PLEN; ++i) mask[i] = (float)(mask[i] * scale);
}

/* ------------------------------------------------------------------ image primitives */

/* cv::cvtColor(CV_RGB2GRAY) on 8-bit data as OpenCV 4.13 computes it (hessian.h:100 applies it
 * to BGR bytes, i.e. the first byte gets the "R" weight -- preserved). */
void orc_gray_u8(const uint8_t* bgr, int w, int h, size_t stride, uint8_t* gray) {
  for (int y = 0; y < h; ++y) {
    const uint8_t* s = bgr + (size_t)y * stride;
    uint8_t* d = gray + (size_t)y * w;
    for (int x = 0; x < w; ++x)
      d[x] = (uint8_t)((9798 * s[3 * x] + 19235 * s[3 * x + 1] + 3735 * s[3 * x + 2] + (1 << 14)) >> 15);
  }
}

static void gray_f32(const uint8_t* bgr, int w, int h, size_t stride, float* dst) {
  /* convertTo(CV_32F, 1./255.) == (float)g * (float)(1./255.)  (hessian.h:101) */
  const float a = (float)(1. / 255.);
  for (int y = 0; y < h; ++y) {
    const uint8_t* s = bgr + (size_t)y * stride;
    float* d = dst + (size_t)y * w;
    for (int x = 0; x < w; ++x) {
      int g = (9798 * s[3 * x] + 19235 * s[3 * x + 1] + 3735 * s[3 * x + 2] + (1 << 14)) >> 15;
      d[x] = (float)g * a;
    }
  }
}

/* cv::GaussianBlur 5x5, BORDER_REFLECT_101, kernel (k2,k1,k0,k1,k2). dst may alias src.
 * OpenCV 4.13's AVX2 build evaluates the separable filter with FMAs in its vector bodies and with
 * plain multiply/add in its scalar tails; which columns are "tail" is a pure function of the width
 * (probed against cv2 column by column, tests/golden/make_golden.py re-runs the probe):
 *   row filter     scalar only for the last column of an odd-width image
 *   column filter  scalar for columns x >= (w/8)*8
 * Both scalar forms are (k0*c + (m1+p1)*k1) + (m2+p2)*k2. */
void orc_gauss5(const float* src, int w, int h, float k0, float k1, float k2, float* dst) {
  float* tmp = (float*)malloc(sizeof(float) * (size_t)w * h);
  const int row_body = (w & 1) ? w - 1 : w, col_body = (w / 8) * 8;
  for (int y = 0; y < h; ++y) {
    const float* s = src + (size_t)y * w;
    float* t = tmp + (size_t)y * w;
    for (int x = 0; x < w; ++x) {
      float m2 = s[reflect101(x - 2, w)], m1 = s[reflect101(x - 1, w)], c = s[x];
      float p1 = s[reflect101(x + 1, w)], p2 = s[reflect101(x + 2, w)];
      float r;
      if (x < row_body) {
        r = (m1 + p1) * k1;
        r = fmaf(k0, c, r);
        r = fmaf(k2, m2 + p2, r);
      } else {
        r = (k0 * c + (m1 + p1) * k1) + (m2 + p2) * k2;
      }
      t[x] = r;
    }
  }
  for (int y = 0; y < h; ++y) {
    const float* r0 = tmp + (size_t)reflect101(y - 2, h) * w;
    const float* r1 = tmp + (size_t)reflect101(y - 1, h) * w;
    const float* r2 = tmp + (size_t)y * w;
    const float* r3 = tmp + (size_t)reflect101(y + 1, h) * w;
    const float* r4 = tmp + (size_t)reflect101(y + 2, h) * w;
    float* d = dst + (size_t)y * w;
    for (int x = 0; x < col_body; ++x) {
      float r = r2[x] * k0;
      r = fmaf(k1, r1[x] + r3[x], r);
      r = fmaf(k2, r0[x] + r4[x], r);
      d[x] = r;
    }
    for (int x = col_body; x < w; ++x) d[x] = (k0 * r2[x] + (r1[x] + r3[x]) * k1) + (r0[x] + r4[x]) * k2;
  }
  free(tmp);
}

/* cv::getGaussianKernel(5, sigma, CV_32F) for the three sigmas the reference uses
 * (hessian.h:102,113  klt.h:119); bit patterns read from OpenCV 4.13. */
static void gauss_taps(double sigma, float* k0, float* k1, float* k2) {
  union { uint32_t u; float f; } a, b, c;
  if (sigma == 1.1) { c.u = 0x3d90edf6u; b.u = 0x3e7a53d4u; a.u = 0x3ebd3532u; }
  else if (sigma == 0.8) { c.u = 0x3cb3a5ccu; b.u = 0x3e69ff17u; a.u = 0x3eff8c30u; }
  else if (sigma == 0.6) { c.u = 0x3b282ed8u; b.u = 0x3e297f46u; a.u = 0x3f29efffu; }
  else {
    /* generic: exp(-x^2/(2 s^2)) normalised, double then float */
    double e1 = exp(-0.5 / (sigma * sigma)), e2 = exp(-2.0 / (sigma * sigma));
    double s = 1.0 + 2 * e1 + 2 * e2;
    a.f = (float)(1.0 / s); b.f = (float)(e1 / s); c.f = (float)(e2 / s);
  }
  *k0 = a.f; *k1 = b.f; *k2 = c.f;
}

void orc_gauss5_sigma(const float* src, int w, int h, double sigma, float* dst) {
  float k0, k1, k2;
  gauss_taps(sigma, &k0, &k1, &k2);
  orc_gauss5(src, w, h, k0, k1, k2, dst);
}

/* cv::pyrDown, BORDER_REFLECT_101 (hessian.h:112).  As with GaussianBlur, OpenCV's vector bodies and
 * scalar paths order the same sum differently, and the split is a function of the width alone (probed
 * against cv2 4.13 column by column):
 *   horizontal pass  border columns (0 and >= width0 = min((w-3)/2+1, dw)) go through the border table,
 *                    columns 1..4*((width0-1)/4) through the 4-wide vector body, the rest of the middle
 *                    through the scalar loop; table and scalar loop compute ((6c + 4(m1+p1)) + m2) + p2,
 *                    the vector body ((m2+p2) + 4(m1+p1)) + 6c
 *   vertical pass    columns x >= (dw/4)*4 are scalar: (((6 r2 + 4(r1+r3)) + r0) + r4) / 256,
 *                    the vector body (4((r1+r3)+r2) + ((r0+r4)+(r2+r2))) / 256 */
int orc_pyrdown_hbody(int w) { /* last column of the horizontal vector body (columns 1..n) */
  int dw = (w + 1) / 2, width0 = (w - 3) / 2 + 1;
  if (w < 3) width0 = 0;
  if (width0 > dw) width0 = dw;
  return width0 >= 1 ? ((width0 - 1) / 4) * 4 : 0;
}
void orc_pyrdown(const float* src, int w, int h, float* dst) {
  int dw = (w + 1) / 2, dh = (h + 1) / 2;
  float* tmp = (float*)malloc(sizeof(float) * (size_t)dw * h);
  const int hbody = orc_pyrdown_hbody(w), vbody = (dw / 4) * 4;
  for (int y = 0; y < h; ++y) {
    const float* s = src + (size_t)y * w;
    float* t = tmp + (size_t)y * dw;
    for (int x = 0; x < dw; ++x) {
      float m2 = s[reflect101(2 * x - 2, w)], m1 = s[reflect101(2 * x - 1, w)], c = s[reflect101(2 * x, w)];
      float p1 = s[reflect101(2 * x + 1, w)], p2 = s[reflect101(2 * x + 2, w)];
      if (x >= 1 && x <= hbody) t[x] = ((m2 + p2) + (m1 + p1) * 4.f) + c * 6.f;
      else t[x] = ((c * 6.f + (m1 + p1) * 4.f) + m2) + p2;
    }
  }
  for (int y = 0; y < dh; ++y) {
    const float* r0 = tmp + (size_t)reflect101(2 * y - 2, h) * dw;
    const float* r1 = tmp + (size_t)reflect101(2 * y - 1, h) * dw;
    const float* r2 = tmp + (size_t)reflect101(2 * y, h) * dw;
    const float* r3 = tmp + (size_t)reflect101(2 * y + 1, h) * dw;
    const float* r4 = tmp + (size_t)reflect101(2 * y + 2, h) * dw;
    float* d = dst + (size_t)y * dw;
    for (int x = 0; x < vbody; ++x)
      d[x] = (((r1[x] + r3[x]) + r2[x]) * 4.f + ((r0[x] + r4[x]) + (r2[x] + r2[x]))) * (1.f / 256.f);
    for (int x = vbody; x < dw; ++x)
      d[x] = (((r2[x] * 6.f + (r1[x] + r3[x]) * 4.f) + r0[x]) + r4[x]) * (1.f / 256.f);
  }
  free(tmp);
}

/* cv::Sobel(.., CV_SCHARR, scale 1/32) for dx and dy (klt.h:105-106). */
void orc_scharr(const float* src, int w, int h, float* gx, float* gy) {
  const float k3 = 3.f / 32.f, k10 = 10.f / 32.f;
  float* tx = (float*)malloc(sizeof(float) * (size_t)w * h);
  float* ty = (float*)malloc(sizeof(float) * (size_t)w * h);
  for (int y = 0; y < h; ++y) {
    const float* s = src + (size_t)y * w;
    for (int x = 0; x < w; ++x) {
      float m = s[reflect101(x - 1, w)], c = s[x], p = s[reflect101(x + 1, w)];
      tx[(size_t)y * w + x] = p - m;
      /* the row filter's scalar tail (last column of an odd-width image) fuses the other product */
      ty[(size_t)y * w + x] = ((w & 1) && x == w - 1) ? fmaf(k3, m + p, c * k10) : fmaf(k10, c, (m + p) * k3);
    }
  }
  for (int y = 0; y < h; ++y) {
    size_t u = (size_t)reflect101(y - 1, h) * w, c = (size_t)y * w, d = (size_t)reflect101(y + 1, h) * w;
    for (int x = 0; x < w; ++x) {
      gx[c + x] = fmaf(k3, tx[u + x] + tx[d + x], tx[c + x] * k10);
      gy[c + x] = ty[d + x] - ty[u + x];
    }
  }
  free(tx);
  free(ty);
}

/* cv::getRectSubPix for CV_32F -> CV_32F, n columns x m rows around (cx,cy).
 * Interior: 4-tap bilinear as one multiply and three FMAs.  OpenCV's border path replicates
 * edge pixels but evaluates overflow columns with a 2-tap vertical form, overflow rows with a
 * 2-tap horizontal form, and -- its one irregularity -- reads column w-2 for right-overflow
 * columns of top-overflow rows.  All of it is reproduced here. */
void orc_rect_subpix(const float* img, int w, int h, int n, int m, float cx, float cy, float* dst,
                     int dst_pitch) {
  cx = cx - (float)(n - 1) * 0.5f;
  cy = cy - (float)(m - 1) * 0.5f;
  int ix = (int)floorf(cx), iy = (int)floorf(cy);
  float a = cx - (float)ix, b = cy - (float)iy;
  float a1 = 1.f - a, b1 = 1.f - b;
  float a11 = a1 * b1, a12 = a * b1, a21 = a1 * b, a22 = a * b;
  for (int i = 0; i < m; ++i) {
    int Y = iy + i;
    int yin = (Y >= 0 && Y + 1 <= h - 1);
    int y0 = clampi(Y, 0, h - 1), y1 = clampi(Y + 1, 0, h - 1);
    for (int j = 0; j < n; ++j) {
      int X = ix + j;
      int xin = (X >= 0 && X + 1 <= w - 1);
      int x0 = clampi(X, 0, w - 1), x1 = clampi(X + 1, 0, w - 1);
      float v;
      if (xin && yin) {
        float s00 = img[(size_t)y0 * w + x0], s01 = img[(size_t)y0 * w + x1];
        float s10 = img[(size_t)y1 * w + x0], s11 = img[(size_t)y1 * w + x1];
        v = fmaf(s11, a22, fmaf(s10, a21, fmaf(s01, a12, s00 * a11)));
      } else if (yin) {
        v = fmaf(img[(size_t)y1 * w + x0], b, img[(size_t)y0 * w + x0] * b1);
      } else if (xin) {
        v = fmaf(img[(size_t)y0 * w + x1], a, img[(size_t)y0 * w + x0] * a1);
      } else {
        int xq = x0;
        if (Y < 0 && X >= w - 1 && w >= 2) xq = w - 2; /* OpenCV top-right quirk */
        float s = img[(size_t)y0 * w + xq];
        v = fmaf(s, b, s * b1);
      }
      dst[i * dst_pitch + j] = v;
    }
  }
}

/* ------------------------------------------------------------------ pyramids */

static void plane_alloc(orc_plane* p, int w, int h) {
  p->w = w; p->h = h;
  p->data = (float*)malloc(sizeof(float) * (size_t)w * h);
}

orc_pyr* orc_pyr_build(const uint8_t* bgr, int w, int h, size_t stride, int depth, int flavor) {
  if (depth < 1 || depth > ORC_MAX_LEVELS) return NULL;
  orc_pyr* p = (orc_pyr*)calloc(1, sizeof(orc_pyr));
  p->depth = depth;
  p->flavor = flavor;
  plane_alloc(&p->img[0], w, h);
  gray_f32(bgr, w, h, stride, p->img[0].data);
  if (flavor == ORC_FLAVOR_HESSIAN) {
    orc_gauss5_sigma(p->img[0].data, w, h, 1.1, p->img[0].data); /* hessian.h:102 */
  } else if (flavor == ORC_FLAVOR_KLT) {
    plane_alloc(&p->gx[0], w, h);
    plane_alloc(&p->gy[0], w, h);
    orc_scharr(p->img[0].data, w, h, p->gx[0].data, p->gy[0].data); /* klt.h:105-106 */
  }
  for (int i = 1; i < depth; ++i) {
    int pw = p->img[i - 1].w, ph = p->img[i - 1].h;
    int cw = (pw + 1) / 2, ch = (ph + 1) / 2; /* hessian.h:108 */
    plane_alloc(&p->img[i], cw, ch);
    orc_pyrdown(p->img[i - 1].data, pw, ph, p->img[i].data);
    if (flavor == ORC_FLAVOR_HESSIAN) {
      orc_gauss5_sigma(p->img[i].data, cw, ch, 0.8, p->img[i].data); /* hessian.h:113 */
    } else if (flavor == ORC_FLAVOR_KLT) {
      orc_gauss5_sigma(p->img[i].data, cw, ch, 0.6, p->img[i].data); /* klt.h:119 */
      plane_alloc(&p->gx[i], cw, ch);
      plane_alloc(&p->gy[i], cw, ch);
      /* klt.h:120-121 Sobel results are overwritten by klt.h:123-124 */
      orc_pyrdown(p->gx[i - 1].data, pw, ph, p->gx[i].data);
      orc_pyrdown(p->gy[i - 1].data, pw, ph, p->gy[i].data);
      for (size_t k = 0; k < (size_t)cw * ch; ++k) {
        p->gx[i].data[k] *= 2.f;
        p->gy[i].data[k] *= 2.f;
      }
    }
  }
  return p;
}

/* Wrap caller-provided planes (dense, row pitch == w) as a pyramid: lets tests run the trackers on
 * planes produced by the real OpenCV.  planes[level*3 + {0,1,2}] = img, gx, gy (gx/gy may be NULL). */
orc_pyr* orc_pyr_from_planes(int depth, int flavor, const int* ws, const int* hs, const float* const* planes) {
  if (depth < 1 || depth > ORC_MAX_LEVELS) return NULL;
  orc_pyr* p = (orc_pyr*)calloc(1, sizeof(orc_pyr));
  p->depth = depth;
  p->flavor = flavor;
  for (int i = 0; i < depth; ++i) {
    orc_plane* dst[3] = {&p->img[i], &p->gx[i], &p->gy[i]};
    for (int k = 0; k < 3; ++k) {
      const float* src = planes[i * 3 + k];
      if (!src) continue;
      plane_alloc(dst[k], ws[i], hs[i]);
      memcpy(dst[k]->data, src, sizeof(float) * (size_t)ws[i] * hs[i]);
    }
  }
  return p;
}

void orc_pyr_free(orc_pyr* p) {
  if (!p) return;
  for (int i = 0; i < ORC_MAX_LEVELS; ++i) {
    free(p->img[i].data);
    free(p->gx[i].data);
    free(p->gy[i].data);
  }
  free(p);
}
int orc_pyr_depth(const orc_pyr* p) { return p->depth; }
int orc_pyr_w(const orc_pyr* p, int level) { return p->img[level].w; }
int orc_pyr_h(const orc_pyr* p, int level) { return p->img[level].h; }
const float* orc_pyr_plane(const orc_pyr* p, int level, int plane) {
  return plane == 0 ? p->img[level].data : (plane == 1 ? p->gx[level].data : p->gy[level].data);
}

/* ------------------------------------------------------------------ patch statistics */

/* hessian.h:85-91: mean = sum/len, sumsq = sum(d*d)/len over ALL 169 entries (zeros included),
 * float accumulators, in the declared lane/tree order. */
static void patch_stats_tree(const float* d, float* mean, float* sumsq, int adjacent) {
  float s[32], q[32];
  for (int l = 0; l < 32; ++l) s[l] = q[l] = 0.f;
  for (int i = 0; i < ORC_PLEN; ++i) {
    int l = i & 31;
    s[l] = s[l] + d[i];
    q[l] = fmaf(d[i], d[i], q[l]);
  }
  *mean = (adjacent ? tree32_adj(s) : tree32(s)) / (float)ORC_PLEN;
  *sumsq = (adjacent ? tree32_adj(q) : tree32(q)) / (float)ORC_PLEN;
}
static void patch_stats(const float* d, float* mean, float* sumsq) { patch_stats_tree(d, mean, sumsq, 0); }  /* klt.h, brute.h */

/* ------------------------------------------------------------------ P1 HessianTracker */

static float g_mask[ORC_PLEN];
static int g_mask_ready = 0;
static const float* mask13(void) {
  if (!g_mask_ready) {
#ifdef _OPENMP
#pragma omp critical(orc_mask_init)
#endif
    {
      if (!g_mask_ready) { orc_mask13(g_mask); g_mask_ready = 1; }
    }
  }
  return g_mask;
}

/* hessian.h:54-93 */
static void hes_get_patch(const orc_plane* g, float px, float py, float* data, float* mean,
                          float* sumsq) {
  const int n = ORC_PATCH;
  memset(data, 0, sizeof(float) * ORC_PLEN);
  int rx = 0, ry = 0, rw = n, rh = n;
  if (px < 0.5 * n) {                       /* :65 (double compare) */
    int d = (int)((0.5 * n - px) + 0.9999); /* :66 */
    px = (float)(px + 0.5 * d);             /* :67 */
    rx = d; rw = n - d;                     /* :68 */
  }
  if (py < 0.5 * n) {                       /* :71 */
    int d = (int)(0.5 * n - py);            /* :72 (no +0.9999: asymmetric, preserved) */
    py = (float)(py + 0.5 * d);
    ry = d; rh = n - d;
  }
  if (rw > 0 && rh > 0)
    orc_rect_subpix(g->data, g->w, g->h, rw, rh, px, py, data + rx + ry * n, n); /* :77-83 */
  patch_stats_tree(data, mean, sumsq, 1);
}

void orc_hes_get_patch(const orc_pyr* p, int level, float x, float y, float* data169, float* mean,
                       float* sumsq) {
  hes_get_patch(&p->img[level], x, y, data169, mean, sumsq);
}

/* hessian.h:129-141 */
float orc_hes_score(const float* p1, float mean1, float sumsq1, const float* p2, float mean2,
                    float sumsq2) {
  const float* mask = mask13();
  float alpha = sqrtf(sumsq1 / sumsq2);
  float beta = mean1 - alpha * mean2;
  float s[32];
  for (int l = 0; l < 32; ++l) s[l] = 0.f;
  for (int i = 0; i < ORC_PLEN; ++i) {
    if (p1[i] == 0 || p2[i] == 0) continue;
    float diff = fmaf(-p2[i], alpha, p1[i]) - beta;
    diff = diff * diff;
    s[i & 31] = fmaf(diff, mask[i], s[i & 31]);
  }
  return tree32_adj(s);
}

/* hessian.h:147-172.  out6 = dx,dy,dxx,dxy,dyx,dyy (each rounded to float as the reference
 * stores them through float*).  Returns sad0. */
static float hes_brute_hessian(const orc_plane* g, const float* patch, float pm, float pq, float x,
                               float y, float* out6, orc_counters* c) {
  const double h = 0.02;
  float xm = (float)(x - h), xp = (float)(x + h), ym = (float)(y - h), yp = (float)(y + h);
  const float px[6] = {x, xm, x, xp, x, xp};
  const float py[6] = {y, y, ym, y, yp, yp};
  double s[6];
  float buf[ORC_PLEN], m, q;
  for (int k = 0; k < 6; ++k) {
    hes_get_patch(g, px[k], py[k], buf, &m, &q);
    s[k] = orc_hes_score(patch, pm, pq, buf, m, q);
  }
  if (c) c->patches += 6;
  double sad0 = s[0], sadn1x = s[1], sadn1y = s[2], sadp1x = s[3], sadp1y = s[4], sadxy = s[5];
  out6[0] = (float)(0.5 * (sadp1x - sadn1x) / h);
  out6[1] = (float)(0.5 * (sadp1y - sadn1y) / h);
  out6[2] = (float)(((sadp1x - sad0) / h - (sad0 - sadn1x) / h) / h);
  out6[5] = (float)(((sadp1y - sad0) / h - (sad0 - sadn1y) / h) / h);
  out6[3] = (float)(((sadxy - sadp1y) / h - (sadp1x - sad0) / h) / h);
  out6[4] = (float)(((sadxy - sadp1x) / h - (sadp1y - sad0) / h) / h);
  return (float)sad0;
}

float orc_hes_brute_hessian(const orc_pyr* p, int level, const float* patch, float mean,
                            float sumsq, float x, float y, float out6[6]) {
  return hes_brute_hessian(&p->img[level], patch, mean, sumsq, x, y, out6, NULL);
}

/* The Newton update shared by hessian.h:209-233 and klt.h:360-392.
 * H.inverse()*g follows Eigen's 2x2 closed form on doubles built from the float entries. */
static inline void newton_step(const float* d6, float* dx_out, float* dy_out) {
  double H00 = d6[2], H01 = d6[3], H10 = d6[4], H11 = d6[5];
  double g0 = d6[0], g1 = d6[1];
  double det = H00 * H11 - H10 * H01;
  double invdet = 1.0 / det;
  double i00 = H11 * invdet, i10 = -H10 * invdet, i01 = -H01 * invdet, i11 = H00 * invdet;
  double j0 = i00 * g0 + i01 * g1;
  double j1 = i10 * g0 + i11 * g1;
  float dx = (float)(-j0), dy = (float)(-j1);
  if ((dx * dx + dy * dy) > 1) {                 /* :224 */
    dx /= sqrtf(dx * dx + dy * dy);              /* :225 */
    dy /= sqrtf(dx * dx + dy * dy);              /* :226 uses the UPDATED dx (quirk, preserved) */
  }
  *dx_out = dx;
  *dy_out = dy;
}

/* hessian.h:185-241 */
static int hes_track(const orc_plane* g, const float* patch, float pm, float pq, float threshold,
                     int max_iterations, float* px, float* py, orc_counters* c) {
  float x = *px, y = *py;
  const float margin = 0.01f;
  for (int it = 0; it < max_iterations; ++it) {
    if (x < margin || y < margin || (x + margin) > (float)g->w || (y + margin) > (float)g->h) {
      *px = x; *py = y;
      return ORC_OUT_OF_BOUNDS;
    }
    float d6[6];
    hes_brute_hessian(g, patch, pm, pq, x, y, d6, c);
    if (c) c->newton_steps += 1;
    float dx, dy;
    newton_step(d6, &dx, &dy);
    x += fmaxf(-1.f, fminf(1.f, dx));
    y += fmaxf(-1.f, fminf(1.f, dy));
    if (fabsf(dx) < threshold && fabsf(dy) < threshold) break;
  }
  *px = x; *py = y;
  return ORC_OK;
}

/* GetPatches (hessian.h:175-183) on tmpl_pyr at (tx,ty) + TrackFeature (hessian.h:243-264) on
 * `search` from the seed (*x,*y).  (*x,*y) is written only on full success (:262). */
int orc_hes_track_feature(const orc_pyr* tmpl_pyr, float tx, float ty, const orc_pyr* search,
                          int levels, float thr, int maxit, float* x, float* y, orc_counters* c) {
  int lv = levels < tmpl_pyr->depth ? levels : tmpl_pyr->depth; /* :176 */
  float patches[ORC_MAX_LEVELS][ORC_PLEN], pm[ORC_MAX_LEVELS], pq[ORC_MAX_LEVELS];
  float qx = tx, qy = ty;
  for (int i = 0; i < lv; ++i) {
    hes_get_patch(&tmpl_pyr->img[i], qx, qy, patches[i], &pm[i], &pq[i]);
    qx *= 0.5f; qy *= 0.5f;
  }
  if (c) c->patches += lv;
  int lvls = search->depth < lv ? search->depth : lv; /* :249 */
  float scale = (float)(1. / (1 << (lvls - 1)));
  float px = *x * scale, py = *y * scale;              /* :251 (power of two: exact) */
  for (int i = lvls - 1; i > 0; --i) {
    int st = hes_track(&search->img[i], patches[i], pm[i], pq[i], thr, maxit, &px, &py, c);
    if (st != ORC_OK) return st;
    px *= 2.f; py *= 2.f;
  }
  int st = hes_track(&search->img[0], patches[0], pm[0], pq[0], thr, maxit, &px, &py, c);
  if (st != ORC_OK) return st;
  *x = px; *y = py;
  return ORC_OK;
}

/* matcher.cpp:173-206 applied to n features. */
int orc_hes_track_fb(const orc_pyr* from, const orc_pyr* to, int n, const float* from_xy,
                     float* to_xy, const int* levels, float thr, int maxit, double fb_max,
                     float* back_xy, int* st_fwd, int* st_bwd, uint8_t* accepted,
                     orc_counters* c, int nthreads) {
  int64_t steps = 0, patches = 0;
  int nacc = 0;
  mask13();
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads) reduction(+ : steps, patches, nacc)
#endif
  for (int i = 0; i < n; ++i) {
    orc_counters lc = {0, 0};
    float fx = from_xy[2 * i], fy = from_xy[2 * i + 1];
    float tx = to_xy[2 * i], ty = to_xy[2 * i + 1];
    int lv = levels ? levels[i] : 3;
    int s1 = orc_hes_track_feature(from, fx, fy, to, lv, thr, maxit, &tx, &ty, &lc); /* :175-176 */
    float bx = fx, by = fy;                                                           /* :181 */
    int s2 = orc_hes_track_feature(to, tx, ty, from, lv, thr, maxit, &bx, &by, &lc);  /* :180-182 */
    int ok = !(s1 || s2);                                                             /* :192 */
    if (ok) {
      float ddx = fx - bx, ddy = fy - by;
      double nrm = sqrt((double)ddx * ddx + (double)ddy * ddy);                       /* cv::norm */
      if (nrm > fb_max) ok = 0;                                               /* :201 */
    }
    to_xy[2 * i] = tx; to_xy[2 * i + 1] = ty;
    if (back_xy) { back_xy[2 * i] = bx; back_xy[2 * i + 1] = by; }
    if (st_fwd) st_fwd[i] = s1;
    if (st_bwd) st_bwd[i] = s2;
    if (accepted) accepted[i] = (uint8_t)ok;
    steps += lc.newton_steps; patches += lc.patches; nacc += ok;
  }
  if (c) { c->newton_steps += steps; c->patches += patches; }
  return nacc;
}

/* ------------------------------------------------------------------ corner seeding */

/* cv::cornerMinEigenVal(gray8, blockSize 3, ksize 3) as cv::goodFeaturesToTrack calls it (matcher.cpp:125 ->
 * featureselect.cpp -> corner.cpp cornerEigenValsVecs), BORDER_REFLECT_101 everywhere:
 *   Dx = Sobel(1,0) * s, Dy = Sobel(0,1) * s with s = 1/(4*3*255); OpenCV scales the COLUMN kernel of Dx
 *        and the ROW kernel of Dy:  Dx = fma(r(-1)+r(+1), s, r(0)*2s), r = g(x+1)-g(x-1)
 *                                   Dy = q(+1)-q(-1),  q = fma(g(x+1), s, fma(g(x), 2s, g(x-1)*s))
 *   cov = (Dx*Dx, Dx*Dy, Dy*Dy) in float;  3x3 box sums accumulate in DOUBLE: row sums (c(-1)+c(0))+c(+1),
 *        then OpenCV's running column sum from the top of the image: S = (0+R(-1))+R(0); D(y) = S+R(y+1);
 *        S = D(y)-R(y-1)  -- history dependent in the last bit of the double, reproduced here;
 *   eig = (a+c) - sqrt((a-c)*(a-c) + b*b), a = 0.5*Sxx, b = Sxy, c = 0.5*Syy, float, no FMA.
 * Probed bit-for-bit against cv2 4.13 for widths that are a multiple of 16 (the right-most W mod 16 columns
 * of cv2 run scalar code whose Dy row filter is not fused; the SIMD formula is used uniformly here). */
void orc_min_eigen_val(const uint8_t* gray, int w, int h, float* eig) {
  const double sd = 1.0 / (4.0 * 3.0 * 255.0);
  const float k0 = (float)sd, k1 = (float)(2.0 * sd);
  float* r = (float*)malloc(sizeof(float) * (size_t)w * (h + 2));  /* horizontal difference rows -1..h */
  float* q = (float*)malloc(sizeof(float) * (size_t)w * (h + 2));  /* horizontally smoothed rows -1..h */
  for (int yy = -1; yy <= h; ++yy) {
    const uint8_t* g = gray + (size_t)reflect101(yy, h) * w;
    float* rr = r + (size_t)(yy + 1) * w;
    float* qq = q + (size_t)(yy + 1) * w;
    for (int x = 0; x < w; ++x) {
      const float a = (float)g[reflect101(x - 1, w)], b = (float)g[x], c = (float)g[reflect101(x + 1, w)];
      rr[x] = c - a;
      qq[x] = fmaf(c, k0, fmaf(b, k1, a * k0));
    }
  }
  float* cov = (float*)malloc(sizeof(float) * 3 * (size_t)w * h);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const float* r0 = r + (size_t)y * w;  /* row y-1 */
      const float* q0 = q + (size_t)y * w;
      const float dx = fmaf(r0[x] + r0[2 * (size_t)w + x], k0, r0[w + x] * k1);
      const float dy = q0[2 * (size_t)w + x] - q0[x];
      float* c = cov + 3 * ((size_t)y * w + x);
      c[0] = dx * dx; c[1] = dx * dy; c[2] = dy * dy;
    }
  free(r); free(q);
  /* box filter: double row sums R(y)[x][ch] for rows -1..h, then the running column sum */
  double* R = (double*)malloc(sizeof(double) * 3 * (size_t)w * (h + 2));
  for (int yy = -1; yy <= h; ++yy) {
    const float* c = cov + 3 * (size_t)reflect101(yy, h) * w;
    double* o = R + 3 * (size_t)(yy + 1) * w;
    for (int x = 0; x < w; ++x)
      for (int ch = 0; ch < 3; ++ch)
        o[3 * x + ch] = ((double)c[3 * reflect101(x - 1, w) + ch] + (double)c[3 * x + ch]) + (double)c[3 * reflect101(x + 1, w) + ch];
  }
  double* S = (double*)calloc(3 * (size_t)w, sizeof(double));
  for (size_t i = 0; i < 3 * (size_t)w; ++i) S[i] = (0.0 + R[i]) + R[3 * (size_t)w + i];
  for (int y = 0; y < h; ++y) {
    const double* Rp = R + 3 * (size_t)(y + 2) * w;  /* row y+1 */
    const double* Rm = R + 3 * (size_t)y * w;        /* row y-1 */
    for (int x = 0; x < w; ++x) {
      float box[3];
      for (int ch = 0; ch < 3; ++ch) {
        const double s = S[3 * x + ch] + Rp[3 * x + ch];
        box[ch] = (float)s;
        S[3 * x + ch] = s - Rm[3 * x + ch];
      }
      const float a = box[0] * 0.5f, b = box[1], c = box[2] * 0.5f;
      const float t = a - c;
      eig[(size_t)y * w + x] = (a + c) - sqrtf(t * t + b * b);
    }
  }
  free(S); free(R); free(cov);
}

typedef struct { float v; int ofs; } orc_cand;
static int cand_cmp(const void* pa, const void* pb) {
  /* featureselect.cpp greaterThanPtr: larger value first, ties by HIGHER address first */
  const orc_cand* a = (const orc_cand*)pa; const orc_cand* b = (const orc_cand*)pb;
  if (a->v > b->v) return -1;
  if (a->v < b->v) return 1;
  return a->ofs > b->ofs ? -1 : (a->ofs < b->ofs ? 1 : 0);
}

/* cv::goodFeaturesToTrack(gray, corners, max_corners, quality, min_distance) with its defaults (blockSize 3,
 * min-eigenvalue response, no mask) -- matcher.cpp:123-130 on the RGB2GRAY image of matcher.cpp:313.
 * corners_xy holds max_corners (x,y) pairs (max_corners must be > 0); returns the number found.  eig_out
 * (w*h, may be NULL) receives the response map, max_out its maximum. */
int orc_good_features(const uint8_t* bgr, int w, int h, size_t stride, int max_corners, double quality,
                      double min_distance, float* corners_xy, float* eig_out, float* max_out) {
  uint8_t* gray = (uint8_t*)malloc((size_t)w * h);
  orc_gray_u8(bgr, w, h, stride, gray);
  float* eig = eig_out ? eig_out : (float*)malloc(sizeof(float) * (size_t)w * h);
  orc_min_eigen_val(gray, w, h, eig);
  free(gray);
  float mx = eig[0];                                   /* minMaxLoc */
  for (size_t i = 1; i < (size_t)w * h; ++i) if (eig[i] > mx) mx = eig[i];
  if (max_out) *max_out = mx;
  const float thr = (float)((double)mx * quality);     /* threshold(..., THRESH_TOZERO) compares in float */
  /* local maxima of the thresholded map under a 3x3 dilation, interior pixels only (featureselect.cpp) */
  orc_cand* cand = (orc_cand*)malloc(sizeof(orc_cand) * (size_t)w * h);
  size_t nc = 0;
#define TZ(v) ((v) > thr ? (v) : 0.f)
  for (int y = 1; y < h - 1; ++y)
    for (int x = 1; x < w - 1; ++x) {
      const float v = TZ(eig[(size_t)y * w + x]);
      if (v == 0.f) continue;
      float m = v;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const float n = TZ(eig[(size_t)(y + dy) * w + x + dx]);
          if (n > m) m = n;
        }
      if (v == m) { cand[nc].v = v; cand[nc].ofs = y * w + x; ++nc; }
    }
#undef TZ
  qsort(cand, nc, sizeof(orc_cand), cand_cmp);
  int n = 0;
  if (min_distance >= 1) {
    const int cell = (int)lrint(min_distance);         /* cvRound */
    const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
    int* head = (int*)malloc(sizeof(int) * (size_t)gw * gh);  /* per cell: linked list through next[] */
    int* next = (int*)malloc(sizeof(int) * (size_t)max_corners);
    for (int i = 0; i < gw * gh; ++i) head[i] = -1;
    const double md2 = min_distance * min_distance;
    for (size_t i = 0; i < nc && n < max_corners; ++i) {
      const int y = cand[i].ofs / w, x = cand[i].ofs - y * w;
      const int xc = x / cell, yc = y / cell;
      const int x1 = xc > 0 ? xc - 1 : 0, y1 = yc > 0 ? yc - 1 : 0;
      const int x2 = xc + 1 < gw ? xc + 1 : gw - 1, y2 = yc + 1 < gh ? yc + 1 : gh - 1;
      int good = 1;
      for (int yy = y1; yy <= y2 && good; ++yy)
        for (int xx = x1; xx <= x2 && good; ++xx)
          for (int j = head[yy * gw + xx]; j >= 0; j = next[j]) {
            const float dx = (float)x - corners_xy[2 * j], dy = (float)y - corners_xy[2 * j + 1];
            if ((double)(dx * dx + dy * dy) < md2) { good = 0; break; }
          }
      if (!good) continue;
      corners_xy[2 * n] = (float)x; corners_xy[2 * n + 1] = (float)y;
      next[n] = head[yc * gw + xc]; head[yc * gw + xc] = n;
      ++n;
    }
    free(head); free(next);
  } else {
    for (size_t i = 0; i < nc && n < max_corners; ++i, ++n) {
      const int y = cand[i].ofs / w, x = cand[i].ofs - y * w;
      corners_xy[2 * n] = (float)x; corners_xy[2 * n + 1] = (float)y;
    }
  }
  free(cand);
  if (!eig_out) free(eig);
  return n;
}

/* ------------------------------------------------------------------ seeding the search */

/* Frame::Project (localmap.cpp:18-26 -> project.h:11-54) in double, operation for operation; the quaternion
 * product is Eigen 3.2's _transformVector: uv = 2*(q.vec x v); v + w*uv + q.vec x uv.  rot = [x,y,z,w].
 * Returns 0 when the point is behind the lens (project.h:27). */
int orc_project(const double* rot, const double* trans, const double* k, const double* pt, double* out2) {
  const double vx = pt[0] - trans[0] * pt[3], vy = pt[1] - trans[1] * pt[3], vz = pt[2] - trans[2] * pt[3];
  const double qx = rot[0], qy = rot[1], qz = rot[2], qw = rot[3];
  double ux = qy * vz - qz * vy, uy = qz * vx - qx * vz, uz = qx * vy - qy * vx;
  ux += ux; uy += uy; uz += uz;
  const double px = (vx + qw * ux) + (qy * uz - qz * uy);
  const double py = (vy + qw * uy) + (qz * ux - qx * uz);
  const double pz = (vz + qw * uz) + (qx * uy - qy * ux);
  if (pz < 0.001 * pt[3]) return 0;
  double xp = px / pz, yp = py / pz;
  const double r2 = xp * xp + yp * yp;
  const double distort = 1.0 + r2 * (k[0] + r2 * (k[1] + r2 * k[2]));
  xp *= distort; yp *= distort;
  xp *= k[3]; yp *= k[4];
  xp += k[5]; yp += k[6];
  out2[0] = xp; out2[1] = yp;
  return 1;
}

/* The per-feature head of the FindMatches loop (matcher.cpp:224-245): levels = uncertainty > 100 ? 6 : 3; the seed is
 * from_pt, or the projection of the map point when uncertainty < 100 and it projects; go = 0 when the seed is out of
 * bounds (note `>` on y, :243, preserved). */
void orc_seed_features(int n, const double* points4, const double* uncertainty, const double* rot, const double* trans,
                       const double* k, const float* from_xy, int cols, int rows, float* seed_xy, int32_t* levels,
                       uint8_t* go) {
  for (int i = 0; i < n; ++i) {
    float x = from_xy[2 * i], y = from_xy[2 * i + 1];
    levels[i] = uncertainty[i] > 100 ? 6 : 3;
    double p[2];
    if (uncertainty[i] < 100 && orc_project(rot, trans, k, points4 + 4 * i, p)) { x = (float)p[0]; y = (float)p[1]; }
    seed_xy[2 * i] = x; seed_xy[2 * i + 1] = y;
    go[i] = !(x < 0 || y < 0 || x >= (float)cols || y > (float)rows);
  }
}

/* ------------------------------------------------------------------ live capture format */

/* The integer YUYV -> BGR conversion of the V4L2 capture path, video.cpp:187-223: `bytes` of YUYV (4 bytes = 2 pixels)
 * to 3 bytes per pixel. */
void orc_yuyv_to_bgr(const uint8_t* in, size_t bytes, uint8_t* out) {
#define ORC_SAT(c) if ((c) & (~255)) { if ((c) < 0) (c) = 0; else (c) = 255; }
  for (size_t i = 0; i + 3 < bytes; i += 4, in += 4) {
    int y1 = in[0];
    int cb = ((in[1] - 128) * 454) >> 8;
    int cg = (in[1] - 128) * 88;
    int y2 = in[2];
    int cr = ((in[3] - 128) * 359) >> 8;
    cg = (cg + (in[3] - 128) * 183) >> 8;
    int r = y1 + cr, b = y1 + cb, g = y1 - cg;
    ORC_SAT(r); ORC_SAT(g); ORC_SAT(b);
    *out++ = (uint8_t)b; *out++ = (uint8_t)g; *out++ = (uint8_t)r;
    r = y2 + cr; b = y2 + cb; g = y2 - cg;
    ORC_SAT(r); ORC_SAT(g); ORC_SAT(b);
    *out++ = (uint8_t)b; *out++ = (uint8_t)g; *out++ = (uint8_t)r;
  }
#undef ORC_SAT
}

/* ------------------------------------------------------------------ P2 KLTTracker */

/* klt.h:59-96: three full 13x13 getRectSubPix calls, no edge clipping. */
static void klt_get_patch(const orc_pyr* p, int level, float x, float y, float* data, float* gx,
                          float* gy, float* mean, float* sumsq) {
  const orc_plane* g = &p->img[level];
  orc_rect_subpix(g->data, g->w, g->h, ORC_PATCH, ORC_PATCH, x, y, data, ORC_PATCH);
  if (gx) orc_rect_subpix(p->gx[level].data, g->w, g->h, ORC_PATCH, ORC_PATCH, x, y, gx, ORC_PATCH);
  if (gy) orc_rect_subpix(p->gy[level].data, g->w, g->h, ORC_PATCH, ORC_PATCH, x, y, gy, ORC_PATCH);
  patch_stats(data, mean, sumsq);
}

/* klt.h:139-149 */
static float klt_sad(const float* p1, const float* p2) {
  const float* mask = mask13();
  float s[32];
  for (int l = 0; l < 32; ++l) s[l] = 0.f;
  for (int i = 0; i < ORC_PLEN; ++i) {
    if (p1[i] == 0 || p2[i] == 0) continue;
    float diff = p1[i] - p2[i];
    s[i & 31] = fmaf(diff * diff, mask[i], s[i & 31]);
  }
  return tree32(s);
}

/* klt.h:181-204: forward differences, h = 0.01 */
static void klt_brute_hessian(const orc_pyr* p, int level, const float* patch, float x, float y,
                              float* out6, orc_counters* c) {
  const double h = 0.01;
  float x1 = (float)(x + h), x2 = (float)(x + 2 * h), y1 = (float)(y + h), y2 = (float)(y + 2 * h);
  const float px[6] = {x, x1, x, x2, x, x1};
  const float py[6] = {y, y, y1, y, y2, y1};
  double s[6];
  float buf[ORC_PLEN], m, q;
  for (int k = 0; k < 6; ++k) {
    klt_get_patch(p, level, px[k], py[k], buf, NULL, NULL, &m, &q);
    s[k] = klt_sad(patch, buf);
  }
  if (c) c->patches += 6;
  double sad0 = s[0], sadx = s[1], sady = s[2], sadxx = s[3], sadyy = s[4], sadxy = s[5];
  out6[0] = (float)((sadx - sad0) / h);
  out6[1] = (float)((sady - sad0) / h);
  out6[2] = (float)(((sadxx - sadx) / h - (sadx - sad0) / h) / h);
  out6[5] = (float)(((sadyy - sady) / h - (sady - sad0) / h) / h);
  out6[3] = (float)(((sadxy - sady) / h - (sadx - sad0) / h) / h);
  out6[4] = (float)(((sadxy - sadx) / h - (sady - sad0) / h) / h);
}

/* klt.h:286-353: the symmetric-KLT system at one point.  These values are computed by the
 * reference and then discarded (klt.h:355-380 overwrites dx,dy); they are restated because
 * north_star names them.  Sums use the declared lane/tree order; out24 layout in oracle.h. */
static void klt_system(const float* I, const float* Igx, const float* Igy, float Im, float Iq,
                       const float* J0, const float* Jgx, const float* Jgy, float Jm, float Jq,
                       float* out24) {
  const float* mask = mask13();
  float alpha = sqrtf(Iq / Jq);        /* :289 */
  float beta = Im - alpha * Jm;        /* :290 */
  float acc[16][32];
  memset(acc, 0, sizeof(acc));
  for (int i = 0; i < ORC_PLEN; ++i) {
    if (J0[i] == 0 || I[i] == 0) continue;  /* :303 */
    int l = i & 31;
    float m = mask[i];
    float Iv = I[i];
    float Jv = fmaf(J0[i], alpha, beta);    /* :308 */
    float gI0 = Igx[i], gI1 = Igy[i];
    float gJ0 = Jgx[i] * alpha, gJ1 = Jgy[i] * alpha; /* :313 */
    float diff = (Iv - Jv) * m;             /* :319 */
    /* A :316 */
    acc[0][l] = fmaf(gI0 * gI0, m, acc[0][l]); acc[1][l] = fmaf(gI0 * gI1, m, acc[1][l]);
    acc[2][l] = fmaf(gI1 * gI0, m, acc[2][l]); acc[3][l] = fmaf(gI1 * gI1, m, acc[3][l]);
    /* B :317 */
    acc[4][l] = fmaf(gI0 * gJ0, m, acc[4][l]); acc[5][l] = fmaf(gI0 * gJ1, m, acc[5][l]);
    acc[6][l] = fmaf(gI1 * gJ0, m, acc[6][l]); acc[7][l] = fmaf(gI1 * gJ1, m, acc[7][l]);
    /* C :318 */
    acc[8][l] = fmaf(gJ0 * gJ0, m, acc[8][l]); acc[9][l] = fmaf(gJ0 * gJ1, m, acc[9][l]);
    acc[10][l] = fmaf(gJ1 * gJ0, m, acc[10][l]); acc[11][l] = fmaf(gJ1 * gJ1, m, acc[11][l]);
    /* RS, VW :320-321 */
    acc[12][l] = fmaf(diff, gI0, acc[12][l]); acc[13][l] = fmaf(diff, gI1, acc[13][l]);
    acc[14][l] = fmaf(diff, gJ0, acc[14][l]); acc[15][l] = fmaf(diff, gJ1, acc[15][l]);
  }
  float A[4], B[4], C[4], RS[2], VW[2];
  for (int k = 0; k < 4; ++k) { A[k] = tree32(acc[k]); B[k] = tree32(acc[4 + k]); C[k] = tree32(acc[8 + k]); }
  RS[0] = tree32(acc[12]); RS[1] = tree32(acc[13]); VW[0] = tree32(acc[14]); VW[1] = tree32(acc[15]);
  const float lambda = .0001f;          /* :274 */
  /* Di = B^T inverse (:326), Eigen 2x2 closed form in float */
  float t00 = B[0], t01 = B[2], t10 = B[1], t11 = B[3];
  float det = t00 * t11 - t10 * t01;
  float invdet = 1.f / det;
  float D[4] = {t11 * invdet, -t01 * invdet, -t10 * invdet, t00 * invdet};
  float Al[4] = {A[0] + lambda, A[1], A[2], A[3] + lambda};
  float M[4] = {Al[0] * D[0] + Al[1] * D[2], Al[0] * D[1] + Al[1] * D[3],
                Al[2] * D[0] + Al[3] * D[2], Al[2] * D[1] + Al[3] * D[3]};
  float U[4] = {(M[0] * C[0] + M[1] * C[2]) - 0.5f * B[0], (M[0] * C[1] + M[1] * C[3]) - 0.5f * B[1],
                (M[2] * C[0] + M[3] * C[2]) - 0.5f * B[2], (M[2] * C[1] + M[3] * C[3]) - 0.5f * B[3]}; /* :330 */
  float e[2] = {(M[0] * VW[0] + M[1] * VW[1]) - 0.5f * RS[0],
                (M[2] * VW[0] + M[3] * VW[1]) - 0.5f * RS[1]};                                         /* :331 */
  /* U.lu().solve(e) (:343): 2x2 partial-pivot LU */
  float u00 = U[0], u01 = U[1], u10 = U[2], u11 = U[3], e0 = e[0], e1 = e[1];
  if (fabsf(u10) > fabsf(u00)) {
    float t;
    t = u00; u00 = u10; u10 = t;
    t = u01; u01 = u11; u11 = t;
    t = e0; e0 = e1; e1 = t;
  }
  float l = u10 / u00;
  float w11 = u11 - l * u01;
  float y1 = e1 - l * e0;
  float d1 = y1 / w11;
  float d0 = (e0 - u01 * d1) / u00;
  memcpy(out24 + 0, A, sizeof(A)); memcpy(out24 + 4, B, sizeof(B)); memcpy(out24 + 8, C, sizeof(C));
  out24[12] = RS[0]; out24[13] = RS[1]; out24[14] = VW[0]; out24[15] = VW[1];
  memcpy(out24 + 16, U, sizeof(U));
  out24[20] = e[0]; out24[21] = e[1]; out24[22] = d0; out24[23] = d1;
}

void orc_klt_system(const orc_pyr* tmpl_pyr, float tx, float ty, const orc_pyr* search, int level,
                    float x, float y, float out24[24]) {
  float I[ORC_PLEN], Igx[ORC_PLEN], Igy[ORC_PLEN], Im, Iq;
  float J[ORC_PLEN], Jgx[ORC_PLEN], Jgy[ORC_PLEN], Jm, Jq;
  klt_get_patch(tmpl_pyr, level, tx, ty, I, Igx, Igy, &Im, &Iq);
  klt_get_patch(search, level, x, y, J, Jgx, Jgy, &Jm, &Jq);
  klt_system(I, Igx, Igy, Im, Iq, J, Jgx, Jgy, Jm, Jq, out24);
}

/* klt.h:258-401 as written: the KLT step is computed and then replaced by the finite-difference
 * Newton step, so only the image plane influences the result. */
static int klt_track(const orc_pyr* p, int level, const float* patch, float threshold,
                     int max_iterations, float* px, float* py, orc_counters* c) {
  float x = *px, y = *py;
  const float margin = 0.1f;                    /* :272 */
  const orc_plane* g = &p->img[level];
  for (int it = 0; it < max_iterations; ++it) {
    if (x < margin || y < margin || (x + margin) > (float)g->w || (y + margin) > (float)g->h) {
      *px = x; *py = y;
      return ORC_OUT_OF_BOUNDS;
    }
    float d6[6];
    klt_brute_hessian(p, level, patch, x, y, d6, c);
    if (c) { c->newton_steps += 1; c->patches += 1; /* the np patch of :286 */ }
    float dx, dy;
    newton_step(d6, &dx, &dy);
    x += fmaxf(-1.f, fminf(1.f, dx));
    y += fmaxf(-1.f, fminf(1.f, dy));
    /* :392 compares float |dx| with the double threshold/10. */
    if ((double)fabsf(dx) < threshold / 10. && (double)fabsf(dy) < threshold / 10.) break;
  }
  *px = x; *py = y;
  return ORC_OK;
}

/* klt.h:249-256 + :403-424 (all levels of the stack, coarse threshold x50) */
int orc_klt_track_feature(const orc_pyr* tmpl_pyr, float tx, float ty, const orc_pyr* search,
                          float thr, int maxit, float* x, float* y, orc_counters* c) {
  int lvls = search->depth;
  if (tmpl_pyr->depth < lvls) lvls = tmpl_pyr->depth;
  float patches[ORC_MAX_LEVELS][ORC_PLEN], m, q;
  float qx = tx, qy = ty;
  for (int i = 0; i < lvls; ++i) {
    klt_get_patch(tmpl_pyr, i, qx, qy, patches[i], NULL, NULL, &m, &q);
    qx *= 0.5f; qy *= 0.5f;
  }
  if (c) c->patches += lvls;
  float scale = (float)(1. / (1 << (lvls - 1)));
  float px = *x * scale, py = *y * scale;
  for (int i = lvls - 1; i > 0; --i) {
    int st = klt_track(search, i, patches[i], thr * 50, maxit, &px, &py, c); /* :413 */
    if (st != ORC_OK) return st;
    px *= 2.f; py *= 2.f;
  }
  int st = klt_track(search, 0, patches[0], thr, maxit, &px, &py, c);
  if (st != ORC_OK) return st;
  *x = px; *y = py;
  return ORC_OK;
}

int orc_klt_track_fb(const orc_pyr* from, const orc_pyr* to, int n, const float* from_xy,
                     float* to_xy, float thr, int maxit, double fb_max, float* back_xy, int* st_fwd,
                     int* st_bwd, uint8_t* accepted, orc_counters* c, int nthreads) {
  int64_t steps = 0, patches = 0;
  int nacc = 0;
  mask13();
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads) reduction(+ : steps, patches, nacc)
#endif
  for (int i = 0; i < n; ++i) {
    orc_counters lc = {0, 0};
    float fx = from_xy[2 * i], fy = from_xy[2 * i + 1];
    float tx = to_xy[2 * i], ty = to_xy[2 * i + 1];
    int s1 = orc_klt_track_feature(from, fx, fy, to, thr, maxit, &tx, &ty, &lc);
    float bx = fx, by = fy;
    int s2 = orc_klt_track_feature(to, tx, ty, from, thr, maxit, &bx, &by, &lc);
    int ok = !(s1 || s2);
    if (ok) {
      float ddx = fx - bx, ddy = fy - by;
      if (sqrt((double)ddx * ddx + (double)ddy * ddy) > fb_max) ok = 0;
    }
    to_xy[2 * i] = tx; to_xy[2 * i + 1] = ty;
    if (back_xy) { back_xy[2 * i] = bx; back_xy[2 * i + 1] = by; }
    if (st_fwd) st_fwd[i] = s1;
    if (st_bwd) st_bwd[i] = s2;
    if (accepted) accepted[i] = (uint8_t)ok;
    steps += lc.newton_steps; patches += lc.patches; nacc += ok;
  }
  if (c) { c->newton_steps += steps; c->patches += patches; }
  return nacc;
}

/* ------------------------------------------------------------------ P3 BruteTracker */

/* brute.h:34-57 */
static void brute_get_patch(const orc_plane* g, float x, float y, float* data, float* mean,
                            float* sumsq) {
  orc_rect_subpix(g->data, g->w, g->h, ORC_PATCH, ORC_PATCH, x, y, data, ORC_PATCH);
  patch_stats(data, mean, sumsq);
}

/* brute.h:82-94: alpha/beta-normalised UNWEIGHTED SSD with zero-skip */
static float brute_sad(const float* p1, float m1, float q1, const float* p2, float m2, float q2) {
  float alpha = sqrtf(q1 / q2);
  float beta = m1 - alpha * m2;
  float s[32];
  for (int l = 0; l < 32; ++l) s[l] = 0.f;
  for (int i = 0; i < ORC_PLEN; ++i) {
    if (p1[i] == 0 || p2[i] == 0) continue;
    float diff = fmaf(-p2[i], alpha, p1[i]) - beta;
    s[i & 31] = fmaf(diff, diff, s[i & 31]);
  }
  return tree32(s);
}

/* brute.h:96-117.  Float loop counters (x += res accumulates rounding, preserved); x is the
 * outer loop; `if (sad > best) continue` => the LAST minimum wins. */
float orc_brute_search_best(const orc_pyr* search, int level, const float* patch, float mean,
                            float sumsq, float window, float res, float* px, float* py,
                            int64_t* npos) {
  const orc_plane* g = &search->img[level];
  float best = 1e6f;
  float cx = *px, cy = *py;
  float buf[ORC_PLEN], m, q;
  int64_t cnt = 0;
  for (float x = -window; x <= window; x += res) {
    for (float y = -window; y <= window; y += res) {
      brute_get_patch(g, cx + x, cy + y, buf, &m, &q);
      float sad = brute_sad(patch, mean, sumsq, buf, m, q);
      ++cnt;
      if (sad > best) continue;
      *px = cx + x;
      *py = cy + y;
      best = sad;
    }
  }
  if (npos) *npos += cnt;
  return best;
}

int orc_brute_track_feature(const orc_pyr* tmpl_pyr, float tx, float ty, const orc_pyr* search,
                            const float* coarse_sched, int n_coarse, const float* fine_sched,
                            int n_fine, float* x, float* y, float* best_sad, int64_t* npos) {
  int lvls = search->depth;
  if (tmpl_pyr->depth < lvls) lvls = tmpl_pyr->depth;
  const float margin = 13;                      /* brute.h:137 */
  int w0 = search->img[0].w, h0 = search->img[0].h;
  if (*x < margin || *y < margin || (*x + margin) > (float)w0 || (*y + margin) > (float)h0)
    return ORC_OUT_OF_BOUNDS;
  float patches[ORC_MAX_LEVELS][ORC_PLEN], pm[ORC_MAX_LEVELS], pq[ORC_MAX_LEVELS];
  float qx = tx, qy = ty;
  for (int i = 0; i < lvls; ++i) {              /* brute.h:120-127 */
    brute_get_patch(&tmpl_pyr->img[i], qx, qy, patches[i], &pm[i], &pq[i]);
    qx *= 0.5f; qy *= 0.5f;
  }
  float scale = (float)(1. / (1 << (lvls - 1)));
  float px = *x * scale, py = *y * scale;
  float sad = 0;
  for (int i = lvls - 1; i > 0; --i) {
    for (int k = 0; k < n_coarse; ++k)
      sad = orc_brute_search_best(search, i, patches[i], pm[i], pq[i], coarse_sched[2 * k],
                                  coarse_sched[2 * k + 1], &px, &py, npos);
    if (sad > 100) {                            /* :149 */
      if (best_sad) *best_sad = sad;
      return ORC_OUT_OF_BOUNDS;
    }
    px *= 2.f; py *= 2.f;
  }
  for (int k = 0; k < n_fine; ++k)
    sad = orc_brute_search_best(search, 0, patches[0], pm[0], pq[0], fine_sched[2 * k],
                                fine_sched[2 * k + 1], &px, &py, npos);
  if (best_sad) *best_sad = sad;
  if (sad > 100) return ORC_OUT_OF_BOUNDS;      /* :159 */
  *x = px; *y = py;
  return ORC_OK;
}

int orc_brute_track(const orc_pyr* from, const orc_pyr* to, int n, const float* from_xy,
                    float* to_xy, const float* coarse_sched, int n_coarse, const float* fine_sched,
                    int n_fine, int* status, float* best_sad, int64_t* npos, int nthreads) {
  int64_t total = 0;
  int nok = 0;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : total, nok)
#endif
  for (int i = 0; i < n; ++i) {
    int64_t np = 0;
    float x = to_xy[2 * i], y = to_xy[2 * i + 1], sad = 0;
    int st = orc_brute_track_feature(from, from_xy[2 * i], from_xy[2 * i + 1], to, coarse_sched,
                                     n_coarse, fine_sched, n_fine, &x, &y, &sad, &np);
    to_xy[2 * i] = x; to_xy[2 * i + 1] = y;
    if (status) status[i] = st;
    if (best_sad) best_sad[i] = sad;
    total += np; nok += (st == ORC_OK);
  }
  if (npos) *npos += total;
  return nok;
}

/* ------------------------------------------------------------------ P4 Hamming */

void orc_hamming256_top2(const uint32_t* q, int nq, const uint32_t* t, int nt, int ratio_num,
                         int ratio_den, int max_dist, int32_t* idx, int32_t* dist, uint8_t* pass,
                         int nthreads) {
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
  for (int i = 0; i < nq; ++i) {
    const uint64_t* a = (const uint64_t*)(q + (size_t)i * 8);
    int d1 = 257, d2 = 257, i1 = -1, i2 = -1;
    for (int j = 0; j < nt; ++j) {
      const uint64_t* b = (const uint64_t*)(t + (size_t)j * 8);
      int d = __builtin_popcountll(a[0] ^ b[0]) + __builtin_popcountll(a[1] ^ b[1]) +
              __builtin_popcountll(a[2] ^ b[2]) + __builtin_popcountll(a[3] ^ b[3]);
      if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = j; }      /* strict <: lowest index wins ties */
      else if (d < d2) { d2 = d; i2 = j; }
    }
    idx[2 * i] = i1; idx[2 * i + 1] = i2;
    dist[2 * i] = d1; dist[2 * i + 1] = d2;
    if (pass) pass[i] = (uint8_t)(i1 >= 0 && d1 <= max_dist && (int64_t)d1 * ratio_den < (int64_t)d2 * ratio_num);
  }
}
