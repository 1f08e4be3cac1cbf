/*
 * oracle.h -- CPU restatement of the slam-robot visual front-end hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and there only as the checker / the timed CPU arm.
 * The product path (slam-robot_b200/csrc) never links or calls it.
 *
 * What is restated (file:line into /root/reference):
 *   P1  hessian.h:11-30 (mask) :54-93 (GetPatch) :95-126 (MakePyramid)
 *       :129-141 (ScorePatchMatch) :147-172 (BruteHessian) :185-241 (Track)
 *       :243-264 (TrackFeature);  matcher.cpp:173-206 (forward/backward TrackFeature)
 *   P2  klt.h:59-96, :98-137, :139-149, :181-204, :258-424
 *   P3  brute.h:34-57, :59-80, :82-117, :129-164
 *   P4  256-bit Hamming top-2 + integer ratio test (no reference counterpart;
 *       specified against cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2), lowest
 *       train index wins ties -- "parity unpinned" by the reference itself).
 *
 * The pixel arithmetic of the reference lives in un-vendored OpenCV (version
 * unpinned, 2.4-era API) and Eigen.  The restatement follows OpenCV 4.13's
 * AVX2/FMA code paths, which were probed bit-for-bit in this container
 * (tests/golden/make_golden.py re-runs the probes):
 *   cvtColor RGB2GRAY   (9798*c0 + 19235*c1 + 3735*c2 + 2^14) >> 15     bit-exact
 *   convertTo(1/255.)   (float)g * (float)(1/255.)                      bit-exact
 *   GaussianBlur 5x5    row: fma(k2,(m2+p2), fma(k0,c, k1*(m1+p1)))
 *                       col: fma(k2,(m2+p2), fma(k1,(m1+p1), k0*c))     bit-exact except
 *                       the (W mod 8) right-most SIMD-tail columns (<=1 ulp)
 *   pyrDown             row: ((m2+p2) + 4(m1+p1)) + 6c
 *                       col: (4((r1+r3)+r2) + ((r0+r4)+(r2+r2))) / 256  bit-exact except
 *                       column 0 and the <=3 right-most columns (<=1 ulp)
 *   getRectSubPix       fma(s11,a22, fma(s10,a21, fma(s01,a12, s00*a11))),
 *                       2-tap forms on overflow rows/columns, incl. OpenCV's
 *                       top-right corner quirk                          bit-exact
 * Reductions over the 169 patch pixels use a DECLARED order:
 * pixel i goes to lane i%32, each lane accumulates its pixels in increasing i,
 * then a balanced pairwise tree over the 32 lanes: strides 16,8,4,2,1 for klt.h and brute.h
 * (tree32, what a __shfl_xor butterfly computes), strides 1,2,4,8,16 -- adjacent lanes first --
 * for hessian.h (tree32_adj, what the live tracker's shared-memory transposition computes).
 * The reference itself is built with -ffast-math, so it has no defined order of its own.
 */
#ifndef SLAMFE_ORACLE_H_
#define SLAMFE_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LEVELS 12
#define ORC_PATCH 13
#define ORC_PLEN 169

enum { ORC_OK = 0, ORC_SMALL_DET = 1, ORC_OUT_OF_BOUNDS = 2 };          /* hessian.h:48-52 */
enum { ORC_FLAVOR_HESSIAN = 0, ORC_FLAVOR_KLT = 1, ORC_FLAVOR_BRUTE = 2 };

typedef struct {
  int w, h;
  float* data; /* dense, row pitch == w */
} orc_plane;

typedef struct {
  int depth;
  int flavor;
  orc_plane img[ORC_MAX_LEVELS];
  orc_plane gx[ORC_MAX_LEVELS]; /* KLT flavour only */
  orc_plane gy[ORC_MAX_LEVELS];
} orc_pyr;

/* per-call work counters (for the samples/s metric) */
typedef struct {
  int64_t newton_steps;
  int64_t patches; /* 13x13 bilinear extractions */
} orc_counters;

/* ---- primitives ---------------------------------------------------------- */
void orc_mask13(float* mask169);                                        /* hessian.h:11-30 */
void orc_gray_u8(const uint8_t* bgr, int w, int h, size_t stride, uint8_t* gray);
void orc_gauss5(const float* src, int w, int h, float sigma_id_k0, float k1, float k2, float* dst);
void orc_gauss5_sigma(const float* src, int w, int h, double sigma, float* dst); /* sigma in {1.1,0.8,0.6} */
void orc_pyrdown(const float* src, int w, int h, float* dst);           /* dst is ((w+1)/2,(h+1)/2) */
int orc_pyrdown_hbody(int w); /* last column of cv::pyrDown's horizontal vector body for input width w */
void orc_scharr(const float* src, int w, int h, float* gx, float* gy);  /* klt.h:105-106, scale 1/32 */
void orc_rect_subpix(const float* img, int w, int h, int n, int m, float cx, float cy,
                     float* dst, int dst_pitch);                        /* cv::getRectSubPix 32f */

/* ---- pyramids ------------------------------------------------------------ */
orc_pyr* orc_pyr_build(const uint8_t* bgr, int w, int h, size_t stride, int depth, int flavor);
orc_pyr* orc_pyr_from_planes(int depth, int flavor, const int* ws, const int* hs,
                             const float* const* planes /* [depth*3]: img,gx,gy */);
void orc_pyr_free(orc_pyr* p);
int orc_pyr_depth(const orc_pyr* p);
int orc_pyr_w(const orc_pyr* p, int level);
int orc_pyr_h(const orc_pyr* p, int level);
const float* orc_pyr_plane(const orc_pyr* p, int level, int plane /*0 img,1 gx,2 gy*/);

/* ---- P1: HessianTracker --------------------------------------------------- */
void orc_hes_get_patch(const orc_pyr* p, int level, float x, float y, float* data169,
                       float* mean, float* sumsq);                      /* hessian.h:54-93 */
float orc_hes_score(const float* p1, float mean1, float sumsq1, const float* p2, float mean2,
                    float sumsq2);                                      /* hessian.h:129-141 */
float orc_hes_brute_hessian(const orc_pyr* p, int level, const float* patch, float mean,
                            float sumsq, float x, float y, float out6[6]); /* hessian.h:147-172 */
int orc_hes_track_feature(const orc_pyr* tmpl_pyr, float tx, float ty, const orc_pyr* search,
                          int levels, float thr, int maxit, float* x, float* y,
                          orc_counters* c);                             /* GetPatches + TrackFeature */
/* matcher.cpp:173-206 for n features; to_xy is in (seed) / out; returns #accepted */
int orc_hes_track_fb(const orc_pyr* from, const orc_pyr* to, int n, const float* from_xy,
                     float* to_xy, const int* levels, float thr, int maxit, double fb_max,
                     float* back_xy, int* st_fwd, int* st_bwd, uint8_t* accepted,
                     orc_counters* c, int nthreads);

/* ---- P2: KLTTracker -------------------------------------------------------- */
/* one klt.h Track iteration's symmetric-KLT quantities at (x,y): A,B,C (row major 2x2),
 * RS, VW, U, e, d  -> out[4+4+4+2+2+4+2+2 = 24] (klt.h:286-343) */
void orc_klt_system(const orc_pyr* tmpl_pyr, float tx, float ty, const orc_pyr* search, int level,
                    float x, float y, float out24[24]);
int orc_klt_track_feature(const orc_pyr* tmpl_pyr, float tx, float ty, const orc_pyr* search,
                          float thr, int maxit, float* x, float* y, orc_counters* c);
int orc_klt_track_fb(const orc_pyr* from, const orc_pyr* to, int n, const float* from_xy,
                     float* to_xy, float thr, int maxit, double fb_max, float* back_xy, int* st_fwd,
                     int* st_bwd, uint8_t* accepted, orc_counters* c, int nthreads);

/* ---- P3: BruteTracker ------------------------------------------------------ */
/* brute.h:96-117; returns best sad, updates (x,y); npos = #window positions visited */
float orc_brute_search_best(const orc_pyr* search, int level, const float* patch, float mean,
                            float sumsq, float window, float res, float* x, float* y,
                            int64_t* npos);
/* brute.h:129-164 with the schedule given explicitly (so the 1601^2 debug pass can be dropped
 * or kept): sched = {window,res} pairs; n_coarse pairs per coarse level, n_fine at level 0 */
int orc_brute_track_feature(const orc_pyr* tmpl_pyr, float tx, float ty, const orc_pyr* search,
                            const float* coarse_sched, int n_coarse, const float* fine_sched,
                            int n_fine, float* x, float* y, float* best_sad, int64_t* npos);
int orc_brute_track(const orc_pyr* from, const orc_pyr* to, int n, const float* from_xy,
                    float* to_xy, const float* coarse_sched, int n_coarse, const float* fine_sched,
                    int n_fine, int* status, float* best_sad, int64_t* npos, int nthreads);

/* ---- P4: Hamming ----------------------------------------------------------- */
/* Seeding (8f rank 3): Frame::Project (localmap.cpp:18-26, project.h:11-54) and the head of the FindMatches loop
 * (matcher.cpp:224-245).  Live capture (8f rank 4): the integer YUYV -> BGR conversion of video.cpp:187-223. */
int orc_project(const double* rot, const double* trans, const double* k, const double* pt, double* out2);
void orc_seed_features(int n, const double* points4, const double* uncertainty, const double* rot, const double* trans,
                       const double* k, const float* from_xy, int cols, int rows, float* seed_xy, int32_t* levels,
                       uint8_t* go);
void orc_yuyv_to_bgr(const uint8_t* in, size_t bytes, uint8_t* out);

/* Corner seeding: cv::cornerMinEigenVal(gray8, 3, 3) and cv::goodFeaturesToTrack with
 * its defaults, as matcher.cpp:123-130 calls them; probed bit-for-bit against cv2 4.13 (see oracle.c). */
void orc_min_eigen_val(const uint8_t* gray, int w, int h, float* eig);
int orc_good_features(const uint8_t* bgr, int w, int h, size_t stride, int max_corners, double quality,
                      double min_distance, float* corners_xy, float* eig_out, float* max_out);

/* q: nq x 8 u32, t: nt x 8 u32.  idx/dist: nq x 2 (best, second); lowest train index wins
 * ties; missing neighbours are idx -1 / dist 257.  pass[i] = d1 <= max_dist &&
 * d1*ratio_den < d2*ratio_num (integer ratio test). */
void orc_hamming256_top2(const uint32_t* q, int nq, const uint32_t* t, int nt, int ratio_num,
                         int ratio_den, int max_dist, int32_t* idx, int32_t* dist, uint8_t* pass,
                         int nthreads);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
