/*
 * slamfe.h -- C ABI of the B200-native visual front-end (libslamfe.so).
 *
 * This is the drop-in boundary for the data-parallel front-end of ywrt/slam-robot:
 * image-pyramid construction, coarse-to-fine patch tracking with forward/backward
 * validation, dense SSD window search, and 256-bit Hamming matching.  Plain pointers and
 * sizes only; no C++ or torch types.  Every entry point names the reference interface it
 * replaces (file:line into the reference tree).  There is NO CPU fallback: every call
 * needs a CUDA device and fails with SFE_ERR_CUDA otherwise.
 *
 * Conventions
 *   - Functions return 0 (SFE_SUCCESS) or an sfe_error; sfe_last_error(ctx) has the text.
 *   - "_dev" variants take DEVICE pointers and only enqueue work on the context's stream
 *     (no synchronisation); the plain variants take HOST pointers, copy in, run, copy out
 *     and return after the results are in the caller's buffers.
 *   - Points are interleaved float x,y pairs in pixel units of pyramid level 0, exactly the
 *     cv::Point2f values the reference passes around (matcher.cpp:173).
 *   - Per-feature status values are the reference's Status enum (hessian.h:48-52).
 *   - A context is bound to one GPU and one stream; calls on one context are stream-ordered
 *     and must not be issued from several host threads at once (the reference's Matcher is
 *     single-threaded too, main.cpp:490).  sfe_set_stream may switch streams between calls:
 *     every tracker launch owns its work-queue counter, but the scratch buffers of the
 *     host-pointer entries are shared, so those still assume one stream at a time.
 */
#ifndef SLAMFE_H_
#define SLAMFE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFE_MAX_LEVELS 12
#define SFE_PATCH 13 /* kWindowSize, matcher.cpp:27 */

typedef struct sfe_ctx sfe_ctx;
typedef struct sfe_pyr sfe_pyr;

typedef enum { SFE_SUCCESS = 0, SFE_ERR_INVALID = 1, SFE_ERR_CUDA = 2, SFE_ERR_NOMEM = 3 } sfe_error;

/* hessian.h:48-52 / klt.h:53-57 / brute.h:28-32 */
typedef enum { SFE_OK = 0, SFE_SMALL_DET = 1, SFE_OUT_OF_BOUNDS = 2 } sfe_status;

/* which tracker's MakePyramid: hessian.h:95-126, klt.h:98-137, brute.h:59-80 */
typedef enum { SFE_HESSIAN = 0, SFE_KLT = 1, SFE_BRUTE = 2 } sfe_flavor;

/* ---- context ------------------------------------------------------------------------- */

/* Replaces the FeatureTracker constructor (matcher.cpp:304 -> hessian.h:11-30): builds the
 * 13x13 weighting mask once and uploads it. */
int sfe_create(int device, sfe_ctx** out);
void sfe_destroy(sfe_ctx* ctx);
const char* sfe_last_error(const sfe_ctx* ctx);
/* Use the caller's CUDA stream (a cudaStream_t passed as void*); NULL restores the
 * context's own stream. */
int sfe_set_stream(sfe_ctx* ctx, void* cuda_stream);
int sfe_sync(sfe_ctx* ctx);
/* Copies the 169 mask weights (hessian.h:11-30) to host memory. */
int sfe_get_mask(sfe_ctx* ctx, float* mask169);
/* Pinned host memory for the host-pointer entry points (optional; any host memory works). */
int sfe_host_alloc(sfe_ctx* ctx, size_t bytes, void** out);
int sfe_host_free(sfe_ctx* ctx, void* p);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t sfe_launch_count(const sfe_ctx* ctx);

/* ---- pyramids ------------------------------------------------------------------------ */

/* Storage for `batch` pyramids of `depth` levels of a w x h frame; level sizes follow
 * ((w+1)/2,(h+1)/2) (hessian.h:108).  Replaces the Pyramid = vector<GradImage> value type
 * (hessian.h:42-46). */
int sfe_pyr_create(sfe_ctx* ctx, int w, int h, int depth, int flavor, int batch, sfe_pyr** out);
void sfe_pyr_destroy(sfe_pyr* pyr);
int sfe_pyr_level_size(const sfe_pyr* pyr, int level, int* w, int* h, int* pitch_floats);
/* Algorithmic HBM bytes of building ONE frame: 3*w*h read + 4*sum(level pixels)*planes written. */
int64_t sfe_pyr_bytes_per_frame(const sfe_pyr* pyr);

/* Replaces MakePyramid(img, depth) (matcher.cpp:317 -> hessian.h:95-126 / klt.h:98-137 /
 * brute.h:59-80) for `count` frames starting at slot `first`.  bgr: 8-bit interleaved
 * 3-channel rows (the CV_8UC3 frames of video.cpp:185), row_stride / frame_stride in bytes. */
int sfe_pyr_build(sfe_ctx* ctx, sfe_pyr* pyr, const uint8_t* bgr_host, size_t row_stride,
                  size_t frame_stride, int first, int count);
int sfe_pyr_build_dev(sfe_ctx* ctx, sfe_pyr* pyr, const uint8_t* bgr_dev, size_t row_stride,
                      size_t frame_stride, int first, int count);
/* Copies one level plane (0 image, 1 gradx, 2 grady) of one frame to a dense w*h host array. */
int sfe_pyr_download(sfe_ctx* ctx, const sfe_pyr* pyr, int frame, int level, int plane,
                     float* host_dst);

/* ---- P1: forward/backward patch tracking (the live path) ------------------------------ */

/* Replaces the free function TrackFeature (matcher.cpp:173-206) -- GetPatches + TrackFeature
 * forward, GetPatches + TrackFeature backward, status and 0.3 px consistency gate -- for a
 * batch of features, i.e. the body of the FindMatches loop (matcher.cpp:218-270) minus the
 * host-side map bookkeeping.
 *
 * Feature i belongs to pair p = i / n_per_pair and is tracked from frame (from_first + p) of
 * `from` to frame (to_first + p) of `to` (the same pyramid object may be passed twice).
 *   from_xy  [n][2]  in   template position in the `from` frame           (from_pt)
 *   to_xy    [n][2]  in   initial guess; out: tracked position, unchanged when the forward
 *                         track fails (hessian.h:262 semantics)             (to_pt)
 *   levels   [n] or NULL  pyramid levels to use per feature (3 or 6, matcher.cpp:227-229), each >= 1 (the host-pointer
 *                         entries reject smaller values, the _dev entries track such a feature with 1 level);
 *                         NULL = default_levels for all
 *   thr, maxit, fb_max    0.001, 10, 0.3 in the reference (matcher.cpp:176,182,201)
 *   back_xy  [n][2]  out  backward-tracked position                       (back_pt)
 *   status_fwd/status_bwd [n] out  sfe_status of each direction           (s1, s2)
 *   accepted [n]     out  1 iff both OK and |from - back| <= fb_max       (return value)
 *   steps    [n] or NULL out  Newton steps spent on the feature (forward + backward)
 */
int sfe_track_fb(sfe_ctx* ctx, const sfe_pyr* from, int from_first, const sfe_pyr* to,
                 int to_first, int n, int n_per_pair, const float* from_xy, float* to_xy,
                 const int32_t* levels, int default_levels, float thr, int maxit, double fb_max,
                 float* back_xy, int32_t* status_fwd, int32_t* status_bwd, uint8_t* accepted,
                 int32_t* steps);
int sfe_track_fb_dev(sfe_ctx* ctx, const sfe_pyr* from, int from_first, const sfe_pyr* to,
                     int to_first, int n, int n_per_pair, const float* from_xy, float* to_xy,
                     const int32_t* levels, int default_levels, float thr, int maxit,
                     double fb_max, float* back_xy, int32_t* status_fwd, int32_t* status_bwd,
                     uint8_t* accepted, int32_t* steps);

/* Replaces one-directional tracker->TrackFeature(stack, GetPatches(tmpl, tmpl_pt, levels), thr, maxit, &pt)
 * (hessian.h:175-183 + :243-264, or the klt.h equivalents for SFE_KLT pyramids) for a batch: the
 * template patches are taken from `tmpl` at tmpl_xy, the search runs on `search` from the seed xy,
 * which is overwritten only on success (hessian.h:262).  Pair/frame addressing as in sfe_track_fb. */
int sfe_track(sfe_ctx* ctx, const sfe_pyr* tmpl, int tmpl_first, const sfe_pyr* search, int search_first,
              int n, int n_per_pair, const float* tmpl_xy, float* xy, const int32_t* levels,
              int default_levels, float thr, int maxit, int32_t* status, int32_t* steps);
int sfe_track_dev(sfe_ctx* ctx, const sfe_pyr* tmpl, int tmpl_first, const sfe_pyr* search,
                  int search_first, int n, int n_per_pair, const float* tmpl_xy, float* xy,
                  const int32_t* levels, int default_levels, float thr, int maxit, int32_t* status,
                  int32_t* steps);

/* Debug/parity accessors, one call per feature list on frame `frame` of `pyr`:
 * GetPatch (hessian.h:54-93): patches [n][169], mean [n], sumsq [n]. */
int sfe_get_patches(sfe_ctx* ctx, const sfe_pyr* pyr, int frame, int level, int n,
                    const float* xy, float* patches, float* mean, float* sumsq);
/* BruteHessian (hessian.h:147-172) of template (tmpl pyr/frame at tmpl_xy) against
 * (search pyr/frame at xy) on `level`: out [n][7] = sad0, dx, dy, dxx, dxy, dyx, dyy. */
int sfe_brute_hessian(sfe_ctx* ctx, const sfe_pyr* tmpl, int tmpl_frame, const sfe_pyr* search,
                      int search_frame, int level, int n, const float* tmpl_xy, const float* xy,
                      float* out7);

/* ---- P2: KLTTracker (klt.h) ------------------------------------------------------------ */

/* klt.h:258-424 forward/backward with the same gate as above; all levels of the pyramid are
 * used (klt.h:409) and the coarse-level threshold is 50x (klt.h:413). */
int sfe_klt_track_fb(sfe_ctx* ctx, const sfe_pyr* from, int from_first, const sfe_pyr* to,
                     int to_first, int n, int n_per_pair, const float* from_xy, float* to_xy,
                     float thr, int maxit, double fb_max, float* back_xy, int32_t* status_fwd,
                     int32_t* status_bwd, uint8_t* accepted, int32_t* steps);
int sfe_klt_track_fb_dev(sfe_ctx* ctx, const sfe_pyr* from, int from_first, const sfe_pyr* to,
                         int to_first, int n, int n_per_pair, const float* from_xy, float* to_xy,
                         float thr, int maxit, double fb_max, float* back_xy, int32_t* status_fwd,
                         int32_t* status_bwd, uint8_t* accepted, int32_t* steps);
/* The symmetric-KLT normal equations of klt.h:286-353 at one point per feature:
 * out [n][24] = A(4) B(4) C(4) RS(2) VW(2) U(4) e(2) d(2), row-major 2x2 blocks. */
int sfe_klt_system(sfe_ctx* ctx, const sfe_pyr* tmpl, int tmpl_frame, const sfe_pyr* search,
                   int search_frame, int level, int n, const float* tmpl_xy, const float* xy,
                   float* out24);

/* ---- P3: BruteTracker dense SSD window search (brute.h) -------------------------------- */

/* brute.h:129-164 with the search schedule passed explicitly as {window,res} pairs
 * (brute.h:147-148 for coarse levels, :154-158 for level 0).  best_sad [n], positions
 * (total window positions evaluated) may be NULL. */
int sfe_brute_track(sfe_ctx* ctx, const sfe_pyr* from, int from_first, const sfe_pyr* to,
                    int to_first, int n, int n_per_pair, const float* from_xy, float* to_xy,
                    const float* coarse_sched, int n_coarse, const float* fine_sched, int n_fine,
                    int32_t* status, float* best_sad, int64_t* positions);
int sfe_brute_track_dev(sfe_ctx* ctx, const sfe_pyr* from, int from_first, const sfe_pyr* to,
                        int to_first, int n, int n_per_pair, const float* from_xy, float* to_xy,
                        const float* coarse_sched, int n_coarse, const float* fine_sched,
                        int n_fine, int32_t* status, float* best_sad);

/* ---- P4: 256-bit Hamming top-2 + ratio test -------------------------------------------- */

/* The descriptor matcher BASELINE.json asks for (no counterpart in the reference; specified
 * against cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2)).  q [nq][8] u32, t [nt][8] u32.
 * idx/dist [nq][2]: best and second-best train row and distance, lowest index wins ties;
 * missing neighbours are -1 / 257.  pass [nq] (may be NULL) = d1 <= max_dist and
 * d1*ratio_den < d2*ratio_num.  `batch` independent (q,t) problems are laid out back to back
 * (q: batch*nq rows, t: batch*nt rows). */
int sfe_match_hamming256(sfe_ctx* ctx, const uint32_t* q, int nq, const uint32_t* t, int nt,
                         int batch, int ratio_num, int ratio_den, int max_dist, int32_t* idx,
                         int32_t* dist, uint8_t* pass);
int sfe_match_hamming256_dev(sfe_ctx* ctx, const uint32_t* q, int nq, const uint32_t* t, int nt,
                             int batch, int ratio_num, int ratio_den, int max_dist, int32_t* idx,
                             int32_t* dist, uint8_t* pass);

/* As sfe_match_hamming256, but only ENQUEUES the uploads, the kernels and the downloads: the
 * result buffers are valid after sfe_sync().  The work is ordered after everything enqueued on
 * the context's stream before the call, but runs on a side stream of the context, so that it
 * overlaps with whatever is enqueued next (e.g. sfe_replay_pairs); only sfe_sync() joins it.
 * The host buffers must stay alive until then (pinned memory from sfe_host_alloc makes the
 * copies truly asynchronous); one asynchronous call may be outstanding per context. */
int sfe_match_hamming256_async(sfe_ctx* ctx, const uint32_t* q, int nq, const uint32_t* t, int nt,
                               int batch, int ratio_num, int ratio_den, int max_dist, int32_t* idx,
                               int32_t* dist, uint8_t* pass);

/* Which kernel the three entries above run: 0 = chosen by problem size (default), 1 = the integer-ALU kernel
 * (LOP3 + POPC, hamming.cu), 2 = the tensor-core kernel (exact int8 tcgen05.mma contraction, hamming_mma.cu).
 * Both are bit-identical to the specification; the selector exists for tests and measurements.  Process-wide
 * (also settable as SFE_HAMMING=alu|mma); an out-of-range value only queries.  Returns the previous setting. */
int sfe_hamming_impl(int impl);

/* ---- batched replay of independent frame pairs ------------------------------------------ */

/* The host-side batching/stream layer: what Matcher::Track does per frame -- MakePyramid
 * (matcher.cpp:317) and the FindMatches tracking loop (matcher.cpp:208-271 -> :173-206) -- for
 * `npairs` independent (from, to) frame pairs of a replayed sequence (BASELINE config 4), host
 * buffers in and out.  The pairs are processed in chunks of `chunk_pairs` (<= 0: automatic, an eighth of the call
 * but at most 128 pairs, so device memory does not grow with the length of the replay) on a
 * stream pipeline (upload of chunk k+2 | pyramids of chunk k+1 | tracking of chunk k | download of
 * chunk k-1), so with pinned host buffers the PCIe transfers hide behind the kernels.
 *   from_bgr, to_bgr  npairs frames each, 8-bit BGR, row_stride / frame_stride in bytes
 *   feature arrays    npairs * n_per_pair entries, pair-major, semantics of sfe_track_fb
 * Returns after all results are in the caller's buffers. */
int sfe_replay_pairs(sfe_ctx* ctx, int w, int h, int depth, int npairs, const uint8_t* from_bgr,
                     const uint8_t* to_bgr, size_t row_stride, size_t frame_stride, int n_per_pair,
                     const float* from_xy, float* to_xy, const int32_t* levels, int default_levels,
                     float thr, int maxit, double fb_max, float* back_xy, int32_t* status_fwd,
                     int32_t* status_bwd, uint8_t* accepted, int32_t* steps, int chunk_pairs);

/* The same pipeline for a replayed SEQUENCE (`./slam --load dir`, main.cpp:446-448): `nframes` consecutive frames in one
 * buffer, pair i = (frame i, frame i + pair_stride) for i < nframes - pair_stride (the reference's two cameras alternate,
 * so same-camera pairs are pair_stride = 2 apart).  Every frame is uploaded and its pyramid built once per chunk it
 * belongs to -- as Matcher::Track builds one pyramid per new frame (matcher.cpp:317) -- instead of once as a `from` and
 * once as a `to` frame.  Feature arrays: (nframes - pair_stride) * n_per_pair entries, pair-major.  Results are
 * identical to sfe_replay_pairs on the pairs spelled out. */
int sfe_replay_sequence(sfe_ctx* ctx, int w, int h, int depth, int nframes, int pair_stride,
                        const uint8_t* frames_bgr, size_t row_stride, size_t frame_stride, int n_per_pair,
                        const float* from_xy, float* to_xy, const int32_t* levels, int default_levels,
                        float thr, int maxit, double fb_max, float* back_xy, int32_t* status_fwd,
                        int32_t* status_bwd, uint8_t* accepted, int32_t* steps, int chunk_pairs);

/* sfe_replay_sequence for frames in the camera's native format: packed YUYV, 2 bytes per pixel (w even), as the V4L2
 * path receives them (video.cpp:136-139).  Every frame crosses PCIe as 2 instead of 3 bytes per pixel and is converted
 * on the device with the integer arithmetic of video.cpp:187-223 (sfe_yuyv_to_bgr) before MakePyramid, so the results
 * equal sfe_replay_sequence on the BGR frames that loop produces.  row_stride / frame_stride in bytes (>= 2 w). */
int sfe_replay_sequence_yuyv(sfe_ctx* ctx, int w, int h, int depth, int nframes, int pair_stride,
                             const uint8_t* frames_yuyv, size_t row_stride, size_t frame_stride, int n_per_pair,
                             const float* from_xy, float* to_xy, const int32_t* levels, int default_levels,
                             float thr, int maxit, double fb_max, float* back_xy, int32_t* status_fwd,
                             int32_t* status_bwd, uint8_t* accepted, int32_t* steps, int chunk_pairs);

/* ---- several GPUs of one box ----------------------------------------------------------------- */

/* The path shards in two places (one sfe_ctx per GPU, one process or host thread per GPU):
 *   frame pairs of a replayed sequence are independent units -- each rank replays the block sfe_shard_range gives it
 *   with sfe_replay_sequence / sfe_replay_pairs, NO collective on the data path; sfe_allgather_rows collects the
 *   per-feature result rows when the caller wants them in one place (the reference's map bookkeeping,
 *   matcher.cpp:253-268, is host code and stays with the caller);
 *   the descriptor matcher has one exchange: train set broadcast, query rows sharded, top-2 rows all-gathered.
 * The reference is a single process with a single Matcher (main.cpp:490); these entries are what its C++ caller
 * binds to spread a replay or a large match over the GPUs of a box.  NCCL (libnccl.so.2) is loaded at run time;
 * the collectives run on the context's stream, ordered with the kernels. */

/* One process per GPU on a multi-socket box: binds the calling thread (and the threads it creates afterwards) to the CPUs
 * next to `device` (NVML's nvmlDeviceGetCpuAffinity, else sysfs local_cpulist), intersected with the affinity it already
 * has.  Call it before allocating pinned host buffers: first-touch then places them on the device's NUMA node, and uploads
 * do not cross the socket interconnect (the reference is a single process on a single-socket robot and has no
 * counterpart; bench.py --gpus N > 1 calls it).  Returns the number of CPUs bound to, 0 when nothing was changed (no
 * topology information, or the intersection is empty); never an error. */
int sfe_bind_host_to_device(int device);

/* Contiguous block [lo, hi) of n units owned by `rank` of `world` (blocks differ by at most one unit). */
int sfe_shard_range(int64_t n, int rank, int world, int64_t* lo, int64_t* hi);
/* Communicator set-up: rank 0 obtains an id (ncclGetUniqueId) and hands the 128 bytes to the other ranks by any
 * means; every rank then calls sfe_dist_init (ncclCommInitRank on the context's device; collective, blocks until
 * all ranks have called).  Alternatively sfe_dist_attach adopts an ncclComm_t the caller created for the context's
 * device (passed as void*; it is not destroyed with the context). */
int sfe_dist_unique_id(uint8_t* id128);
int sfe_dist_init(sfe_ctx* ctx, const uint8_t* id128, int rank, int world);
int sfe_dist_attach(sfe_ctx* ctx, void* nccl_comm, int rank, int world);
int sfe_dist_shutdown(sfe_ctx* ctx);
/* All-gather of row blocks of unequal length: this rank contributes rows [lo, hi) = sfe_shard_range(n_total) from
 * local_dev (or NULL when they already sit at all_dev + lo * row_bytes); afterwards all_dev holds all n_total rows on
 * every rank.  Device pointers; enqueued on the context's stream. */
int sfe_allgather_rows_dev(sfe_ctx* ctx, const void* local_dev, size_t row_bytes, int64_t n_total, void* all_dev);
/* sfe_match_hamming256 for nq_total query rows sharded over the ranks (BASELINE config 5): q_local holds THIS rank's
 * rows [lo, hi) = sfe_shard_range(nq_total); t holds the nt train rows on rank train_root and is overwritten with them
 * on the other ranks (ncclBroadcast); idx_all / dist_all [nq_total][2] and pass_all [nq_total] (may be NULL) receive the
 * rows of ALL ranks (all-gather), identical to a single-GPU sfe_match_hamming256 of the whole problem.  The _dev
 * variant takes device pointers and only enqueues; the host variant copies in, runs, copies out and synchronises
 * (t may be NULL on ranks other than train_root).  Collective: every rank of the communicator must call it.  A rank
 * whose own match cannot be launched still takes part in the all-gathers with its rows marked invalid (idx = dist = -1,
 * pass = 0) and then returns the error, so that its peers are not left waiting; argument errors are returned before any
 * collective is entered -- treat a non-zero return on any rank as fatal for the communicator, as with NCCL itself. */
int sfe_match_hamming256_sharded_dev(sfe_ctx* ctx, const uint32_t* q_local, int64_t nq_total, uint32_t* t, int nt,
                                     int train_root, int ratio_num, int ratio_den, int max_dist, int32_t* idx_all,
                                     int32_t* dist_all, uint8_t* pass_all);
int sfe_match_hamming256_sharded(sfe_ctx* ctx, const uint32_t* q_local, int64_t nq_total, const uint32_t* t, int nt,
                                 int train_root, int ratio_num, int ratio_den, int max_dist, int32_t* idx_all,
                                 int32_t* dist_all, uint8_t* pass_all);

/* ---- corner seeding (the step after tracking on keyframes) -------------------------------- */

/* Replaces cv::cvtColor(img, grey, CV_RGB2GRAY) + cv::goodFeaturesToTrack(grey, corners, max_corners,
 * quality, min_distance) of matcher.cpp:313 and :123-130 (defaults: min-eigenvalue response, blockSize 3,
 * 3x3 Sobel, no mask) for `count` BGR frames.  corners [count][max_corners][2] (x,y as cv::Point2f,
 * strongest first), ncorners [count].  The reference calls it with 120, 0.01, 20.
 * eig_out (host variant only, may be NULL): the cornerMinEigenVal response maps [count][h][w]. */
int sfe_good_features(sfe_ctx* ctx, const uint8_t* bgr_host, int w, int h, size_t row_stride,
                      size_t frame_stride, int count, int max_corners, double quality,
                      double min_distance, float* corners, int32_t* ncorners, float* eig_out);
int sfe_good_features_dev(sfe_ctx* ctx, const uint8_t* bgr_dev, int w, int h, size_t row_stride,
                          size_t frame_stride, int count, int max_corners, double quality,
                          double min_distance, float* corners, int32_t* ncorners);

/* ---- seeding the search (the step before tracking) ------------------------------------------ */

/* Replaces the head of the FindMatches loop, matcher.cpp:224-245, for n features of one target frame:
 *   levels[i]  = uncertainty[i] > 100 ? 6 : 3                                     (matcher.cpp:227-229)
 *   seed_xy[i] = from_xy[i], or Frame::Project(point) (localmap.cpp:18-26 -> project.h:11-54) when
 *                uncertainty[i] < 100 and the point projects                      (matcher.cpp:233-239)
 *   go[i]      = 0 when the seed is out of bounds (`continue`, matcher.cpp:243), else 1
 * points4 [n][4] homogeneous map points (TrackedPoint::location()), rot4 = the frame's rotation
 * quaternion coefficients [x,y,z,w], trans3 its translation, k7 = Camera::k [k1,k2,k3,fx,fy,cx,cy]
 * (localmap.h:28-30); cols/rows = size of the target frame.  The pose arguments are host arrays in
 * both variants (call parameters). */
int sfe_seed_features(sfe_ctx* ctx, int n, const double* points4, const double* uncertainty,
                      const double* rot4, const double* trans3, const double* k7, const float* from_xy,
                      int cols, int rows, float* seed_xy, int32_t* levels, uint8_t* go);
int sfe_seed_features_dev(sfe_ctx* ctx, int n, const double* points4, const double* uncertainty,
                          const double* rot4, const double* trans3, const double* k7,
                          const float* from_xy, int cols, int rows, float* seed_xy, int32_t* levels,
                          uint8_t* go);

/* ---- live capture format --------------------------------------------------------------------- */

/* Replaces the integer YUYV -> BGR loop of the V4L2 capture path, video.cpp:187-223: npixels pixels
 * (a multiple of 4), 2 bytes per pixel in, 3 bytes per pixel out (the CV_8UC3 frame MakePyramid takes). */
int sfe_yuyv_to_bgr(sfe_ctx* ctx, const uint8_t* yuyv_host, size_t npixels, uint8_t* bgr_host);
int sfe_yuyv_to_bgr_dev(sfe_ctx* ctx, const uint8_t* yuyv_dev, size_t npixels, uint8_t* bgr_dev);

#ifdef __cplusplus
}
#endif
#endif /* SLAMFE_H_ */
