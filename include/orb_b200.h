/*
 * orb_b200.h -- C ABI of the B200-native ORB front end (liborb_b200.so).
 *
 * Drop-in boundary for the two data-parallel hot paths of
 * WangHewei16/ORB-SLAM-System (an ORB-SLAM2 fork): ORBextractor::operator()
 * and the ORBmatcher / Frame Hamming searches.  The reference has no FFI layer;
 * the boundary is the C++ ABI of two classes inside libORB_SLAM2.so.  A thin
 * C++ adapter (orb_slam_system_b200/adapter/) keeps those class signatures and
 * calls the entry points below; INTEGRATION.md shows the binding.
 *
 * Conventions: plain C types only; every function returns an orb_status
 * (0 = ok); no exceptions cross the boundary; outputs go to caller-owned
 * buffers with explicit capacities; one handle per calling thread (each handle
 * owns a CUDA stream and its device memory), mirroring the reference's
 * one-extractor-instance-per-camera use (reference src/Tracking.cc:76-82,
 * src/Frame.cc:58-61).  There is no CPU fallback: if no CUDA device is usable
 * the create calls fail with ORB_ERR_CUDA.
 *
 * All file:line citations are relative to the reference repository root.
 */
#ifndef ORB_B200_H
#define ORB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum orb_status {
    ORB_OK = 0,
    ORB_ERR_INVALID = -1,     /* bad argument */
    ORB_ERR_SHAPE = -2,       /* image shape on which the reference itself faults */
    ORB_ERR_CAPACITY = -3,    /* caller buffer too small; counts hold the required sizes */
    ORB_ERR_CUDA = -4,        /* CUDA runtime/driver error; see orb_last_error() */
    ORB_ERR_UNSEPARABLE = -5  /* octree cannot terminate (the reference would loop forever) */
} orb_status;

/* Same 28-byte layout as cv::KeyPoint (reference include/Frame.h:117-118 holds
 * std::vector<cv::KeyPoint>): pt.x, pt.y, size, angle, response, octave, class_id. */
typedef struct orb_keypoint {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} orb_keypoint;

/* Arguments of ORBextractor::ORBextractor (include/ORBextractor.h:31-32); read from the
 * settings YAML keys ORBextractor.{nFeatures,scaleFactor,nLevels,iniThFAST,minThFAST}
 * (src/Tracking.cc:70-74). */
typedef struct orb_params {
    int32_t nfeatures;
    float scale_factor;
    int32_t nlevels; /* 1..16 */
    int32_t ini_th_fast;
    int32_t min_th_fast;
} orb_params;

typedef struct orb_extractor orb_extractor; /* opaque */
typedef struct orb_matcher orb_matcher;     /* opaque */

/* ---- extractor: replaces ORBextractor (include/ORBextractor.h:26-93) -------------------- */

/* ORBextractor::ORBextractor (src/ORBextractor.cc:116-170).  Device buffers are sized for
 * max_batch frames of up to max_rows x max_cols.  device = CUDA ordinal. */
int orb_extractor_create(const orb_params* params, int max_rows, int max_cols, int max_batch,
                         int device, orb_extractor** out);
void orb_extractor_destroy(orb_extractor* h);

/* GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (include/ORBextractor.h:43-63) and mnFeaturesPerLevel.
 * Each array receives nlevels entries; NULL pointers are skipped. */
int orb_extractor_tables(const orb_extractor* h, float* scale, float* inv_scale, float* sigma2,
                         float* inv_sigma2, int32_t* features_per_level);

/* Upper bound on keypoints per frame for rows x cols images (the fork's octree returns up
 * to <4x the per-level quota); use it to size kps/desc buffers. */
int orb_extractor_keypoint_bound(const orb_extractor* h, int rows, int cols, int* bound);

/* ORBextractor::operator() (src/ORBextractor.cc:442-495) on one host image (8-bit gray,
 * `stride` bytes per row).  Writes *count keypoints (cv::KeyPoint layout) and count x 32
 * descriptor bytes, rows in keypoint order.  Empty image (rows or cols 0, or img NULL):
 * returns ORB_OK with *count = 0, like the reference's silent return (:444-445).
 * Synchronous: results are in the host buffers on return. */
int orb_extract(orb_extractor* h, const uint8_t* img, int rows, int cols, size_t stride,
                orb_keypoint* kps, uint8_t* desc, int cap, int* count);

/* n same-shape host frames in one launch sequence (frame f at imgs + f*frame_stride);
 * outputs for frame f at kps + f*cap, desc + f*cap*32, counts[f].  Host buffers should be
 * pinned for full PCIe rate; copies overlap the kernels of neighbouring sub-batches. */
int orb_extract_batch(orb_extractor* h, int n, const uint8_t* imgs, int rows, int cols,
                      size_t stride, size_t frame_stride, orb_keypoint* kps, uint8_t* desc,
                      int cap, int* counts);

/* Asynchronous form of orb_extract_batch for throughput serving: submit enqueues the copies and
 * kernels of one batch and returns a ticket; wait blocks until that batch's keypoints, descriptors
 * and counts are in the caller's buffers.  Up to three batches may be in flight per handle (ORB_MAX_IN_FLIGHT): the H2D
 * copies of batch i+1 run while batch i computes and the D2H copies of batch i run while batch
 * i+1 computes.  The buffers must stay valid (and the inputs unchanged) until wait returns.
 * orb_extract_batch == submit + wait. */
#define ORB_MAX_IN_FLIGHT 3
int orb_extract_batch_submit(orb_extractor* h, int n, const uint8_t* imgs, int rows, int cols,
                             size_t stride, size_t frame_stride, orb_keypoint* kps, uint8_t* desc,
                             int cap, int* counts, int* ticket);
int orb_extract_batch_wait(orb_extractor* h, int ticket);

/* Same, with every buffer resident in device memory (inputs already in HBM; no copies).
 * Asynchronous on the handle's stream; call orb_extractor_sync before reading results
 * from another stream.  d_imgs needs stride % 16 == 0 and 16-byte aligned base. */
int orb_extract_batch_device(orb_extractor* h, int n, const uint8_t* d_imgs, int rows, int cols,
                             size_t stride, size_t frame_stride, orb_keypoint* d_kps,
                             uint8_t* d_desc, int cap, int* d_counts);
int orb_extractor_sync(orb_extractor* h);
/* The handle's CUDA stream (cudaStream_t) so callers can order their own work after it. */
void* orb_extractor_stream(orb_extractor* h);

/* ---- image ingest: the step before the path, fused into the level-0 load -------------------------
 * What the reference does to a raw camera frame before ORBextractor::operator() sees it:
 *   cv::remap(raw, rect, M1, M2, cv::INTER_LINEAR) with the CV_32FC1 maps of cv::initUndistortRectifyMap
 *   (Examples/Stereo/stereo_euroc.cc:97-98, :136-137; default BORDER_CONSTANT 0), and
 *   cv::cvtColor(im, mImGray, CV_RGB2GRAY / CV_BGR2GRAY) when channels() > 1, in
 *   Tracking::GrabImageStereo / GrabImageRGBD / GrabImageMonocular (src/Tracking.cc:118-126, :136-141, :155-160).
 * With an ingest configuration set, the orb_ingest_* calls take RAW frames (src_rows x src_cols x channels, 8 bit)
 * and one kernel writes remap -> gray straight into the extractor's level-0 buffer (the rectified / gray image
 * never makes its own round trip through HBM); level 0 of orb_get_pyramid_level is then the reference's mImGray.
 * The maps are computed once per sequence by the caller (cv::initUndistortRectifyMap is setup code, not on the
 * per-frame path) and stay resident on the device.  remap applies to each channel before the gray conversion,
 * the reference's order (remap in the example driver, cvtColor inside Tracking). */
typedef struct orb_ingest_config {
    int32_t src_rows, src_cols; /* raw frame */
    int32_t channels;           /* 1 (gray: no conversion), 3 or 4 (alpha ignored) */
    int32_t bgr;                /* first channel is blue (mbRGB false = Camera.RGB: 0 in the YAML, src/Tracking.cc:68, :119-125) */
    int32_t gray_variant;       /* 4: OpenCV 4.x coefficients (R*9798 + G*19235 + B*3735 + 2^14) >> 15, pinned against
                                   cv2 4.13; 3: OpenCV 2.4 - 3.4 (R*4899 + G*9617 + B*1868 + 2^13) >> 14, what the
                                   reference's 3.4.15 build computes */
    int32_t dst_rows, dst_cols; /* size of the maps = the image the extractor sees; without maps must equal src */
    const float* map_x;         /* host, dst_rows x dst_cols floats, dense; NULL (both): no remap */
    const float* map_y;
} orb_ingest_config;

/* Installs (cfg != NULL) or clears (cfg == NULL) the ingest configuration; uploads the maps. */
int orb_extractor_set_ingest(orb_extractor* h, const orb_ingest_config* cfg);

/* orb_extract_batch / _submit / _device on raw frames: frame f at raw + f*frame_stride, `stride` bytes per raw
 * row (>= src_cols * channels).  Outputs exactly as orb_extract_batch; tickets are waited on with
 * orb_extract_batch_wait.  ORB_ERR_INVALID without a configuration. */
int orb_ingest_extract_batch(orb_extractor* h, int n, const uint8_t* raw, size_t stride, size_t frame_stride,
                             orb_keypoint* kps, uint8_t* desc, int cap, int* counts);
int orb_ingest_extract_batch_submit(orb_extractor* h, int n, const uint8_t* raw, size_t stride, size_t frame_stride,
                                    orb_keypoint* kps, uint8_t* desc, int cap, int* counts, int* ticket);
int orb_ingest_extract_batch_device(orb_extractor* h, int n, const uint8_t* d_raw, size_t stride, size_t frame_stride,
                                    orb_keypoint* d_kps, uint8_t* d_desc, int cap, int* d_counts);

/* Pyramid level `level` of frame `frame` of the last call, tightly cropped (no border), to a
 * host buffer -- what refills the public ORBextractor::mvImagePyramid
 * (include/ORBextractor.h:65, read by src/Frame.cc:453,543-560).  dst may be NULL to
 * query the size only. */
int orb_get_pyramid_level(orb_extractor* h, int frame, int level, uint8_t* dst, size_t dst_stride,
                          int* rows, int* cols);

/* All levels of frame `frame` in one go: dst[l] / dst_stride[l] for l < nlevels of the extractor (dst[l] NULL: level skipped).
 * The copies are enqueued together and waited for once -- what the adapter uses to refill mvImagePyramid
 * (2 x nlevels synchronous calls otherwise). */
int orb_get_pyramid_levels(orb_extractor* h, int frame, uint8_t* const* dst, const size_t* dst_stride);

/* Counters of the last call for frame `frame` (arrays of nlevels; NULL skipped):
 * FAST candidates handed to the octree and keypoints kept, per level. */
int orb_extractor_level_stats(orb_extractor* h, int frame, int32_t* candidates, int32_t* kept);

/* Per-stage device timing (CUDA events on the handle's stream around each kernel stage:
 * 0 pyramid, 1 detect, 2 octree, 3 blur, 4 describe).  set_profiling(1) clears the records and
 * starts recording every following extract call; stage_times synchronises and returns the
 * summed milliseconds per stage and the number of calls recorded.  While profiling is on the
 * stages run back to back on one stream (each duration is the kernel's own); with profiling
 * off the blur runs on a second stream concurrently with detect + octree. */
#define ORB_NUM_STAGES 5
int orb_extractor_set_profiling(orb_extractor* h, int on);
int orb_extractor_stage_times(orb_extractor* h, double* ms_sum, int* ncalls);
const char* orb_stage_name(int stage);

/* ---- matcher: replaces the Hamming scans of ORBmatcher / Frame --------------------------- */

int orb_matcher_create(int device, orb_matcher** out);
void orb_matcher_destroy(orb_matcher* m);

/* ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:896-908) is the unit.  The scan every
 * search shares (src/ORBmatcher.cc:49-55, :225-231, :321-327): candidates in list order,
 * `if d<best {second=best; best=d; idx=i} else if d<second {second=d}`, both starting at
 * INT_MAX.  Outputs per query: best_idx (-1 if no candidate), best_dist, second_dist
 * (INT_MAX when absent).  Accept rules (:58, :235, :329) stay with the caller. */

/* Brute force: every query against all train rows in index order (BASELINE config 4).
 * Host buffers; q is nq x 32 bytes, t is nt x 32 bytes. */
int orb_match_all(orb_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt,
                  int32_t* best_idx, int32_t* best_dist, int32_t* second_dist);

/* npairs independent (query set, train set) pairs with fixed strides, device or host
 * buffers (on_device != 0: all pointers are device pointers, asynchronous).
 * Pair p: q + p*q_stride (nq[p] rows), t + p*t_stride (nt[p] rows); outputs at p*out_stride. */
int orb_match_all_batch(orb_matcher* m, int npairs, const uint8_t* q, const int32_t* nq,
                        size_t q_stride, const uint8_t* t, const int32_t* nt, size_t t_stride,
                        int32_t* best_idx, int32_t* best_dist, int32_t* second_dist,
                        size_t out_stride, int on_device);

typedef enum orb_tie_rule {
    ORB_TIE_FIRST_MIN = 0, /* the shared scan above */
    ORB_TIE_LAST_MIN = 1   /* SearchForTriangulation (src/ORBmatcher.cc:404-419): best starts at
                              max_dist, accept d <= max_dist && d <= best, ties -> last */
} orb_tie_rule;

/* Windowed search: query i scans train rows cand[offsets[i] .. offsets[i+1]) in that order
 * (candidate lists from Frame::GetFeaturesInArea src/Frame.cc:307-360, DBoW2 node buckets,
 * ...).  max_dist is used by ORB_TIE_LAST_MIN only. */
int orb_match_csr(orb_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt,
                  const int32_t* offsets, const int32_t* cand, int tie_rule, int max_dist,
                  int32_t* best_idx, int32_t* best_dist, int32_t* second_dist);

/* Every candidate distance of a windowed search: dist[c] = DescriptorDistance(query i, train cand[c])
 * for c in [offsets[i], offsets[i+1]).  For the searches whose candidate eligibility depends on
 * earlier accepts -- vbMatched2 in SearchByBoW (src/ORBmatcher.cc:316,331), vMatchedDistance in
 * SearchForInitialization (:224), claimed features in SearchByProjection (:38,59) -- the adapter
 * replays the reference's sequential scan over these numbers on the host. */
int orb_distances_csr(orb_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt,
                      const int32_t* offsets, const int32_t* cand, int32_t* dist);

/* Hamming part of Frame::ComputeStereoMatches (src/Frame.cc:446-529): row-band candidate
 * table, octave and disparity gates, first minimum below TH_HIGH=100.  best_r[i] = right
 * keypoint index or -1; best_dist[i] = its distance (100 when none).  scale = the
 * extractor's mvScaleFactor (nlevels floats); rows = image rows; bf, fx from the YAML. */
int orb_stereo_match(orb_matcher* m, const orb_keypoint* kps_left, const uint8_t* desc_left,
                     int n_left, const orb_keypoint* kps_right, const uint8_t* desc_right,
                     int n_right, const float* scale, int nlevels, int rows, float bf, float fx,
                     int32_t* best_r, int32_t* best_dist);

/* Frame::ComputeStereoMatches whole (src/Frame.cc:446-619): the Hamming search above, the 11x11 SAD sliding
 * window on the pyramid level of the left keypoint (:531-575), the parabola fit (:577-586), the disparity
 * gate (:588-603) and the median * 2.1 cut (:606-619).  The pyramids are read where the two extractor
 * handles left them on the device (frame `frame_left` of ex_left's last call, `frame_right` of ex_right's;
 * the two may be the same handle holding the pair in one batch), so mvImagePyramid never travels to the host.
 * u_right / depth = Frame::mvuRight / mvDepth (n_left floats, -1 where there is no match).  Inputs on which
 * the reference itself faults (row band outside the image, SAD window outside its level) -> ORB_ERR_SHAPE. */
int orb_compute_stereo_matches(orb_matcher* m, orb_extractor* ex_left, int frame_left, orb_extractor* ex_right,
                               int frame_right, const orb_keypoint* kps_left, const uint8_t* desc_left, int n_left,
                               const orb_keypoint* kps_right, const uint8_t* desc_right, int n_right, float bf,
                               float fx, float* u_right, float* depth);

/* The same with Frame::mb given as the reference's function reads it (minZ = mb, maxD = mbf / minZ, src/Frame.cc:476-478)
 * instead of derived from fx: inside the stereo Frame constructor ComputeStereoMatches runs BEFORE the static fx and
 * `mb = mbf / fx` are assigned (:70 vs :86-94), so the drop-in member function (adapter/Frame_stereo_b200.cc) must take
 * mb from the object exactly as the reference does. */
int orb_compute_stereo_matches_mb(orb_matcher* m, orb_extractor* ex_left, int frame_left, orb_extractor* ex_right,
                                  int frame_right, const orb_keypoint* kps_left, const uint8_t* desc_left, int n_left,
                                  const orb_keypoint* kps_right, const uint8_t* desc_right, int n_right, float bf,
                                  float mb, float* u_right, float* depth);

/* Frame::ComputeStereoMatches for every stereo pair of ONE extractor batch (BASELINE config 2 / 3 batched): pair p = frames
 * (2p, 2p + 1) of ex's last call (left = even, right = odd frame), keypoints / descriptors / counts in the batch layout of
 * orb_extract_batch ([frame][cap] rows).  u_right / depth: [npairs][cap] floats (Frame::mvuRight / mvDepth of pair p in row p).
 * on_device != 0: every pointer is a device pointer (the outputs of orb_extract_batch_device, untouched), asynchronous on the
 * matcher's stream; status (optional, npairs ints) receives a non-zero flag word for a pair on which the reference faults
 * (its rows are then all -1).  on_device == 0: host buffers, synchronous; status (optional) receives ORB_OK / ORB_ERR_SHAPE /
 * ORB_ERR_INVALID per pair, without it the first refused pair is the call's error.  The four steps (row-band table, Hamming
 * search, SAD refinement, median cut) run as four launches over all pairs. */
int orb_compute_stereo_matches_batch(orb_matcher* m, orb_extractor* ex, int npairs, const orb_keypoint* kps, const uint8_t* desc,
                                     int cap, const int32_t* counts, float bf, float fx, float* u_right, float* depth,
                                     int32_t* status, int on_device);

int orb_matcher_sync(orb_matcher* m);
void* orb_matcher_stream(orb_matcher* m);

/* ---- the reference's search methods, whole: windows and distances on the GPU, the method's own
 *      accept rule and greedy state replayed in query order on the host ----------------------
 * Everything a method needs from Frame / KeyFrame / MapPoint objects arrives as plain arrays; the
 * adapter gathers them.  3-D projection of map points (cv::Mat pose arithmetic)
 * stays with the caller: queries arrive as projected pixel positions. */

/* What Frame::GetFeaturesInArea (src/Frame.cc:307-360) and KeyFrame::GetFeaturesInArea
 * (src/KeyFrame.cc:549-588) read.  The 64 x 48 grid of Frame::AssignFeaturesToGrid
 * (src/Frame.cc:210-225, PosInGrid :362-372) is rebuilt on the device from keys_un. */
typedef struct orb_frame_view {
    const orb_keypoint* keys_un; /* mvKeysUn, n entries */
    const uint8_t* desc;         /* mDescriptors, n x 32 */
    int32_t n;
    float min_x, min_y;           /* mnMinX, mnMinY */
    float grid_w_inv, grid_h_inv; /* mfGridElementWidthInv, mfGridElementHeightInv */
} orb_frame_view;

/* GetFeaturesInArea(x[i], y[i], r[i], min_level[i], max_level[i]) for nq queries, candidate order as
 * in the reference (cells ix outer / iy inner, ascending index inside a cell), plus
 * dist[c] = DescriptorDistance(qdesc row i, F.desc row cand[c]).  min_level / max_level may be NULL
 * (= -1: no level check, the KeyFrame form).  offsets gets nq+1 entries; *total the candidate count.
 * If total > cap: ORB_ERR_CAPACITY, offsets and *total are valid, cand/dist untouched. */
int orb_window_search(orb_matcher* m, const orb_frame_view* F, int nq, const uint8_t* qdesc, const float* x,
                      const float* y, const float* r, const int32_t* min_level, const int32_t* max_level,
                      int32_t* offsets, int32_t* cand, int32_t* dist, int cap, int* total);

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:19-65).
 * Queries = the map points that pass `:24` (non-null, not bad, mbTrackInView), in vector order:
 * proj_x/proj_y/proj_xr = mTrackProjX/Y/XR, level = mnTrackScaleLevel, view_cos = mTrackViewCos,
 * q_observed[i] = (Observations() > 0) of that point (NULL: all observed).  u_right = F.mvuRight
 * (NULL: monocular, all -1).  occupied[idx] (in/out) = F.mvpMapPoints[idx] && Observations() > 0 (:38).
 * feature_of_query[i] = index the point was written to (F.mvpMapPoints[best] = pMP, :59) or -1. */
int orb_search_by_projection_map(orb_matcher* m, const orb_frame_view* F, const float* u_right, uint8_t* occupied,
                                 const float* scale_factors, int nlevels, int nq, const uint8_t* qdesc,
                                 const float* proj_x, const float* proj_y, const float* proj_xr, const int32_t* level,
                                 const float* view_cos, const uint8_t* q_observed, float th, float nnratio,
                                 int32_t* feature_of_query, int* nmatches);

typedef enum orb_rot_mode {
    ORB_ROT_NONE = 0,   /* no rotation histogram (mbCheckOrientation false, or the method has none) */
    ORB_ROT_WRAP = 1,   /* rot < 0 -> rot += 360; bin = round(rot/30) % 30 (src/ORBmatcher.cc:873-876) */
    ORB_ROT_NOWRAP = 2  /* SearchByProjection(Current, Last): no wrap (:796-797); a negative bin
                           indexes rotHist out of bounds in the reference -> ORB_ERR_SHAPE here */
} orb_rot_mode;

/* The best-only projection searches: SearchByProjection(Frame&, const Frame&, th, bMono) (:732-818),
 * SearchByProjection(Frame&, KeyFrame*, set&, th, ORBdist) (:820-894), SearchByProjection(KeyFrame*, Scw,
 * ...) (:121-195), SearchBySim3 (:636-730) and the search part of Fuse (:504-634).  Query i: window
 * GetFeaturesInArea(u, v, radius, min_level, max_level); candidates whose claimed[idx] != 0 are skipped
 * (claimed == NULL: no such state -- SearchBySim3 / Fuse); first minimum; accepted iff best <= max_dist;
 * an accept sets claimed[best].  q_angle / rot_mode: rotation-consistency histogram over
 * q_angle[i] - F.keys_un[best].angle ... the reference uses mvKeys angles, which equal mvKeysUn's
 * (UndistortKeyPoints copies the keypoint and moves pt only, src/Frame.cc:384-414).  Matches in the discarded bins are reset
 * (feature_of_query = -1, claimed cleared) and subtracted from *nmatches. */
int orb_search_by_projection_best(orb_matcher* m, const orb_frame_view* F, uint8_t* claimed, int nq,
                                  const uint8_t* qdesc, const float* u, const float* v, const float* radius,
                                  const int32_t* min_level, const int32_t* max_level, const float* q_angle,
                                  int rot_mode, int max_dist, int32_t* feature_of_query, int* nmatches);

/* ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:197-276).  keys1/desc1/n1 = F1.mvKeysUn /
 * mDescriptors; prev_matched = vbPrevMatched as n1 (x, y) float pairs, updated like :270-273;
 * matches12 = vnMatches12 (n1 ints). */
int orb_search_for_initialization(orb_matcher* m, const orb_keypoint* keys1, const uint8_t* desc1, int n1,
                                  const orb_frame_view* F2, float* prev_matched, int window_size, float nnratio,
                                  int check_ori, int32_t* matches12, int* nmatches);

/* DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>) as CSR: ascending node ids, and for
 * node k the feature indices idx[off[k] .. off[k+1]) in stored order. */
typedef struct orb_feature_vector {
    const int32_t* nodes;
    const int32_t* off;
    const int32_t* idx;
    int32_t n_nodes;
} orb_feature_vector;

/* ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (src/ORBmatcher.cc:278-366).
 * has_mp1/has_mp2[i] = the keypoint holds a map point that is not bad (:306-307, :315-316); angle =
 * mvKeysUn[i].angle.  matches12[i1] = i2 (the caller maps it to vpMapPoints2[i2]) or -1. */
int orb_search_by_bow(orb_matcher* m, const uint8_t* desc1, const float* angle1, const uint8_t* has_mp1, int n1,
                      const uint8_t* desc2, const float* angle2, const uint8_t* has_mp2, int n2,
                      const orb_feature_vector* fv1, const orb_feature_vector* fv2, float nnratio, int check_ori,
                      int32_t* matches12, int* nmatches);

/* ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:368-467) with bOnlyStereo = false, the only
 * way the fork calls it (src/LocalMapping.cc:187).  keys = mvKeysUn; has_mp = GetMapPoint(i) != NULL;
 * f12 = the 3x3 fundamental matrix row-major; sigma2 = KF2's mvLevelSigma2 (nlevels floats).
 * matches12[i1] = i2 or -1 (vMatchedPairs = the pairs with i2 >= 0 in i1 order, :456-462). */
int orb_search_for_triangulation(orb_matcher* m, const orb_keypoint* keys1, const uint8_t* desc1,
                                 const uint8_t* has_mp1, int n1, const orb_keypoint* keys2, const uint8_t* desc2,
                                 const uint8_t* has_mp2, int n2, const orb_feature_vector* fv1,
                                 const orb_feature_vector* fv2, const float* f12, const float* sigma2, int nlevels,
                                 int check_ori, int32_t* matches12, int* nmatches);

/* ---- DBoW2 vocabulary transform: replaces ORBVocabulary::transform as called by Frame::ComputeBoW
 *      (src/Frame.cc:375-382) and KeyFrame::ComputeBoW (src/KeyFrame.cc:39-47) ------------------------
 * The vocabulary tree (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h, m_nodes) arrives flattened: node 0 is the
 * root; node i has children children[child_off[i] .. child_off[i+1]) in m_nodes[i].children order, a 32-byte
 * descriptor, a weight and (leaves) a word id.  It is uploaded once and stays resident on the device. */
typedef struct orb_vocabulary orb_vocabulary; /* opaque */

typedef enum orb_voc_weighting { ORB_VOC_TF_IDF = 0, ORB_VOC_TF = 1, ORB_VOC_IDF = 2, ORB_VOC_BINARY = 3 } orb_voc_weighting; /* BowVector.h:36-42 */
typedef enum orb_voc_scoring {  /* BowVector.h:45-53; decides the normalisation (ScoringObject.h:74-89) */
    ORB_VOC_L1_NORM = 0, ORB_VOC_L2_NORM = 1, ORB_VOC_CHI_SQUARE = 2, ORB_VOC_KL = 3, ORB_VOC_BHATTACHARYYA = 4, ORB_VOC_DOT_PRODUCT = 5
} orb_voc_scoring;

/* depth_l = m_L.  node_weight: n_nodes doubles; node_word: n_nodes ints (meaningful for leaves). */
int orb_vocabulary_create(int device, int n_nodes, const int32_t* child_off, const int32_t* children,
                          const uint8_t* node_desc, const double* node_weight, const int32_t* node_word, int depth_l,
                          int weighting, int scoring, orb_vocabulary** out);
void orb_vocabulary_destroy(orb_vocabulary* v);

/* transform(features, BowVector&, FeatureVector&, levelsup) (TemplatedVocabulary.h:1126-1187): the tree descent
 * of all n descriptors runs on the device; the two std::map results are assembled on the host in feature order
 * with the reference's double arithmetic.  BowVector -> (bow_ids ascending, bow_values), *bow_n entries;
 * FeatureVector -> orb_feature_vector layout (fv_nodes ascending, fv_off, fv_idx), *fv_n nodes.
 * word_of_feature / node_of_feature (n ints each, may be NULL): per-feature word id and node id.
 * Capacities too small -> ORB_ERR_CAPACITY with *bow_n / *fv_n holding the needed sizes (fv_idx needs n). */
int orb_vocabulary_transform(orb_vocabulary* v, const uint8_t* desc, int n, int levelsup, int32_t* word_of_feature,
                             int32_t* node_of_feature, int32_t* bow_ids, double* bow_values, int bow_cap, int* bow_n,
                             int32_t* fv_nodes, int32_t* fv_off, int32_t* fv_idx, int fv_cap, int* fv_n);

/* ---- misc ------------------------------------------------------------------------------- */

/* Page-locked host memory for the caller's frame and result buffers: copies from / to it run at full PCIe rate and
 * asynchronously, pageable memory is staged by the driver.  (cudaHostAlloc / cudaFreeHost behind a C name, so that the
 * adapter classes and other callers need not link the CUDA runtime themselves.) */
int orb_host_alloc(size_t bytes, void** out);
void orb_host_free(void* p);

/* Thread-local description of the last error on this thread ("" if none). */
const char* orb_last_error(void);
/* Number of kernels this library has launched in this process (for bench accounting). */
uint64_t orb_kernel_launch_count(void);
const char* orb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ORB_B200_H */
