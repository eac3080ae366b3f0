// CPU oracle for the ORB front end -- TEST INFRASTRUCTURE ONLY.
//
// This is a restatement, in OpenCV-free C++, of what the reference fork
// (WangHewei16/ORB-SLAM-System) computes on its extract + describe + Hamming
// match path, plus exact integer/float models of the OpenCV primitives that
// path calls (OpenCV 3.4.15 in the reference build; cv2 4.13.0 is what the
// models are pinned against here, see tests/test_oracle_cv2.py and
// tests/golden/).  Every function cites the reference file:line it follows.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this.  The product (orb_slam_system_b200)
// never links, imports or calls anything in oracle/.
//
// Parity pinning: the reference ships no tests or golden vectors,
// so the oracle is pinned three ways: (1) every OpenCV primitive model against
// cv2 4.13.0 in this container, with the resulting vectors committed under
// tests/golden/; (2) the reference's own src/ORBextractor.cc compiled
// unmodified against oracle/cvshim (oracle/_ref/libref_orb.so) and compared
// end to end with this restatement; (3) the known-answer counts in SURVEY A.8.
#pragma once
#include <cstdint>
#include <vector>

namespace orb_oracle {

// Same 28-byte layout as cv::KeyPoint.
struct KeyPoint {
    float x, y;
    float size;
    float angle;
    float response;
    int octave;
    int class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct Image {  // non-owning 8-bit view
    const uint8_t* data;
    int rows, cols, step;
    const uint8_t* row(int y) const { return data + (size_t)y * step; }
};

// ---- OpenCV primitive models ----------------------
int cv_round_f(float v);   // cvRound(float): round-half-even
int cv_round_d(double v);  // cvRound(double)
float fast_atan2(float y, float x);  // cv::fastAtan2, A.4
// cv::resize(..., INTER_LINEAR) for 8UC1, A.2. dst is dcols x drows, dstep pitch.
void resize_linear_u8(const Image& src, uint8_t* dst, int drows, int dcols, int dstep);
// cv::GaussianBlur(Size(7,7), 2, 2, BORDER_REFLECT_101) for 8UC1, A.3. In-place safe.
void gaussian_blur7_u8(const Image& src, uint8_t* dst, int dstep);
// cv::FAST(img, kps, threshold, true) TYPE_9_16, A.1.  Appends (x, y, score) in
// row-major order; coordinates relative to the view.
struct FastPoint { int x, y, score; };
void fast9_16_nms(const Image& img, int threshold, std::vector<FastPoint>& out);
// FAST corner score m-1 of one pixel (needs a 3-px margin); 0 when m<=threshold.
int fast_score_at(const Image& img, int x, int y, int threshold);

// ---- extractor (reference src/ORBextractor.cc) --------------------------
struct Params {
    int nfeatures;
    float scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
};

struct Tables {  // what ORBextractor::ORBextractor builds (ORBextractor.cc:116-170)
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> features_per_level;
    std::vector<int> umax;
    std::vector<int8_t> pattern;  // 1024 ints: 728 from the table, rest 0
};
Tables make_tables(const Params& p);

// ORBextractor::DistributeOctTree (ORBextractor.cc:228-286) with DivideNode
// (:178-225).  Returns indices into `keys` in the reference's output order.
// Returns false (and an empty result) if the reference would not terminate.
struct Cand { float x, y; float response; };
bool distribute_octree(const std::vector<Cand>& keys, int minX, int maxX, int minY, int maxY,
                       int N, std::vector<int>& out_idx);

struct ExtractResult {
    std::vector<KeyPoint> keypoints;
    std::vector<uint8_t> descriptors;           // K x 32
    std::vector<std::vector<uint8_t>> pyramid;  // level images, tightly packed (no 19-px border)
    std::vector<int> level_rows, level_cols;
    std::vector<int> n_candidates;              // FAST candidates handed to the octree per level
    std::vector<int> n_keypoints;               // kept per level
    int n_retry_cells = 0;                      // cells that fell back to minThFAST
};
// ORBextractor::operator() (ORBextractor.cc:442-495). Returns 0 on success,
// <0 for inputs on which the reference has undefined behaviour (see .cpp).
int extract(const Params& p, const Image& img, ExtractResult& out, bool keep_pyramid);

// ---- matcher (reference src/ORBmatcher.cc, src/Frame.cc) -----------------
int descriptor_distance(const uint8_t* a, const uint8_t* b);  // ORBmatcher.cc:896-908

// ---- synthetic frames ---------------------------------------
void synth_frame(uint8_t* dst, int rows, int cols, int step, uint64_t seed, uint64_t frame,
                 int variant, int right);

}  // namespace orb_oracle

// ---- C ABI for ctypes (tests / smoke / cpu_baseline) -----------------------
extern "C" {
int orc_extract(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
                const uint8_t* img, int rows, int cols, int step, orb_oracle::KeyPoint* kps,
                uint8_t* desc, int cap, int* count, int* per_level_candidates,
                int* per_level_kept, int* n_retry_cells);
int orc_tables(int nfeatures, float scaleFactor, int nlevels, float* scale, float* inv_scale,
               float* sigma2, float* inv_sigma2, int* nfeat_per_level, int* umax16);
int orc_pyramid(float scaleFactor, int nlevels, const uint8_t* img, int rows, int cols, int step,
                int level, uint8_t* dst, int dst_cap, int* lrows, int* lcols);
void orc_resize(const uint8_t* src, int srows, int scols, int sstep, uint8_t* dst, int drows,
                int dcols, int dstep);
void orc_blur7(const uint8_t* src, int rows, int cols, int step, uint8_t* dst, int dstep);
int orc_fast(const uint8_t* img, int rows, int cols, int step, int threshold, int* xys, int cap);
float orc_fast_atan2(float y, float x);
int orc_octree(const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int* out_idx,
               int cap);
float orc_ic_angle(const uint8_t* img, int rows, int cols, int step, int x, int y);
void orc_describe(const uint8_t* blurred, int rows, int cols, int step, int x, int y, float angle,
                  uint8_t* desc32);
int orc_distance(const uint8_t* a, const uint8_t* b);
void orc_match_all(const uint8_t* q, int nq, const uint8_t* t, int nt, int* best_idx,
                   int* best_dist, int* second_dist);
void orc_match_csr(const uint8_t* q, int nq, const uint8_t* t, int nt, const int* offsets,
                   const int* cand, int tie_last, int max_dist, int* best_idx, int* best_dist,
                   int* second_dist);
int orc_stereo_match(const orb_oracle::KeyPoint* kl, const uint8_t* dl, int nl,
                     const orb_oracle::KeyPoint* kr, const uint8_t* dr, int nr, const float* scale,
                     int nlevels, int rows, float bf, float fx, int* best_r, int* best_dist);
int orc_search_by_bow_kf(const uint8_t* d1, const float* ang1, const uint8_t* has1, int n1, const uint8_t* d2, const float* ang2,
                         const uint8_t* has2, int n2, const int* nodes1, const int* off1, const int* idx1, int nn1, const int* nodes2,
                         const int* off2, const int* idx2, int nn2, float nnratio, int checkOri, int* matches12);
int orc_search_for_triangulation(const uint8_t* d1, const float* x1, const float* y1, const float* ang1, const uint8_t* has1, int n1,
                                 const uint8_t* d2, const float* x2, const float* y2, const float* ang2, const int* oct2,
                                 const uint8_t* has2, int n2, const int* nodes1, const int* off1, const int* idx1, int nn1,
                                 const int* nodes2, const int* off2, const int* idx2, int nn2, const float* F12, const float* sigma2,
                                 int checkOri, int* matches12);
// Image ingest models: cv::cvtColor RGB/BGR(A) -> gray and cv::remap INTER_LINEAR with CV_32FC1 maps (BORDER_CONSTANT 0).
void orc_cvt_gray(const uint8_t* src, int rows, int cols, int step, int channels, int bgr, int variant, uint8_t* dst, int dstep);
void orc_remap_linear(const uint8_t* src, int srows, int scols, int sstep, int channels, const float* mapx, const float* mapy,
                      int mstep, int drows, int dcols, uint8_t* dst, int dstep);
void orc_synth_frame(uint8_t* dst, int rows, int cols, int step, uint64_t seed, uint64_t frame,
                     int variant, int right);
// Multi-threaded CPU baseline: n frames of the synthetic generator, one frame
// per worker thread at a time; returns seconds spent extracting (generation excluded).
double orc_extract_many(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
                        int rows, int cols, uint64_t seed, int first_frame, int nframes,
                        int nthreads, long long* total_keypoints);
}
