// Launch interface of the Hamming-search kernels (match_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct orb_kp28 {  // cv::KeyPoint layout
    float x, y, size, angle, response;
    int octave, class_id;
};

// Pyramid levels of a stereo pair as the refinement kernel reads them (device pointers, level l of the left and
// the right image have the same shape and pitch).
struct OrbStereoLevels {
    const uint8_t* left[16];
    const uint8_t* right[16];
    int pitch[16], rows[16], cols[16];
    float scale[16], inv_scale[16];
    unsigned long long plane[16];  // bytes between the frames of a level buffer (batched form; unused by the single-pair kernels)
};

// Function attributes of the matcher kernels for the current device (called once per matcher handle).
cudaError_t orbk_match_init_device();
// max_nt: largest train set, or -1 when the counts live on the device only.
cudaError_t orbk_match_all(const uint8_t* q, const int* nq, size_t q_stride, const uint8_t* t, const int* nt,
                           size_t t_stride, int npairs, int max_nq, int max_nt, int* best_idx, int* best_dist, int* second_dist,
                           size_t out_stride, cudaStream_t st);
cudaError_t orbk_match_csr(const uint8_t* q, int nq, const uint8_t* t, const int* offsets, const int* cand, int tie_last,
                           int max_dist, int* best_idx, int* best_dist, int* second_dist, cudaStream_t st);
cudaError_t orbk_dist_csr(const uint8_t* q, int nq, const uint8_t* t, const int* offsets, const int* cand, int* out, cudaStream_t st);
cudaError_t orbk_stereo(const orb_kp28* kl, const uint8_t* dl, int nl, const orb_kp28* kr, const uint8_t* dr, int nr,
                        const float* d_scale, int4* d_rinfo, float maxD, int* best_r, int* best_dist, cudaStream_t st);
// Frame grid (Frame::AssignFeaturesToGrid) and Frame::GetFeaturesInArea windows with their distances.
// cell_start: 64*48+1 ints, members: n ints.  window_count writes counts[nq] and offsets[nq+1].
cudaError_t orbk_grid_assign(const orb_kp28* keys, int n, float minX, float minY, float wInv, float hInv, int* cell_start, int* members,
                             cudaStream_t st);
cudaError_t orbk_window_count(const orb_kp28* keys, const int* cell_start, const int* members, float minX, float minY, float wInv,
                              float hInv, int nq, const float* qx, const float* qy, const float* qr, const int* qmin, const int* qmax,
                              int* counts, int* offsets, cudaStream_t st);
cudaError_t orbk_window_fill(const orb_kp28* keys, const uint8_t* tdesc, const int* cell_start, const int* members, float minX,
                             float minY, float wInv, float hInv, int nq, const uint8_t* qdesc, const float* qx, const float* qy,
                             const float* qr, const int* qmin, const int* qmax, const int* offsets, int* cand, int* dist,
                             cudaStream_t st);
// Frame::ComputeStereoMatches after the Hamming search (src/Frame.cc:531-603): SAD refinement, parabola, disparity gate.
cudaError_t orbk_stereo_refine(const orb_kp28* kl, int nl, const orb_kp28* kr, const int* best_r, const int* best_dist,
                               const OrbStereoLevels& lv, float mbf, float maxD, float* u_right, float* depth, int* sad, int* flags,
                               cudaStream_t st);
// Frame::ComputeStereoMatches for npairs stereo pairs of one extractor batch (pair p = frames 2p, 2p + 1; kps / desc in the
// batch layout [frame][cap]); scratch: d_rinfo, d_best_r, d_best_dist, d_sad [npairs][cap], d_flags [npairs].
cudaError_t orbk_stereo_batch(const orb_kp28* kps, const uint8_t* desc, const int* counts, int cap, int npairs, int nlevels, int rows,
                              const float* d_scale, const OrbStereoLevels& lv, float mbf, float maxD, int4* d_rinfo, int* d_best_r,
                              int* d_best_dist, int* d_sad, int* d_flags, float* u_right, float* depth, cudaStream_t st);
// DBoW2 TemplatedVocabulary::transform descent for n features: leaf node and the node at level nid_level of each.
cudaError_t orbk_voc_descent(const uint8_t* feat, int n, const int* child_off, const int* children, const uint8_t* node_desc,
                             int nid_level, int* leaf_node, int* node_at_level, cudaStream_t st);
