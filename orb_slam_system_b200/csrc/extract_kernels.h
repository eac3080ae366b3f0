// Launch interface of the extractor kernels (extract_kernels.cu) used by the C ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "orb_plan.h"

// Device-side twin of orb_keypoint / cv::KeyPoint (28 bytes).
struct orb_keypoint_dev {
    float x, y, size, angle, response;
    int octave, class_id;
};

cudaError_t orbk_init_device();
// Runs pyramid -> detect -> octree -> blur -> describe for `nframes` frames on stream `st`.
// plan.lv[0].img must point at the level-0 frames.  Outputs: d_kps [nframes][cap],
// d_desc [nframes][cap][32], d_counts [nframes] (device memory).
// `ev` (optional): ORB_STAGES+1 events recorded before the first stage and after each stage.
#define ORB_STAGES 5
cudaError_t orbk_run_extract(const OrbPlan& plan, int nframes, orb_keypoint_dev* d_kps, uint8_t* d_desc, int cap,
                             int* d_counts, cudaStream_t st, cudaEvent_t* ev = nullptr);
void orbk_build_ic_table(int2* out /* 4*31*9 */);
unsigned long long orbk_launch_count();
void orbk_count_launch(int n);
