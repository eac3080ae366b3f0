// Launch interface of the extractor kernels (extract_kernels.cu) used by the C ABI layer.
#pragma once
#include <cuda.h>
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
// Two streams: `st` runs pyramid -> detect -> octree -> describe, `st2` runs the blur (which only
// needs the pyramid) concurrently with detect + octree; `fork` / `join` are the events that order
// them.  `ev` (optional, ORB_EVENTS entries): ev[0..5] on `st` before the first stage and after
// pyramid, detect, octree, the join, describe; ev[6], ev[7] around the blur on `st2`.
#define ORB_STAGES 5
#define ORB_EVENTS 8
// TMA descriptors of the level images (x, y, frame) for k_detect's tile staging: encoded on the
// host by orbk_encode_level_map, copied to device global memory, read by the TMA unit from there.
struct DetectMaps {
    CUtensorMap m[ORB_MAX_LEVELS];     // raw level, 256 x boxH box (k_detect)
    CUtensorMap raw[ORB_MAX_LEVELS];   // raw level, DSC_BOX_W x DSC_BOX_H box (k_describe_tile: IC_Angle)
    CUtensorMap blur[ORB_MAX_LEVELS];  // blurred level, same box (k_describe_tile: rBRIEF)
    CUtensorMap blr[ORB_MAX_LEVELS];   // raw level, 256 x 70 box (k_blur: 224 x 64 tile + halo)
    CUtensorMap rsz[ORB_MAX_LEVELS];   // entry l: the SOURCE image of level l (level l-1), 256 x RSZ_BOX_H box (k_resize_tile)
};
// Encodes the (cols x rows x frames) uint8 tensor of one level with a boxW x boxH x 1 box.
// Returns cudaSuccess or an error (the driver entry point is looked up at run time).
cudaError_t orbk_encode_level_map(CUtensorMap* out, const uint8_t* base, int cols, int rows, int frames, int pitch,
                                  unsigned long long plane, int boxW, int boxH);

struct OrbStreams {
    cudaStream_t st, st2;
    cudaEvent_t fork, join;
    cudaEvent_t pyr;  // the pyramid is complete (recorded on st2 when it builds the pyramid next to the level-0 detect)
};
cudaError_t orbk_run_extract(const OrbPlan& plan, int nframes, orb_keypoint_dev* d_kps, uint8_t* d_desc, int cap,
                             int* d_counts, const OrbStreams& ss, const DetectMaps* d_maps, cudaEvent_t* ev = nullptr);
void orbk_build_ic_table(int2* out /* 8*32 */);
void orbk_build_pair_table(float4* out /* 182 */);
// Dense frames (row stride == cols) frame0 .. frame0 + nframes of the landing buffer `dense` (4-byte aligned base of
// the whole buffer) -> pitched level-0 layout at dst.
cudaError_t orbk_repitch(const uint8_t* dense, int frame0, int nframes, int rows, int cols, uint8_t* dst, int pitch,
                         unsigned long long plane, cudaStream_t st);
unsigned long long orbk_launch_count();
void orbk_count_launch(int n);
// Ingest fused into the level-0 load: cv::remap (INTER_LINEAR, CV_32FC1 maps; mapx == NULL: none) and / or
// cv::cvtColor to gray of nframes raw frames, written into the pitched level-0 buffer.
cudaError_t orbk_ingest(const uint8_t* raw, int nframes, int srows, int scols, size_t sstride, size_t sframe, int channels, int bgr,
                        int variant, const float* mapx, const float* mapy, int drows, int dcols, uint8_t* dst, int dpitch,
                        unsigned long long dplane, cudaStream_t st);
