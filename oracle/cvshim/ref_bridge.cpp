// C entry points over the reference's own ORB_SLAM2::ORBextractor, compiled from
// /root/reference/src/ORBextractor.cc against cvshim.hpp -- TEST INFRASTRUCTURE ONLY.
#include <chrono>
#include <atomic>
#include <thread>

#include "ORBextractor.h"  // the reference's header (include path set by oracle/Makefile)

extern "C" {

// ORBextractor::operator() of the reference.  Returns 0, -2 if cv::Exception was thrown.
int ref_extract(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const uint8_t* img,
                int rows, int cols, int step, orb_oracle::KeyPoint* kps, uint8_t* desc, int cap,
                int* count) {
    try {
        ORB_SLAM2::ORBextractor ex(nfeatures, scaleFactor, nlevels, iniTh, minTh);
        cv::Mat im(rows, cols, CV_8UC1);
        for (int y = 0; y < rows; ++y) memcpy(im.data + (size_t)y * im.step, img + (size_t)y * step, cols);
        std::vector<cv::KeyPoint> k;
        cv::Mat d;
        ex(im, cv::Mat(), k, d);
        int n = d.empty() ? 0 : (int)k.size();
        *count = n;
        if (n > cap) return -10;
        for (int i = 0; i < n; ++i) {
            kps[i] = orb_oracle::KeyPoint{k[i].pt.x, k[i].pt.y, k[i].size, k[i].angle, k[i].response,
                                          k[i].octave, k[i].class_id};
            memcpy(desc + 32 * (size_t)i, d.ptr<uint8_t>(i), 32);
        }
        return 0;
    } catch (const cv::Exception&) {
        return -2;
    } catch (const std::exception&) {  // e.g. std::length_error from vector(nIni<0)
        return -1;
    }
}

// Getter tables of the reference ctor (ORBextractor.h:43-63).
int ref_tables(int nfeatures, float scaleFactor, int nlevels, float* scale, float* inv_scale, float* sigma2,
               float* inv_sigma2) {
    ORB_SLAM2::ORBextractor ex(nfeatures, scaleFactor, nlevels, 20, 7);
    auto a = ex.GetScaleFactors(), b = ex.GetInverseScaleFactors(), c = ex.GetScaleSigmaSquares(),
         d = ex.GetInverseScaleSigmaSquares();
    for (int i = 0; i < nlevels; ++i) {
        scale[i] = a[i];
        inv_scale[i] = b[i];
        sigma2[i] = c[i];
        inv_sigma2[i] = d[i];
    }
    return ex.GetLevels();
}

// One padded pyramid level (mvImagePyramid[level] is the ROI inside a 19-px border).
int ref_pyramid_level(float scaleFactor, int nlevels, const uint8_t* img, int rows, int cols, int step,
                      int level, uint8_t* dst, int cap, int* lrows, int* lcols) {
    ORB_SLAM2::ORBextractor ex(1000, scaleFactor, nlevels, 20, 7);
    cv::Mat im(rows, cols, CV_8UC1);
    for (int y = 0; y < rows; ++y) memcpy(im.data + (size_t)y * im.step, img + (size_t)y * step, cols);
    std::vector<cv::KeyPoint> k;
    cv::Mat d;
    ex(im, cv::Mat(), k, d);
    const cv::Mat& l = ex.mvImagePyramid[level];
    *lrows = l.rows;
    *lcols = l.cols;
    if ((size_t)l.rows * l.cols > (size_t)cap) return -10;
    for (int y = 0; y < l.rows; ++y) memcpy(dst + (size_t)y * l.cols, l.ptr<uint8_t>(y), l.cols);
    return 0;
}

// CPU baseline: the reference extractor, one instance and one frame at a time per thread
// (the reference's own concurrency model is one instance per camera, Frame.cc:58-61).
double ref_extract_many(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, int rows,
                        int cols, uint64_t seed, int first_frame, int nframes, int nthreads,
                        long long* total_keypoints) {
    std::vector<cv::Mat> frames(nframes);
    {
        std::atomic<int> next{0};
        std::vector<std::thread> th;
        for (int w = 0; w < nthreads; ++w)
            th.emplace_back([&] {
                for (int f; (f = next.fetch_add(1)) < nframes;) {
                    frames[f].create(rows, cols, CV_8UC1);
                    orb_oracle::synth_frame(frames[f].data, rows, cols, cols, seed, (uint64_t)(first_frame + f), 0, 0);
                }
            });
        for (auto& t : th) t.join();
    }
    std::atomic<int> next{0};
    std::atomic<long long> total{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int w = 0; w < nthreads; ++w)
        th.emplace_back([&] {
            ORB_SLAM2::ORBextractor ex(nfeatures, scaleFactor, nlevels, iniTh, minTh);
            for (int f; (f = next.fetch_add(1)) < nframes;) {
                std::vector<cv::KeyPoint> k;
                cv::Mat d;
                ex(frames[f], cv::Mat(), k, d);
                total += (long long)k.size();
            }
        });
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (total_keypoints) *total_keypoints = total.load();
    return std::chrono::duration<double>(t1 - t0).count();
}
}
