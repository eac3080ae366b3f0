// C entry points over the reference's own ORB_SLAM2::Frame, compiled UNMODIFIED from /root/reference/src/Frame.cc against
// oracle/mshim/frame_objects.h -- TEST INFRASTRUCTURE ONLY.
//   refm_compute_stereo_matches: the stereo constructor Frame(imLeft, imRight, ...) (src/Frame.cc:41-97) runs ExtractORB (the
//     stub extractors hand out the given keypoints / descriptors / pyramids), UndistortKeyPoints (no distortion),
//     ComputeStereoMatches (:446-619) and AssignFeaturesToGrid; mvuRight / mvDepth are read back.
//   refm_frame_features_in_area: the monocular constructor (:154-208: ExtractORB, UndistortKeyPoints, ComputeImageBounds,
//     AssignFeaturesToGrid), then Frame::GetFeaturesInArea (:307-360) per query.
// Frame::mb is not initialised before ComputeStereoMatches reads it (`mb = mbf / fx` comes after the call, :94; in the
// running system the temporary Frame lands on the stack slot of the previous one and finds its mb there): the bridge
// constructs the Frame in storage whose mb already holds mbf / fx.
#include "Frame.h"

#include <cstddef>
#include <cstring>
#include <memory>
#include <new>

using namespace ORB_SLAM2;
typedef orb_oracle::KeyPoint OKP;

namespace {
cv::Mat image_mat(const uint8_t* p, int rows, int cols, int step) {
    cv::Mat m(rows, cols, CV_8U);
    for (int y = 0; y < rows; ++y) memcpy(m.data + (size_t)y * m.step, p + (size_t)y * step, cols);
    return m;
}
void load(ORBextractor& ex, const OKP* k, const uint8_t* d, int n, const float* scale, int nlevels) {
    for (int i = 0; i < n; ++i) {
        cv::KeyPoint c;
        c.pt = cv::Point2f(k[i].x, k[i].y);
        c.size = k[i].size;
        c.angle = k[i].angle;
        c.response = k[i].response;
        c.octave = k[i].octave;
        c.class_id = k[i].class_id;
        ex.keys.push_back(c);
    }
    ex.desc = cv::Mat(std::max(n, 1), 32, CV_8U);
    if (n) memcpy(ex.desc.data, d, (size_t)n * 32);
    for (int l = 0; l < nlevels; ++l) {
        ex.scale.push_back(scale[l]);
        ex.invScale.push_back(1.0f / scale[l]);
        ex.sigma2.push_back(scale[l] * scale[l]);
        ex.invSigma2.push_back(1.0f / (scale[l] * scale[l]));
    }
}
struct FrameBox {  // a Frame constructed in storage whose `mb` was pre-set
    alignas(Frame) unsigned char raw[sizeof(Frame)];
    Frame* f = nullptr;
    ~FrameBox() {
        if (f) f->~Frame();
    }
};
}  // namespace

extern "C" {

// pyrL / pyrR: the nlevels pyramid levels of the two extractors, level l = rows[l] x cols[l] bytes, dense, back to back.
int refm_compute_stereo_matches(const OKP* kl, const uint8_t* dl, int nl, const OKP* kr, const uint8_t* dr, int nr, const float* scale,
                                int nlevels, const uint8_t* pyrL, const uint8_t* pyrR, const int* rows, const int* cols, float bf,
                                float fx, float* uRight, float* depth) {
    ORBextractor exL, exR;
    load(exL, kl, dl, nl, scale, nlevels);
    load(exR, kr, dr, nr, scale, nlevels);
    size_t off = 0;
    for (int l = 0; l < nlevels; ++l) {
        exL.mvImagePyramid.push_back(image_mat(pyrL + off, rows[l], cols[l], cols[l]));
        exR.mvImagePyramid.push_back(image_mat(pyrR + off, rows[l], cols[l], cols[l]));
        off += (size_t)rows[l] * cols[l];
    }
    cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
    K.at<float>(0, 0) = fx;
    K.at<float>(1, 1) = fx;
    K.at<float>(0, 2) = cols[0] * 0.5f;
    K.at<float>(1, 2) = rows[0] * 0.5f;
    cv::Mat dist = cv::Mat::zeros(4, 1, CV_32F);
    ORBVocabulary voc;
    Frame::mbInitialComputations = true;
    FrameBox box;
    memset(box.raw, 0, sizeof box.raw);
    const float mb = bf / fx;
    memcpy(box.raw + offsetof(Frame, mb), &mb, sizeof mb);
    try {
        box.f = new (box.raw) Frame(exL.mvImagePyramid[0], exR.mvImagePyramid[0], 0.0, &exL, &exR, &voc, K, dist, bf, 40.0f);
    } catch (const std::exception&) {
        return -2;
    }
    for (int i = 0; i < nl; ++i) {
        uRight[i] = box.f->mvuRight[i];
        depth[i] = box.f->mvDepth[i];
    }
    return 0;
}

#ifdef ADAPTER_FRAME
// The adapter arm (oracle/Makefile adapterframe): the reference's stereo constructor, compiled unmodified, with
// Frame::ComputeStereoMatches supplied by orb_slam_system_b200/adapter/Frame_stereo_b200.cc (the reference's own definition
// is weakened in the object file).  Two C-ABI extractors extract the two images on the GPU -- their pyramids stay on the
// device --, the stub extractors hand the resulting keypoints / descriptors to the constructor and carry the handles.
// Returns the number of left keypoints (<= cap) or a negative error; kl_out / dl_out receive them.
}  // extern "C"
#include "orb_b200.h"
extern "C" {
int adpf_stereo_frame(const uint8_t* imgL, const uint8_t* imgR, int rows, int cols, int nfeatures, float scaleFactor, int nlevels,
                      int iniTh, int minTh, float bf, float fx, float* uRight, float* depth, OKP* kl_out, uint8_t* dl_out, int cap) {
    orb_params prm = {nfeatures, scaleFactor, nlevels, iniTh, minTh};
    orb_extractor *gl = nullptr, *gr = nullptr;
    if (orb_extractor_create(&prm, rows, cols, 1, 0, &gl) != ORB_OK || orb_extractor_create(&prm, rows, cols, 1, 0, &gr) != ORB_OK) return -4;
    int bound = 0;
    orb_extractor_keypoint_bound(gl, rows, cols, &bound);
    std::vector<orb_keypoint> kl(bound), kr(bound);
    std::vector<uint8_t> dl((size_t)bound * 32), dr((size_t)bound * 32);
    int nl = 0, nr = 0, rc = -1;
    if (orb_extract(gl, imgL, rows, cols, cols, kl.data(), dl.data(), bound, &nl) == ORB_OK &&
        orb_extract(gr, imgR, rows, cols, cols, kr.data(), dr.data(), bound, &nr) == ORB_OK && nl <= cap) {
        std::vector<float> scale(nlevels);
        orb_extractor_tables(gl, scale.data(), nullptr, nullptr, nullptr, nullptr);
        ORBextractor exL, exR;
        static_assert(sizeof(OKP) == sizeof(orb_keypoint), "keypoint layouts");
        load(exL, reinterpret_cast<const OKP*>(kl.data()), dl.data(), nl, scale.data(), nlevels);
        load(exR, reinterpret_cast<const OKP*>(kr.data()), dr.data(), nr, scale.data(), nlevels);
        exL.gpu = gl;
        exR.gpu = gr;
        cv::Mat im = image_mat(imgL, rows, cols, cols);
        cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
        K.at<float>(0, 0) = fx;
        K.at<float>(1, 1) = fx;
        K.at<float>(0, 2) = cols * 0.5f;
        K.at<float>(1, 2) = rows * 0.5f;
        cv::Mat dist = cv::Mat::zeros(4, 1, CV_32F);
        ORBVocabulary voc;
        Frame::mbInitialComputations = true;
        FrameBox box;
        memset(box.raw, 0, sizeof box.raw);
        const float mb = bf / fx;  // as in refm_compute_stereo_matches: the storage's mb holds the running system's value
        memcpy(box.raw + offsetof(Frame, mb), &mb, sizeof mb);
        try {
            box.f = new (box.raw) Frame(im, im, 0.0, &exL, &exR, &voc, K, dist, bf, 40.0f);
            for (int i = 0; i < nl; ++i) {
                uRight[i] = box.f->mvuRight[i];
                depth[i] = box.f->mvDepth[i];
            }
            memcpy(kl_out, kl.data(), (size_t)nl * sizeof(OKP));
            memcpy(dl_out, dl.data(), (size_t)nl * 32);
            rc = nl;
        } catch (const std::exception&) {
            rc = -2;
        }
    }
    orb_extractor_destroy(gl);
    orb_extractor_destroy(gr);
    return rc;
}
#endif

// Frame::AssignFeaturesToGrid + Frame::GetFeaturesInArea of the reference on a frame of `cols` x `rows` pixels with the given
// (undistorted = raw) keypoints; CSR result like orc_features_in_area.  Returns the total number of candidates.
int refm_frame_features_in_area(const OKP* keys, int n, int rows, int cols, int nq, const float* x, const float* y, const float* r,
                                const int* minLevel, const int* maxLevel, int* offsets, int* cand, int cap) {
    ORBextractor exL, exR;
    const float scale[8] = {1.f, 1.f, 1.2f, 1.44f, 1.728f, 2.0736f, 2.48832f, 2.985984f};
    std::vector<uint8_t> zeros((size_t)std::max(n, 1) * 32, 0);
    load(exL, keys, zeros.data(), n, scale, 8);
    load(exR, keys, zeros.data(), 0, scale, 8);
    cv::Mat img(rows, cols, CV_8U);
    for (int l = 0; l < 8; ++l) {
        exL.mvImagePyramid.push_back(img);
        exR.mvImagePyramid.push_back(img);
    }
    cv::Mat K = cv::Mat::eye(3, 3, CV_32F), dist = cv::Mat::zeros(4, 1, CV_32F);
    ORBVocabulary voc;
    Frame::mbInitialComputations = true;
    FrameBox box;
    memset(box.raw, 0, sizeof box.raw);
    const float mb = 1.0f;
    memcpy(box.raw + offsetof(Frame, mb), &mb, sizeof mb);
    box.f = new (box.raw) Frame(img, 0.0, &exL, &voc, K, dist, 1.0f, 40.0f);  // the monocular constructor (:154-208): no stereo step
    int total = 0;
    offsets[0] = 0;
    for (int i = 0; i < nq; ++i) {
        const std::vector<size_t> v = box.f->GetFeaturesInArea(x[i], y[i], r[i], minLevel ? minLevel[i] : -1, maxLevel ? maxLevel[i] : -1);
        for (size_t idx : v) {
            if (total < cap) cand[total] = (int)idx;
            ++total;
        }
        offsets[i + 1] = total;
    }
    return total;
}
}
