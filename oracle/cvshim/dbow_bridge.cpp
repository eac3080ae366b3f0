// C bridge to the reference's own DBoW2 (Thirdparty/DBoW2, compiled from where it lies against cvshim):
// TemplatedVocabulary<FORB::TDescriptor, FORB> = ORB_SLAM2::ORBVocabulary (include/ORBVocabulary.h).
// TEST INFRASTRUCTURE ONLY: pins the oracle's restatement of transform() and is never shipped.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> RefVocabulary;

extern "C" {

void* refvoc_load_text(const char* filename) {
    RefVocabulary* v = new RefVocabulary();
    if (!v->loadFromTextFile(filename)) {
        delete v;
        return nullptr;
    }
    return v;
}

void refvoc_free(void* v) { delete (RefVocabulary*)v; }

int refvoc_size(void* v) { return (int)((RefVocabulary*)v)->size(); }

// Frame::ComputeBoW (src/Frame.cc:375-382): Converter::toDescriptorVector (rows of mDescriptors) then
// transform(vCurrentDesc, mBowVec, mFeatVec, levelsup).  Outputs in map order.  Returns 0, or -1 if a capacity is too small.
int refvoc_transform(void* vp, const uint8_t* desc, int n, int levelsup, int* bow_ids, double* bow_vals, int bow_cap, int* bow_n,
                     int* fv_nodes, int* fv_off, int* fv_idx, int fv_cap, int* fv_n) {
    RefVocabulary* v = (RefVocabulary*)vp;
    std::vector<cv::Mat> feats(n);
    for (int i = 0; i < n; ++i) {
        feats[i].create(1, 32, CV_8U);
        memcpy(feats[i].data, desc + 32 * (size_t)i, 32);
    }
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    v->transform(feats, bv, fv, levelsup);
    *bow_n = (int)bv.size();
    *fv_n = (int)fv.size();
    if ((int)bv.size() > bow_cap || (int)fv.size() > fv_cap) return -1;
    int k = 0;
    for (auto& e : bv) {
        bow_ids[k] = (int)e.first;
        bow_vals[k] = e.second;
        ++k;
    }
    k = 0;
    int c = 0;
    fv_off[0] = 0;
    for (auto& e : fv) {
        fv_nodes[k] = (int)e.first;
        for (unsigned idx : e.second) fv_idx[c++] = (int)idx;
        fv_off[++k] = c;
    }
    return 0;
}
}
