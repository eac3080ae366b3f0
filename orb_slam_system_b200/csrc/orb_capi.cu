// C ABI of liborb_b200.so (include/orb_b200.h): handles, the per-shape plan, device
// memory and streams.  No torch types, no exceptions across the boundary, no CPU fallback.
//
// Host-side arithmetic that decides shapes (ctor tables, level sizes, cell grid, resize
// coefficient tables) follows reference src/ORBextractor.cc:116-170, :497-515, :288-331 and
// the cv::resize model float op by float op; this file is compiled with
// -ffp-contract=off for that reason.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/orb_b200.h"
#include "capi_internal.h"
#include "extract_kernels.h"
#include "match_kernels.h"
#include "orb_plan.h"

static_assert(sizeof(orb_keypoint) == 28, "orb_keypoint must match cv::KeyPoint");
static_assert(sizeof(orb_keypoint_dev) == 28 && sizeof(orb_kp28) == 28, "device keypoint layout");

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string t_err;
int orb_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_err = buf;
    return code;
}
#define fail orb_fail

extern "C" const char* orb_last_error(void) { return t_err.c_str(); }
extern "C" int orb_host_alloc(size_t bytes, void** out) {
    if (!out) return orb_fail(ORB_ERR_INVALID, "null argument");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) return orb_fail(ORB_ERR_CUDA, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    return ORB_OK;
}
extern "C" void orb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
extern "C" uint64_t orb_kernel_launch_count(void) { return orbk_launch_count(); }
extern "C" const char* orb_version(void) { return "orb_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------
// ctor tables (reference src/ORBextractor.cc:116-151)
// ------------------------------------------------------------------------------------------
struct HostTables {
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> nfeat;
};

static int round_half_even_f(float v) { return (int)lrintf(v); }  // cvRound

static HostTables build_tables(const orb_params& p) {
    HostTables t;
    const int n = p.nlevels;
    const double sf = (double)p.scale_factor;  // the member is a double (ORBextractor.h:79)
    t.scale.assign(n, 1.0f);
    if (n >= 2) {  // partial_sum over the vector itself, shifted by one (:120-124)
        float acc = t.scale[0];
        t.scale[1] = acc;
        for (int i = 1; i <= n - 2; ++i) {
            acc = (float)((double)acc * sf);
            t.scale[i + 1] = acc;
        }
    }
    t.sigma2.resize(n);
    t.inv_scale.resize(n);
    t.inv_sigma2.resize(n);
    for (int i = 0; i < n; ++i) {
        t.sigma2[i] = t.scale[i] * t.scale[i];
        t.inv_scale[i] = 1.0f / t.scale[i];
        t.inv_sigma2[i] = 1.0f / t.sigma2[i];
    }
    t.nfeat.assign(n, 0);
    const float factor = (float)(1.0 / sf);
    float desired = (float)((double)((float)p.nfeatures * (1 - factor)) / (1.0 - pow((double)factor, (double)n)));
    int sum = 0;
    for (int i = 0; i < n - 1; ++i) {
        const int cur = round_half_even_f(desired);
        sum += cur;
        desired *= factor;
        t.nfeat[i] = cur;
    }
    t.nfeat[n - 1] = std::max(p.nfeatures - sum, 0);
    return t;
}

// cv::resize INTER_LINEAR coefficient tables of one axis
static void linear_axis(int ssize, int dsize, bool is_x, std::vector<int>& tab, std::vector<int>& coef) {
    tab.resize(dsize);
    coef.resize(dsize);
    const double inv_scale = (double)dsize / ssize;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        int s0, s1;
        if (is_x) {  // x: coefficients collapse at the edges
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
            s0 = s;
            s1 = std::min(s + 1, ssize - 1);
        } else {  // y: coefficients kept, row indices clipped
            s0 = std::min(std::max(s, 0), ssize - 1);
            s1 = std::min(std::max(s + 1, 0), ssize - 1);
        }
        const int c0 = round_half_even_f((1.f - f) * 2048.f);
        const int c1 = round_half_even_f(f * 2048.f);
        tab[d] = s0 | (s1 << 16);
        coef[d] = (c0 & 0xffff) | (c1 << 16);
    }
}

// ------------------------------------------------------------------------------------------
// extractor handle
// ------------------------------------------------------------------------------------------
struct orb_extractor {
    orb_params params;
    HostTables tab;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // blur runs here, concurrently with detect + octree
    cudaEvent_t evFork = nullptr, evJoin = nullptr, evPyr = nullptr;
    // second kernel lane: the device-resident path runs the two halves of a batch concurrently
    enum { MAX_LANES = 4 };
    cudaStream_t laneSt[MAX_LANES] = {}, laneSt2[MAX_LANES] = {};
    cudaEvent_t laneFork[MAX_LANES] = {}, laneJoin[MAX_LANES] = {}, laneMerge[MAX_LANES] = {}, lanePyr[MAX_LANES] = {}, evSplit = nullptr;
    OrbStreams lane(int i) const { return i == 0 ? streams() : OrbStreams{laneSt[i], laneSt2[i], laneFork[i], laneJoin[i], lanePyr[i]}; }
    int lanes = 2;
    // host-buffer pipeline: copies in and out run on their own streams, chunk by chunk
    cudaStream_t streamIn = nullptr, streamOut = nullptr, streamCnt = nullptr;
    enum { MAX_CHUNKS = 8, NUM_SLOTS = 3 };
    cudaEvent_t evFree = nullptr;
    // One in-flight host batch: its landing / result staging buffers, its events, and where the results go.
    // One chunk's kernel sequence (pitch conversion, extraction, status copy) as an instantiated CUDA graph, replayed while
    // the arguments it was captured with stay the same: one cudaGraphLaunch instead of ~20 launch calls.  Used for batches of
    // one chunk (a frame or a stereo pair per call, the way ORB_SLAM2::Frame drives the extractor): 5-10 % less time per call.
    // Multi-chunk batches keep their stream launches: replayed graphs measured 3 % slower there.
    struct ChunkGraph {
        cudaGraphExec_t exec = nullptr;
        OrbPlan plan;
        const void* ptr[8] = {};
        long long num[8] = {};
        int launches = 0;
    };
    struct HostSlot {
        bool busy = false;
        ChunkGraph graph[MAX_CHUNKS];
        uint8_t* d_dense = nullptr;  // landing buffer for densely packed host frames (one linear copy per chunk)
        size_t dense_cap = 0;
        orb_keypoint_dev* d_kps = nullptr;
        uint8_t* d_desc = nullptr;
        int* d_counts = nullptr;
        int out_cap = 0;
        cudaEvent_t evIn[MAX_CHUNKS] = {}, evDone[MAX_CHUNKS] = {}, evCnt[MAX_CHUNKS] = {};  // evCnt: the chunk's counts are on the host
        cudaEvent_t evOut = nullptr;  // counts, status flags and the speculative result copies of every chunk are in host memory
        // the submitted call
        int n = 0, cap = 0, nchunks = 0, per = 0;
        int spec_w = 0;  // keypoint rows per frame whose copies were enqueued at submit time
        orb_keypoint* kps = nullptr;
        uint8_t* desc = nullptr;
        int* counts = nullptr;
        int* h_status = nullptr;  // pinned, max_batch ints: octree status flags of the submitted batch
    } slot[NUM_SLOTS];
    // Result rows per frame to copy back before the counts are known on the host (see submit_impl): the largest count
    // of the previous batch of this shape plus a margin; 0 = nothing known yet (copy up to the caller's capacity).
    int spec_rows = 0;
    enum { DEV_GRAPHS = 8 };
    ChunkGraph devGraph[MAX_LANES][DEV_GRAPHS];  // the device-resident path's lanes: a few argument sets each (rotating input buffers)
    int devGraphNext[MAX_LANES] = {};
    int graph_mode = 2;  // ORB_B200_GRAPH: 0 off, 1 every chunk, 2 (default) single-chunk batches only (the latency form)
    // ORB_B200_EAGER_D2H (default 1): result copies are enqueued by submit behind the chunk's kernels; 0: by wait, once the
    // chunk's counts are on the host (one host round trip per chunk, exact row counts).
    bool eager_out = true;
    // image ingest (orb_extractor_set_ingest): raw frames -> remap -> gray, fused into the level-0 load
    struct Ingest {
        bool on = false;
        int srows = 0, scols = 0, channels = 1, bgr = 0, variant = 4, drows = 0, dcols = 0;
        float *d_mapx = nullptr, *d_mapy = nullptr;
    } ing;
    OrbStreams streams() const { return OrbStreams{stream, stream2, evFork, evJoin, evPyr}; }
    int max_batch = 1;
    OrbPlan plan;            // current shape (plan.rows == 0: none)
    DetectMaps maps;         // TMA descriptors of the internal level buffers (host copy)
    DetectMaps maps_user;    // same with level 0 pointing at the caller's device frames
    DetectMaps* d_maps = nullptr;        // device copies (global memory)
    DetectMaps* d_maps_user = nullptr;
    const uint8_t* user_base = nullptr;
    int user_n = 0;          // frames the user maps were encoded for
    std::vector<void*> allocs;  // everything the plan points at
    uint8_t* level0 = nullptr;  // internal level-0 buffer (host-input path)
    const uint8_t* last_l0 = nullptr;  // where level 0 of the last call lives (the caller's frames when read in place)
    int last_n = 0;
    std::vector<int> h_status;
    // per-stage CUDA-event records (orb_extractor_set_profiling)
    bool profiling = false;
    std::vector<cudaEvent_t> events;  // ORB_EVENTS per recorded call
    cudaEvent_t* next_events() {
        if (!profiling) return nullptr;
        const size_t base = events.size();
        for (int i = 0; i < ORB_EVENTS; ++i) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) {
                events.resize(base);
                return nullptr;
            }
            events.push_back(e);
        }
        return &events[base];
    }
    void clear_events() {
        for (cudaEvent_t e : events) cudaEventDestroy(e);
        events.clear();
    }
};

static void free_plan(orb_extractor* h) {
    for (void* p : h->allocs) cudaFree(p);
    h->allocs.clear();
    h->level0 = nullptr;
    h->d_maps = nullptr;
    h->d_maps_user = nullptr;
    memset(&h->plan, 0, sizeof h->plan);
}

template <typename T>
static cudaError_t dev_alloc(orb_extractor* h, T** out, size_t count) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 256));
    if (e != cudaSuccess) return e;
    h->allocs.push_back(p);
    *out = (T*)p;
    return cudaSuccess;
}

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// Geometry of one level; returns an orb_status.
static int level_geometry(const orb_params& prm, const HostTables& tab, int rows, int cols, int l, OrbLevel& L) {
    memset(&L, 0, sizeof L);
    const float s = tab.inv_scale[l];
    L.cols = round_half_even_f(cols * s);  // Size sz(cvRound(cols*scale), cvRound(rows*scale)) (:502)
    L.rows = round_half_even_f(rows * s);
    if (L.cols <= 0 || L.rows <= 0) return fail(ORB_ERR_SHAPE, "level %d is empty (%dx%d)", l, L.cols, L.rows);
    L.pitch = align_up(L.cols, 64);
    L.plane = (unsigned long long)L.pitch * L.rows;
    L.src = l;
    L.scale = tab.scale[l];
    L.patchSize = (int)(31 * tab.scale[l]);
    L.nFeat = tab.nfeat[l];
    const int minB = ORB_MINB, maxBX = L.cols - ORB_MINB, maxBY = L.rows - ORB_MINB;
    L.W = maxBX - minB;
    L.H = maxBY - minB;
    if (L.H == 0) return fail(ORB_ERR_SHAPE, "level %d: reference divides by zero (rows == 32)", l);
    L.nIni = L.W / L.H;  // integer division
    if (L.nIni < 0) return fail(ORB_ERR_SHAPE, "level %d: reference throws (negative root count)", l);
    if (L.nIni > 63) return fail(ORB_ERR_SHAPE, "level %d: aspect ratio above 63:1 is not supported", l);
    L.hX = (float)L.W / L.nIni;  // inf/NaN when nIni == 0, as in the reference; then no key passes the root test
    const float width = (float)L.W, height = (float)L.H;
    const int nCols = (int)(width / ORB_CELL), nRows = (int)(height / ORB_CELL);
    if (nCols > 0 && nRows > 0) {
        L.nCols = nCols;
        L.nRows = nRows;
        L.wCell = (int)ceilf(width / nCols);
        L.hCell = (int)ceilf(height / nRows);
        // the reference's Mat::operator()(Rect) throws on a negative cell extent
        if (minB + (nCols - 1) * L.wCell > maxBX || minB + (nRows - 1) * L.hCell > maxBY)
            return fail(ORB_ERR_SHAPE, "level %d (%dx%d): reference throws cv::Exception (negative cell ROI)", l, L.cols, L.rows);
        // tile width + 6-px halo + up to 15 bytes of TMA start alignment must fit the 256-byte box
        int tc = std::max(1, (DET_TILE_W - 6 - 15) / L.wCell);
        L.boxH = L.hCell + 6;
        L.tilesX = (nCols + tc - 1) / tc;
        L.tileCells = (nCols + L.tilesX - 1) / L.tilesX;
        L.nTiles = L.tilesX * nRows;
        // worst-case candidates: NMS survivors are never 8-adjacent inside a cell
        unsigned long long cap = 0;
        for (int i = 0; i < nRows; ++i) {
            const int ch = std::max(0, std::min(ORB_EDGE + (i + 1) * L.hCell, maxBY - 3) - (ORB_EDGE + i * L.hCell));
            for (int j = 0; j < nCols; ++j) {
                const int cw = std::max(0, std::min(ORB_EDGE + (j + 1) * L.wCell, maxBX - 3) - (ORB_EDGE + j * L.wCell));
                cap += (unsigned long long)((cw + 1) / 2) * ((ch + 1) / 2);
            }
        }
        if (cap >= (1ull << 24)) return fail(ORB_ERR_SHAPE, "level %d too large", l);
        L.candCap = (unsigned)std::max<unsigned long long>(cap, 1);
    } else {
        L.candCap = 1;
    }
    L.kmax = (int)std::min<long long>(L.candCap, std::max(std::max(4 * L.nIni, 4 * L.nFeat), 1));
    unsigned sc = 2;
    while (sc < L.candCap) sc <<= 1;
    L.sortCap = sc;
    return ORB_OK;
}

static int build_plan(orb_extractor* h, int rows, int cols) {
    if (h->plan.rows == rows && h->plan.cols == cols) return ORB_OK;
    if (rows > ORB_MAX_DIM || cols > ORB_MAX_DIM) return fail(ORB_ERR_SHAPE, "image larger than %d", ORB_MAX_DIM);
    OrbPlan P;
    memset(&P, 0, sizeof P);
    P.nlevels = h->params.nlevels;
    P.rows = rows;
    P.cols = cols;
    P.iniTh = std::min(std::max(h->params.ini_th_fast, 0), 255);  // cv::FAST clamps its threshold
    P.minTh = std::min(std::max(h->params.min_th_fast, 0), 255);
    P.lowTh = std::min(P.iniTh, P.minTh);
    P.batch = h->max_batch;
    for (int l = 0; l < P.nlevels; ++l) {
        int rc = level_geometry(h->params, h->tab, rows, cols, l, P.lv[l]);
        if (rc != ORB_OK) return rc;
        // a same-size resize is a byte copy: share pixels and candidates with the source level
        if (l > 0 && P.lv[l].rows == P.lv[l - 1].rows && P.lv[l].cols == P.lv[l - 1].cols) P.lv[l].src = P.lv[l - 1].src;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    free_plan(h);
    const int B = h->max_batch;
    int tileBase = 0, keptBase = 0, descTileBase = 0;
    for (int l = 0; l < P.nlevels; ++l) {
        OrbLevel& L = P.lv[l];
        L.keptBase = keptBase;
        keptBase += L.kmax;
        CUDA_TRY(dev_alloc(h, &L.kept, (size_t)B * L.kmax));
        if (L.src != l) {  // an alias runs its own octree problem: own scratch, sized like its source's
            L.sortCap = P.lv[L.src].sortCap;
            CUDA_TRY(dev_alloc(h, &L.sortScratch, (size_t)B * 2 * L.sortCap));
            continue;
        }
        L.tileBase = tileBase;
        tileBase += L.nTiles;
        // describe tiles over the keypoint area [19, cols - 19) x [19, rows - 19)
        if (L.cols > 2 * ORB_EDGE && L.rows > 2 * ORB_EDGE) {
            L.dTilesX = (L.cols - 2 * ORB_EDGE + DSC_W - 1) / DSC_W;
            L.dTiles = L.dTilesX * ((L.rows - 2 * ORB_EDGE + DSC_H - 1) / DSC_H);
        }
        L.dTileBase = descTileBase;
        descTileBase += L.dTiles;
        CUDA_TRY(dev_alloc(h, &L.img, (size_t)B * L.plane + 256));
        CUDA_TRY(dev_alloc(h, &L.blur, (size_t)B * L.plane + 256));
        CUDA_TRY(dev_alloc(h, &L.cand, (size_t)B * L.candCap));
        CUDA_TRY(dev_alloc(h, &L.sortScratch, (size_t)B * 2 * L.sortCap));
        if (l > 0) {
            const OrbLevel& S = P.lv[P.lv[l - 1].src];
            std::vector<int> xt, xc, yt, yc;
            linear_axis(S.cols, L.cols, true, xt, xc);
            linear_axis(S.rows, L.rows, false, yt, yc);
            int *dxt, *dxc, *dyt, *dyc;
            CUDA_TRY(dev_alloc(h, &dxt, xt.size()));
            CUDA_TRY(dev_alloc(h, &dxc, xc.size()));
            CUDA_TRY(dev_alloc(h, &dyt, yt.size()));
            CUDA_TRY(dev_alloc(h, &dyc, yc.size()));
            CUDA_TRY(cudaMemcpy(dxt, xt.data(), xt.size() * 4, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(dxc, xc.data(), xc.size() * 4, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(dyt, yt.data(), yt.size() * 4, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(dyc, yc.data(), yc.size() * 4, cudaMemcpyHostToDevice));
            L.xtab = dxt;
            L.xcoef = dxc;
            L.ytab = dyt;
            L.ycoef = dyc;
            // k_resize4 descriptors: per group of 4 output columns, gather 8 source bytes
            // (S[s_k], S[s1_k]) from three aligned words starting at word wb
            const int ng = (L.cols + 3) / 4;
            std::vector<int4> xg(ng), xc4(ng);
            bool fits = true;
            for (int g = 0; g < ng && fits; ++g) {
                int sA[4], sB[4], cf[4];
                for (int k = 0; k < 4; ++k) {
                    const int x = std::min(4 * g + k, L.cols - 1);
                    sA[k] = xt[x] & 0xffff;
                    sB[k] = xt[x] >> 16;
                    cf[k] = xc[x];
                }
                int lo = sA[0];
                for (int k = 0; k < 4; ++k) lo = std::min(lo, std::min(sA[k], sB[k]));
                const int wb = lo >> 2;
                unsigned sel01[2] = {0, 0}, sel2[2] = {0, 0};
                for (int i = 0; i < 8; ++i) {  // output byte i: word i/4, position i%4: A0 B0 A1 B1 | A2 B2 A3 B3
                    const int k = i >> 1, j = ((i & 1) ? sB[k] : sA[k]) - 4 * wb;
                    if (j < 0 || j >= 12) { fits = false; break; }
                    const int w = i >> 2, pos = i & 3;
                    if (j < 8) {
                        sel01[w] |= (unsigned)j << (4 * pos);            // stage 1a: from (W0, W1)
                        sel2[w] |= (unsigned)pos << (4 * pos);           // stage 2: take t0.byte[pos]
                    } else {
                        sel01[w] |= (unsigned)(j - 4) << (16 + 4 * pos); // stage 1b: from (W1, W2)
                        sel2[w] |= (unsigned)(4 + pos) << (4 * pos);     // stage 2: take t1.byte[pos]
                    }
                }
                xg[g] = make_int4(wb, (int)sel01[0], (int)sel01[1], (int)(sel2[0] | (sel2[1] << 16)));
                xc4[g] = make_int4(cf[0], cf[1], cf[2], cf[3]);
            }
            if (fits && (S.pitch & 3) == 0) {
                int4 *dxg, *dxc4;
                CUDA_TRY(dev_alloc(h, &dxg, xg.size()));
                CUDA_TRY(dev_alloc(h, &dxc4, xc4.size()));
                CUDA_TRY(cudaMemcpy(dxg, xg.data(), xg.size() * sizeof(int4), cudaMemcpyHostToDevice));
                CUDA_TRY(cudaMemcpy(dxc4, xc4.data(), xc4.size() * sizeof(int4), cudaMemcpyHostToDevice));
                L.xgrp = dxg;
                L.xcoef4 = dxc4;
                // k_resize_tile: does every output tile's source fit one 16-byte aligned 256 x RSZ_BOX_H box?
                bool tiled = (S.pitch & 15) == 0;
                for (int g0 = 0; g0 < ng && tiled; g0 += RSZ_W / 4) {
                    const int g1 = std::min(ng, g0 + RSZ_W / 4) - 1;
                    const int bx0 = (4 * xg[g0].x) & ~15;
                    for (int g = g0; g <= g1; ++g)
                        if (4 * xg[g].x < bx0 || 4 * (xg[g].x + 3) > bx0 + 256) tiled = false;  // words wb .. wb + 2 inside the box
                }
                for (int y0 = 0; y0 < L.rows && tiled; y0 += RSZ_H) {
                    const int y1 = std::min(L.rows, y0 + RSZ_H) - 1;
                    const int by0 = yt[y0] & 0xffff;
                    for (int y = y0; y <= y1; ++y)
                        if ((yt[y] & 0xffff) < by0 || (yt[y] >> 16) >= by0 + RSZ_BOX_H) tiled = false;
                }
                L.rszTiled = tiled ? 1 : 0;
            }
        }
    }
    // aliases share their source's buffers
    for (int l = 0; l < P.nlevels; ++l) {
        OrbLevel& L = P.lv[l];
        if (L.src == l) continue;
        const OrbLevel& S = P.lv[L.src];
        L.img = S.img;
        L.blur = S.blur;
        L.cand = S.cand;
        L.candCap = S.candCap;
        L.kmax = std::min<int>(L.kmax, (int)S.candCap);
    }
    P.totalTiles = tileBase;
    P.totalDescTiles = descTileBase;
    P.detRows = 8;  // k_detect sizes its shared memory by the tallest tile
    for (int l = 0; l < P.nlevels; ++l)
        if (P.lv[l].src == l && P.lv[l].nTiles > 0) P.detRows = std::max(P.detRows, P.lv[l].boxH);
    if (P.detRows > DET_TILE_H) return fail(ORB_ERR_SHAPE, "detect tile of %d rows exceeds %d", P.detRows, DET_TILE_H);
    P.totalKmax = keptBase;
    {
        // tile tables of the detect / blur / describe launches (same tile order as the launches' tile ids)
        auto pack = [](int l, int ty, int tx) { return (unsigned)l | ((unsigned)ty << 4) | ((unsigned)tx << 18); };
        std::vector<unsigned> det, blr, dsc;
        for (int l = 0; l < P.nlevels; ++l) {
            const OrbLevel& L = P.lv[l];
            if (L.src != l) continue;
            for (int t = 0; t < L.nTiles; ++t) det.push_back(pack(l, t / L.tilesX, t % L.tilesX));
            const int bx = (L.cols + BLUR_TW - 1) / BLUR_TW, by = (L.rows + BLUR_TH - 1) / BLUR_TH;
            for (int t = 0; t < bx * by; ++t) blr.push_back(pack(l, t / bx, t % bx));
            for (int t = 0; t < L.dTiles; ++t) dsc.push_back(pack(l, t / L.dTilesX, t % L.dTilesX));
        }
        if ((int)det.size() != P.totalTiles || (int)dsc.size() != P.totalDescTiles) return fail(ORB_ERR_INVALID, "tile table mismatch");
        P.blurTiles = (int)blr.size();
        unsigned *dDet, *dBlr, *dDsc;
        CUDA_TRY(dev_alloc(h, &dDet, det.size() + 1));
        CUDA_TRY(dev_alloc(h, &dBlr, blr.size() + 1));
        CUDA_TRY(dev_alloc(h, &dDsc, dsc.size() + 1));
        if (!det.empty()) CUDA_TRY(cudaMemcpy(dDet, det.data(), det.size() * 4, cudaMemcpyHostToDevice));
        if (!blr.empty()) CUDA_TRY(cudaMemcpy(dBlr, blr.data(), blr.size() * 4, cudaMemcpyHostToDevice));
        if (!dsc.empty()) CUDA_TRY(cudaMemcpy(dDsc, dsc.data(), dsc.size() * 4, cudaMemcpyHostToDevice));
        P.detTileTab = dDet;
        P.blurTileTab = dBlr;
        P.descTileTab = dDsc;
    }
    CUDA_TRY(dev_alloc(h, &P.candCount, (size_t)B * ORB_MAX_LEVELS));
    CUDA_TRY(dev_alloc(h, &P.keptCount, (size_t)B * ORB_MAX_LEVELS));
    CUDA_TRY(dev_alloc(h, &P.status, (size_t)B));
    {
        std::vector<int2> ic(8 * 32);
        orbk_build_ic_table(ic.data());
        int2* d_ic;
        CUDA_TRY(dev_alloc(h, &d_ic, ic.size()));
        CUDA_TRY(cudaMemcpy(d_ic, ic.data(), ic.size() * sizeof(int2), cudaMemcpyHostToDevice));
        P.icTab = d_ic;
        std::vector<float4> pt(182);
        orbk_build_pair_table(pt.data());
        float4* d_pt;
        CUDA_TRY(dev_alloc(h, &d_pt, pt.size()));
        CUDA_TRY(cudaMemcpy(d_pt, pt.data(), pt.size() * sizeof(float4), cudaMemcpyHostToDevice));
        P.pairTab = d_pt;
    }
    CUDA_TRY(dev_alloc(h, &P.needGeneric, (size_t)B * ORB_MAX_LEVELS));
    CUDA_TRY(cudaMemset(P.needGeneric, 0, sizeof(int) * B * ORB_MAX_LEVELS));
    CUDA_TRY(cudaMemset(P.keptCount, 0, sizeof(int) * B * ORB_MAX_LEVELS));
    CUDA_TRY(cudaMemset(P.candCount, 0, sizeof(int) * B * ORB_MAX_LEVELS));
    memset(&h->maps, 0, sizeof h->maps);
    for (int l = 0; l < P.nlevels; ++l) {
        const OrbLevel& L = P.lv[l];
        if (L.src != l) continue;
        if (L.nTiles > 0) CUDA_TRY(orbk_encode_level_map(&h->maps.m[l], L.img, L.cols, L.rows, B, L.pitch, L.plane, DET_TILE_W, L.boxH));
        CUDA_TRY(orbk_encode_level_map(&h->maps.blr[l], L.img, L.cols, L.rows, B, L.pitch, L.plane, 256, BLUR_BOX_H));
        if (l > 0 && L.rszTiled) {
            const OrbLevel& S = P.lv[P.lv[l - 1].src];
            CUDA_TRY(orbk_encode_level_map(&h->maps.rsz[l], S.img, S.cols, S.rows, B, S.pitch, S.plane, 256, RSZ_BOX_H));
        }
        if (L.dTiles > 0) {
            CUDA_TRY(orbk_encode_level_map(&h->maps.raw[l], L.img, L.cols, L.rows, B, L.pitch, L.plane, DSC_BOX_W, DSC_BOX_H));
            CUDA_TRY(orbk_encode_level_map(&h->maps.blur[l], L.blur, L.cols, L.rows, B, L.pitch, L.plane, DSC_BOX_W, DSC_BOX_H));
        }
    }
    h->maps_user = h->maps;
    h->user_base = nullptr;
    CUDA_TRY(dev_alloc(h, &h->d_maps, (size_t)1));
    CUDA_TRY(dev_alloc(h, &h->d_maps_user, (size_t)1));
    CUDA_TRY(cudaMemcpy(h->d_maps, &h->maps, sizeof(DetectMaps), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d_maps_user, &h->maps, sizeof(DetectMaps), cudaMemcpyHostToDevice));
    h->level0 = P.lv[0].img;
    h->plan = P;
    return ORB_OK;
}

static int ensure_out(orb_extractor* h, orb_extractor::HostSlot& S, int cap) {
    // counts, and behind them this batch's own copy of the octree status flags
    if (!S.d_counts) CUDA_TRY(cudaMalloc((void**)&S.d_counts, sizeof(int) * 2 * h->max_batch));
    if (!S.h_status) CUDA_TRY(cudaMallocHost((void**)&S.h_status, sizeof(int) * h->max_batch));
    if (cap <= S.out_cap) return ORB_OK;
    CUDA_TRY(cudaStreamSynchronize(h->streamOut));
    if (S.d_kps) cudaFree(S.d_kps);
    if (S.d_desc) cudaFree(S.d_desc);
    S.d_kps = nullptr;
    S.d_desc = nullptr;
    S.out_cap = 0;
    CUDA_TRY(cudaMalloc((void**)&S.d_kps, (size_t)h->max_batch * cap * sizeof(orb_keypoint_dev)));
    CUDA_TRY(cudaMalloc((void**)&S.d_desc, (size_t)h->max_batch * cap * 32));
    S.out_cap = cap;
    return ORB_OK;
}

extern "C" int orb_extractor_create(const orb_params* params, int max_rows, int max_cols, int max_batch, int device,
                                    orb_extractor** out) {
    if (!params || !out) return fail(ORB_ERR_INVALID, "null argument");
    if (params->nlevels < 1 || params->nlevels > ORB_MAX_LEVELS) return fail(ORB_ERR_INVALID, "nlevels must be 1..%d", ORB_MAX_LEVELS);
    if (max_batch < 1) return fail(ORB_ERR_INVALID, "max_batch must be >= 1");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ORB_ERR_CUDA, "no CUDA device: %s (liborb_b200 has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(ORB_ERR_INVALID, "device %d out of range (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(orbk_init_device());
    orb_extractor* h = new orb_extractor();
    h->params = *params;
    h->tab = build_tables(*params);
    h->device = device;
    h->max_batch = max_batch;
    memset(&h->plan, 0, sizeof h->plan);
    // The second stream of every lane (pyramid chain + blur: small dependent launches) gets the highest stream priority, so
    // its blocks are scheduled ahead of the queued blocks of the other lane's large kernels: 99.1 k -> 99.7 k frames/s.
    // ORB_B200_PRIO=0 switches it off, 2 gives the priority to the first stream instead (96.9 k).
    int prioLo = 0, prioHi = 0;
    cudaDeviceGetStreamPriorityRange(&prioLo, &prioHi);
    const int prioMode = getenv("ORB_B200_PRIO") ? atoi(getenv("ORB_B200_PRIO")) : 1;
    const int prio2 = prioMode == 1 ? prioHi : prioLo, prio1 = prioMode == 2 ? prioHi : prioLo;
    cudaError_t ce = cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio1);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, prio2);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->evJoin, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->evPyr, cudaEventDisableTiming);
    for (int i = 1; i < orb_extractor::MAX_LANES && ce == cudaSuccess; ++i) {
        ce = cudaStreamCreateWithPriority(&h->laneSt[i], cudaStreamNonBlocking, prio1);
        if (ce == cudaSuccess) ce = cudaStreamCreateWithPriority(&h->laneSt2[i], cudaStreamNonBlocking, prio2);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->laneFork[i], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->laneJoin[i], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->laneMerge[i], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->lanePyr[i], cudaEventDisableTiming);
    }
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->evSplit, cudaEventDisableTiming);
    if (const char* e = getenv("ORB_B200_EAGER_D2H")) h->eager_out = atoi(e) != 0;
    if (const char* e = getenv("ORB_B200_GRAPH")) h->graph_mode = atoi(e);
    if (const char* e = getenv("ORB_B200_LANES")) h->lanes = std::min<int>(orb_extractor::MAX_LANES, std::max(1, atoi(e)));
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->streamIn, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->streamOut, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->streamCnt, cudaStreamNonBlocking);
    for (int k = 0; k < orb_extractor::NUM_SLOTS; ++k)
        for (int i = 0; i < orb_extractor::MAX_CHUNKS && ce == cudaSuccess; ++i) {
            ce = cudaEventCreateWithFlags(&h->slot[k].evIn[i], cudaEventDisableTiming);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->slot[k].evDone[i], cudaEventDisableTiming);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->slot[k].evCnt[i], cudaEventDisableTiming);
        }
    for (int k = 0; k < orb_extractor::NUM_SLOTS && ce == cudaSuccess; ++k)
        ce = cudaEventCreateWithFlags(&h->slot[k].evOut, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->evFree, cudaEventDisableTiming);
    if (ce != cudaSuccess) {
        orb_extractor_destroy(h);
        return fail(ORB_ERR_CUDA, "stream/event/alloc setup: %s", cudaGetErrorString(ce));
    }
    if (max_rows > 0 && max_cols > 0) {
        int rc = build_plan(h, max_rows, max_cols);
        if (rc != ORB_OK) {
            orb_extractor_destroy(h);
            return rc;
        }
    }
    *out = h;
    return ORB_OK;
}

static void destroy_graphs(orb_extractor* h) {
    for (int k = 0; k < orb_extractor::NUM_SLOTS; ++k)
        for (auto& g : h->slot[k].graph)
            if (g.exec) cudaGraphExecDestroy(g.exec), g.exec = nullptr;
    for (auto& lane : h->devGraph)
        for (auto& g : lane)
            if (g.exec) cudaGraphExecDestroy(g.exec), g.exec = nullptr;
}

extern "C" void orb_extractor_destroy(orb_extractor* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_plan(h);
    h->clear_events();
    destroy_graphs(h);
    for (int k = 0; k < orb_extractor::NUM_SLOTS; ++k) {
        orb_extractor::HostSlot& S = h->slot[k];
        if (S.d_kps) cudaFree(S.d_kps);
        if (S.d_desc) cudaFree(S.d_desc);
        if (S.d_counts) cudaFree(S.d_counts);
        if (S.d_dense) cudaFree(S.d_dense);
        if (S.h_status) cudaFreeHost(S.h_status);
        for (int i = 0; i < orb_extractor::MAX_CHUNKS; ++i) {
            if (S.evIn[i]) cudaEventDestroy(S.evIn[i]);
            if (S.evDone[i]) cudaEventDestroy(S.evDone[i]);
            if (S.evCnt[i]) cudaEventDestroy(S.evCnt[i]);
        }
        if (S.evOut) cudaEventDestroy(S.evOut);
    }
    if (h->evFree) cudaEventDestroy(h->evFree);
    if (h->ing.d_mapx) cudaFree(h->ing.d_mapx);
    if (h->ing.d_mapy) cudaFree(h->ing.d_mapy);
    for (int i = 1; i < orb_extractor::MAX_LANES; ++i) {
        if (h->laneFork[i]) cudaEventDestroy(h->laneFork[i]);
        if (h->laneJoin[i]) cudaEventDestroy(h->laneJoin[i]);
        if (h->laneMerge[i]) cudaEventDestroy(h->laneMerge[i]);
        if (h->lanePyr[i]) cudaEventDestroy(h->lanePyr[i]);
        if (h->laneSt2[i]) cudaStreamDestroy(h->laneSt2[i]);
        if (h->laneSt[i]) cudaStreamDestroy(h->laneSt[i]);
    }
    if (h->evSplit) cudaEventDestroy(h->evSplit);
    if (h->streamIn) cudaStreamDestroy(h->streamIn);
    if (h->streamOut) cudaStreamDestroy(h->streamOut);
    if (h->streamCnt) cudaStreamDestroy(h->streamCnt);
    if (h->evFork) cudaEventDestroy(h->evFork);
    if (h->evJoin) cudaEventDestroy(h->evJoin);
    if (h->evPyr) cudaEventDestroy(h->evPyr);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int orb_extractor_tables(const orb_extractor* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                                    int32_t* features_per_level) {
    if (!h) return fail(ORB_ERR_INVALID, "null handle");
    const int n = h->params.nlevels;
    for (int i = 0; i < n; ++i) {
        if (scale) scale[i] = h->tab.scale[i];
        if (inv_scale) inv_scale[i] = h->tab.inv_scale[i];
        if (sigma2) sigma2[i] = h->tab.sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->tab.inv_sigma2[i];
        if (features_per_level) features_per_level[i] = h->tab.nfeat[i];
    }
    return ORB_OK;
}

extern "C" int orb_extractor_keypoint_bound(const orb_extractor* h, int rows, int cols, int* bound) {
    if (!h || !bound) return fail(ORB_ERR_INVALID, "null argument");
    long long tot = 0;
    for (int l = 0; l < h->params.nlevels; ++l) {
        OrbLevel L;
        int rc = level_geometry(h->params, h->tab, rows, cols, l, L);
        if (rc != ORB_OK) return rc;
        tot += L.kmax;
    }
    *bound = (int)tot;
    return ORB_OK;
}

// The plan with every per-frame pointer advanced to frame f0 (one chunk of a batch).
static OrbPlan plan_slice(const OrbPlan& P, int f0) {
    OrbPlan Q = P;
    if (f0 == 0) return Q;
    Q.frameBase = P.frameBase + f0;
    for (int l = 0; l < P.nlevels; ++l) {
        OrbLevel& L = Q.lv[l];
        L.img += (size_t)f0 * L.plane;
        L.blur += (size_t)f0 * L.plane;
        L.cand += (size_t)f0 * L.candCap;
        L.kept += (size_t)f0 * L.kmax;
        L.sortScratch += (size_t)f0 * 2 * L.sortCap;
    }
    Q.candCount += (size_t)f0 * ORB_MAX_LEVELS;
    Q.keptCount += (size_t)f0 * ORB_MAX_LEVELS;
    Q.needGeneric += (size_t)f0 * ORB_MAX_LEVELS;
    Q.status += f0;
    return Q;
}

static int check_status(orb_extractor* h, int n) {
    h->h_status.assign(n, 0);
    CUDA_TRY(cudaMemcpyAsync(h->h_status.data(), h->plan.status, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n; ++i)
        if (h->h_status[i]) return fail(ORB_ERR_UNSEPARABLE, "frame %d: octree cannot separate its keys (reference would not terminate)", i);
    return ORB_OK;
}

// The kernels of one chunk on lane `ls`: [pitch conversion of densely packed frames] -> extraction -> [status flags to the
// slot].  With graphs on, the sequence is captured once per distinct argument set and replayed.
struct ChunkArgs {
    OrbPlan P;
    int nf;
    orb_keypoint_dev* d_kps;
    uint8_t* d_desc;
    int cap;
    int* d_counts;
    const DetectMaps* maps;
    const uint8_t* dense;  // nullptr: no pitch conversion
    int f0, rows, cols;
    uint8_t* l0;
    int pitch;
    size_t plane;
    int* status_dst;  // nullptr: no status copy
    const int* status_src;
};
static cudaError_t enqueue_chunk_kernels(const ChunkArgs& a, const OrbStreams& ls, cudaEvent_t* ev) {
    cudaError_t e = cudaSuccess;
    if (a.dense) e = orbk_repitch(a.dense, a.f0, a.nf, a.rows, a.cols, a.l0, a.pitch, a.plane, ls.st);
    if (e == cudaSuccess) e = orbk_run_extract(a.P, a.nf, a.d_kps, a.d_desc, a.cap, a.d_counts, ls, a.maps, ev);
    if (e == cudaSuccess && a.status_dst)
        e = cudaMemcpyAsync(a.status_dst, a.status_src, sizeof(int) * a.nf, cudaMemcpyDeviceToDevice, ls.st);
    return e;
}
static void graph_key(const ChunkArgs& a, const OrbStreams& ls, const void** ptr, long long* num) {
    const void* p[8] = {a.d_kps, a.d_desc, a.d_counts, a.dense, a.l0, a.status_dst, a.status_src, ls.st};
    const long long n[8] = {a.nf, a.cap, a.f0, a.rows, a.cols, a.pitch, (long long)a.plane, (long long)(size_t)a.maps};
    memcpy(ptr, p, sizeof p);
    memcpy(num, n, sizeof n);
}
static bool graph_matches(const orb_extractor::ChunkGraph& G, const ChunkArgs& a, const void* const* ptr, const long long* num) {
    return G.exec && !memcmp(&G.plan, &a.P, sizeof(OrbPlan)) && !memcmp(G.ptr, ptr, sizeof G.ptr) && !memcmp(G.num, num, sizeof G.num);
}
// the device path's graph for these arguments, or the entry to replace
static orb_extractor::ChunkGraph& pick_graph(orb_extractor* h, int lane, const ChunkArgs& a) {
    const void* ptr[8];
    long long num[8];
    graph_key(a, h->lane(lane), ptr, num);
    for (int k = 0; k < orb_extractor::DEV_GRAPHS; ++k)
        if (graph_matches(h->devGraph[lane][k], a, ptr, num)) return h->devGraph[lane][k];
    int& nx = h->devGraphNext[lane];
    nx = (nx + 1) % orb_extractor::DEV_GRAPHS;
    return h->devGraph[lane][nx];
}
static cudaError_t run_chunk(orb_extractor* h, orb_extractor::ChunkGraph& G, const ChunkArgs& a, const OrbStreams& ls, bool graphed) {
    if (!graphed || h->profiling) return enqueue_chunk_kernels(a, ls, h->next_events());
    const void* ptr[8];
    long long num[8];
    graph_key(a, ls, ptr, num);
    if (!graph_matches(G, a, ptr, num)) {
        if (G.exec) cudaGraphExecDestroy(G.exec);
        G.exec = nullptr;
        const unsigned long long before = orbk_launch_count();
        cudaError_t e = cudaStreamBeginCapture(ls.st, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) {
            cudaGetLastError();
            h->graph_mode = 0;
            return enqueue_chunk_kernels(a, ls, h->next_events());
        }
        e = enqueue_chunk_kernels(a, ls, nullptr);
        cudaGraph_t g = nullptr;
        const cudaError_t e2 = cudaStreamEndCapture(ls.st, &g);
        if (e == cudaSuccess) e = e2;
        if (e == cudaSuccess) e = cudaGraphInstantiate(&G.exec, g, 0);
        if (g) cudaGraphDestroy(g);
        if (e != cudaSuccess) {
            // capture or instantiation refused (nothing has run: a capture only records): this handle goes back to plain
            // stream launches for good
            G.exec = nullptr;
            cudaGetLastError();
            orbk_count_launch(-(int)(orbk_launch_count() - before));
            h->graph_mode = 0;
            return enqueue_chunk_kernels(a, ls, h->next_events());
        }
        G.launches = (int)(orbk_launch_count() - before);
        orbk_count_launch(-G.launches);  // counted when the graph runs
        G.plan = a.P;
        memcpy(G.ptr, ptr, sizeof ptr);
        memcpy(G.num, num, sizeof num);
    }
    orbk_count_launch(G.launches);
    return cudaGraphLaunch(G.exec, ls.st);
}

static int extract_device_impl(orb_extractor* h, int n, const uint8_t* d_imgs, int rows, int cols, size_t stride, size_t frame_stride,
                               orb_keypoint* d_kps, uint8_t* d_desc, int cap, int* d_counts, bool ingest) {
    if (!h || !d_kps || !d_desc || !d_counts) return fail(ORB_ERR_INVALID, "null argument");
    if (n < 0 || n > h->max_batch) return fail(ORB_ERR_INVALID, "n=%d exceeds max_batch=%d", n, h->max_batch);
    if (n == 0) return ORB_OK;
    if (!d_imgs || rows <= 0 || cols <= 0) {  // empty image: silent return (ORBextractor.cc:444-445)
        CUDA_TRY(cudaSetDevice(h->device));
        CUDA_TRY(cudaMemsetAsync(d_counts, 0, sizeof(int) * n, h->stream));
        h->last_n = 0;
        return ORB_OK;
    }
    if (cap <= 0) return fail(ORB_ERR_INVALID, "cap must be positive");
    if (!ingest && ((stride & 3) || ((uintptr_t)d_imgs & 3) || (frame_stride & 3) || stride < (size_t)cols))
        return fail(ORB_ERR_INVALID, "device frames need 4-byte aligned base and strides >= cols");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = build_plan(h, rows, cols);
    if (rc != ORB_OK) return rc;
    OrbPlan P = h->plan;
    const DetectMaps* maps = h->d_maps;
    if (ingest) {
        const orb_extractor::Ingest& I = h->ing;
        if (stride < (size_t)I.scols * I.channels) return fail(ORB_ERR_INVALID, "stride < src_cols * channels");
        h->last_l0 = h->level0;
        CUDA_TRY(orbk_ingest(d_imgs, n, I.srows, I.scols, stride, frame_stride, I.channels, I.bgr, I.variant, I.d_mapx, I.d_mapy, rows, cols,
                             h->level0, h->plan.lv[0].pitch, h->plan.lv[0].plane, h->stream));
    } else if ((int)stride == h->plan.lv[0].pitch && frame_stride == h->plan.lv[0].plane && ((uintptr_t)d_imgs & 15) == 0) {
        // same layout as the internal level-0 buffer: read the caller's frames in place
        for (int l = 0; l < P.nlevels; ++l)
            if (P.lv[l].src == 0) P.lv[l].img = const_cast<uint8_t*>(d_imgs);
        if (h->user_base != d_imgs || n > h->user_n) {
            const OrbLevel& L0 = P.lv[0];
            if (L0.nTiles > 0)
                CUDA_TRY(orbk_encode_level_map(&h->maps_user.m[0], d_imgs, L0.cols, L0.rows, n, L0.pitch, L0.plane, DET_TILE_W, L0.boxH));
            if (L0.dTiles > 0)
                CUDA_TRY(orbk_encode_level_map(&h->maps_user.raw[0], d_imgs, L0.cols, L0.rows, n, L0.pitch, L0.plane, DSC_BOX_W, DSC_BOX_H));
            CUDA_TRY(orbk_encode_level_map(&h->maps_user.blr[0], d_imgs, L0.cols, L0.rows, n, L0.pitch, L0.plane, 256, BLUR_BOX_H));
            CUDA_TRY(cudaMemcpyAsync(&h->d_maps_user->blr[0], &h->maps_user.blr[0], sizeof(CUtensorMap), cudaMemcpyHostToDevice, h->stream));
            for (int l = 1; l < P.nlevels; ++l)  // the level(s) resized from level 0
                if (P.lv[l].src == l && P.lv[l].rszTiled && P.lv[l - 1].src == 0) {
                    CUDA_TRY(orbk_encode_level_map(&h->maps_user.rsz[l], d_imgs, L0.cols, L0.rows, n, L0.pitch, L0.plane, 256, RSZ_BOX_H));
                    CUDA_TRY(cudaMemcpyAsync(&h->d_maps_user->rsz[l], &h->maps_user.rsz[l], sizeof(CUtensorMap), cudaMemcpyHostToDevice, h->stream));
                }
            // stream-ordered update of the device copy (pageable source: staged before the call returns)
            CUDA_TRY(cudaMemcpyAsync(&h->d_maps_user->m[0], &h->maps_user.m[0], sizeof(CUtensorMap), cudaMemcpyHostToDevice, h->stream));
            CUDA_TRY(cudaMemcpyAsync(&h->d_maps_user->raw[0], &h->maps_user.raw[0], sizeof(CUtensorMap), cudaMemcpyHostToDevice, h->stream));
            h->user_base = d_imgs;
            h->user_n = n;
        }
        maps = h->d_maps_user;
        h->last_l0 = d_imgs;
    } else if (stride == (size_t)cols && frame_stride == (size_t)rows * cols && cols >= 4) {
        // densely packed device frames: one pitch-conversion kernel into the internal level-0 buffer
        h->last_l0 = h->level0;
        CUDA_TRY(orbk_repitch(d_imgs, 0, n, rows, cols, h->level0, h->plan.lv[0].pitch, h->plan.lv[0].plane, h->stream));
    } else {
        h->last_l0 = h->level0;
        // pitch conversion into the internal level-0 buffer (blur/describe share its pitch)
        for (int f = 0; f < n; ++f)
            CUDA_TRY(cudaMemcpy2DAsync(h->level0 + f * h->plan.lv[0].plane, h->plan.lv[0].pitch, d_imgs + f * frame_stride,
                                       stride, cols, rows, cudaMemcpyDeviceToDevice, h->stream));
    }
    h->last_n = n;
    const int lanes = h->profiling ? 1 : std::min(h->lanes, std::max(1, n / 16));
    if (lanes < 2) {
        CUDA_TRY(orbk_run_extract(P, n, reinterpret_cast<orb_keypoint_dev*>(d_kps), d_desc, cap, d_counts, h->streams(), maps, h->next_events()));
        return ORB_OK;
    }
    // independent slices of the batch on concurrent kernel lanes: one slice's launch gaps, wave tails and
    // latency-bound phases (octree) are filled by the other slices' kernels
    CUDA_TRY(cudaEventRecord(h->evSplit, h->stream));
    const int per = (n + lanes - 1) / lanes;
    for (int i = 0; i < lanes; ++i) {
        const int f0 = i * per, nf = std::min(per, n - f0);
        if (nf <= 0) break;
        const OrbStreams ls = h->lane(i);
        if (i > 0) CUDA_TRY(cudaStreamWaitEvent(ls.st, h->evSplit, 0));
        ChunkArgs ca;
        ca.P = plan_slice(P, f0);
        ca.nf = nf;
        ca.d_kps = reinterpret_cast<orb_keypoint_dev*>(d_kps) + (size_t)f0 * cap;
        ca.d_desc = d_desc + (size_t)f0 * cap * 32;
        ca.cap = cap;
        ca.d_counts = d_counts + f0;
        ca.maps = maps;
        ca.dense = nullptr;
        ca.f0 = f0;
        ca.rows = rows;
        ca.cols = cols;
        ca.l0 = nullptr;
        ca.pitch = 0;
        ca.plane = 0;
        ca.status_dst = nullptr;
        ca.status_src = nullptr;
        CUDA_TRY(run_chunk(h, pick_graph(h, i, ca), ca, ls, h->graph_mode == 1));
        if (i > 0) {
            CUDA_TRY(cudaEventRecord(h->laneMerge[i], ls.st));
            CUDA_TRY(cudaStreamWaitEvent(h->stream, h->laneMerge[i], 0));
        }
    }
    return ORB_OK;
}

extern "C" int orb_extract_batch_device(orb_extractor* h, int n, const uint8_t* d_imgs, int rows, int cols, size_t stride,
                                        size_t frame_stride, orb_keypoint* d_kps, uint8_t* d_desc, int cap, int* d_counts) {
    return extract_device_impl(h, n, d_imgs, rows, cols, stride, frame_stride, d_kps, d_desc, cap, d_counts, false);
}

extern "C" int orb_extractor_sync(orb_extractor* h) {
    if (!h) return fail(ORB_ERR_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->last_n > 0 && h->plan.status) return check_status(h, h->last_n);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ORB_OK;
}

extern "C" void* orb_extractor_stream(orb_extractor* h) { return h ? (void*)h->stream : nullptr; }

extern "C" int orb_extractor_set_profiling(orb_extractor* h, int on) {
    if (!h) return fail(ORB_ERR_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->clear_events();
    h->profiling = on != 0;
    return ORB_OK;
}

extern "C" int orb_extractor_stage_times(orb_extractor* h, double* ms_sum, int* ncalls) {
    if (!h || !ms_sum || !ncalls) return fail(ORB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream2));
    for (int i = 0; i < ORB_STAGES; ++i) ms_sum[i] = 0.0;
    const size_t per = ORB_EVENTS;
    *ncalls = (int)(h->events.size() / per);
    // event pairs of the stages: pyramid, detect, octree on the main stream; blur on the second
    // stream (it overlaps detect + octree); describe after the join
    static const int first[ORB_STAGES] = {0, 1, 2, 6, 4}, last[ORB_STAGES] = {1, 2, 3, 7, 5};
    for (size_t c = 0; c < h->events.size() / per; ++c)
        for (int i = 0; i < ORB_STAGES; ++i) {
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, h->events[c * per + first[i]], h->events[c * per + last[i]]));
            ms_sum[i] += ms;
        }
    return ORB_OK;
}

extern "C" const char* orb_stage_name(int stage) {
    static const char* names[ORB_STAGES] = {"pyramid", "detect", "octree", "blur", "describe"};
    return stage >= 0 && stage < ORB_STAGES ? names[stage] : "";
}

// Everything of one host batch that is enqueued on the device: called by submit_impl, which cleans up after a failure.
// rows x cols: the image the extractor sees.  ingest: imgs holds raw frames described by h->ing (stride / frame_stride are
// theirs), which land in the slot's dense buffer and reach level 0 through k_ingest instead of k_repitch.
static int enqueue_batch(orb_extractor* h, orb_extractor::HostSlot& S, int n, const uint8_t* imgs, int rows, int cols, size_t stride,
                         size_t frame_stride, orb_keypoint* kps, uint8_t* desc, int cap, int* counts, bool ingest, bool pipelined) {
    const orb_extractor::Ingest& I = h->ing;
    const size_t rowBytes = ingest ? (size_t)I.scols * I.channels : (size_t)cols;
    const int srcRows = ingest ? I.srows : rows;
    const OrbLevel& L0 = h->plan.lv[0];
    // Pipeline over chunks of frames: H2D copy (streamIn) -> kernels (kernel lanes) -> counts, status flags and result
    // rows D2H (streamOut), all enqueued here, so that PCIe in, compute and PCIe out of neighbouring chunks -- and of the
    // neighbouring batches in flight -- overlap and the host never sits between a chunk's kernels and its result copy.
    // The number of result rows per frame is not known on the host at this point: the copies take the largest count of
    // the previous batch plus a margin (h->spec_rows; the caller's capacity before anything is known), and
    // orb_extract_batch_wait tops up the rare frame that produced more.  While per-stage profiling is on, one chunk is
    // used (stage times then describe whole-batch launches).
    // Fine chunking only pays when this batch has nothing else to overlap with: with another batch in flight the
    // copies of one batch already hide behind the kernels of the other, so one chunk per kernel lane keeps the
    // launches large and the host-side enqueue cost low (measured on B200 at 64 frames, three batches in flight:
    // 74.7k frames/s with two 32-frame chunks, 71.3k with one chunk, 51.7k with four 16-frame chunks).
    // Measured too: while the copy engines move this data the kernels run ~15 % slower than on resident frames (a
    // copy by SM loads from mapped host memory was worse still), which is what keeps e2e below the resident rate.
    const int chunkFrames = getenv("ORB_B200_CHUNK") ? std::max(1, atoi(getenv("ORB_B200_CHUNK"))) : (pipelined ? std::max(16, (n + h->lanes - 1) / std::max(1, h->lanes)) : 16);
    int nchunks = h->profiling ? 1 : std::min<int>(orb_extractor::MAX_CHUNKS, std::max(1, n / chunkFrames));
    const int per = (n + nchunks - 1) / nchunks;
    nchunks = (n + per - 1) / per;
    const size_t landFrame = rowBytes * srcRows;  // one frame in the dense landing buffer
    const bool linear = stride == rowBytes && frame_stride == landFrame;
    const bool dense = ingest || (linear && cols >= 4);
    if (dense && S.dense_cap < (size_t)n * landFrame + 16) {
        CUDA_TRY(cudaStreamSynchronize(h->streamIn));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        if (S.d_dense) cudaFree(S.d_dense);
        S.d_dense = nullptr;
        S.dense_cap = 0;
        const size_t want = (size_t)h->max_batch * landFrame + 16;
        CUDA_TRY(cudaMalloc((void**)&S.d_dense, want));
        S.dense_cap = want;
    }
    const int spec = h->eager_out ? std::min(cap, h->spec_rows > 0 ? h->spec_rows : cap) : 0;
    // everything enqueued so far on the main stream (the previous batch's kernels: every lane joins there)
    // finishes before this batch touches the shared level / scratch buffers; a batch's status flags are copied
    // into its own slot at the end of its kernels, so the next batch does not have to wait for the result copies
    h->last_l0 = h->level0;
    CUDA_TRY(cudaEventRecord(h->evFree, h->stream));
    if (!dense) CUDA_TRY(cudaStreamWaitEvent(h->streamIn, h->evFree, 0));  // non-dense copies write level 0 directly
    h->last_n = n;
    for (int c = 0; c < nchunks; ++c) {
        const int f0 = c * per, nf = std::min(per, n - f0);
        const OrbStreams ls = h->lane(nchunks > 1 ? c % std::max(1, h->lanes) : 0);
        bool repitchHere = false;  // the pitch conversion joins the chunk's kernel sequence (run_chunk)
        if (ls.st != h->stream) CUDA_TRY(cudaStreamWaitEvent(ls.st, h->evFree, 0));
        if (dense) {
            // densely packed frames: one linear copy into this batch's own landing buffer (it may run while the
            // previous batch still computes), then the pitch conversion on the kernel lane
            uint8_t* land = S.d_dense + (size_t)f0 * landFrame;
            if (linear) {
                CUDA_TRY(cudaMemcpyAsync(land, imgs + (size_t)f0 * frame_stride, (size_t)nf * landFrame, cudaMemcpyHostToDevice, h->streamIn));
            } else {  // raw frames with padded rows: packed on the way in
                for (int f = 0; f < nf; ++f)
                    CUDA_TRY(cudaMemcpy2DAsync(land + (size_t)f * landFrame, rowBytes, imgs + (size_t)(f0 + f) * frame_stride, stride, rowBytes,
                                               srcRows, cudaMemcpyHostToDevice, h->streamIn));
            }
            CUDA_TRY(cudaEventRecord(S.evIn[c], h->streamIn));
            CUDA_TRY(cudaStreamWaitEvent(ls.st, S.evIn[c], 0));
            if (ingest) {
                CUDA_TRY(orbk_ingest(land, nf, I.srows, I.scols, rowBytes, landFrame, I.channels, I.bgr, I.variant, I.d_mapx, I.d_mapy, rows, cols,
                                     h->level0 + f0 * L0.plane, L0.pitch, L0.plane, ls.st));
            } else {
                // the chunk's first byte need not be word aligned (odd-area frames): the kernel reads relative to the
                // aligned base of the landing buffer
                repitchHere = true;
            }
        } else {
            if (nf == 1 || frame_stride == stride * (size_t)rows) {
                // frames are consecutive rows on both sides (device plane == pitch * rows): one 2D copy
                CUDA_TRY(cudaMemcpy2DAsync(h->level0 + f0 * L0.plane, L0.pitch, imgs + f0 * frame_stride, stride, cols, (size_t)rows * nf,
                                           cudaMemcpyHostToDevice, h->streamIn));
            } else {
                for (int f = f0; f < f0 + nf; ++f)
                    CUDA_TRY(cudaMemcpy2DAsync(h->level0 + f * L0.plane, L0.pitch, imgs + f * frame_stride, stride, cols, rows,
                                               cudaMemcpyHostToDevice, h->streamIn));
            }
            CUDA_TRY(cudaEventRecord(S.evIn[c], h->streamIn));
            CUDA_TRY(cudaStreamWaitEvent(ls.st, S.evIn[c], 0));
        }
        ChunkArgs ca;
        ca.P = plan_slice(h->plan, f0);
        ca.nf = nf;
        ca.d_kps = S.d_kps + (size_t)f0 * S.out_cap;
        ca.d_desc = S.d_desc + (size_t)f0 * S.out_cap * 32;
        ca.cap = S.out_cap;
        ca.d_counts = S.d_counts + f0;
        ca.maps = h->d_maps;
        ca.dense = repitchHere ? S.d_dense : nullptr;
        ca.f0 = f0;
        ca.rows = rows;
        ca.cols = cols;
        ca.l0 = h->level0 + f0 * L0.plane;
        ca.pitch = L0.pitch;
        ca.plane = L0.plane;
        ca.status_dst = S.d_counts + h->max_batch + f0;
        ca.status_src = h->plan.status + f0;
        CUDA_TRY(run_chunk(h, S.graph[c], ca, ls, h->graph_mode == 1 || (h->graph_mode == 2 && nchunks == 1)));
        CUDA_TRY(cudaEventRecord(S.evDone[c], ls.st));
        if (ls.st != h->stream) CUDA_TRY(cudaStreamWaitEvent(h->stream, S.evDone[c], 0));  // the main stream stays the join point
        // results of chunk c: counts and status flags first (small), then `spec` rows of every frame
        CUDA_TRY(cudaStreamWaitEvent(h->streamOut, S.evDone[c], 0));
        CUDA_TRY(cudaMemcpyAsync(counts + f0, S.d_counts + f0, sizeof(int) * nf, cudaMemcpyDeviceToHost, h->streamOut));
        CUDA_TRY(cudaMemcpyAsync(S.h_status + f0, S.d_counts + h->max_batch + f0, sizeof(int) * nf, cudaMemcpyDeviceToHost, h->streamOut));
        if (spec > 0) {
            CUDA_TRY(cudaMemcpy2DAsync(kps + (size_t)f0 * cap, (size_t)cap * sizeof(orb_keypoint), S.d_kps + (size_t)f0 * S.out_cap,
                                       (size_t)S.out_cap * sizeof(orb_keypoint), (size_t)spec * sizeof(orb_keypoint), nf, cudaMemcpyDeviceToHost,
                                       h->streamOut));
            CUDA_TRY(cudaMemcpy2DAsync(desc + (size_t)f0 * cap * 32, (size_t)cap * 32, S.d_desc + (size_t)f0 * S.out_cap * 32,
                                       (size_t)S.out_cap * 32, (size_t)spec * 32, nf, cudaMemcpyDeviceToHost, h->streamOut));
        }
        CUDA_TRY(cudaEventRecord(S.evCnt[c], h->streamOut));
    }
    CUDA_TRY(cudaEventRecord(S.evOut, h->streamOut));
    S.n = n;
    S.cap = cap;
    S.nchunks = nchunks;
    S.per = per;
    S.spec_w = spec;
    S.kps = kps;
    S.desc = desc;
    S.counts = counts;
    return ORB_OK;
}

static int submit_impl(orb_extractor* h, int n, const uint8_t* imgs, int rows, int cols, size_t stride, size_t frame_stride,
                       orb_keypoint* kps, uint8_t* desc, int cap, int* counts, int* ticket, bool ingest) {
    if (!h || !counts || !ticket) return fail(ORB_ERR_INVALID, "null argument");
    *ticket = -1;
    if (n < 0 || n > h->max_batch) return fail(ORB_ERR_INVALID, "n=%d exceeds max_batch=%d", n, h->max_batch);
    if (n == 0) return ORB_OK;
    if (!imgs || rows <= 0 || cols <= 0) {  // empty image: silent return (ORBextractor.cc:444-445)
        for (int f = 0; f < n; ++f) counts[f] = 0;
        return ORB_OK;
    }
    if (!kps || !desc || cap <= 0) return fail(ORB_ERR_INVALID, "null output buffer");
    const orb_extractor::Ingest& I = h->ing;
    if (stride < (ingest ? (size_t)I.scols * I.channels : (size_t)cols))
        return fail(ORB_ERR_INVALID, ingest ? "stride < src_cols * channels" : "stride < cols");
    int slot = -1;
    bool pipelined = false;
    for (int k = 0; k < orb_extractor::NUM_SLOTS; ++k) {
        if (h->slot[k].busy) pipelined = true;
        else if (slot < 0) slot = k;
    }
    if (slot < 0) return fail(ORB_ERR_INVALID, "three batches are already in flight: call orb_extract_batch_wait first");
    orb_extractor::HostSlot& S = h->slot[slot];
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->plan.rows != rows || h->plan.cols != cols) {
        if (pipelined) return fail(ORB_ERR_INVALID, "cannot change the image shape while a batch is in flight");
        h->spec_rows = 0;
    }
    int rc = build_plan(h, rows, cols);
    if (rc != ORB_OK) return rc;
    rc = ensure_out(h, S, cap);
    if (rc != ORB_OK) return rc;
    rc = enqueue_batch(h, S, n, imgs, rows, cols, stride, frame_stride, kps, desc, cap, counts, ingest, pipelined);
    if (rc != ORB_OK) {
        // part of the batch may be enqueued and still reading / writing the caller's buffers: drain it before the
        // error returns (the message of the first failure is kept)
        const std::string msg = orb_last_error();
        cudaDeviceSynchronize();
        cudaGetLastError();
        return fail(rc, "%s", msg.c_str());
    }
    S.busy = true;
    *ticket = slot;
    return ORB_OK;
}

extern "C" int orb_extract_batch_submit(orb_extractor* h, int n, const uint8_t* imgs, int rows, int cols, size_t stride,
                                        size_t frame_stride, orb_keypoint* kps, uint8_t* desc, int cap, int* counts, int* ticket) {
    return submit_impl(h, n, imgs, rows, cols, stride, frame_stride, kps, desc, cap, counts, ticket, false);
}

// ------------------------------------------------------------------------------------------
// image ingest (reference src/Tracking.cc:118-126, Examples/Stereo/stereo_euroc.cc:136-137)
// ------------------------------------------------------------------------------------------
extern "C" int orb_extractor_set_ingest(orb_extractor* h, const orb_ingest_config* cfg) {
    if (!h) return fail(ORB_ERR_INVALID, "null handle");
    for (int k = 0; k < orb_extractor::NUM_SLOTS; ++k)
        if (h->slot[k].busy) return fail(ORB_ERR_INVALID, "cannot change the ingest configuration while a batch is in flight");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    orb_extractor::Ingest& I = h->ing;
    if (I.d_mapx) cudaFree(I.d_mapx);
    if (I.d_mapy) cudaFree(I.d_mapy);
    I = orb_extractor::Ingest();
    if (!cfg) return ORB_OK;
    if (cfg->src_rows <= 0 || cfg->src_cols <= 0 || cfg->dst_rows <= 0 || cfg->dst_cols <= 0)
        return fail(ORB_ERR_INVALID, "ingest: sizes must be positive");
    if (cfg->channels != 1 && cfg->channels != 3 && cfg->channels != 4) return fail(ORB_ERR_INVALID, "ingest: channels must be 1, 3 or 4");
    if (cfg->gray_variant != 3 && cfg->gray_variant != 4) return fail(ORB_ERR_INVALID, "ingest: gray_variant must be 3 or 4");
    if ((cfg->map_x == nullptr) != (cfg->map_y == nullptr)) return fail(ORB_ERR_INVALID, "ingest: map_x and map_y go together");
    if (!cfg->map_x && (cfg->dst_rows != cfg->src_rows || cfg->dst_cols != cfg->src_cols))
        return fail(ORB_ERR_INVALID, "ingest: without maps dst must equal src");
    if (cfg->map_x) {
        const size_t bytes = sizeof(float) * (size_t)cfg->dst_rows * cfg->dst_cols;
        CUDA_TRY(cudaMalloc((void**)&I.d_mapx, bytes));
        CUDA_TRY(cudaMalloc((void**)&I.d_mapy, bytes));
        CUDA_TRY(cudaMemcpy(I.d_mapx, cfg->map_x, bytes, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(I.d_mapy, cfg->map_y, bytes, cudaMemcpyHostToDevice));
    }
    I.srows = cfg->src_rows;
    I.scols = cfg->src_cols;
    I.channels = cfg->channels;
    I.bgr = cfg->bgr != 0;
    I.variant = cfg->gray_variant;
    I.drows = cfg->dst_rows;
    I.dcols = cfg->dst_cols;
    I.on = true;
    return ORB_OK;
}

extern "C" int orb_ingest_extract_batch_submit(orb_extractor* h, int n, const uint8_t* raw, size_t stride, size_t frame_stride,
                                               orb_keypoint* kps, uint8_t* desc, int cap, int* counts, int* ticket) {
    if (!h) return fail(ORB_ERR_INVALID, "null handle");
    if (!h->ing.on) return fail(ORB_ERR_INVALID, "no ingest configuration: call orb_extractor_set_ingest first");
    return submit_impl(h, n, raw, h->ing.drows, h->ing.dcols, stride, frame_stride, kps, desc, cap, counts, ticket, true);
}

extern "C" int orb_ingest_extract_batch(orb_extractor* h, int n, const uint8_t* raw, size_t stride, size_t frame_stride,
                                        orb_keypoint* kps, uint8_t* desc, int cap, int* counts) {
    int ticket = -1;
    int rc = orb_ingest_extract_batch_submit(h, n, raw, stride, frame_stride, kps, desc, cap, counts, &ticket);
    if (rc != ORB_OK) return rc;
    return orb_extract_batch_wait(h, ticket);
}

extern "C" int orb_ingest_extract_batch_device(orb_extractor* h, int n, const uint8_t* d_raw, size_t stride, size_t frame_stride,
                                               orb_keypoint* d_kps, uint8_t* d_desc, int cap, int* d_counts) {
    if (!h) return fail(ORB_ERR_INVALID, "null handle");
    if (!h->ing.on) return fail(ORB_ERR_INVALID, "no ingest configuration: call orb_extractor_set_ingest first");
    return extract_device_impl(h, n, d_raw, h->ing.drows, h->ing.dcols, stride, frame_stride, d_kps, d_desc, cap, d_counts, true);
}

extern "C" int orb_extract_batch_wait(orb_extractor* h, int ticket) {
    if (!h) return fail(ORB_ERR_INVALID, "null handle");
    if (ticket < 0) return ORB_OK;  // nothing was submitted (empty batch)
    if (ticket >= orb_extractor::NUM_SLOTS || !h->slot[ticket].busy) return fail(ORB_ERR_INVALID, "bad ticket");
    orb_extractor::HostSlot& S = h->slot[ticket];
    S.busy = false;
    CUDA_TRY(cudaSetDevice(h->device));
    // eager copies: counts, status flags and the first spec_w rows of every frame arrive together
    if (S.spec_w > 0) CUDA_TRY(cudaEventSynchronize(S.evOut));
    int maxAll = 0, bad = -1;
    bool topup = false;
    for (int c = 0; c < S.nchunks; ++c) {
        const int f0 = c * S.per, nf = std::min(S.per, S.n - f0);
        if (S.spec_w == 0) CUDA_TRY(cudaEventSynchronize(S.evCnt[c]));  // counts of chunk c are on the host, its kernels are done
        int maxc = 0;
        for (int f = f0; f < f0 + nf; ++f) {
            maxc = std::max(maxc, S.counts[f]);
            if (S.h_status[f] && bad < 0) bad = f;
        }
        maxAll = std::max(maxAll, maxc);
        const int w = std::min(maxc, S.cap);
        if (w > S.spec_w) {  // a frame kept more rows than were copied ahead of the counts: fetch the rest of the chunk
            const size_t o = (size_t)S.spec_w, m = (size_t)(w - S.spec_w);
            CUDA_TRY(cudaMemcpy2DAsync(S.kps + (size_t)f0 * S.cap + o, (size_t)S.cap * sizeof(orb_keypoint),
                                       S.d_kps + (size_t)f0 * S.out_cap + o, (size_t)S.out_cap * sizeof(orb_keypoint),
                                       m * sizeof(orb_keypoint), nf, cudaMemcpyDeviceToHost, h->streamCnt));
            CUDA_TRY(cudaMemcpy2DAsync(S.desc + ((size_t)f0 * S.cap + o) * 32, (size_t)S.cap * 32, S.d_desc + ((size_t)f0 * S.out_cap + o) * 32,
                                       (size_t)S.out_cap * 32, m * 32, nf, cudaMemcpyDeviceToHost, h->streamCnt));
            topup = true;
        }
    }
    if (topup) CUDA_TRY(cudaStreamSynchronize(h->streamCnt));
    // next batch: this batch's largest count + ~3 %, in whole 64-row steps
    h->spec_rows = std::max(64, (maxAll + maxAll / 32 + 63) / 64 * 64);
    if (bad >= 0) return fail(ORB_ERR_UNSEPARABLE, "frame %d: octree cannot separate its keys (reference would not terminate)", bad);
    if (maxAll > S.cap) return fail(ORB_ERR_CAPACITY, "frame needs %d keypoints, cap is %d", maxAll, S.cap);
    return ORB_OK;
}

extern "C" int orb_extract_batch(orb_extractor* h, int n, const uint8_t* imgs, int rows, int cols, size_t stride,
                                 size_t frame_stride, orb_keypoint* kps, uint8_t* desc, int cap, int* counts) {
    int ticket = -1;
    int rc = orb_extract_batch_submit(h, n, imgs, rows, cols, stride, frame_stride, kps, desc, cap, counts, &ticket);
    if (rc != ORB_OK) return rc;
    return orb_extract_batch_wait(h, ticket);
}

extern "C" int orb_extract(orb_extractor* h, const uint8_t* img, int rows, int cols, size_t stride, orb_keypoint* kps,
                           uint8_t* desc, int cap, int* count) {
    if (!count) return fail(ORB_ERR_INVALID, "null count");
    return orb_extract_batch(h, 1, img, rows, cols, stride, stride * (size_t)std::max(rows, 0), kps, desc, cap, count);
}

extern "C" int orb_get_pyramid_level(orb_extractor* h, int frame, int level, uint8_t* dst, size_t dst_stride, int* rows, int* cols) {
    if (!h) return fail(ORB_ERR_INVALID, "null handle");
    if (h->plan.rows == 0) return fail(ORB_ERR_INVALID, "no frame extracted yet");
    if (level < 0 || level >= h->plan.nlevels || frame < 0 || frame >= h->max_batch) return fail(ORB_ERR_INVALID, "bad frame/level");
    const OrbLevel& L = h->plan.lv[level];
    if (rows) *rows = L.rows;
    if (cols) *cols = L.cols;
    if (!dst) return ORB_OK;
    if (dst_stride < (size_t)L.cols) return fail(ORB_ERR_INVALID, "dst_stride < cols");
    CUDA_TRY(cudaSetDevice(h->device));
    const uint8_t* base = (L.src == 0 && h->last_l0) ? h->last_l0 : L.img;
    CUDA_TRY(cudaMemcpy2DAsync(dst, dst_stride, base + (size_t)frame * L.plane, L.pitch, L.cols, L.rows, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ORB_OK;
}

extern "C" int orb_get_pyramid_levels(orb_extractor* h, int frame, uint8_t* const* dst, const size_t* dst_stride) {
    if (!h || !dst || !dst_stride) return fail(ORB_ERR_INVALID, "null argument");
    if (h->plan.rows == 0) return fail(ORB_ERR_INVALID, "no frame extracted yet");
    if (frame < 0 || frame >= h->max_batch) return fail(ORB_ERR_INVALID, "bad frame");
    CUDA_TRY(cudaSetDevice(h->device));
    for (int l = 0; l < h->plan.nlevels; ++l) {
        if (!dst[l]) continue;
        const OrbLevel& L = h->plan.lv[l];
        if (dst_stride[l] < (size_t)L.cols) return fail(ORB_ERR_INVALID, "dst_stride < cols (level %d)", l);
        const uint8_t* base = (L.src == 0 && h->last_l0) ? h->last_l0 : L.img;
        CUDA_TRY(cudaMemcpy2DAsync(dst[l], dst_stride[l], base + (size_t)frame * L.plane, L.pitch, L.cols, L.rows, cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ORB_OK;
}

extern "C" int orb_extractor_level_stats(orb_extractor* h, int frame, int32_t* candidates, int32_t* kept) {
    if (!h || h->plan.rows == 0) return fail(ORB_ERR_INVALID, "no frame extracted yet");
    if (frame < 0 || frame >= h->max_batch) return fail(ORB_ERR_INVALID, "bad frame");
    CUDA_TRY(cudaSetDevice(h->device));
    int c[ORB_MAX_LEVELS], k[ORB_MAX_LEVELS];
    CUDA_TRY(cudaMemcpyAsync(c, h->plan.candCount + frame * ORB_MAX_LEVELS, sizeof c, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(k, h->plan.keptCount + frame * ORB_MAX_LEVELS, sizeof k, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int l = 0; l < h->plan.nlevels; ++l) {
        if (candidates) candidates[l] = c[h->plan.lv[l].src];
        if (kept) kept[l] = k[l];
    }
    return ORB_OK;
}

// ------------------------------------------------------------------------------------------
// matcher handle
// ------------------------------------------------------------------------------------------
int orb_matcher_scratch(orb_matcher* m, int slot, size_t bytes, void** out) {
    if (bytes > m->cap[slot]) {
        CUDA_TRY(cudaStreamSynchronize(m->stream));
        if (m->buf[slot]) cudaFree(m->buf[slot]);
        m->buf[slot] = nullptr;
        m->cap[slot] = 0;
        size_t want = std::max<size_t>(bytes + bytes / 4, 4096);
        CUDA_TRY(cudaMalloc(&m->buf[slot], want));
        m->cap[slot] = want;
    }
    *out = m->buf[slot];
    return ORB_OK;
}
#define scratch orb_matcher_scratch

extern "C" int orb_matcher_create(int device, orb_matcher** out) {
    if (!out) return fail(ORB_ERR_INVALID, "null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ORB_ERR_CUDA, "no CUDA device: %s (liborb_b200 has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(ORB_ERR_INVALID, "device %d out of range", device);
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(orbk_match_init_device());
    orb_matcher* m = new orb_matcher();
    m->device = device;
    cudaError_t ce = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) {
        delete m;
        return fail(ORB_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(ce));
    }
    *out = m;
    return ORB_OK;
}

extern "C" void orb_matcher_destroy(orb_matcher* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream);
    for (int i = 0; i < orb_matcher::kSlots; ++i)
        if (m->buf[i]) cudaFree(m->buf[i]);
    cudaStreamDestroy(m->stream);
    delete m;
}

extern "C" int orb_matcher_sync(orb_matcher* m) {
    if (!m) return fail(ORB_ERR_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(m->device));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return ORB_OK;
}
extern "C" void* orb_matcher_stream(orb_matcher* m) { return m ? (void*)m->stream : nullptr; }

extern "C" int orb_match_all_batch(orb_matcher* m, int npairs, const uint8_t* q, const int32_t* nq, size_t q_stride,
                                   const uint8_t* t, const int32_t* nt, size_t t_stride, int32_t* best_idx,
                                   int32_t* best_dist, int32_t* second_dist, size_t out_stride, int on_device) {
    if (!m || !nq || !nt || !best_idx || !best_dist || !second_dist) return fail(ORB_ERR_INVALID, "null argument");
    if (npairs <= 0) return ORB_OK;
    if ((q_stride & 15) || (t_stride & 15)) return fail(ORB_ERR_INVALID, "descriptor strides must be multiples of 16 bytes");
    CUDA_TRY(cudaSetDevice(m->device));
    if (on_device) {
        if (((uintptr_t)q & 15) || ((uintptr_t)t & 15)) return fail(ORB_ERR_INVALID, "descriptor arrays must be 16-byte aligned");
        // max query count is not known on the host: launch for the stride
        const int max_nq = (int)(out_stride);
        CUDA_TRY(orbk_match_all(q, nq, q_stride, t, nt, t_stride, npairs, max_nq, -1, best_idx, best_dist, second_dist, out_stride, m->stream));
        return ORB_OK;
    }
    int max_nq = 0, max_nt = 0;
    for (int p = 0; p < npairs; ++p) {
        if (nq[p] < 0 || nt[p] < 0) return fail(ORB_ERR_INVALID, "negative count");
        max_nq = std::max(max_nq, nq[p]);
        max_nt = std::max(max_nt, nt[p]);
    }
    if ((size_t)max_nq > out_stride) return fail(ORB_ERR_INVALID, "out_stride < nq");
    if ((size_t)max_nq * 32 > q_stride && npairs > 1) return fail(ORB_ERR_INVALID, "q_stride < nq*32");
    if ((size_t)max_nt * 32 > t_stride && npairs > 1) return fail(ORB_ERR_INVALID, "t_stride < nt*32");
    void *dq, *dt, *dn, *dout;
    const size_t qbytes = (size_t)(npairs - 1) * q_stride + (size_t)max_nq * 32;
    const size_t tbytes = (size_t)(npairs - 1) * t_stride + (size_t)max_nt * 32;
    const size_t obytes = ((size_t)(npairs - 1) * out_stride + max_nq) * sizeof(int);
    int rc;
    if ((rc = scratch(m, 0, qbytes + 32, &dq)) || (rc = scratch(m, 1, tbytes + 32, &dt)) ||
        (rc = scratch(m, 2, sizeof(int) * 2 * npairs, &dn)) || (rc = scratch(m, 3, obytes * 3, &dout)))
        return rc;
    if (qbytes) CUDA_TRY(cudaMemcpyAsync(dq, q, qbytes, cudaMemcpyHostToDevice, m->stream));
    if (tbytes) CUDA_TRY(cudaMemcpyAsync(dt, t, tbytes, cudaMemcpyHostToDevice, m->stream));
    int* dnq = (int*)dn;
    int* dnt = dnq + npairs;
    CUDA_TRY(cudaMemcpyAsync(dnq, nq, sizeof(int) * npairs, cudaMemcpyHostToDevice, m->stream));
    CUDA_TRY(cudaMemcpyAsync(dnt, nt, sizeof(int) * npairs, cudaMemcpyHostToDevice, m->stream));
    int* o0 = (int*)dout;
    int* o1 = (int*)((char*)dout + obytes);
    int* o2 = (int*)((char*)dout + 2 * obytes);
    CUDA_TRY(orbk_match_all((const uint8_t*)dq, dnq, q_stride, (const uint8_t*)dt, dnt, t_stride, npairs, max_nq, max_nt, o0, o1, o2, out_stride, m->stream));
    if (obytes) {
        CUDA_TRY(cudaMemcpyAsync(best_idx, o0, obytes, cudaMemcpyDeviceToHost, m->stream));
        CUDA_TRY(cudaMemcpyAsync(best_dist, o1, obytes, cudaMemcpyDeviceToHost, m->stream));
        CUDA_TRY(cudaMemcpyAsync(second_dist, o2, obytes, cudaMemcpyDeviceToHost, m->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return ORB_OK;
}

extern "C" int orb_match_all(orb_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* best_idx,
                             int32_t* best_dist, int32_t* second_dist) {
    if (nq < 0 || nt < 0) return fail(ORB_ERR_INVALID, "negative count");
    if (nq == 0) return ORB_OK;
    const int32_t a = nq, b = nt;
    return orb_match_all_batch(m, 1, q, &a, (size_t)(nq + 1) * 32, t, &b, (size_t)(nt + 1) * 32, best_idx, best_dist, second_dist,
                               (size_t)nq, 0);
}

extern "C" int orb_match_csr(orb_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, const int32_t* offsets,
                             const int32_t* cand, int tie_rule, int max_dist, int32_t* best_idx, int32_t* best_dist,
                             int32_t* second_dist) {
    if (!m || !offsets || !best_idx || !best_dist || !second_dist) return fail(ORB_ERR_INVALID, "null argument");
    if (nq < 0 || nt < 0) return fail(ORB_ERR_INVALID, "negative count");
    if (nq == 0) return ORB_OK;
    if (tie_rule != ORB_TIE_FIRST_MIN && tie_rule != ORB_TIE_LAST_MIN) return fail(ORB_ERR_INVALID, "bad tie_rule");
    const int ncand = offsets[nq];
    if (ncand < 0 || offsets[0] != 0) return fail(ORB_ERR_INVALID, "offsets must start at 0 and be non-decreasing");
    for (int i = 0; i < nq; ++i)
        if (offsets[i + 1] < offsets[i]) return fail(ORB_ERR_INVALID, "offsets must be non-decreasing");
    for (int c = 0; c < ncand; ++c)
        if (cand[c] < 0 || cand[c] >= nt) return fail(ORB_ERR_INVALID, "candidate %d out of range", c);
    CUDA_TRY(cudaSetDevice(m->device));
    void *dq, *dt, *dofs, *dout;
    int rc;
    if ((rc = scratch(m, 0, (size_t)nq * 32 + 32, &dq)) || (rc = scratch(m, 1, (size_t)nt * 32 + 32, &dt)) ||
        (rc = scratch(m, 2, sizeof(int) * ((size_t)nq + 1 + ncand), &dofs)) || (rc = scratch(m, 3, sizeof(int) * 3 * (size_t)nq, &dout)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
    if (nt) CUDA_TRY(cudaMemcpyAsync(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice, m->stream));
    int* d_off = (int*)dofs;
    int* d_cand = d_off + nq + 1;
    CUDA_TRY(cudaMemcpyAsync(d_off, offsets, sizeof(int) * ((size_t)nq + 1), cudaMemcpyHostToDevice, m->stream));
    if (ncand) CUDA_TRY(cudaMemcpyAsync(d_cand, cand, sizeof(int) * (size_t)ncand, cudaMemcpyHostToDevice, m->stream));
    int* o = (int*)dout;
    CUDA_TRY(orbk_match_csr((const uint8_t*)dq, nq, (const uint8_t*)dt, d_off, d_cand, tie_rule == ORB_TIE_LAST_MIN, max_dist, o, o + nq,
                            o + 2 * (size_t)nq, m->stream));
    CUDA_TRY(cudaMemcpyAsync(best_idx, o, sizeof(int) * nq, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaMemcpyAsync(best_dist, o + nq, sizeof(int) * nq, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaMemcpyAsync(second_dist, o + 2 * (size_t)nq, sizeof(int) * nq, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return ORB_OK;
}

extern "C" int orb_distances_csr(orb_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, const int32_t* offsets,
                                 const int32_t* cand, int32_t* dist) {
    if (!m || !offsets) return fail(ORB_ERR_INVALID, "null argument");
    if (nq < 0 || nt < 0) return fail(ORB_ERR_INVALID, "negative count");
    if (nq == 0) return ORB_OK;
    const int ncand = offsets[nq];
    if (ncand < 0 || offsets[0] != 0) return fail(ORB_ERR_INVALID, "offsets must start at 0 and be non-decreasing");
    for (int i = 0; i < nq; ++i)
        if (offsets[i + 1] < offsets[i]) return fail(ORB_ERR_INVALID, "offsets must be non-decreasing");
    if (ncand == 0) return ORB_OK;
    if (!cand || !dist) return fail(ORB_ERR_INVALID, "null argument");
    for (int c = 0; c < ncand; ++c)
        if (cand[c] < 0 || cand[c] >= nt) return fail(ORB_ERR_INVALID, "candidate %d out of range", c);
    CUDA_TRY(cudaSetDevice(m->device));
    void *dq, *dt, *dofs, *dout;
    int rc;
    if ((rc = scratch(m, 0, (size_t)nq * 32 + 32, &dq)) || (rc = scratch(m, 1, (size_t)nt * 32 + 32, &dt)) ||
        (rc = scratch(m, 2, sizeof(int) * ((size_t)nq + 1 + ncand), &dofs)) || (rc = scratch(m, 3, sizeof(int) * (size_t)ncand, &dout)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
    CUDA_TRY(cudaMemcpyAsync(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice, m->stream));
    int* d_off = (int*)dofs;
    int* d_cand = d_off + nq + 1;
    CUDA_TRY(cudaMemcpyAsync(d_off, offsets, sizeof(int) * ((size_t)nq + 1), cudaMemcpyHostToDevice, m->stream));
    CUDA_TRY(cudaMemcpyAsync(d_cand, cand, sizeof(int) * (size_t)ncand, cudaMemcpyHostToDevice, m->stream));
    CUDA_TRY(orbk_dist_csr((const uint8_t*)dq, nq, (const uint8_t*)dt, d_off, d_cand, (int*)dout, m->stream));
    CUDA_TRY(cudaMemcpyAsync(dist, dout, sizeof(int) * (size_t)ncand, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return ORB_OK;
}

extern "C" int orb_stereo_match(orb_matcher* m, const orb_keypoint* kl, const uint8_t* dl, int nl, const orb_keypoint* kr,
                                const uint8_t* dr, int nr, const float* scale, int nlevels, int rows, float bf, float fx,
                                int32_t* best_r, int32_t* best_dist) {
    if (!m || !best_r || !best_dist || !scale) return fail(ORB_ERR_INVALID, "null argument");
    if (nl < 0 || nr < 0 || nlevels < 1 || nlevels > ORB_MAX_LEVELS) return fail(ORB_ERR_INVALID, "bad count");
    (void)rows;
    if (nl == 0) return ORB_OK;
    for (int i = 0; i < nr; ++i)
        if (kr[i].octave < 0 || kr[i].octave >= nlevels) return fail(ORB_ERR_INVALID, "right keypoint %d: octave out of range", i);
    CUDA_TRY(cudaSetDevice(m->device));
    // mb = mbf/fx (src/Frame.cc:215); maxD = mbf/minZ with minZ = mb (:476-478)
    const float mb = bf / fx;
    const float maxD = bf / mb;
    void *dkl, *ddl, *dkr, *ddr, *dsc, *dri, *dout;
    int rc;
    if ((rc = scratch(m, 0, (size_t)nl * 32 + 32, &ddl)) || (rc = scratch(m, 1, (size_t)nr * 32 + 32, &ddr)) ||
        (rc = scratch(m, 4, (size_t)nl * 28 + 32, &dkl)) || (rc = scratch(m, 5, (size_t)nr * 28 + 32, &dkr)) ||
        (rc = scratch(m, 6, sizeof(float) * ORB_MAX_LEVELS, &dsc)) || (rc = scratch(m, 7, sizeof(int4) * (size_t)(nr + 1), &dri)) ||
        (rc = scratch(m, 3, sizeof(int) * 2 * (size_t)nl, &dout)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(dkl, kl, (size_t)nl * 28, cudaMemcpyHostToDevice, m->stream));
    CUDA_TRY(cudaMemcpyAsync(ddl, dl, (size_t)nl * 32, cudaMemcpyHostToDevice, m->stream));
    if (nr) {
        CUDA_TRY(cudaMemcpyAsync(dkr, kr, (size_t)nr * 28, cudaMemcpyHostToDevice, m->stream));
        CUDA_TRY(cudaMemcpyAsync(ddr, dr, (size_t)nr * 32, cudaMemcpyHostToDevice, m->stream));
    }
    CUDA_TRY(cudaMemcpyAsync(dsc, scale, sizeof(float) * nlevels, cudaMemcpyHostToDevice, m->stream));
    int* o = (int*)dout;
    CUDA_TRY(orbk_stereo((const orb_kp28*)dkl, (const uint8_t*)ddl, nl, (const orb_kp28*)dkr, (const uint8_t*)ddr, nr, (const float*)dsc,
                         (int4*)dri, maxD, o, o + nl, m->stream));
    CUDA_TRY(cudaMemcpyAsync(best_r, o, sizeof(int) * nl, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaMemcpyAsync(best_dist, o + nl, sizeof(int) * nl, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return ORB_OK;
}

// Frame::ComputeStereoMatches whole (src/Frame.cc:446-619).  The pyramids stay where the extractors left them.
extern "C" int orb_compute_stereo_matches(orb_matcher* m, orb_extractor* ex_left, int frame_left, orb_extractor* ex_right,
                                          int frame_right, const orb_keypoint* kl, const uint8_t* dl, int nl, const orb_keypoint* kr,
                                          const uint8_t* dr, int nr, float bf, float fx, float* u_right, float* depth) {
    // mb = mbf / fx (src/Frame.cc:94, :215)
    return orb_compute_stereo_matches_mb(m, ex_left, frame_left, ex_right, frame_right, kl, dl, nl, kr, dr, nr, bf, bf / fx, u_right, depth);
}

extern "C" int orb_compute_stereo_matches_mb(orb_matcher* m, orb_extractor* ex_left, int frame_left, orb_extractor* ex_right,
                                             int frame_right, const orb_keypoint* kl, const uint8_t* dl, int nl, const orb_keypoint* kr,
                                             const uint8_t* dr, int nr, float bf, float mb, float* u_right, float* depth) {
    if (!m || !ex_left || !ex_right || !u_right || !depth) return fail(ORB_ERR_INVALID, "null argument");
    if (nl < 0 || nr < 0) return fail(ORB_ERR_INVALID, "bad count");
    if (ex_left->plan.rows == 0 || ex_right->plan.rows == 0) return fail(ORB_ERR_INVALID, "no frame extracted yet");
    const OrbPlan& PL = ex_left->plan;
    const OrbPlan& PR = ex_right->plan;
    if (PL.rows != PR.rows || PL.cols != PR.cols || PL.nlevels != PR.nlevels || ex_left->params.scale_factor != ex_right->params.scale_factor)
        return fail(ORB_ERR_INVALID, "left and right extractors differ in shape or pyramid");
    if (ex_left->device != m->device || ex_right->device != m->device) return fail(ORB_ERR_INVALID, "handles live on different devices");
    if (frame_left < 0 || frame_left >= ex_left->max_batch || frame_right < 0 || frame_right >= ex_right->max_batch)
        return fail(ORB_ERR_INVALID, "bad frame index");
    if (nl == 0) return ORB_OK;
    if (!kl || !dl || (nr && (!kr || !dr))) return fail(ORB_ERR_INVALID, "null argument");
    const int nlevels = PL.nlevels, rows = PL.rows;
    const std::vector<float>& scale = ex_left->tab.scale;
    for (int i = 0; i < nl; ++i)
        if (kl[i].octave < 0 || kl[i].octave >= nlevels) return fail(ORB_ERR_INVALID, "left keypoint %d: octave out of range", i);
    for (int i = 0; i < nr; ++i) {
        if (kr[i].octave < 0 || kr[i].octave >= nlevels) return fail(ORB_ERR_INVALID, "right keypoint %d: octave out of range", i);
        // vRowIndices[yi] for yi in [floor(y - r), ceil(y + r)] (:463-473): out of bounds in the reference when it leaves the image
        const float r = 2.0f * scale[kr[i].octave];
        if ((int)floorf(kr[i].y - r) < 0 || (int)ceilf(kr[i].y + r) >= rows)
            return fail(ORB_ERR_SHAPE, "right keypoint %d: row band leaves the image (the reference indexes vRowIndices out of bounds)", i);
    }
    for (int i = 0; i < nl; ++i)
        if (!(kl[i].y >= 0.0f && kl[i].y < (float)rows)) return fail(ORB_ERR_SHAPE, "left keypoint %d: row outside the image", i);
    CUDA_TRY(cudaSetDevice(m->device));
    const float maxD = bf / mb;  // minZ = mb, maxD = mbf / minZ (:476-478)
    void *dkl, *ddl, *dkr, *ddr, *dsc, *dri, *dout, *dres;
    int rc;
    if ((rc = scratch(m, 0, (size_t)nl * 32 + 32, &ddl)) || (rc = scratch(m, 1, (size_t)nr * 32 + 32, &ddr)) ||
        (rc = scratch(m, 4, (size_t)nl * 28 + 32, &dkl)) || (rc = scratch(m, 5, (size_t)nr * 28 + 32, &dkr)) ||
        (rc = scratch(m, 6, sizeof(float) * ORB_MAX_LEVELS, &dsc)) || (rc = scratch(m, 7, sizeof(int4) * (size_t)(nr + 1), &dri)) ||
        (rc = scratch(m, 3, sizeof(int) * 2 * (size_t)nl, &dout)) || (rc = scratch(m, 2, sizeof(int) * (3 * (size_t)nl + 1), &dres)))
        return rc;
    cudaStream_t st = m->stream;
    // the pyramids were written on the extractors' streams
    cudaEvent_t ev;
    CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e1 = cudaEventRecord(ev, ex_left->stream);
    if (e1 == cudaSuccess) e1 = cudaStreamWaitEvent(st, ev, 0);
    if (e1 == cudaSuccess && ex_right != ex_left) {
        e1 = cudaEventRecord(ev, ex_right->stream);
        if (e1 == cudaSuccess) e1 = cudaStreamWaitEvent(st, ev, 0);
    }
    cudaEventDestroy(ev);
    CUDA_TRY(e1);
    CUDA_TRY(cudaMemcpyAsync(dkl, kl, (size_t)nl * 28, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ddl, dl, (size_t)nl * 32, cudaMemcpyHostToDevice, st));
    if (nr) {
        CUDA_TRY(cudaMemcpyAsync(dkr, kr, (size_t)nr * 28, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ddr, dr, (size_t)nr * 32, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(dsc, scale.data(), sizeof(float) * nlevels, cudaMemcpyHostToDevice, st));
    int* o = (int*)dout;
    CUDA_TRY(orbk_stereo((const orb_kp28*)dkl, (const uint8_t*)ddl, nl, (const orb_kp28*)dkr, (const uint8_t*)ddr, nr, (const float*)dsc,
                         (int4*)dri, maxD, o, o + nl, st));
    OrbStereoLevels lv;
    memset(&lv, 0, sizeof lv);
    for (int l = 0; l < nlevels; ++l) {
        const OrbLevel& A = PL.lv[l];
        const OrbLevel& B = PR.lv[l];
        const uint8_t* la = (A.src == 0 && ex_left->last_l0) ? ex_left->last_l0 : A.img;
        const uint8_t* lb = (B.src == 0 && ex_right->last_l0) ? ex_right->last_l0 : B.img;
        lv.left[l] = la + (size_t)frame_left * A.plane;
        lv.right[l] = lb + (size_t)frame_right * B.plane;
        lv.pitch[l] = A.pitch;
        lv.rows[l] = A.rows;
        lv.cols[l] = A.cols;
        lv.scale[l] = scale[l];
        lv.inv_scale[l] = ex_left->tab.inv_scale[l];
    }
    float* d_ur = (float*)dres;
    float* d_dep = d_ur + nl;
    int* d_sad = (int*)(d_dep + nl);
    int* d_flags = d_sad + nl;
    CUDA_TRY(cudaMemsetAsync(d_flags, 0, sizeof(int), st));
    CUDA_TRY(orbk_stereo_refine((const orb_kp28*)dkl, nl, (const orb_kp28*)dkr, o, o + nl, lv, bf, maxD, d_ur, d_dep, d_sad, d_flags, st));
    std::vector<int> sad((size_t)nl + 1);
    CUDA_TRY(cudaMemcpyAsync(u_right, d_ur, sizeof(float) * nl, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(depth, d_dep, sizeof(float) * nl, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(sad.data(), d_sad, sizeof(int) * ((size_t)nl + 1), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (sad[nl] & 1) return fail(ORB_ERR_SHAPE, "a SAD window leaves its pyramid level (cv::Exception in the reference)");
    // median cut (:606-619): vDistIdx sorted by (dist, iL); median = the element at size/2; everything with
    // dist >= 1.5f * 1.4f * median is reset.  An empty list is left alone (the reference reads past an empty vector).
    std::vector<int> d;
    d.reserve(nl);
    for (int i = 0; i < nl; ++i)
        if (sad[i] >= 0) d.push_back(sad[i]);
    if (!d.empty()) {
        std::nth_element(d.begin(), d.begin() + d.size() / 2, d.end());
        const float median = (float)d[d.size() / 2];
        const float thDist = 1.5f * 1.4f * median;
        for (int i = 0; i < nl; ++i)
            if (sad[i] >= 0 && !((float)sad[i] < thDist)) {
                u_right[i] = -1;
                depth[i] = -1;
            }
    }
    return ORB_OK;
}

// Frame::ComputeStereoMatches for the stereo pairs of one extractor batch.  Pair p = frames (2p, 2p + 1) of ex's last call
// (left = even, right = odd frame).
extern "C" int orb_compute_stereo_matches_batch(orb_matcher* m, orb_extractor* ex, int npairs, const orb_keypoint* kps, const uint8_t* desc,
                                                int cap, const int32_t* counts, float bf, float fx, float* u_right, float* depth,
                                                int32_t* status, int on_device) {
    if (!m || !ex || !kps || !desc || !counts || !u_right || !depth) return fail(ORB_ERR_INVALID, "null argument");
    if (npairs < 0 || cap <= 0) return fail(ORB_ERR_INVALID, "bad count");
    if (npairs == 0) return ORB_OK;
    if (ex->plan.rows == 0) return fail(ORB_ERR_INVALID, "no frame extracted yet");
    if (2 * npairs > ex->max_batch) return fail(ORB_ERR_INVALID, "%d pairs exceed the extractor's max_batch=%d", npairs, ex->max_batch);
    if (ex->device != m->device) return fail(ORB_ERR_INVALID, "handles live on different devices");
    const OrbPlan& P = ex->plan;
    const int nlevels = P.nlevels;
    CUDA_TRY(cudaSetDevice(m->device));
    const float mb = bf / fx;    // src/Frame.cc:94
    const float maxD = bf / mb;  // :476-478
    const size_t rowsN = (size_t)npairs * cap, frames = 2 * (size_t)npairs;
    void *dri, *dwork, *dsc, *dk = nullptr, *dd = nullptr, *dc = nullptr, *dres = nullptr;
    int rc;
    if ((rc = scratch(m, 7, sizeof(int4) * rowsN, &dri)) || (rc = scratch(m, 3, sizeof(int) * (3 * rowsN + npairs), &dwork)) ||
        (rc = scratch(m, 6, sizeof(float) * ORB_MAX_LEVELS, &dsc)))
        return rc;
    cudaStream_t st = m->stream;
    // the pyramids were written on the extractor's stream
    cudaEvent_t ev;
    CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e1 = cudaEventRecord(ev, ex->stream);
    if (e1 == cudaSuccess) e1 = cudaStreamWaitEvent(st, ev, 0);
    cudaEventDestroy(ev);
    CUDA_TRY(e1);
    const orb_kp28* d_kps = (const orb_kp28*)kps;
    const uint8_t* d_desc = desc;
    const int* d_counts = counts;
    float *d_ur = u_right, *d_dep = depth;
    if (!on_device) {
        if ((rc = scratch(m, 4, frames * cap * 28, &dk)) || (rc = scratch(m, 0, frames * cap * 32, &dd)) ||
            (rc = scratch(m, 5, sizeof(int) * frames, &dc)) || (rc = scratch(m, 2, sizeof(float) * 2 * rowsN, &dres)))
            return rc;
        int maxc = 0;
        for (size_t f = 0; f < frames; ++f) maxc = std::max(maxc, std::min(counts[f], cap));
        if (maxc > 0) {
            CUDA_TRY(cudaMemcpy2DAsync(dk, (size_t)cap * 28, kps, (size_t)cap * 28, (size_t)maxc * 28, frames, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpy2DAsync(dd, (size_t)cap * 32, desc, (size_t)cap * 32, (size_t)maxc * 32, frames, cudaMemcpyHostToDevice, st));
        }
        CUDA_TRY(cudaMemcpyAsync(dc, counts, sizeof(int) * frames, cudaMemcpyHostToDevice, st));
        d_kps = (const orb_kp28*)dk;
        d_desc = (const uint8_t*)dd;
        d_counts = (const int*)dc;
        d_ur = (float*)dres;
        d_dep = d_ur + rowsN;
    }
    CUDA_TRY(cudaMemcpyAsync(dsc, ex->tab.scale.data(), sizeof(float) * nlevels, cudaMemcpyHostToDevice, st));
    OrbStereoLevels lv;
    memset(&lv, 0, sizeof lv);
    for (int l = 0; l < nlevels; ++l) {
        const OrbLevel& A = P.lv[l];
        const uint8_t* base = (A.src == 0 && ex->last_l0) ? ex->last_l0 : A.img;
        lv.left[l] = lv.right[l] = base;
        lv.plane[l] = A.plane;
        lv.pitch[l] = A.pitch;
        lv.rows[l] = A.rows;
        lv.cols[l] = A.cols;
        lv.scale[l] = ex->tab.scale[l];
        lv.inv_scale[l] = ex->tab.inv_scale[l];
    }
    int* w = (int*)dwork;
    int *d_best_r = w, *d_best_dist = w + rowsN, *d_sad = w + 2 * rowsN, *d_flags = w + 3 * rowsN;
    CUDA_TRY(orbk_stereo_batch(d_kps, d_desc, d_counts, cap, npairs, nlevels, P.rows, (const float*)dsc, lv, bf, maxD, (int4*)dri, d_best_r,
                               d_best_dist, d_sad, d_flags, d_ur, d_dep, st));
    if (on_device) {
        // status[p] (device, optional): the flag word of pair p, 0 = ok, non-zero = the reference faults on this pair (rows reset to -1)
        if (status) CUDA_TRY(cudaMemcpyAsync(status, d_flags, sizeof(int) * npairs, cudaMemcpyDeviceToDevice, st));
        return ORB_OK;
    }
    std::vector<int> flags(npairs);
    int maxl = 0;
    for (int p = 0; p < npairs; ++p) maxl = std::max(maxl, std::min(counts[2 * p], cap));
    if (maxl > 0) {
        CUDA_TRY(cudaMemcpy2DAsync(u_right, (size_t)cap * 4, d_ur, (size_t)cap * 4, (size_t)maxl * 4, npairs, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpy2DAsync(depth, (size_t)cap * 4, d_dep, (size_t)cap * 4, (size_t)maxl * 4, npairs, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaMemcpyAsync(flags.data(), d_flags, sizeof(int) * npairs, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    int bad = -1;
    for (int p = 0; p < npairs; ++p) {
        if (status) status[p] = flags[p] ? ((flags[p] & 2) ? ORB_ERR_INVALID : ORB_ERR_SHAPE) : ORB_OK;
        if (flags[p] && bad < 0) bad = p;
    }
    if (bad >= 0 && !status)
        return fail((flags[bad] & 2) ? ORB_ERR_INVALID : ORB_ERR_SHAPE, "pair %d: %s", bad,
                    (flags[bad] & 2) ? "keypoint octave out of range" : "a row band or SAD window leaves the image (the reference faults here)");
    return ORB_OK;
}
