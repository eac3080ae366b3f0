// Host/device shared plan of one extractor configuration: per-level geometry of the
// pyramid, the FAST cell grid, the octree roots and the device buffer layout.
// Mirrors the arithmetic of reference src/ORBextractor.cc:116-170 (ctor tables),
// :497-515 (level sizes) and :288-357 (cell grid), bit for bit.
#pragma once
#include <stdint.h>
#include <vector_types.h>

#define ORB_MAX_LEVELS 16
#define ORB_EDGE 19          // EDGE_THRESHOLD (ORBextractor.cc:17)
#define ORB_MINB 16          // minBorderX/Y = EDGE_THRESHOLD-3 (ORBextractor.cc:300-303)
#define ORB_CELL 30.f        // W (ORBextractor.cc:290)
#define ORB_OCT_DEPTH 13     // quadtree depth carried in a 32-bit path code (6 root bits + 2*13)
#define ORB_MAX_DIM 8192     // rows/cols limit so that 13 halvings reach 1-px nodes

// detect tile: up to DET_TILE_W x DET_TILE_H level pixels staged in shared memory
#define DET_TILE_W 256
#define DET_TILE_H 72
#define DET_SP 272           // smem pitch of the image tile (bytes)
#define DET_THREADS 256

// blur tile (k_blur)
#define BLUR_TW 224                    // output columns per tile: 16 + 224 + 3 halo columns fit the 256-byte TMA box
#define BLUR_TH 128
#define BLUR_RPT 32                    // output rows per thread
#define BLUR_SW 64                     // smem words per row = the box: 16 bytes left of the tile, the tile, right halo
#define BLUR_BOX_H (BLUR_TH + 6)
#define BLUR_THREADS ((BLUR_TW / 4) * (BLUR_TH / BLUR_RPT))

// resize tile: RSZ_W x RSZ_H output pixels of level l from a 256 x RSZ_BOX_H box of level l-1 staged by TMA
#define RSZ_W 192
#define RSZ_H 32
#define RSZ_BOX_H 48
#define RSZ_THREADS ((RSZ_W / 4) * (RSZ_H / 4))

// describe tile: keypoints whose level position lies in a DSC_W x DSC_H core; the raw and the blurred level are
// staged with an 18-px halo (the rBRIEF pattern reaches 18 px after rotation, IC_Angle 15) by one TMA box each,
// one after the other into the same shared-memory buffer.
// Keypoints sit at x, y >= 19, so the box of tile (tx, ty) starts at level pixel (DSC_W * tx, 1 + DSC_H * ty):
// 16-byte aligned for the TMA unit without any slack.  The 224-byte box pitch (56 words) puts four consecutive
// tile rows into four disjoint 8-bank groups, which makes the IC_Angle loads conflict-free.
#define DSC_W 160
#define DSC_H 128
#define DSC_HALO 18
#define DSC_BOX_W 224
#define DSC_BOX_H (DSC_H + 2 * DSC_HALO)
#define DSC_THREADS 256
#define DSC_LIST 256         // keypoints of one tile handled per round

struct OrbLevel {
    // image
    int rows, cols;
    int pitch;               // bytes per row of this level's device buffers
    int src;                 // level whose pixels (and FAST candidates) this level shares
    unsigned long long plane;  // bytes per frame of this level's device buffers
    uint8_t* img;            // [batch][rows][pitch] (level 0: set per call)
    uint8_t* blur;           // [batch][rows][pitch] Gaussian-blurred copy
    // resize tables (level = resize(level-1)), device pointers; null when src != self or level 0
    const int* xtab;         // [cols] packed: s | s1<<16
    const int* xcoef;        // [cols] packed: a0 | a1<<16
    const int* ytab;         // [rows] packed: sy0 | sy1<<16
    const int* ycoef;        // [rows] packed: b0 | b1<<16
    const int4* xgrp;        // [ceil(cols/4)] k_resize4 gather descriptors (null: use k_resize)
    const int4* xcoef4;      // [ceil(cols/4)] a0 | a1<<16 of the 4 columns of a group
    int rszTiled;            // every RSZ_W x RSZ_H output tile's source pixels fit one 256 x RSZ_BOX_H TMA box (k_resize_tile)
    // FAST cell grid over [16, cols-16) x [16, rows-16)
    int W, H;                // maxBorder - minBorder
    int nCols, nRows, wCell, hCell;  // 0 cells when the level is smaller than one cell
    int tileCells, tilesX;   // detect tiles: tileCells cells wide, one cell row high
    int boxH;                // rows of the TMA box that stages one detect tile (hCell + 6)
    int tileBase;            // first tile id of this level in the detect launch (src == self)
    int nTiles;
    int dTilesX, dTiles, dTileBase;  // describe tiles of this level's pixels (src == self)
    // octree
    int nIni;
    float hX;
    int nFeat;               // mnFeaturesPerLevel[level]
    int kmax;                // bound on kept keypoints
    unsigned candCap;        // worst-case FAST candidates per frame
    uint2* cand;             // [batch][candCap]: x|y<<16 (relative to minBorder), score
    unsigned long long* sortScratch;  // [batch][sortCap] global fallback for big problems
    unsigned sortCap;        // pow2 >= candCap
    uint2* kept;             // [batch][kmax]: x|y<<16 (level coords), score
    int keptBase;            // offset of this level in describe launch space
    float scale;             // mvScaleFactor[level]
    int patchSize;           // int(31*scale)
};

struct OrbPlan {
    int nlevels;
    int rows, cols;          // level-0 shape this plan was built for
    int iniTh, minTh, lowTh; // clamped to [0,255]; lowTh = min(ini,min)
    int batch;               // frames per launch the buffers are sized for
    int frameBase;           // first frame of this launch inside the batch buffers (TMA z offset)
    int totalTiles;          // detect tiles per frame
    int detRows;             // tallest detect tile (max boxH over the levels), sizes k_detect's shared memory
    int totalKmax;           // sum of kmax
    int totalDescTiles;      // describe tiles per frame
    int* candCount;          // [batch][ORB_MAX_LEVELS]
    int* keptCount;          // [batch][ORB_MAX_LEVELS]
    int* status;             // [batch] octree status flags (non-zero: unseparable keys)
    int* needGeneric;        // [batch][ORB_MAX_LEVELS] problems the table-based octree handed over
    // tile id -> (level, tile row, tile column) of the detect / blur / describe launches, packed level | row << 4 | col << 18:
    // one load instead of every thread of every CTA walking the level table
    const unsigned* detTileTab;   // [totalTiles]
    const unsigned* blurTileTab;  // [blurTiles]
    const unsigned* descTileTab;  // [totalDescTiles]
    int blurTiles;           // blur tiles per frame
    const int2* icTab;       // [8][32] IC_Angle dp4a weights: step it, lane (row it*4 + lane/8, word lane%8)
    const float4* pairTab;   // [182] rBRIEF test pairs (x0, y0, x1, y1)
    OrbLevel lv[ORB_MAX_LEVELS];
};
