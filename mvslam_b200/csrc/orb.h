// orb.h — feature extraction stage (cv::ORB as VisualFeature::extract uses it, reference
// source/vision/visual-feature.cpp:9-17,40-49).  Internal interface between api.cu and orb.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/mvslam_b200.h"

namespace mvs {

constexpr int kOrbLevels = 8;          // cv::ORB default nlevels
constexpr int kOrbEdge = 31;           // edgeThreshold == patchSize
constexpr int kOrbHalfPatch = 15;
constexpr int kOrbFastThreshold = 20;
constexpr int kOrbSortCap = 4096;      // keypoints one pyramid level of one image may keep (quota + ties at the cut-off)

struct OrbLevel {
    int w, h, pitch;       // level image, rows padded to 16 bytes
    int off;               // byte offset of the level inside one image's pyramid slab
    int quota;             // nfeaturesPerLevel
    int cand_off, cand_cap;  // slice of the per-image candidate list (every 3x3 maximum fits)
    int tile_off;          // first FAST tile / first blur tile of this level in the flattened tile index
    int blur_tile_off;
    int tab_off;           // resize tables of this level (int32, packed offset | weight << 16): x[w padded to 4], y[h]
    float scale;           // layerScale
};

struct OrbGeom {
    OrbLevel lv[kOrbLevels];
    int slab;              // bytes per image in the pyramid (and blurred pyramid) buffer
    int cand_total;        // candidate slots per image
    int fast_tiles, blur_tiles;
    float harris_scale4;   // (1 / (4 * 7 * 255))^4 in float, orb.cpp HarrisResponses
    int umax[kOrbHalfPatch + 1];
};

struct OrbBuffers {
    uint8_t *pyr;          // [images][slab]
    uint8_t *blur;         // [images][slab]
    const int32_t *tabs;   // resize tables, shared by all images
    uint32_t *cand_xy;     // [images][cand_total]  x | y << 16
    float *cand_val;       // [images][cand_total]  FAST score, then Harris response (-inf = dropped)
    int32_t *cand_cnt;     // [images][8]
    int32_t *hist;         // [images][8][256] FAST score histogram of the 3x3 maxima
    uint32_t *kept_idx;    // [images][8][kOrbSortCap] indices into the level's candidate slice, raster order
    int32_t *kept_cnt;     // [images][8]; -1 = more than kOrbSortCap keypoints tie at the cut-off
};

struct OrbDescribeArgs {
    const int32_t *img_off;   // [images] first output slot of each image
    mvs_keypoint *kp;         // compact outputs (may be nullptr)
    uint8_t *desc;            // [total][32]
    float2 *frame_kp;         // optional second copy of pt for the resident frame table
};

// Host side: geometry of the pyramid for a w x h image and nfeatures; fills tabs_host (resize tables).
bool orb_make_geometry(int w, int h, int nfeatures, OrbGeom &g, std::vector<int32_t> &tabs_host);

void launch_orb_pyramid(const OrbGeom &g, const OrbBuffers &b, const uint8_t *stage, int stride, int n_images, cudaStream_t s);
void launch_orb_fast(const OrbGeom &g, const OrbBuffers &b, int n_images, cudaStream_t s);
void launch_orb_harris(const OrbGeom &g, const OrbBuffers &b, int n_images, cudaStream_t s);
void launch_orb_select(const OrbGeom &g, const OrbBuffers &b, int n_images, cudaStream_t s);
void launch_orb_blur(const OrbGeom &g, const OrbBuffers &b, int n_images, cudaStream_t s);
void launch_orb_describe(const OrbGeom &g, const OrbBuffers &b, const OrbDescribeArgs &d, int n_images, int max_keypoints,
                         cudaStream_t s);

}  // namespace mvs
