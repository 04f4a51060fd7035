// roi_align helpers shared by the crop kernels (hpb_crop.cu: per-lane gathers; hpb_crop_tma.cu: TMA-fed shared-memory ring).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "hpb_common.cuh"

namespace hpbc {

constexpr int CROP_SPAN = 4;       // dense tap span handled by the fast path
constexpr int CROP_BAND_MAX = 120;  // most output rows one CTA handles (fewer when the batch is small)
constexpr int CROP_MAX_THREADS = 320;  // output columns per CTA (wider outputs use several column tiles, blockIdx.z)

struct CropPixParams {
    const float *images;
    const int32_t *im_ids;
    const float *boxes;  // [b,4]
    int n_im, C, H, W, b, h, w;
    float *crops;
    long long crops_bs;
    int band;  // output rows per CTA
    const float4 *packed;  // [n_im,H,W] pixel-interleaved copy of `images` (r,g,b,depth|0) or nullptr
    const uint2 *packed_h;  // same in fp16 (r,g,b,0): 8-byte taps (hpb_set_crop_tap_precision(ctx, 16), RGB frames)
    uint2 *crops_h;         // OUTFMT 1: [b][h][w] (r,g,b,0) bfloat16 pixels, batch stride crops_bs pixels (hpb_crop_bf16x4)
};

__device__ __forceinline__ uint2 pack_bf16x4(float r, float g, float b) {
    const __nv_bfloat162 rg = __floats2bfloat162_rn(r, g);
    uint2 v;
    v.x = *reinterpret_cast<const unsigned *>(&rg);
    v.y = (unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(b));
    return v;
}

// One roi_align sample coordinate along one axis (torchvision roi_align bilinear_interpolate, aligned=False).
struct AxisTap {
    int lo, hi;
    float wlo, whi;
    bool valid;
};

__device__ __forceinline__ AxisTap axis_tap(float c, int n) {
    AxisTap t;
    t.valid = !(c < -1.0f || c > (float)n);
    if (c <= 0.0f) c = 0.0f;
    int lo = (int)c;
    int hi;
    if (lo >= n - 1) {
        lo = hi = n - 1;
        c = (float)lo;
    } else {
        hi = lo + 1;
    }
    const float l = c - (float)lo;
    t.lo = lo; t.hi = hi; t.whi = l; t.wlo = 1.0f - l;
    return t;
}

// Dense tap weights of the 4 samples of output bin `i` along one axis.  Returns false when the span exceeds
// CROP_SPAN (heavy down-sampling): the caller then takes the generic path.
__device__ __forceinline__ bool axis_weights(float start, float bin, int i, int n, int &base, float (&wt)[CROP_SPAN]) {
    AxisTap t0 = axis_tap(start + (float)i * bin + (0.0f + 0.5f) * bin / 4.0f, n);
    AxisTap t1 = axis_tap(start + (float)i * bin + (1.0f + 0.5f) * bin / 4.0f, n);
    AxisTap t2 = axis_tap(start + (float)i * bin + (2.0f + 0.5f) * bin / 4.0f, n);
    AxisTap t3 = axis_tap(start + (float)i * bin + (3.0f + 0.5f) * bin / 4.0f, n);
    int lo_min = 0x7fffffff, hi_max = -1;
    if (t0.valid) { lo_min = min(lo_min, t0.lo); hi_max = max(hi_max, t0.hi); }
    if (t1.valid) { lo_min = min(lo_min, t1.lo); hi_max = max(hi_max, t1.hi); }
    if (t2.valid) { lo_min = min(lo_min, t2.lo); hi_max = max(hi_max, t2.hi); }
    if (t3.valid) { lo_min = min(lo_min, t3.lo); hi_max = max(hi_max, t3.hi); }
    float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f, w3 = 0.0f;
    wt[0] = wt[1] = wt[2] = wt[3] = 0.0f;
    if (hi_max < 0) {  // no valid sample: all-zero weights
        base = 0;
        return true;
    }
    if (hi_max - lo_min >= CROP_SPAN) return false;
    base = lo_min;
    // scatter-add of the 8 (tap, weight) pairs written as selects on scalars: keeps everything in registers (an indexed
    // wt[lo - base] += ... is turned into local-memory loads/stores by the compiler)
    auto add = [&](int d, float v) {
        w0 += d == 0 ? v : 0.0f;
        w1 += d == 1 ? v : 0.0f;
        w2 += d == 2 ? v : 0.0f;
        w3 += d == 3 ? v : 0.0f;
    };
    if (t0.valid) { add(t0.lo - lo_min, 0.25f * t0.wlo); add(t0.hi - lo_min, 0.25f * t0.whi); }
    if (t1.valid) { add(t1.lo - lo_min, 0.25f * t1.wlo); add(t1.hi - lo_min, 0.25f * t1.whi); }
    if (t2.valid) { add(t2.lo - lo_min, 0.25f * t2.wlo); add(t2.hi - lo_min, 0.25f * t2.whi); }
    if (t3.valid) { add(t3.lo - lo_min, 0.25f * t3.wlo); add(t3.hi - lo_min, 0.25f * t3.whi); }
    wt[0] = w0; wt[1] = w1; wt[2] = w2; wt[3] = w3;
    return true;
}

}  // namespace hpbc
