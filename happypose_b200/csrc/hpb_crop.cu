// Perspective crop + resize of the observed frame (PosePredictor.crop_inputs) for sm_100a.
//
// Replaces happypose/pose_estimators/megapose/models/pose_rigid.py:199-277:
//   project_points_robust + boxes_from_uv      toolbox/lib3d/camera_geometry.py:40-67
//   deepim_boxes / deepim_crops_robust          toolbox/lib3d/cropping.py:27-75,113-152
//   crop_images -> torchvision.ops.roi_align    toolbox/lib3d/cropping.py:155-197 (sampling_ratio=4, aligned=False)
//   get_K_crop_resize                           toolbox/lib3d/camera_geometry.py:70-122
//
// Two kernels:
//   hpb_crop_boxes_kernel   one CTA per hypothesis: projects the object's point set, reduces min/max, builds
//                           boxes_rend, boxes_crop and K_crop on the device.
//   hpb_crop_pixels_kernel  one CTA per (hypothesis, band of output rows), one THREAD per output column.  roi_align's
//                           4x4 bilinear samples per output pixel are separable: the 4 row samples and the 4 column
//                           samples of a bin fold into dense 4-tap row / column weights; a thread keeps a rolling
//                           window of horizontally filtered source rows in registers while it marches down its
//                           column, so a pixel costs ~bin_h * 4 loads per channel instead of 64.  The frame is
//                           indexed by im_id: no per-hypothesis copy of the frame (the reference materialises
//                           images[batch_im_ids], pose_estimator.py:390).  Stores are coalesced along x.  RGB-D frames also resample the depth-validity map and
//                           zero depth where validity < 0.99 (cropping.py:181-195).
#include "hpb_common.cuh"

namespace {

constexpr int CROP_SPAN = 4;       // dense tap span handled by the fast path
constexpr int CROP_BAND_MAX = 48;
constexpr int CROP_CHUNK = 8;    // output rows per chunk of the column march
constexpr int CROP_SROWS = 20;   // filtered source rows staged per chunk (8 rows * bin_h <= 1.9 + 4 taps)  // most output rows one CTA handles (fewer when the batch is small)
constexpr int CROP_MAX_THREADS = 384;  // output widths beyond this take the generic path

struct CropBoxParams {
    const float *points;
    const int32_t *obj_ids;
    const float *K, *TCO, *tCR;
    int n_pts, b;
    int H, W, h, w;
    float lamb;
    float *K_crop, *boxes_rend, *boxes_crop;
};

__global__ void __launch_bounds__(256) hpb_crop_boxes_kernel(const CropBoxParams p) {
    const int n = blockIdx.x;
    const int tid = threadIdx.x;
    __shared__ float sP[12];
    __shared__ float red[4][8];
    const float *K = p.K + (size_t)n * 9;
    const float *T = p.TCO + (size_t)n * 16;
    if (tid < 12) {
        const int i = tid / 4, j = tid % 4;  // P = K @ TCO[:3]
        sP[tid] = fmaf(K[i * 3 + 2], T[8 + j], fmaf(K[i * 3 + 1], T[4 + j], K[i * 3] * T[j]));
    }
    __syncthreads();
    const float *pts = p.points + (size_t)p.obj_ids[n] * p.n_pts * 3;
    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int i = tid; i < p.n_pts; i += blockDim.x) {
        const float x = __ldg(pts + 3 * i), y = __ldg(pts + 3 * i + 1), z = __ldg(pts + 3 * i + 2);
        const float su = fmaf(sP[2], z, fmaf(sP[1], y, fmaf(sP[0], x, sP[3])));
        const float sv = fmaf(sP[6], z, fmaf(sP[5], y, fmaf(sP[4], x, sP[7])));
        float sz = fmaf(sP[10], z, fmaf(sP[9], y, fmaf(sP[8], x, sP[11])));
        sz = fmaxf(0.1f, sz);  // project_points_robust z_min (camera_geometry.py:53-54)
        const float u = su / sz, v = sv / sz;
        mnx = fminf(mnx, u); mxx = fmaxf(mxx, u);
        mny = fminf(mny, v); mxy = fmaxf(mxy, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = mnx; red[1][tid >> 5] = mny; red[2][tid >> 5] = mxx; red[3][tid >> 5] = mxy;
    }
    __syncthreads();
    if (tid == 0) {
        const int nw = blockDim.x >> 5;
        for (int k = 1; k < nw; ++k) {
            mnx = fminf(mnx, red[0][k]); mny = fminf(mny, red[1][k]);
            mxx = fmaxf(mxx, red[2][k]); mxy = fmaxf(mxy, red[3][k]);
        }
        // reference point projection: TCR = TCO with translation tCR, point (0,0,0) (cropping.py:131-137)
        const float *c = p.tCR + (size_t)n * 3;
        const float su = fmaf(K[2], c[2], fmaf(K[1], c[1], K[0] * c[0]));
        const float sv = fmaf(K[5], c[2], fmaf(K[4], c[1], K[3] * c[0]));
        float sz = fmaf(K[8], c[2], fmaf(K[7], c[1], K[6] * c[0]));
        sz = fmaxf(0.1f, sz);
        const float xc = su / sz, yc = sv / sz;
        // deepim_boxes (cropping.py:27-75); obs box == rend box on this path
        const float r = (float)max(p.H, p.W) / (float)min(p.H, p.W);
        const float xdist = fmaxf(fabsf(mnx - xc), fabsf(mxx - xc));
        const float ydist = fmaxf(fabsf(mny - yc), fabsf(mxy - yc));
        const float width = fmaxf(xdist, ydist * r) * 2.0f * p.lamb;
        const float height = fmaxf(xdist / r, ydist) * 2.0f * p.lamb;
        const float x1 = xc - width / 2.0f, y1 = yc - height / 2.0f, x2 = xc + width / 2.0f, y2 = yc + height / 2.0f;
        float *br = p.boxes_rend + (size_t)n * 4;
        br[0] = mnx; br[1] = mny; br[2] = mxx; br[3] = mxy;
        float *bc = p.boxes_crop + (size_t)n * 4;
        bc[0] = x1; bc[1] = y1; bc[2] = x2; bc[3] = y2;
        // get_K_crop_resize (camera_geometry.py:70-122)
        const float final_w = (float)max(p.h, p.w), final_h = (float)min(p.h, p.w);
        const float cw = x2 - x1, ch = y2 - y1;
        const float ccj = (x1 + x2) / 2.0f, cci = (y1 + y2) / 2.0f;
        const float cx = K[2] + (cw - 1.0f) / 2.0f - ccj;
        const float cy = K[5] + (ch - 1.0f) / 2.0f - cci;
        const float dcx = cx - (cw - 1.0f) / 2.0f, dcy = cy - (ch - 1.0f) / 2.0f;
        const float sx = final_w / cw, sy = final_h / ch;
        float *ko = p.K_crop + (size_t)n * 9;
        for (int k = 0; k < 9; ++k) ko[k] = K[k];
        ko[0] = sx * K[0];
        ko[4] = sy * K[4];
        ko[2] = (final_w - 1.0f) / 2.0f + sx * dcx;
        ko[5] = (final_h - 1.0f) / 2.0f + sy * dcy;
    }
}

struct CropPixParams {
    const float *images;
    const int32_t *im_ids;
    const float *boxes;  // [b,4]
    int n_im, C, H, W, b, h, w;
    float *crops;
    long long crops_bs;
    int band;  // output rows per CTA
    const float4 *packed;  // [n_im,H,W] pixel-interleaved copy of `images` (r,g,b,depth|0) or nullptr
};

// Planar [n_im,C,H,W] -> pixel-interleaved float4 [n_im,H,W]: the crop then fetches all channels of a tap with ONE
// 16-byte load.  Done per hpb_crop call when few distinct frames serve many hypotheses (the usual case: 1 frame).
template <int C>
__global__ void hpb_pack_frames_kernel(const float *images, long long n_px_per_im, long long total, float4 *out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const long long im = i / n_px_per_im, px = i - im * n_px_per_im;
        const float *src = images + im * C * n_px_per_im + px;
        float4 v;
        v.x = __ldg(src); v.y = __ldg(src + n_px_per_im); v.z = __ldg(src + 2 * n_px_per_im);
        v.w = C == 4 ? __ldg(src + 3 * n_px_per_im) : 0.0f;
        out[i] = v;
    }
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
#pragma unroll
    for (int k = 0; k < CROP_SPAN; ++k) wt[k] = 0.0f;
    AxisTap taps[4];
    int lo_min = 0x7fffffff, hi_max = -1;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const float c = start + (float)i * bin + ((float)s + 0.5f) * bin / 4.0f;
        taps[s] = axis_tap(c, n);
        if (taps[s].valid) {
            lo_min = min(lo_min, taps[s].lo);
            hi_max = max(hi_max, taps[s].hi);
        }
    }
    if (hi_max < 0) {  // no valid sample: all-zero weights
        base = 0;
        return true;
    }
    if (hi_max - lo_min >= CROP_SPAN) return false;
    base = lo_min;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (!taps[s].valid) continue;
#pragma unroll
        for (int k = 0; k < CROP_SPAN; ++k) {
            if (taps[s].lo - lo_min == k) wt[k] += 0.25f * taps[s].wlo;
            if (taps[s].hi - lo_min == k) wt[k] += 0.25f * taps[s].whi;
        }
    }
    return true;
}

// One thread per output COLUMN, marching down the band's rows.  roi_align's 4x4 samples per output pixel are separable:
// the thread keeps its 4 column-tap weights in registers for the whole band, and a rolling window of 4 horizontally
// filtered source rows (per channel); every output pixel is then a 4-tap vertical combination of the window.  Going
// down one output row advances the window by floor/ceil(bin_h) source rows, so (when up-sampling, the usual case) a
// pixel costs < 1 new filtered row = 4 loads per channel, instead of 64 taps per channel.
template <int C, bool PACKED>
__global__ void __launch_bounds__(CROP_MAX_THREADS, 3) hpb_crop_pixels_kernel(const CropPixParams p) {
    constexpr int NCH = C == 4 ? 5 : C;  // RGB-D: the depth-validity map is resampled as a 5th channel
    const int n = blockIdx.y;
    const int band = p.band;
    const int row0 = blockIdx.x * band;
    const int tid = threadIdx.x;
    extern __shared__ float sm[];
    float *sWY = sm;                                             // [band][CROP_SPAN]
    int *sBY = reinterpret_cast<int *>(sWY + band * CROP_SPAN);  // [band]
    __shared__ int sGeneric;

    const float *bx = p.boxes + (size_t)n * 4;
    const float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
    const float roi_w = fmaxf(x2 - x1, 1.0f), roi_h = fmaxf(y2 - y1, 1.0f);
    const float bin_w = roi_w / (float)p.w, bin_h = roi_h / (float)p.h;
    if (tid == 0) sGeneric = 0;
    __syncthreads();
    const int rows = min(band, p.h - row0);
    for (int i = tid; i < rows; i += blockDim.x) {
        float wt[CROP_SPAN];
        int base = 0;
        if (!axis_weights(y1, bin_h, row0 + i, p.H, base, wt)) sGeneric = 1;
        sBY[i] = base;
        *reinterpret_cast<float4 *>(sWY + i * CROP_SPAN) = make_float4(wt[0], wt[1], wt[2], wt[3]);
    }
    // column taps of this thread's columns (registers); a span over CROP_SPAN anywhere sends the CTA to the generic path
    float wx[CROP_SPAN] = {0.f, 0.f, 0.f, 0.f};
    int bxx = 0;
    const int j0 = tid;
    if (j0 < p.w) {
        if (!axis_weights(x1, bin_w, j0, p.W, bxx, wx)) sGeneric = 1;
    }
    if (tid + (int)blockDim.x < p.w || p.W < CROP_SPAN) sGeneric = 1;  // more columns than threads / tiny frames: generic path
    __syncthreads();
    const bool generic = sGeneric != 0;
    const int im = p.im_ids[n];
    const float *img = p.images + (size_t)im * C * p.H * p.W;
    float *out = p.crops + (size_t)n * p.crops_bs;
    const size_t plane_in = (size_t)p.H * p.W, plane_out = (size_t)p.h * p.w;

    if (!generic) {
        if (j0 >= p.w) return;
        // Re-base the 4 column taps so they are 4 CONSECUTIVE pixels inside the frame: taps clamped to the last column
        // (roi_align clamps samples in (W-1, W] to W-1) have their weights folded onto that column.  One base pointer
        // per thread then serves all taps with immediate offsets.
        {
            const int over = max(0, bxx + CROP_SPAN - 1 - (p.W - 1));
            if (over > 0) {
                float w2[CROP_SPAN] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < CROP_SPAN; ++k) {
                    const int jn = min(bxx + k, p.W - 1) - (bxx - over);
#pragma unroll
                    for (int q = 0; q < CROP_SPAN; ++q)
                        if (q == jn) w2[q] += wx[k];
                }
#pragma unroll
                for (int k = 0; k < CROP_SPAN; ++k) wx[k] = w2[k];
                bxx -= over;
            }
        }
        const int xo0 = bxx, xo1 = bxx + 1, xo2 = bxx + 2, xo3 = bxx + 3;
        const float4 *col4 = PACKED ? p.packed + (size_t)im * p.H * p.W + xo0 : nullptr;  // taps at col4[0..3]
        float hwin[CROP_SPAN][NCH];  // horizontally filtered source rows win_base .. win_base+3
#pragma unroll
        for (int r = 0; r < CROP_SPAN; ++r)
#pragma unroll
            for (int c = 0; c < NCH; ++c) hwin[r][c] = 0.f;
        int win_base = -0x40000000;
        auto filter_row = [&](int yy, float (&dst)[NCH]) {
            if (PACKED) {
                const float4 *src = col4 + (size_t)min(yy, p.H - 1) * p.W;
                const float4 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2), v3 = __ldg(src + 3);
                dst[0] = fmaf(wx[3], v3.x, fmaf(wx[2], v2.x, fmaf(wx[1], v1.x, wx[0] * v0.x)));
                dst[1] = fmaf(wx[3], v3.y, fmaf(wx[2], v2.y, fmaf(wx[1], v1.y, wx[0] * v0.y)));
                dst[2] = fmaf(wx[3], v3.z, fmaf(wx[2], v2.z, fmaf(wx[1], v1.z, wx[0] * v0.z)));
                if (C == 4) {
                    dst[3] = fmaf(wx[3], v3.w, fmaf(wx[2], v2.w, fmaf(wx[1], v1.w, wx[0] * v0.w)));
                    dst[NCH - 1] = fmaf(wx[3], v3.w > 0.f ? 1.f : 0.f, fmaf(wx[2], v2.w > 0.f ? 1.f : 0.f,
                                        fmaf(wx[1], v1.w > 0.f ? 1.f : 0.f, wx[0] * (v0.w > 0.f ? 1.f : 0.f))));
                }
                return;
            }
            const float *src = img + (size_t)min(yy, p.H - 1) * p.W;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float *pl = src + c * plane_in;
                const float v0 = __ldg(pl + xo0), v1 = __ldg(pl + xo1), v2 = __ldg(pl + xo2), v3 = __ldg(pl + xo3);
                dst[c] = fmaf(wx[3], v3, fmaf(wx[2], v2, fmaf(wx[1], v1, wx[0] * v0)));
                if (C == 4 && c == 3)
                    dst[NCH - 1] = fmaf(wx[3], v3 > 0.f ? 1.f : 0.f, fmaf(wx[2], v2 > 0.f ? 1.f : 0.f,
                                        fmaf(wx[1], v1 > 0.f ? 1.f : 0.f, wx[0] * (v0 > 0.f ? 1.f : 0.f))));
            }
        };
        float *o = out + (size_t)row0 * p.w + j0;
        for (int i = 0; i < rows; ++i, o += p.w) {
            const float4 wy = *reinterpret_cast<const float4 *>(sWY + i * CROP_SPAN);
            float acc[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) acc[c] = 0.f;
            if (wy.x != 0.f || wy.y != 0.f || wy.z != 0.f || wy.w != 0.f) {  // uniform over the CTA
                const int by = sBY[i];
                int shift = by - win_base;
                if (shift < 0 || shift >= CROP_SPAN) {  // (re)fill the whole window
#pragma unroll
                    for (int r = 0; r < CROP_SPAN; ++r) filter_row(by + r, hwin[r]);
                } else {
                    for (; shift > 0; --shift) {
#pragma unroll
                        for (int r = 0; r + 1 < CROP_SPAN; ++r)
#pragma unroll
                            for (int c = 0; c < NCH; ++c) hwin[r][c] = hwin[r + 1][c];
                        filter_row(by - shift + CROP_SPAN, hwin[CROP_SPAN - 1]);
                    }
                }
                win_base = by;
#pragma unroll
                for (int c = 0; c < NCH; ++c)
                    acc[c] = fmaf(wy.w, hwin[3][c], fmaf(wy.z, hwin[2][c], fmaf(wy.y, hwin[1][c], wy.x * hwin[0][c])));
            }
            if (C == 4 && acc[NCH - 1] < 0.99f) acc[3] = 0.0f;  // cropping.py:191-195
#pragma unroll
            for (int c = 0; c < C; ++c) __stcs(o + c * plane_out, acc[c]);
        }
        return;
    }

    // generic roi_align: 4x4 samples, 4 taps each (heavy down-sampling: crop box wider than ~1.7x the output)
    const int npx = rows * p.w;
    for (int q = tid; q < npx; q += blockDim.x) {
        const int i = q / p.w, j = q - i * p.w;
        const int oy = row0 + i;
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.f;
        float accv = 0.f;
        for (int sy = 0; sy < 4; ++sy) {
            const AxisTap ty = axis_tap(y1 + (float)oy * bin_h + ((float)sy + 0.5f) * bin_h / 4.0f, p.H);
            if (!ty.valid) continue;
            for (int sx = 0; sx < 4; ++sx) {
                const AxisTap tx = axis_tap(x1 + (float)j * bin_w + ((float)sx + 0.5f) * bin_w / 4.0f, p.W);
                if (!tx.valid) continue;
                const float w1 = ty.wlo * tx.wlo, w2 = ty.wlo * tx.whi, w3 = ty.whi * tx.wlo, w4 = ty.whi * tx.whi;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float *pl = img + c * plane_in;
                    const float v1 = __ldg(pl + (size_t)ty.lo * p.W + tx.lo), v2 = __ldg(pl + (size_t)ty.lo * p.W + tx.hi);
                    const float v3 = __ldg(pl + (size_t)ty.hi * p.W + tx.lo), v4 = __ldg(pl + (size_t)ty.hi * p.W + tx.hi);
                    acc[c] += 0.0625f * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
                    if (c == 3)
                        accv += 0.0625f * (w1 * (v1 > 0.f) + w2 * (v2 > 0.f) + w3 * (v3 > 0.f) + w4 * (v4 > 0.f));
                }
            }
        }
        if (C == 4 && accv < 0.99f) acc[C - 1] = 0.0f;  // cropping.py:191-195
        float *o = out + (size_t)oy * p.w + j;
#pragma unroll
        for (int c = 0; c < C; ++c) __stcs(o + c * plane_out, acc[c]);
    }
}

}  // namespace

int hpb_launch_crop_boxes(hpb_ctx *ctx, int H, int W, const float *points, int n_pts, const int32_t *obj_ids,
                          const float *K, const float *TCO, const float *tCR, int b, int h, int w, float lamb,
                          float *K_crop, float *boxes_rend, float *boxes_crop, cudaStream_t stream) {
    if (b == 0) return HPB_OK;
    CropBoxParams p;
    p.points = points; p.obj_ids = obj_ids; p.K = K; p.TCO = TCO; p.tCR = tCR;
    p.n_pts = n_pts; p.b = b; p.H = H; p.W = W; p.h = h; p.w = w; p.lamb = lamb;
    p.K_crop = K_crop; p.boxes_rend = boxes_rend; p.boxes_crop = boxes_crop;
    hpb_crop_boxes_kernel<<<b, 256, 0, stream>>>(p);
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}

int hpb_launch_crop_pixels(hpb_ctx *ctx, const float *images, int n_im, int C, int H, int W, const int32_t *im_ids,
                           const float *boxes, int b, int h, int w, float *crops, int64_t crops_bs,
                           cudaStream_t stream) {
    if (b == 0) return HPB_OK;
    if (C != 3 && C != 4) {
        hpb_set_error("hpb_crop: C must be 3 or 4 (got %d)", C);
        return HPB_EINVAL;
    }
    CropPixParams p;
    p.images = images; p.im_ids = im_ids; p.boxes = boxes;
    p.n_im = n_im; p.C = C; p.H = H; p.W = W; p.b = b; p.h = h; p.w = w;
    p.crops = crops; p.crops_bs = crops_bs;
    p.packed = nullptr;
    // few distinct frames, many hypotheses: interleave the frames once so a tap is one 16-byte load
    const long long n_px = (long long)H * W, total = n_px * n_im;
    if ((long long)n_im * 8 <= b && total * 16 <= (1ll << 28)) {
        if (ctx->frame_pack_bytes < (size_t)total * 16) {
            if (ctx->frame_pack) HPB_CUDA_OK(cudaFree(ctx->frame_pack));
            ctx->frame_pack = nullptr;
            ctx->frame_pack_bytes = 0;
            HPB_CUDA_OK(cudaMalloc(&ctx->frame_pack, (size_t)total * 16));
            ctx->frame_pack_bytes = (size_t)total * 16;
        }
        const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
        if (C == 3) hpb_pack_frames_kernel<3><<<blocks, 256, 0, stream>>>(images, n_px, total, (float4 *)ctx->frame_pack);
        else hpb_pack_frames_kernel<4><<<blocks, 256, 0, stream>>>(images, n_px, total, (float4 *)ctx->frame_pack);
        HPB_CUDA_OK(cudaGetLastError());
        ctx->launches++;
        p.packed = (const float4 *)ctx->frame_pack;
    }
    // rows per CTA: enough CTAs to fill the machine for small batches, long bands (window fill amortised) otherwise
    int band = CROP_BAND_MAX;
    while (band > 8 && (long long)b * ((h + band - 1) / band) < 2ll * ctx->sm_count) band /= 2;
    p.band = band;
    const int threads = w >= CROP_MAX_THREADS ? CROP_MAX_THREADS : ((w + 31) / 32) * 32;  // one thread per output column
    const size_t smem = (size_t)band * CROP_SPAN * sizeof(float) + (size_t)band * sizeof(int);
    dim3 grid((h + band - 1) / band, b);
    if (C == 3) {
        if (p.packed) hpb_crop_pixels_kernel<3, true><<<grid, threads, smem, stream>>>(p);
        else hpb_crop_pixels_kernel<3, false><<<grid, threads, smem, stream>>>(p);
    } else {
        if (p.packed) hpb_crop_pixels_kernel<4, true><<<grid, threads, smem, stream>>>(p);
        else hpb_crop_pixels_kernel<4, false><<<grid, threads, smem, stream>>>(p);
    }
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}
