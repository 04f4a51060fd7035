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
//   hpb_crop_pixels_kernel  one CTA per (hypothesis, band of output rows[, column tile]), one THREAD per output column.
//                           roi_align's 4x4 bilinear samples per output pixel are separable: the 4 row samples and the 4
//                           column samples of a bin fold into dense 4-tap row / column weights; a thread keeps a window of
//                           4 horizontally filtered source rows in registers (source row s in slot s & 3, the per-row
//                           vertical weights stored pre-permuted, next row's taps prefetched) while it marches down its
//                           column, so a pixel costs ~bin_h * 4 loads per channel instead of 64.  The frame is indexed by
//                           im_id: no per-hypothesis copy of the frame (the reference materialises images[batch_im_ids],
//                           pose_estimator.py:390).  Stores are coalesced along x.  RGB-D frames also resample the
//                           depth-validity map and zero depth where validity < 0.99 (cropping.py:181-195).
#include "hpb_crop_math.cuh"
#include "hpb_pose_math.cuh"

namespace {

using namespace hpbc;

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
    __shared__ float sP[12];
    __shared__ float red[4][8];
    hpbm::crop_boxes_cta(p.K + (size_t)n * 9, p.TCO + (size_t)n * 16, p.tCR + (size_t)n * 3,
                         p.points + (size_t)p.obj_ids[n] * p.n_pts * 3, p.n_pts, p.H, p.W, p.h, p.w, p.lamb, p.K_crop + (size_t)n * 9,
                         p.boxes_rend + (size_t)n * 4, p.boxes_crop + (size_t)n * 4, sP, red);
}

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

// fp16 variant (RGB frames): a tap is one 8-byte load, i.e. half the L1 sectors per tap request -- the crop kernel is
// bound by L1 sector look-ups.  Values in [0,1] keep an absolute error <= 2.5e-4 (BASELINE bar for
// crops: 1e-3); opt-in, used by the bf16 network path only.
__global__ void hpb_pack_frames_half_kernel(const float *images, long long n_px_per_im, long long total, uint2 *out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const long long im = i / n_px_per_im, px = i - im * n_px_per_im;
        const float *src = images + im * 3 * n_px_per_im + px;
        const __half2 rg = __floats2half2_rn(__ldg(src), __ldg(src + n_px_per_im));
        const __half2 b0 = __floats2half2_rn(__ldg(src + 2 * n_px_per_im), 0.0f);
        uint2 v;
        v.x = *reinterpret_cast<const unsigned *>(&rg);
        v.y = *reinterpret_cast<const unsigned *>(&b0);
        out[i] = v;
    }
}

// One thread per output COLUMN, marching down the band's rows.  roi_align's 4x4 samples per output pixel are separable:
// the thread keeps its 4 column-tap weights in registers for the whole band and a window of 4 horizontally filtered
// source rows (per channel); every output pixel is then a 4-tap vertical combination of the window.  Going down one
// output row advances the window by floor/ceil(bin_h) source rows, so (when up-sampling, the usual case) a pixel costs
// < 1 new filtered row = 4 loads per channel, instead of 64 taps per channel.
//
// The window never moves in registers: source row s always lives in slot (s & 3), and the prologue stores every output
// row's 4 vertical weights already permuted to slot order, so the vertical combination is the same straight-line code
// for every row.  The raw taps of the NEXT source row are prefetched one step ahead.  ncu (profiles/): the kernel is
// bound by L1 bandwidth (4 overlapping 16-byte taps per thread per source row), not by issue slots or DRAM.
// MODE 0: taps from the planar float32 frame; 1: from the pixel-interleaved float4 copy; 2: from the fp16 copy (C == 3)
// OUTFMT 0: planar float32 crops; 1 (C == 3): pixel-interleaved bfloat16 (r,g,b,0), one 8-byte store per pixel
template <int C, int MODE, int OUTFMT>
__global__ void __launch_bounds__(CROP_MAX_THREADS, 3) hpb_crop_pixels_kernel(const CropPixParams p) {
    constexpr bool PACKED = MODE != 0;
    constexpr int NCH = C == 4 ? 5 : C;  // RGB-D: the depth-validity map is resampled as a 5th channel
    constexpr int NRAW = PACKED ? 4 : C;
    const int n = blockIdx.y;
    const int band = p.band;
    const int row0 = blockIdx.x * band;
    const int tid = threadIdx.x;
    extern __shared__ __align__(16) float sm[];
    float *sWY = sm;                                             // [band][CROP_SPAN], in SLOT order
    int *sBY = reinterpret_cast<int *>(sWY + band * CROP_SPAN);  // [band] first source row of the window, -1 = all-zero row
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
        const bool zero = wt[0] == 0.f && wt[1] == 0.f && wt[2] == 0.f && wt[3] == 0.f;
        sBY[i] = zero ? -1 : base;
        float ws[CROP_SPAN];
#pragma unroll
        for (int q = 0; q < CROP_SPAN; ++q) {  // slot q holds source row base + ((q - base) & 3)
            const int r = (q - base) & 3;
            ws[q] = r == 0 ? wt[0] : r == 1 ? wt[1] : r == 2 ? wt[2] : wt[3];
        }
        *reinterpret_cast<float4 *>(sWY + i * CROP_SPAN) = make_float4(ws[0], ws[1], ws[2], ws[3]);
    }
    // column taps of this thread's column (registers); a span over CROP_SPAN anywhere sends the CTA to the generic path
    float wx[CROP_SPAN] = {0.f, 0.f, 0.f, 0.f};
    int bxx = 0;
    const int j0 = blockIdx.z * blockDim.x + tid;
    if (j0 < p.w) {
        if (!axis_weights(x1, bin_w, j0, p.W, bxx, wx)) sGeneric = 1;
    }
    if (p.W < CROP_SPAN) sGeneric = 1;  // tiny frames: generic path
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
        const float4 *base4 = MODE == 1 ? p.packed + (size_t)im * p.H * p.W + bxx : nullptr;
        const uint2 *base8 = MODE == 2 ? p.packed_h + (size_t)im * p.H * p.W + bxx : nullptr;
        const uint2 *pp8 = base8;
        const float *base1 = img + bxx;
        const float4 *pp4 = base4;  // first tap of row min(prow, H-1)
        const float *pp1 = base1;
        const int Wst = p.W, Hm1 = p.H - 1;
        float raw[CROP_SPAN][NRAW];  // raw taps of source row `prow` (prefetched)
        float hw[CROP_SPAN][NCH];    // horizontally filtered source rows, slot = row & 3
#pragma unroll
        for (int r = 0; r < CROP_SPAN; ++r)
#pragma unroll
            for (int c = 0; c < NCH; ++c) hw[r][c] = 0.f;
        int prow = 0;                   // row held in `raw`
        int loaded_hi = -0x40000000;    // last row filtered into the window
        auto fetch = [&]() {
            if (MODE == 2) {
#pragma unroll
                for (int k = 0; k < CROP_SPAN; ++k) {
                    const uint2 v = __ldg(pp8 + k);
                    const float2 rg = __half22float2(*reinterpret_cast<const __half2 *>(&v.x));
                    const float2 b0 = __half22float2(*reinterpret_cast<const __half2 *>(&v.y));
                    raw[k][0] = rg.x; raw[k][1] = rg.y; raw[k][2] = b0.x; raw[k][3] = 0.f;
                }
            } else if (MODE == 1) {
#pragma unroll
                for (int k = 0; k < CROP_SPAN; ++k) {
                    const float4 v = __ldg(pp4 + k);
                    raw[k][0] = v.x; raw[k][1] = v.y; raw[k][2] = v.z; raw[k][3] = v.w;
                }
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int k = 0; k < CROP_SPAN; ++k) raw[k][c] = __ldg(pp1 + c * plane_in + k);
            }
        };
        auto filter_into = [&](float (&dst)[NCH]) {
#pragma unroll
            for (int c = 0; c < C; ++c)
                dst[c] = fmaf(wx[3], raw[3][c], fmaf(wx[2], raw[2][c], fmaf(wx[1], raw[1][c], wx[0] * raw[0][c])));
            if (C == 4)
                dst[NCH - 1] = fmaf(wx[3], raw[3][3] > 0.f ? 1.f : 0.f, fmaf(wx[2], raw[2][3] > 0.f ? 1.f : 0.f,
                                    fmaf(wx[1], raw[1][3] > 0.f ? 1.f : 0.f, wx[0] * (raw[0][3] > 0.f ? 1.f : 0.f))));
        };
        float *optr[C];  // one output pointer per plane, bumped by one row per iteration
#pragma unroll
        for (int c = 0; c < C; ++c) optr[c] = out + c * plane_out + (size_t)row0 * p.w + j0;
        uint2 *optr_h = OUTFMT == 1 ? p.crops_h + (size_t)n * p.crops_bs + (size_t)row0 * p.w + j0 : nullptr;
        const int wout = p.w;
        for (int i = 0; i < rows; ++i) {
            const int by = sBY[i];  // uniform over the CTA
            float acc[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) acc[c] = 0.f;
            if (by >= 0) {
                const float4 wq = *reinterpret_cast<const float4 *>(sWY + i * CROP_SPAN);
                if (by > loaded_hi + 1 || by < loaded_hi - 3) {  // first row of the band / a jump: restart the window at `by`
                    prow = by;
                    pp4 = base4 + (size_t)min(by, Hm1) * Wst;
                    pp8 = base8 + (size_t)min(by, Hm1) * Wst;
                    pp1 = base1 + (size_t)min(by, Hm1) * Wst;
                    fetch();
                    loaded_hi = by - 1;
                }
                while (loaded_hi < by + 3) {
                    ++loaded_hi;  // == prow
                    switch (loaded_hi & 3) {
                        case 0: filter_into(hw[0]); break;
                        case 1: filter_into(hw[1]); break;
                        case 2: filter_into(hw[2]); break;
                        default: filter_into(hw[3]); break;
                    }
                    const int step = prow < Hm1 ? Wst : 0;  // rows past the frame repeat the last row (their weights are zero)
                    if (MODE == 2) pp8 += step;
                    else if (MODE == 1) pp4 += step;
                    else pp1 += step;
                    ++prow;
                    fetch();
                }
#pragma unroll
                for (int c = 0; c < NCH; ++c)
                    acc[c] = fmaf(wq.w, hw[3][c], fmaf(wq.z, hw[2][c], fmaf(wq.y, hw[1][c], wq.x * hw[0][c])));
            }
            if (C == 4 && acc[NCH - 1] < 0.99f) acc[3] = 0.0f;  // cropping.py:191-195
            if (OUTFMT == 1) {
                __stcs(optr_h, pack_bf16x4(acc[0], acc[1], acc[2]));
                optr_h += wout;
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    __stcs(optr[c], acc[c]);
                    optr[c] += wout;
                }
            }
        }
        return;
    }

    // generic roi_align: 4x4 samples, 4 taps each (heavy down-sampling: crop box wider than ~1.7x the output)
    const int jlo = blockIdx.z * blockDim.x, ncol = min((int)blockDim.x, p.w - jlo);  // this CTA's column tile
    const int npx = rows * ncol;
    for (int q = tid; q < npx; q += blockDim.x) {
        const int i = q / ncol, j = jlo + (q - i * ncol);
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
        if (OUTFMT == 1) {
            __stcs(p.crops_h + (size_t)n * p.crops_bs + (size_t)oy * p.w + j, pack_bf16x4(acc[0], acc[1], acc[2]));
        } else {
            float *o = out + (size_t)oy * p.w + j;
#pragma unroll
            for (int c = 0; c < C; ++c) __stcs(o + c * plane_out, acc[c]);
        }
    }
}

}  // namespace

int hpb_launch_crop_tma(hpb_ctx *ctx, const hpbc::CropPixParams &p0, int band, cudaStream_t stream);

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
                           const float *boxes, int b, int h, int w, float *crops, int64_t crops_bs, int tap_bits,
                           cudaStream_t stream, void *crops_bf16x4) {
    if (b == 0) return HPB_OK;
    if (C != 3 && C != 4) {
        hpb_set_error("hpb_crop: C must be 3 or 4 (got %d)", C);
        return HPB_EINVAL;
    }
    CropPixParams p;
    p.images = images; p.im_ids = im_ids; p.boxes = boxes;
    p.n_im = n_im; p.C = C; p.H = H; p.W = W; p.b = b; p.h = h; p.w = w;
    p.crops = crops; p.crops_bs = crops_bs;
    p.crops_h = reinterpret_cast<uint2 *>(crops_bf16x4);
    p.packed = nullptr;
    p.packed_h = nullptr;
    // few distinct frames, many hypotheses: interleave the frames once so a tap is one 16-byte (fp16 copy: 8-byte) load
    const long long n_px = (long long)H * W, total = n_px * n_im;
    const bool half_taps = tap_bits == 16 && C == 3;  // per launch, not context state: concurrent callers cannot race on it
    const bool use_pack = (long long)n_im * 8 <= b && total * 16 <= (1ll << 28);
    if (use_pack) {
        const int rc = hpb_crop_reserve(ctx, total, stream);
        if (rc != HPB_OK) return rc;
    }
    {
        const int rc = hpb_stream_enter(ctx, stream);
        if (rc != HPB_OK) return rc;
    }
    if (use_pack) {
        const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
        if (half_taps) hpb_pack_frames_half_kernel<<<blocks, 256, 0, stream>>>(images, n_px, total, (uint2 *)ctx->frame_pack);
        else if (C == 3) hpb_pack_frames_kernel<3><<<blocks, 256, 0, stream>>>(images, n_px, total, (float4 *)ctx->frame_pack);
        else hpb_pack_frames_kernel<4><<<blocks, 256, 0, stream>>>(images, n_px, total, (float4 *)ctx->frame_pack);
        HPB_CUDA_OK(cudaGetLastError());
        ctx->launches++;
        if (half_taps) p.packed_h = (const uint2 *)ctx->frame_pack;
        else p.packed = (const float4 *)ctx->frame_pack;
    }
    // rows per CTA: enough CTAs to fill the machine for small batches, long bands (window fill amortised) otherwise
    int band = CROP_BAND_MAX;
    while (band > 8 && (long long)b * ((h + band - 1) / band) < 2ll * ctx->sm_count) band /= 2;
    p.band = band;
    const int threads = w >= CROP_MAX_THREADS ? CROP_MAX_THREADS : ((w + 31) / 32) * 32;  // one thread per output column
    const size_t smem = (size_t)band * CROP_SPAN * sizeof(float) + (size_t)band * sizeof(int);
    dim3 grid((h + band - 1) / band, b, (w + threads - 1) / threads);
    if (C == 3 && p.packed_h && ctx->crop_tma) {  // fp16 taps of an RGB frame: source rows streamed by TMA (hpb_crop_tma.cu)
        const int rc = hpb_launch_crop_tma(ctx, p, band, stream);
        if (rc != HPB_ENOTFOUND) {
            if (rc == HPB_OK) hpb_stream_leave(ctx, stream);
            return rc;
        }
    }
    if (C == 3 && p.crops_h) {
        if (p.packed_h) hpb_crop_pixels_kernel<3, 2, 1><<<grid, threads, smem, stream>>>(p);
        else if (p.packed) hpb_crop_pixels_kernel<3, 1, 1><<<grid, threads, smem, stream>>>(p);
        else hpb_crop_pixels_kernel<3, 0, 1><<<grid, threads, smem, stream>>>(p);
    } else if (C == 3) {
        if (p.packed_h) hpb_crop_pixels_kernel<3, 2, 0><<<grid, threads, smem, stream>>>(p);
        else if (p.packed) hpb_crop_pixels_kernel<3, 1, 0><<<grid, threads, smem, stream>>>(p);
        else hpb_crop_pixels_kernel<3, 0, 0><<<grid, threads, smem, stream>>>(p);
    } else {
        if (p.packed) hpb_crop_pixels_kernel<4, 1, 0><<<grid, threads, smem, stream>>>(p);
        else hpb_crop_pixels_kernel<4, 0, 0><<<grid, threads, smem, stream>>>(p);
    }
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    hpb_stream_leave(ctx, stream);
    return HPB_OK;
}

int hpb_crop_reserve(hpb_ctx *ctx, int64_t frame_pixels, cudaStream_t stream) {
    return hpb_ws_grow(ctx, &ctx->frame_pack, &ctx->frame_pack_bytes, (size_t)frame_pixels * 16, stream, "crop frame-pack");
}
