// Perspective crop (roi_align, sampling_ratio 4) of an RGB frame with the source rows streamed through a TMA-fed
// shared-memory ring.  Same arithmetic, same results as hpb_crop_pixels_kernel<3, 2, *> (hpb_crop.cu): the separable
// 4-tap row / column weights, a register window of 4 horizontally filtered source rows per thread, one thread per output
// column marching down a band of rows.  What changes is where the taps come from.
//
// hpb_crop.cu gathers every tap with a per-lane 8-byte global load: ncu showed that kernel waiting on those loads
// (long-scoreboard 4.8 per issue, L1 tag stage busy).  Here a producer thread asks the TMA unit for the band's source window
// two rows at a time (cp.async.bulk.tensor.3d boxes of 128 pixels x 2 rows, completion on an mbarrier), four stages
// ahead of the consumers, and the 320 consumer threads read their taps from shared memory: no per-lane address generation
// or tag look-up on the global side, every source row of a band crosses L2 -> SM exactly once.
//
// Source = the context's pixel-interleaved fp16 copy of the frames ([n_im, H, W] pixels of 8 bytes (r, g, b, 0)), described to
// the TMA as a 3-D tensor of 64-bit elements so that rows past the last line of a frame are zero-filled per frame (they
// only ever meet zero weights).  CTAs whose crop box is too wide for the 4-tap fast path take the generic 64-tap loop, as in
// hpb_crop.cu.
#include <cuda.h>

#include "hpb_crop_math.cuh"

namespace {

using namespace hpbc;

constexpr int RING_STAGES = 4;
constexpr int BOX_W = 128, BOX_H = 2;                        // one TMA box: 128 pixels x 2 rows x 8 B = 2 KB
constexpr int BOXES_MAX = 7;                                 // 896 pixels: every fast-path crop (bin_w < 2.7) fits
constexpr int STAGE_BYTES = BOXES_MAX * BOX_W * BOX_H * 8;   // 14 KB
constexpr int N_CONSUMERS = CROP_MAX_THREADS;                // 320: one per output column
constexpr int TMA_THREADS = N_CONSUMERS + 32;                // + the producer warp

__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int OUTFMT>
__global__ void __launch_bounds__(TMA_THREADS, 3) hpb_crop_tma_kernel(const __grid_constant__ CUtensorMap tmap, const CropPixParams p) {
    extern __shared__ __align__(128) unsigned char dsm[];  // [RING_STAGES][STAGE_BYTES] | sWY [band][4] | sBY [band]
    __shared__ __align__(8) unsigned long long full_bar[RING_STAGES], empty_bar[RING_STAGES];
    __shared__ int sGeneric, sR0, sR1, sC0, sC1;
    const int n = blockIdx.y;
    const int band = p.band;
    const int row0 = blockIdx.x * band;
    const int tid = threadIdx.x;
    const bool consumer = tid < N_CONSUMERS;
    float *sWY = reinterpret_cast<float *>(dsm + RING_STAGES * STAGE_BYTES);
    int *sBY = reinterpret_cast<int *>(sWY + band * CROP_SPAN);

    const float *bx = p.boxes + (size_t)n * 4;
    const float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
    const float roi_w = fmaxf(x2 - x1, 1.0f), roi_h = fmaxf(y2 - y1, 1.0f);
    const float bin_w = roi_w / (float)p.w, bin_h = roi_h / (float)p.h;
    const int rows = min(band, p.h - row0);
    const int jlo = blockIdx.z * N_CONSUMERS, ncol = min(N_CONSUMERS, p.w - jlo);  // this CTA's column tile
    if (tid == 0) {
        sGeneric = 0;
        sR0 = 0x7fffffff; sR1 = -1; sC0 = 0x7fffffff; sC1 = -1;
        for (int s = 0; s < RING_STAGES; ++s) {
            mbar_init((unsigned)__cvta_generic_to_shared(&full_bar[s]), 1);
            mbar_init((unsigned)__cvta_generic_to_shared(&empty_bar[s]), ncol);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // ---- vertical weights of the band's rows (slot order), as in hpb_crop.cu ----
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
    // ---- column taps of this thread's column ----
    float wx[CROP_SPAN] = {0.f, 0.f, 0.f, 0.f};
    int bxx = 0;
    const int j0 = jlo + tid;
    const bool active = consumer && tid < ncol;
    if (active) {
        if (!axis_weights(x1, bin_w, j0, p.W, bxx, wx)) sGeneric = 1;
        const int over = max(0, bxx + CROP_SPAN - 1 - (p.W - 1));  // re-base so that the 4 taps are 4 pixels inside the frame
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
    if (p.W < CROP_SPAN) sGeneric = 1;
    __syncthreads();
    // ---- the band's source window: rows [R0, R1] (contiguous, consumed in order), columns [C0, C1] ----
    if (tid == 0) {
        int r0 = 0x7fffffff, r1 = -1;
        for (int i = 0; i < rows; ++i) {
            const int by = sBY[i];
            if (by >= 0) { r0 = min(r0, by); r1 = max(r1, by + CROP_SPAN - 1); }
        }
        sR0 = r0; sR1 = r1;
    }
    if (active) {  // base columns are non-decreasing in j, but the right-edge re-basing can pull late ones back: reduce
        atomicMin(&sC0, bxx);
        atomicMax(&sC1, bxx + CROP_SPAN - 1);
    }
    __syncthreads();
    // the TMA wants the first pixel of a box row on a 16-byte boundary: 8-byte pixels -> an even column
    const int R0 = sR0, R1 = sR1, C0 = sC0 & ~1;
    const int nbx = sC1 >= C0 ? ((sC1 - C0) / BOX_W + 1) : 0;
    const bool generic = sGeneric != 0 || nbx > BOXES_MAX;
    const int im = p.im_ids[n];
    const size_t plane_out = (size_t)p.h * p.w;

    if (generic) {
        // 4x4 samples, 4 taps each, from the planar float32 frame (heavy down-sampling), exactly as hpb_crop.cu
        if (!consumer) return;
        const float *img = p.images + (size_t)im * 3 * p.H * p.W;
        const size_t plane_in = (size_t)p.H * p.W;
        const int npx = rows * ncol;
        for (int q = tid; q < npx; q += N_CONSUMERS) {
            const int i = q / ncol, j = jlo + (q - i * ncol);
            const int oy = row0 + i;
            float acc[3] = {0.f, 0.f, 0.f};
            for (int sy = 0; sy < 4; ++sy) {
                const AxisTap ty = axis_tap(y1 + (float)oy * bin_h + ((float)sy + 0.5f) * bin_h / 4.0f, p.H);
                if (!ty.valid) continue;
                for (int sx = 0; sx < 4; ++sx) {
                    const AxisTap tx = axis_tap(x1 + (float)j * bin_w + ((float)sx + 0.5f) * bin_w / 4.0f, p.W);
                    if (!tx.valid) continue;
                    const float w1 = ty.wlo * tx.wlo, w2 = ty.wlo * tx.whi, w3 = ty.whi * tx.wlo, w4 = ty.whi * tx.whi;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float *pl = img + c * plane_in;
                        const float v1 = __ldg(pl + (size_t)ty.lo * p.W + tx.lo), v2 = __ldg(pl + (size_t)ty.lo * p.W + tx.hi);
                        const float v3 = __ldg(pl + (size_t)ty.hi * p.W + tx.lo), v4 = __ldg(pl + (size_t)ty.hi * p.W + tx.hi);
                        acc[c] += 0.0625f * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
                    }
                }
            }
            if (OUTFMT == 1) {
                __stcs(p.crops_h + (size_t)n * p.crops_bs + (size_t)oy * p.w + j, pack_bf16x4(acc[0], acc[1], acc[2]));
            } else {
                float *o = p.crops + (size_t)n * p.crops_bs + (size_t)oy * p.w + j;
#pragma unroll
                for (int c = 0; c < 3; ++c) __stcs(o + c * plane_out, acc[c]);
            }
        }
        return;
    }

    const int n_tiles = R1 >= R0 ? (R1 - R0) / BOX_H + 1 : 0;
    const unsigned ring = (unsigned)__cvta_generic_to_shared(dsm);
    const unsigned full0 = (unsigned)__cvta_generic_to_shared(&full_bar[0]), empty0 = (unsigned)__cvta_generic_to_shared(&empty_bar[0]);

    if (!consumer) {
        // ---- producer: one elected thread keeps the ring RING_STAGES tiles ahead of the consumers ----
        if (tid == N_CONSUMERS) {
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % RING_STAGES;
                if (t >= RING_STAGES) mbar_wait(empty0 + 8 * s, ((t / RING_STAGES) - 1) & 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + 8 * s), "r"(nbx * BOX_W * BOX_H * 8) : "memory");
                for (int j = 0; j < nbx; ++j)
                    asm volatile(
                        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                        ::"r"(ring + s * STAGE_BYTES + j * (BOX_W * BOX_H * 8)), "l"(&tmap), "r"(C0 + j * BOX_W), "r"(R0 + t * BOX_H), "r"(im),
                        "r"(full0 + 8 * s)
                        : "memory");
            }
        }
        return;
    }
    if (!active) return;  // columns past the output width (the empty barriers count `ncol` arrivals)

    // ---- consumers ----
    unsigned co[CROP_SPAN];  // byte offset of this thread's 4 taps inside a stage row
#pragma unroll
    for (int k = 0; k < CROP_SPAN; ++k) {
        const int xr = bxx - C0 + k;
        co[k] = (unsigned)((xr / BOX_W) * (BOX_W * BOX_H * 8) + (xr % BOX_W) * 8);
    }
    float hw[CROP_SPAN][3];
#pragma unroll
    for (int r = 0; r < CROP_SPAN; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) hw[r][c] = 0.f;
    auto filter_row = [&](int r, float (&dst)[3]) {  // source row r: wait for its tile, filter its 4 taps, release the tile
        const int rel = r - R0;
        const int t = rel / BOX_H, s = t % RING_STAGES;
        if ((rel % BOX_H) == 0) mbar_wait(full0 + 8 * s, (t / RING_STAGES) & 1);
        const unsigned char *row = dsm + s * STAGE_BYTES + (rel % BOX_H) * (BOX_W * 8);
        float raw[CROP_SPAN][3];
#pragma unroll
        for (int k = 0; k < CROP_SPAN; ++k) {
            const uint2 v = *reinterpret_cast<const uint2 *>(row + co[k]);
            const float2 rg = __half22float2(*reinterpret_cast<const __half2 *>(&v.x));
            const float2 b0 = __half22float2(*reinterpret_cast<const __half2 *>(&v.y));
            raw[k][0] = rg.x; raw[k][1] = rg.y; raw[k][2] = b0.x;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            dst[c] = fmaf(wx[3], raw[3][c], fmaf(wx[2], raw[2][c], fmaf(wx[1], raw[1][c], wx[0] * raw[0][c])));
        if ((rel % BOX_H) == BOX_H - 1 || r == R1) mbar_arrive(empty0 + 8 * s);
    };
    float *optr[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) optr[c] = p.crops ? p.crops + (size_t)n * p.crops_bs + c * plane_out + (size_t)row0 * p.w + j0 : nullptr;
    uint2 *optr_h = OUTFMT == 1 ? p.crops_h + (size_t)n * p.crops_bs + (size_t)row0 * p.w + j0 : nullptr;
    const int wout = p.w;
    int loaded_hi = R0 - 1;  // last source row filtered into the window; rows are consumed strictly in order R0 .. R1
    for (int i = 0; i < rows; ++i) {
        const int by = sBY[i];  // uniform over the CTA
        float acc[3] = {0.f, 0.f, 0.f};
        if (by >= 0) {
            const float4 wq = *reinterpret_cast<const float4 *>(sWY + i * CROP_SPAN);
            while (loaded_hi < by + CROP_SPAN - 1) {
                ++loaded_hi;
                switch (loaded_hi & 3) {
                    case 0: filter_row(loaded_hi, hw[0]); break;
                    case 1: filter_row(loaded_hi, hw[1]); break;
                    case 2: filter_row(loaded_hi, hw[2]); break;
                    default: filter_row(loaded_hi, hw[3]); break;
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
                acc[c] = fmaf(wq.w, hw[3][c], fmaf(wq.z, hw[2][c], fmaf(wq.y, hw[1][c], wq.x * hw[0][c])));
        }
        if (OUTFMT == 1) {
            __stcs(optr_h, pack_bf16x4(acc[0], acc[1], acc[2]));
            optr_h += wout;
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                __stcs(optr[c], acc[c]);
                optr[c] += wout;
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn crop_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *q = nullptr;
        cudaDriverEntryPointQueryResult res;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &res) == cudaSuccess && res == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(q);
        else
            cudaGetLastError();
    }
    return fn;
}

}  // namespace

// Launches the TMA crop on the fp16 packed frames `packed_h` ([n_im, H, W] pixels of 8 bytes).  Returns HPB_ENOTFOUND when the
// shape is not served (odd frame width: the TMA needs 16-byte row strides), the caller then uses hpb_crop_pixels_kernel.
int hpb_launch_crop_tma(hpb_ctx *ctx, const hpbc::CropPixParams &p0, int band, cudaStream_t stream) {
    if ((p0.W & 1) || !p0.packed_h || p0.C != 3) return HPB_ENOTFOUND;
    EncodeTiledFn encode = crop_encode_fn();
    if (!encode) return HPB_ENOTFOUND;
    CUtensorMap tmap;
    const cuuint64_t dims[3] = {(cuuint64_t)p0.W, (cuuint64_t)p0.H, (cuuint64_t)p0.n_im};
    const cuuint64_t strides[2] = {(cuuint64_t)p0.W * 8, (cuuint64_t)p0.H * p0.W * 8};
    const cuuint32_t box[3] = {(cuuint32_t)BOX_W, (cuuint32_t)BOX_H, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint2 *>(p0.packed_h), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return HPB_ENOTFOUND;
    hpbc::CropPixParams p = p0;
    p.band = band;
    const size_t smem = (size_t)RING_STAGES * STAGE_BYTES + (size_t)band * CROP_SPAN * sizeof(float) + (size_t)band * sizeof(int);
    static bool attr_set = false;
    if (!attr_set) {
        HPB_CUDA_OK(cudaFuncSetAttribute(hpb_crop_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RING_STAGES * STAGE_BYTES + 4096));
        HPB_CUDA_OK(cudaFuncSetAttribute(hpb_crop_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, RING_STAGES * STAGE_BYTES + 4096));
        attr_set = true;
    }
    dim3 grid((p.h + band - 1) / band, p.b, (p.w + N_CONSUMERS - 1) / N_CONSUMERS);
    if (p.crops_h) hpb_crop_tma_kernel<1><<<grid, TMA_THREADS, smem, stream>>>(tmap, p);
    else hpb_crop_tma_kernel<0><<<grid, TMA_THREADS, smem, stream>>>(tmap, p);
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}
