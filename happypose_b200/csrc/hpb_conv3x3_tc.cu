// 3x3 / stride 1 / pad 1 convolution, 64 -> 64 channels, + bias (+ residual) + ReLU on bf16 NHWC activations as a tcgen05
// implicit GEMM: the BasicBlock convolutions of ResNet layer1 (torchvision_resnet.py:59-83, batch-norm folded:
// relu(bn2(conv2(y)) + identity) and relu(bn1(conv1(x)))).
//
// Same construction as hpb_stem_tc.cu (read its header first): persistent CTAs, tile = 16 rows x 8 columns of output pixels
// (M = 128), N = 64; ONE 4-D TMA box per tile {64 ch, 16 px, 18 rows} = the tile and its 1-pixel halo, origin (-1, -1)
// relative to the tile -- the zero padding of the convolution is TMA's out-of-bounds fill; the A operand of tap (kh, kw) is
// the box at start offset kh * 2 KB + kw * 128 B; the 9 weight taps (72 KB) stay resident; 36 tcgen05.mma per tile into
// double-buffered TMEM accumulators; for the block's second convolution the identity tensor is added on the tensor core too
// (a tenth A operand times an identity tap, see C3Cfg); epilogue adds the float32 bias, clamps at 0, rounds to bf16, stages the
// tile and TMA-stores it.
// cuDNN serves the plain conv+bias+ReLU of these layers with a weight-stationary kernel at 1.27 PFLOP/s but the
// conv+bias+ADD+ReLU form only with a generic implicit-GEMM tile at 0.77 PFLOP/s: the residual variant is where this kernel
// is used (megapose/fast_resnet.py).
#include <cuda.h>
#include <cuda_bf16.h>

#include "hpb_common.cuh"
#include "hpb_tc_common.cuh"

namespace {

using namespace hpbtc;

constexpr int C3_N = 64, C3_KC = 64, C3_TAPS = 9;
constexpr int C3_TILE_W = 8, C3_TILE_H = 16;
constexpr int C3_BOX_W = 16, C3_BOX_H = C3_TILE_H + 2;
constexpr unsigned C3_A_BYTES = (unsigned)C3_BOX_W * C3_BOX_H * 128u;  // 36 KB
constexpr unsigned C3_B_TAP_BYTES = (unsigned)C3_N * C3_KC * 2u;       // 8 KB
constexpr unsigned C3_OUT_BYTES = 128u * C3_N * 2u;                    // 16 KB: the staging tile, and one residual tile
constexpr int C3_THREADS = 256;
constexpr int C3_MAX_STAGES = 3;

// RES: the residual is added ON THE TENSOR CORE: its tile (128 pixels x 64 channels, one TMA box per tile into the stage) is a
// tenth A operand multiplied by an identity matrix the host appends to the weights as a tenth tap -- bf16 -> fp32 and
// x * 1.0 are exact, so the accumulator receives exactly conv + residual in float32 and the epilogue stays the plain one.
// (A first version loaded each thread's residual pixel from global memory in the epilogue: 8 scattered 16-byte loads per
// thread through the L1 data pipe this kernel is bound by -- 0.327 ms against cuDNN's 0.278; see profiles/r2_conv3x3_tc.jsonl.)
template <bool RES>
struct C3Cfg {
    static constexpr int TAPS_B = RES ? C3_TAPS + 1 : C3_TAPS;
    static constexpr int STAGES = RES ? 2 : 3;
    static constexpr unsigned B_BYTES = TAPS_B * C3_B_TAP_BYTES;                          // 80 / 72 KB
    static constexpr unsigned STAGE_BYTES = C3_A_BYTES + (RES ? C3_OUT_BYTES : 0u);       // input box (+ residual tile)
    static constexpr unsigned OFF_B = 0u;
    static constexpr unsigned OFF_A = OFF_B + B_BYTES;
    static constexpr unsigned OFF_OUT = OFF_A + STAGES * STAGE_BYTES;
    static constexpr unsigned OFF_BAR = OFF_OUT + C3_OUT_BYTES;
    static constexpr unsigned SMEM = OFF_BAR + 128u + 1024u;
    static_assert(SMEM <= 232448u, "shared memory budget");
    static_assert(STAGE_BYTES % 1024u == 0 && OFF_A % 1024u == 0 && OFF_OUT % 1024u == 0, "1 KB swizzle alignment");
};

template <bool RES>
__global__ void __launch_bounds__(C3_THREADS, 1)
hpb_conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res, const float *__restrict__ bias,
                      int tiles_x, int tiles_per_image, int n_tiles) {
    using Cfg = C3Cfg<RES>;
    constexpr int C3_STAGES = Cfg::STAGES;
    constexpr unsigned C3_OFF_B = Cfg::OFF_B, C3_OFF_A = Cfg::OFF_A, C3_OFF_OUT = Cfg::OFF_OUT, C3_OFF_BAR = Cfg::OFF_BAR;
    constexpr unsigned C3_STAGE_BYTES = Cfg::STAGE_BYTES, C3_B_BYTES = Cfg::B_BYTES;
    extern __shared__ unsigned char smem_raw[];
    const unsigned base = ((unsigned)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;
    unsigned char *base_ptr = smem_raw + (base - (unsigned)__cvta_generic_to_shared(smem_raw));
    const unsigned sB = base + C3_OFF_B, sA = base + C3_OFF_A, sOut = base + C3_OFF_OUT;
    const unsigned bars = base + C3_OFF_BAR;
    // barriers (8 bytes each): full[s] 0..2, empty[s] 3..5, weights 6, tmem_full[a] 7..8, tmem_empty[a] 9..10; slot 11: TMEM base
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (C3_MAX_STAGES + s); };
    const unsigned w_bar = bars + 8u * (2 * C3_MAX_STAGES);
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * C3_MAX_STAGES + 1 + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * C3_MAX_STAGES + 3 + a); };
    const unsigned tmem_slot = bars + 8u * (2 * C3_MAX_STAGES + 5);
    volatile unsigned *tmem_slot_ptr = reinterpret_cast<volatile unsigned *>(base_ptr + C3_OFF_BAR + 8u * (2 * C3_MAX_STAGES + 5));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < C3_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(w_bar, 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            mbar_expect_tx(w_bar, C3_B_BYTES);
            for (int tap = 0; tap < Cfg::TAPS_B; ++tap) tma_load_2d(sB + tap * C3_B_TAP_BYTES, &map_w, tap * C3_KC, 0, w_bar);
        }
        __syncwarp();
        int s = 0;
        unsigned ph = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int n = t / tiles_per_image, r = t - n * tiles_per_image;
            const int ty = r / tiles_x, tx = r - ty * tiles_x;
            mbar_wait(empty_bar(s), ph ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(full_bar(s), C3_STAGE_BYTES);
                // origin one pixel up and left of the tile: rows / columns -1 and >= H / W are zero-filled = the padding
                tma_load_4d(sA + s * C3_STAGE_BYTES, &map_in, 0, tx * C3_TILE_W - 1, ty * C3_TILE_H - 1, n, full_bar(s));
                if constexpr (RES) tma_load_4d(sA + s * C3_STAGE_BYTES + C3_A_BYTES, &map_res, 0, tx * C3_TILE_W, ty * C3_TILE_H, n, full_bar(s));
            }
            __syncwarp();
            if (++s == C3_STAGES) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        mbar_wait(w_bar, 0);
        int s = 0, acc = 0;
        unsigned ph = 0, acc_ph = 0;
        const unsigned long long b_desc0 = umma_desc(sB, 1024u);
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait(tempty_bar(acc), acc_ph ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned d_tmem = tmem_base + (unsigned)(acc * C3_N);
            mbar_wait(full_bar(s), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned long long a_desc0 = umma_desc(sA + s * C3_STAGE_BYTES, 2048u);
            if (elect_one()) {
#pragma unroll
                for (int tap = 0; tap < C3_TAPS; ++tap) {
#pragma unroll
                    for (int k = 0; k < C3_KC / 16; ++k) {
                        const unsigned long long a_off = (unsigned long long)(((tap / 3) * 2048 + (tap % 3) * 128 + k * 32) >> 4);
                        const unsigned long long b_off = (unsigned long long)((tap * (int)C3_B_TAP_BYTES + k * 32) >> 4);
                        umma_bf16(d_tmem, a_desc0 + a_off, b_desc0 + b_off, (tap | k) != 0 ? 1u : 0u);
                    }
                }
                if constexpr (RES) {  // + residual tile (128 rows of 128 B, 8-row groups 1 KB apart) x identity (tap 9)
                    const unsigned long long r_desc0 = umma_desc(sA + s * C3_STAGE_BYTES + C3_A_BYTES, 1024u);
#pragma unroll
                    for (int k = 0; k < C3_KC / 16; ++k)
                        umma_bf16(d_tmem, r_desc0 + (unsigned long long)((k * 32) >> 4),
                                  b_desc0 + (unsigned long long)((C3_TAPS * (int)C3_B_TAP_BYTES + k * 32) >> 4), 1u);
                }
                umma_commit(empty_bar(s));
                umma_commit(tfull_bar(acc));
            }
            __syncwarp();
            if (++s == C3_STAGES) { s = 0; ph ^= 1u; }
            if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int q = warp & 3;
        const int p = q * 32 + lane;  // pixel of the tile = accumulator row
        const int et = tid - 128;
        int acc = 0;
        unsigned acc_ph = 0;
        float breg[C3_N];
#pragma unroll
        for (int c = 0; c < C3_N; ++c) breg[c] = __ldg(bias + c);
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int n = t / tiles_per_image, r = t - n * tiles_per_image;
            const int ty = r / tiles_x, tx = r - ty * tiles_x;
            mbar_wait(tfull_bar(acc), acc_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned v[64];
            const unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(acc * C3_N);
            tmem_ld16(taddr, v);
            tmem_ld16(taddr + 16u, v + 16);
            tmem_ld16(taddr + 32u, v + 32);
            tmem_ld16(taddr + 48u, v + 48);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(tempty_bar(acc));
            if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous store has read the staging tile
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const unsigned row = sOut + (unsigned)p * 128u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const unsigned w0 = bias_relu_pack(v[8 * j + 0], v[8 * j + 1], breg[8 * j + 0], breg[8 * j + 1]);
                const unsigned w1 = bias_relu_pack(v[8 * j + 2], v[8 * j + 3], breg[8 * j + 2], breg[8 * j + 3]);
                const unsigned w2 = bias_relu_pack(v[8 * j + 4], v[8 * j + 5], breg[8 * j + 4], breg[8 * j + 5]);
                const unsigned w3 = bias_relu_pack(v[8 * j + 6], v[8 * j + 7], breg[8 * j + 6], breg[8 * j + 7]);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (unsigned)((j ^ (p & 7)) << 4)), "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                             : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) {
                tma_store_4d(&map_out, 0, tx * C3_TILE_W, ty * C3_TILE_H, n, sOut);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
        }
        if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

template <bool RES>
int launch_conv3x3(hpb_ctx *ctx, const void *x, int b, int H, int W, const void *w, const float *bias, const void *res, void *out,
                   cudaStream_t stream) {
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return HPB_ENOTFOUND;
    using Cfg = C3Cfg<RES>;
    CUtensorMap map_in, map_w, map_out, map_res;
    const cuuint32_t estr4[4] = {1u, 1u, 1u, 1u};
    const cuuint64_t dims[4] = {(cuuint64_t)C3_KC, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)b};
    const cuuint64_t strides[3] = {(cuuint64_t)C3_KC * 2, (cuuint64_t)W * C3_KC * 2, (cuuint64_t)H * W * C3_KC * 2};
    {
        const cuuint32_t box[4] = {(cuuint32_t)C3_KC, (cuuint32_t)C3_BOX_W, (cuuint32_t)C3_BOX_H, 1u};
        if (encode(&map_in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(x), dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return HPB_ENOTFOUND;
    }
    {
        const cuuint64_t wd[2] = {(cuuint64_t)Cfg::TAPS_B * C3_KC, (cuuint64_t)C3_N};
        const cuuint64_t ws[1] = {(cuuint64_t)Cfg::TAPS_B * C3_KC * 2};
        const cuuint32_t box[2] = {(cuuint32_t)C3_KC, (cuuint32_t)C3_N};
        const cuuint32_t estr[2] = {1u, 1u};
        if (encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(w), wd, ws, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return HPB_ENOTFOUND;
    }
    {
        const cuuint32_t box[4] = {(cuuint32_t)C3_N, (cuuint32_t)C3_TILE_W, (cuuint32_t)C3_TILE_H, 1u};
        if (encode(&map_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return HPB_ENOTFOUND;
    }
    map_res = map_out;
    if (RES) {  // the residual tile: the same box as the output tile, loaded instead of stored
        const cuuint32_t box[4] = {(cuuint32_t)C3_N, (cuuint32_t)C3_TILE_W, (cuuint32_t)C3_TILE_H, 1u};
        if (encode(&map_res, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(res), dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return HPB_ENOTFOUND;
    }
    HPB_CUDA_OK(cudaFuncSetAttribute(hpb_conv3x3_tc_kernel<RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    const int tiles_x = (W + C3_TILE_W - 1) / C3_TILE_W, tiles_per_image = tiles_x * ((H + C3_TILE_H - 1) / C3_TILE_H);
    const long long n_tiles = (long long)tiles_per_image * b;
    if (n_tiles > 0x7fffffffll) return HPB_ENOTFOUND;
    const int grid = (int)(n_tiles < ctx->sm_count ? n_tiles : ctx->sm_count);
    hpb_conv3x3_tc_kernel<RES><<<grid, C3_THREADS, Cfg::SMEM, stream>>>(map_in, map_w, map_out, map_res, bias, tiles_x, tiles_per_image, (int)n_tiles);
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}

}  // namespace

// x, res, out [b][H][W][64] bf16; bias [64] f32; res may be NULL.  w: without res [64 out][9 taps][64] bf16 (= the channels_last
// [64,64,3,3] weight); WITH res [64 out][10 taps][64]: the same nine taps followed by a 64 x 64 identity (out == channel).  Returns HPB_ENOTFOUND for shapes it does not serve (C or O != 64, H < 18, W < 16).
int hpb_launch_conv3x3_tc(hpb_ctx *ctx, const void *x, int b, int H, int W, int C, const void *w, const float *bias, int O, const void *res,
                          void *out, cudaStream_t stream) {
    if (C != C3_KC || O != C3_N || b < 1 || H < C3_BOX_H || W < C3_BOX_W) return HPB_ENOTFOUND;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(res)) & 15u)
        return HPB_ENOTFOUND;
    return res ? launch_conv3x3<true>(ctx, x, b, H, W, w, bias, res, out, stream) : launch_conv3x3<false>(ctx, x, b, H, W, w, bias, res, out, stream);
}
