// The ResNet stem as a tcgen05 implicit GEMM: 4x4 / stride-1 convolution of the space-to-depth network input (64 bf16
// channels per cell, written by the rasteriser's fused hand-off) + bias + ReLU, bf16 NHWC out.
//
// It is the 7x7 / stride-2 `conv1` + `bn1` + `relu` of torchvision_resnet.py:197-214 after batch-norm folding and the
// space-to-depth rewrite (megapose/fast_resnet.py: s2d_weight), i.e. the single largest kernel of a MegaPose step
// (cuDNN: 1.44 ms of 10.7 for the 576-row coarse batch; this kernel 1.0-1.2 ms).  GEMM view per CTA tile:
//   D[128 pixels, 64 out-channels] = sum over the 16 taps (kh, kw) of  A_tap[128 pixels, 64 channels] * W_tap[64 ch, 64 out]
// * persistent CTAs, one per SM; a tile is 16 rows x 8 columns of output pixels (HALO scheme, below);
// * ONE 4-D TMA box per tile -- the tile and its 3-cell halo, 128-byte swizzled -- feeds all 16 taps: the A operand of a tap is
//   a start-address offset into the box, so im2col never exists and every input cell crosses L2 -> SM once;
// * all 16 weight taps (128 KB) stay resident in shared memory for the life of the CTA;
// * one elected lane issues the tcgen05.mma (M 128, N 64, K 16, fp32 accumulate in tensor memory): 4 per tap, minus the
//   weight slices that are identically zero (KMASK: 49 of 64 remain for a 7x7 kernel), fully unrolled with immediate
//   descriptor offsets, and hands stages / accumulators on with tcgen05.commit; accumulators are double-buffered in TMEM
//   (2 x 64 columns) so the epilogue of tile i overlaps the MMAs of tile i + 1;
// * epilogue warps read their TMEM lane quarter (tcgen05.ld 32x32b), add the bias (registers), clamp at 0, round to bf16,
//   write the 128-byte pixel rows into a swizzled staging tile and one thread stores it with a 4-D TMA store (edge tiles
//   are clipped by TMA: any output size is served).
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue.  Every mbarrier wait is bounded
// (a protocol error traps instead of hanging the device).
// Bound (ncu, profiles/r2_stem_tc_ncu_full.txt): the shared-memory data pipe -- N = 64 needs 4 KB of A + 2 KB of B per
// 131 k MAC -- 82 % busy with tensor-core operand reads plus the epilogue's accesses; tensor pipe 47 % busy.
//
// Two operand-feeding schemes (template HALO):
// * HALO = false, "box per tap" (the first version, kept as the cross-check): tile 8 x 16, 16 TMA boxes of 16 KB per tile.
//   Correct but every input cell crosses L2 -> SM 16 times: 22 GB for the 576-row batch, 2.3-3.2 ms.
// * HALO = true: the tile is 16 rows x 8 columns and ONE box {64 ch, 16 px, 19 rows} (38 KB: the tile + its 3-cell halo,
//   rows padded to 16 cells = 2 KB so that consecutive image rows are exactly two swizzle atoms apart) is loaded per tile.
//   The A operand of tap (kh, kw) is then just a different START ADDRESS into that box: 8-row groups (the 8 cells of one
//   tile row) 2 KB apart (the descriptor's stride-dimension offset), start = box + kh * 2 KB + kw * 128 B.  That start is
//   not 1 KB aligned for kw > 0; with the descriptor's base_offset field left 0 the tensor core un-swizzles by the
//   ABSOLUTE shared-memory address bits -- the phase TMA used when it wrote the box -- which is what this needs (setting
//   base_offset = (start >> 7) & 7 makes it use the row index relative to the start instead: decoded on the device with
//   identity weights, scripts/debug_stem_tc.py).
#include <cuda.h>
#include <cuda_bf16.h>

#include "hpb_common.cuh"
#include "hpb_tc_common.cuh"

namespace {

using namespace hpbtc;

constexpr int TC_N = 64;                      // output channels = the MMA's N
constexpr int TC_KC = 64;                     // input channels per tap: one 128-byte swizzle row
constexpr int TC_TAPS = 16;
constexpr unsigned TC_B_TAP_BYTES = (unsigned)TC_N * TC_KC * 2u;  // 8 KB per tap
constexpr unsigned TC_B_BYTES = TC_TAPS * TC_B_TAP_BYTES;  // 128 KB
constexpr unsigned TC_OUT_BYTES = 128u * TC_N * 2u;        // 16 KB per staging tile
constexpr int TC_THREADS = 256;
constexpr int TC_MAX_STAGES = 4;

template <bool HALO>
struct StemCfg {
    static constexpr int TILE_W = HALO ? 8 : 16, TILE_H = HALO ? 16 : 8;  // output pixels per tile: 128 = the MMA's M
    static constexpr int STAGES = HALO ? 2 : 4;
    static constexpr int BOX_W = 16, BOX_H = HALO ? 19 : 8;               // input box (cells)
    static constexpr unsigned A_BYTES = (unsigned)BOX_W * BOX_H * 128u;    // 38 KB (tile + halo) / 16 KB (one tap)
    static constexpr unsigned SBO = HALO ? 2048u : 1024u;                 // bytes between 8-row groups of the A operand
    static constexpr int OUT_BUFS = HALO ? 1 : 2;
    static constexpr unsigned OFF_B = 0u;
    static constexpr unsigned OFF_A = OFF_B + TC_B_BYTES;
    static constexpr unsigned OFF_OUT = OFF_A + STAGES * A_BYTES;
    static constexpr unsigned OFF_BAR = OFF_OUT + OUT_BUFS * TC_OUT_BYTES;
    static constexpr unsigned SMEM = OFF_BAR + 128u + 1024u;  // + slack for the manual 1024-byte alignment
    static_assert(SMEM <= 232448u, "shared memory budget");
    static_assert(A_BYTES % 1024u == 0, "stages must keep the 1 KB swizzle alignment");
};

// KMASK: bit 4 * tap + k set = the 16-channel weight slice k of tap (kh, kw) is multiplied (compile time: the issue loop is
// fully unrolled with immediate descriptor offsets).  Bit 0 must be set.
template <bool HALO, unsigned long long KMASK>
__global__ void __launch_bounds__(TC_THREADS, 1)
hpb_stem_tc_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_w,
                   const __grid_constant__ CUtensorMap map_out, const float *__restrict__ bias, int tiles_x, int tiles_per_image,
                   int n_tiles, uint4 *__restrict__ out, int Hc, int Wc, int direct_store) {
    static_assert(KMASK & 1ull, "slice 0 initialises the accumulator");
    constexpr int FIRST_SLICE = 0;
    extern __shared__ unsigned char smem_raw[];
    const unsigned base = ((unsigned)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;
    unsigned char *base_ptr = smem_raw + (base - (unsigned)__cvta_generic_to_shared(smem_raw));
    using Cfg = StemCfg<HALO>;
    constexpr int TC_STAGES = Cfg::STAGES, TC_TILE_W = Cfg::TILE_W, TC_TILE_H = Cfg::TILE_H;
    constexpr unsigned TC_A_BYTES = Cfg::A_BYTES, OFF_BAR = Cfg::OFF_BAR;
    const unsigned sB = base + Cfg::OFF_B, sA = base + Cfg::OFF_A, sOut = base + Cfg::OFF_OUT;
    const unsigned bars = base + OFF_BAR;
    // barriers (8 bytes each): full[s] 0..3, empty[s] 4..7, weights 8, tmem_full[a] 9..10, tmem_empty[a] 11..12; word 13*8: TMEM base
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (TC_MAX_STAGES + s); };
    const unsigned w_bar = bars + 8u * (2 * TC_MAX_STAGES);
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * TC_MAX_STAGES + 1 + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * TC_MAX_STAGES + 3 + a); };
    const unsigned tmem_slot = bars + 8u * (2 * TC_MAX_STAGES + 5);
    volatile unsigned *tmem_slot_ptr = reinterpret_cast<volatile unsigned *>(base_ptr + OFF_BAR + 8u * (2 * TC_MAX_STAGES + 5));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
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
    if (warp == 2) {  // 128 TMEM columns: two 64-column fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = *tmem_slot_ptr;

    // Producer and MMA warps run their loops with ALL lanes (warp-uniform control flow and operands, so the descriptors live in
    // uniform registers and ptxas emits a bare UTCHMMA / UTMALDG); one elected lane issues.  (With `if (lane == 0)` around
    // the loops ptxas wrapped every tcgen05.mma in a vote / elect "waterfall" loop of ~28 instructions: the issuing thread,
    // not the tensor pipe, was the bottleneck -- ncu showed the pipe 18 % busy.)
    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            mbar_expect_tx(w_bar, TC_B_BYTES);
            for (int tap = 0; tap < TC_TAPS; ++tap) tma_load_2d(sB + tap * TC_B_TAP_BYTES, &map_w, tap * TC_KC, 0, w_bar);
        }
        __syncwarp();
        int s = 0;
        unsigned ph = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int n = t / tiles_per_image, r = t - n * tiles_per_image;
            const int ty = r / tiles_x, tx = r - ty * tiles_x;
            if constexpr (HALO) {  // one box per tile: the tile and its halo
                mbar_wait(empty_bar(s), ph ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(full_bar(s), TC_A_BYTES);
                    tma_load_4d(sA + s * TC_A_BYTES, &map_in, 0, tx * TC_TILE_W, ty * TC_TILE_H, n, full_bar(s));
                }
                __syncwarp();
                if (++s == TC_STAGES) { s = 0; ph ^= 1u; }
            } else {
                for (int tap = 0; tap < TC_TAPS; ++tap) {
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(full_bar(s), TC_A_BYTES);
                        tma_load_4d(sA + s * TC_A_BYTES, &map_in, 0, tx * TC_TILE_W + (tap & 3), ty * TC_TILE_H + (tap >> 2), n, full_bar(s));
                    }
                    __syncwarp();
                    if (++s == TC_STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        mbar_wait(w_bar, 0);
        int s = 0, acc = 0;
        unsigned ph = 0, acc_ph = 0;
        const unsigned long long b_desc0 = umma_desc(sB, 1024u);
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait(tempty_bar(acc), acc_ph ^ 1u);  // the epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned d_tmem = tmem_base + (unsigned)(acc * TC_N);
            if constexpr (HALO) {
                mbar_wait(full_bar(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned long long a_desc0 = umma_desc(sA + s * TC_A_BYTES, Cfg::SBO);
                if (elect_one()) {
#pragma unroll
                    for (int tap = 0; tap < TC_TAPS; ++tap) {
#pragma unroll
                        for (int k = 0; k < TC_KC / 16; ++k) {
                            if (!((KMASK >> (tap * 4 + k)) & 1ull)) continue;  // compile-time: an all-zero weight slice
                            // the window of tap (kh, kw) inside the box: kh image rows (2 KB) down, kw cells (128 B) right, then
                            // 32 bytes per K step inside the swizzle row; all in the descriptor's 16-byte address units
                            const unsigned long long a_off = (unsigned long long)(((tap >> 2) * 2048 + (tap & 3) * 128 + k * 32) >> 4);
                            const unsigned long long b_off = (unsigned long long)((tap * (int)TC_B_TAP_BYTES + k * 32) >> 4);
                            umma_bf16(d_tmem, a_desc0 + a_off, b_desc0 + b_off, (tap * 4 + k) != FIRST_SLICE ? 1u : 0u);
                        }
                    }
                    umma_commit(empty_bar(s));  // the box is free once all MMAs have read it
                    umma_commit(tfull_bar(acc));  // accumulator complete
                }
                __syncwarp();
                if (++s == TC_STAGES) { s = 0; ph ^= 1u; }
            } else {
                for (int tap = 0; tap < TC_TAPS; ++tap) {
                    mbar_wait(full_bar(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned long long a0 = umma_desc(sA + s * TC_A_BYTES, Cfg::SBO), b0 = b_desc0 + (unsigned long long)((tap * (int)TC_B_TAP_BYTES) >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < TC_KC / 16; ++k)  // 32 bytes further along K inside the swizzle row = +2 in the address field
                            umma_bf16(d_tmem, a0 + 2ull * k, b0 + 2ull * k, (tap | k) != 0 ? 1u : 0u);
                        umma_commit(empty_bar(s));  // the stage is free once these MMAs have read it
                        if (tap == TC_TAPS - 1) umma_commit(tfull_bar(acc));  // accumulator complete
                    }
                    __syncwarp();
                    if (++s == TC_STAGES) { s = 0; ph ^= 1u; }
                }
            }
            if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> bias + ReLU -> bf16 -> swizzled staging tile -> TMA store =====
        const int q = warp & 3;         // this warp's TMEM lane quarter
        const int p = q * 32 + lane;    // pixel of the tile = accumulator row
        const int et = tid - 128;
        int acc = 0, ob = 0;
        unsigned acc_ph = 0;
        // The shared-memory data pipe is this kernel's bottleneck (ncu: 82 % busy with tensor-core operand reads + 17 % with the
        // epilogue's own accesses), so the bias lives in registers (a persistent CTA loads it once: 64 broadcast LDS per thread
        // and tile gone).  direct_store (variant 2) also skips the staging tile and writes each pixel's 128 bytes straight to
        // global memory: measured 8 % SLOWER than staging + TMA store (32 scattered 16-byte pieces per store instruction
        // cost the same L1 data pipe more than 4 shared-memory wavefronts + one bulk read), so it is not the default.
        float breg[TC_N];
#pragma unroll
        for (int c = 0; c < TC_N; ++c) breg[c] = __ldg(bias + c);
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int n = t / tiles_per_image, r = t - n * tiles_per_image;
            const int ty = r / tiles_x, tx = r - ty * tiles_x;
            mbar_wait(tfull_bar(acc), acc_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned v[64];
            const unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(acc * TC_N);
            tmem_ld16(taddr, v);
            tmem_ld16(taddr + 16u, v + 16);
            tmem_ld16(taddr + 32u, v + 32);
            tmem_ld16(taddr + 48u, v + 48);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(tempty_bar(acc));  // the MMA warp may overwrite this accumulator
            if (direct_store) {
                const int oy = ty * TC_TILE_H + p / TC_TILE_W, ox = tx * TC_TILE_W + p % TC_TILE_W;
                if (oy < Hc && ox < Wc) {
                    uint4 *dst = out + (((size_t)n * Hc + oy) * Wc + ox) * (TC_N / 8);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        uint4 o;
                        o.x = bias_relu_pack(v[8 * j + 0], v[8 * j + 1], breg[8 * j + 0], breg[8 * j + 1]);
                        o.y = bias_relu_pack(v[8 * j + 2], v[8 * j + 3], breg[8 * j + 2], breg[8 * j + 3]);
                        o.z = bias_relu_pack(v[8 * j + 4], v[8 * j + 5], breg[8 * j + 4], breg[8 * j + 5]);
                        o.w = bias_relu_pack(v[8 * j + 6], v[8 * j + 7], breg[8 * j + 6], breg[8 * j + 7]);
                        dst[j] = o;
                    }
                }
                if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
                continue;
            }
            // the TMA store that last read staging tile `ob` must have finished reading it
            if (et == 0) {
                if constexpr (Cfg::OUT_BUFS == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const unsigned row = sOut + ob * TC_OUT_BYTES + (unsigned)p * 128u;
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
                tma_store_4d(&map_out, 0, tx * TC_TILE_W, ty * TC_TILE_H, n, sOut + ob * TC_OUT_BYTES);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if constexpr (Cfg::OUT_BUFS == 2) ob ^= 1;
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

// The s2d form of a 7x7 kernel: its 8x8 footprint has an empty last row and column, i.e. sub-pixel blocks with s = 1 (k odd) of
// taps kw = 3 and blocks with r = 1 (k >= 2) of taps kh = 3 are all zero: 49 of the 64 slices remain.
constexpr unsigned long long stem_mask_7x7() {
    unsigned long long m = 0;
    for (int tap = 0; tap < 16; ++tap)
        for (int k = 0; k < 4; ++k)
            if (!(((tap & 3) == 3 && (k & 1)) || ((tap >> 2) == 3 && (k & 2)))) m |= 1ull << (tap * 4 + k);
    return m;
}
constexpr unsigned long long TC_MASK_7X7 = stem_mask_7x7();

template <bool HALO, unsigned long long KMASK>
int launch_stem(hpb_ctx *ctx, const void *z, int b, int Hz, int Wz, int C, const void *w, const float *bias, int O, void *out,
                cudaStream_t stream) {
    using Cfg = StemCfg<HALO>;
    const int Hc = Hz - 3, Wc = Wz - 3;
    if (HALO) {  // partial tiles are fine: TMA zero-fills loads and clips stores at the tensor's edge
        if (Hz < Cfg::BOX_H || Wz < Cfg::BOX_W) return HPB_ENOTFOUND;
    } else if (Hc < Cfg::TILE_H || Wc < Cfg::TILE_W || Hc % Cfg::TILE_H || Wc % Cfg::TILE_W) {
        return HPB_ENOTFOUND;
    }
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return HPB_ENOTFOUND;
    CUtensorMap map_in, map_w, map_out;
    const cuuint32_t estr4[4] = {1u, 1u, 1u, 1u};
    {
        const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wz, (cuuint64_t)Hz, (cuuint64_t)b};
        const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)Wz * C * 2, (cuuint64_t)Hz * Wz * C * 2};
        const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)Cfg::BOX_W, (cuuint32_t)Cfg::BOX_H, 1u};
        if (encode(&map_in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(z), dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return HPB_ENOTFOUND;
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)TC_TAPS * C, (cuuint64_t)O};
        const cuuint64_t strides[1] = {(cuuint64_t)TC_TAPS * C * 2};
        const cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)O};
        const cuuint32_t estr[2] = {1u, 1u};
        if (encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return HPB_ENOTFOUND;
    }
    {
        const cuuint64_t dims[4] = {(cuuint64_t)O, (cuuint64_t)Wc, (cuuint64_t)Hc, (cuuint64_t)b};
        const cuuint64_t strides[3] = {(cuuint64_t)O * 2, (cuuint64_t)Wc * O * 2, (cuuint64_t)Hc * Wc * O * 2};
        const cuuint32_t box[4] = {(cuuint32_t)O, (cuuint32_t)Cfg::TILE_W, (cuuint32_t)Cfg::TILE_H, 1u};
        if (encode(&map_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return HPB_ENOTFOUND;
    }
    // per launch (a microsecond): the attribute is per device, and a process may hold contexts on several
    HPB_CUDA_OK(cudaFuncSetAttribute(hpb_stem_tc_kernel<HALO, KMASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    const int tiles_x = (Wc + Cfg::TILE_W - 1) / Cfg::TILE_W, tiles_per_image = tiles_x * ((Hc + Cfg::TILE_H - 1) / Cfg::TILE_H);
    const long long n_tiles = (long long)tiles_per_image * b;
    if (n_tiles > 0x7fffffffll) return HPB_ENOTFOUND;
    const int grid = (int)(n_tiles < ctx->sm_count ? n_tiles : ctx->sm_count);
    hpb_stem_tc_kernel<HALO, KMASK><<<grid, TC_THREADS, Cfg::SMEM, stream>>>(map_in, map_w, map_out, bias, tiles_x, tiles_per_image, (int)n_tiles,
                                                                                reinterpret_cast<uint4 *>(out), Hc, Wc, ctx->stem_tc_halo == 2 ? 1 : 0);
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}

}  // namespace

// Returns HPB_OK when launched, HPB_ENOTFOUND when the shape is not served (the caller keeps its cuDNN convolution):
// z [b][Hz][Wz][64] bf16, w [64 out][4][4][64] bf16 (= the channels_last [64,64,4,4] weight), bias [64] f32,
// out [b][Hz-3][Wz-3][64] bf16.  ctx->stem_tc_halo selects the operand-feeding scheme (see the top of the file).
int hpb_launch_stem_tc(hpb_ctx *ctx, const void *z, int b, int Hz, int Wz, int C, const void *w, const float *bias, int O, void *out,
                       unsigned long long kmask, cudaStream_t stream) {
    if (C != TC_KC || O != TC_N || b < 1 || Hz < 4 || Wz < 4) return HPB_ENOTFOUND;
    if ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) & 15u) return HPB_ENOTFOUND;
    // a slice the caller's mask clears is all-zero, so multiplying it anyway is still exact: the specialised kernel is used when
    // every slice IT skips is cleared in the caller's mask, the dense one otherwise
    const bool skip7 = (kmask & ~TC_MASK_7X7) == 0ull;
    if (!ctx->stem_tc_halo) return launch_stem<false, ~0ull>(ctx, z, b, Hz, Wz, C, w, bias, O, out, stream);
    return skip7 ? launch_stem<true, TC_MASK_7X7>(ctx, z, b, Hz, Wz, C, w, bias, O, out, stream)
                 : launch_stem<true, ~0ull>(ctx, z, b, Hz, Wz, C, w, bias, O, out, stream);
}
