// 3x3 / stride 2 / pad 1 max-pool on bfloat16 NHWC activations with the input tile staged by TMA.
//
// The ResNet stem's nn.MaxPool2d(3, 2, 1) (torchvision_resnet.py:215) reads a 1.4 GB tensor per 576-row coarse batch and
// writes a quarter of it: pure streaming, bound by HBM.  The first version (hpb_maxpool_kernel) issues 9 overlapping
// 16-byte loads per output and leans on L1 for the 2.25x window overlap; it reached 4.2 TB/s.  Here one CTA owns a
// TW x TH tile of outputs: ONE elected thread asks the TMA unit for the (2 TW + 1) x (2 TH + 1) x C input box
// (cp.async.bulk.tensor.4d, completion on an mbarrier), every input element crosses L2 -> SM once, and the 9 taps come from
// shared memory.  Out-of-range taps (the -inf padding of the reference) are masked by index, so the TMA's zero fill never
// reaches a result and the output is bit-identical to torch.nn.functional.max_pool2d for ANY input, not only after a ReLU.
#include <cuda.h>
#include <cuda_bf16.h>

#include "hpb_common.cuh"

namespace {

constexpr int MP_TW = 16, MP_TH = 8;                 // output tile
constexpr int MP_BW = 2 * MP_TW + 1, MP_BH = 2 * MP_TH + 1;  // input box (pixels)

__device__ __forceinline__ uint4 bf16x8_max_nan(uint4 a, uint4 b) {
    uint4 r;
    __nv_bfloat162 x, y, z;
#define HPB_MAX_LANE(f)                                    \
    x = *reinterpret_cast<__nv_bfloat162 *>(&a.f);          \
    y = *reinterpret_cast<__nv_bfloat162 *>(&b.f);          \
    z = __hmax2_nan(x, y);                                  \
    r.f = *reinterpret_cast<unsigned *>(&z);
    HPB_MAX_LANE(x) HPB_MAX_LANE(y) HPB_MAX_LANE(z) HPB_MAX_LANE(w)
#undef HPB_MAX_LANE
    return r;
}

// One thread = 8 channels (one 16-byte vector) of one output pixel; blockDim.x = MP_TW * MP_TH * C8.
__global__ void __launch_bounds__(1024) hpb_maxpool_tma_kernel(const __grid_constant__ CUtensorMap tmap, int H, int W, int C8, int Ho, int Wo,
                                                               uint4 *out) {
    extern __shared__ __align__(128) unsigned char smem[];  // [MP_BH][MP_BW][C8] uint4, dense (no swizzle)
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x;
    const int ow0 = blockIdx.x * MP_TW, oh0 = blockIdx.y * MP_TH, n = blockIdx.z;
    const int w0 = 2 * ow0 - 1, h0 = 2 * oh0 - 1;  // input coordinates of the box origin (may be -1: the padding column / row)
    const unsigned bar_addr = (unsigned)__cvta_generic_to_shared(&bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned bytes = (unsigned)(MP_BH * MP_BW * C8 * 16);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(&tmap), "r"(0), "r"(w0), "r"(h0), "r"(n), "r"(bar_addr)
            : "memory");
    }
    // everybody waits for the bytes to land (phase 0 of a freshly initialised barrier)
    {
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n"
                : "=r"(done)
                : "r"(bar_addr), "r"(0)
                : "memory");
        }
    }
    const int c = tid % C8;
    const int t = tid / C8;
    const int tx = t % MP_TW, ty = t / MP_TW;
    const int ow = ow0 + tx, oh = oh0 + ty;
    if (ow >= Wo || oh >= Ho) return;
    const uint4 *tile = reinterpret_cast<const uint4 *>(smem);
    const unsigned NEG = 0xff80ff80u;  // bf16 -inf pair
    uint4 m = make_uint4(NEG, NEG, NEG, NEG);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int y = 2 * oh - 1 + dy;
        if (y < 0 || y >= H) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int x = 2 * ow - 1 + dx;
            if (x < 0 || x >= W) continue;
            m = bf16x8_max_nan(m, tile[((2 * ty + dy) * MP_BW + (2 * tx + dx)) * C8 + c]);
        }
    }
    __stcs(out + (((size_t)n * Ho + oh) * Wo + ow) * C8 + c, m);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

}  // namespace

// Returns HPB_OK when launched, HPB_ENOTFOUND when this shape is not served by the TMA kernel (the caller then uses the
// plain kernel): C must give exactly 1024 threads per tile group (C = 64) and the box must fit shared memory.
int hpb_launch_maxpool_tma(hpb_ctx *ctx, const void *in, int b, int H, int W, int C, void *out, cudaStream_t stream) {
    const int C8 = C / 8;
    if (C != 64 || MP_TW * MP_TH * C8 != 1024 || b > 65535) return HPB_ENOTFOUND;
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) return HPB_ENOTFOUND;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    CUtensorMap tmap;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)b};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};  // bytes, dims 1..3
    const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)MP_BW, (cuuint32_t)MP_BH, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(in), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return HPB_ENOTFOUND;
    const size_t smem = (size_t)MP_BH * MP_BW * C8 * 16;
    static bool attr_set = false;
    if (!attr_set) {
        HPB_CUDA_OK(cudaFuncSetAttribute(hpb_maxpool_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    dim3 grid((Wo + MP_TW - 1) / MP_TW, (Ho + MP_TH - 1) / MP_TH, b);
    hpb_maxpool_tma_kernel<<<grid, MP_TW * MP_TH * C8, smem, stream>>>(tmap, H, W, C8, Ho, Wo, reinterpret_cast<uint4 *>(out));
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}
