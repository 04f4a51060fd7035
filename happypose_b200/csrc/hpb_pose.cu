// Small per-hypothesis pose kernels for sm_100a (one launch each; they replace chains of ~10-30 eager torch ops
// and, for the multiview placement, a per-sample CPU loop over Panda3D scene-graph nodes).
//
//   hpb_normalize_T_kernel   toolbox/lib3d/transform_ops.py:107-120 (normalize_T) + rotations.py:22-36 (ortho6d)
//   hpb_pose_update_kernel   megapose/models/pose_rigid.py:339-350 -> toolbox/lib3d/cosypose_ops.py:34-62;
//                            cosypose/lib3d/cosypose_ops.py:18-42 + cosypose/models/pose.py:95-106
//   hpb_tco_init_kernel      toolbox/lib3d/cosypose_ops.py:159-181,184-238,241-283
//   hpb_multiview_kernel     toolbox/lib3d/multiview.py:28-92,166-251 (Panda3D NodePath.lookAt in closed form)
//   hpb_normalize_depth_kernel  megapose/models/pose_rigid.py:510-544
#include "hpb_pose_math.cuh"

namespace {

using namespace hpbm;

__global__ void hpb_normalize_T_kernel(const float *T, int b, float *out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= b) return;
    float o[16];
    normalize_T_row(T + (size_t)n * 16, o);
    for (int k = 0; k < 16; ++k) out[(size_t)n * 16 + k] = o[k];
}

// xyzw quaternion -> R through the angle-axis route of the reference (rotations.py:196-229: quat2mat ->
// quaternion_to_angle_axis -> angle_axis_to_rotation_matrix); algebraically the standard unit-quaternion matrix.
__device__ __forceinline__ void quat_to_R(const float q[4], float R[9]) {
    const float nq = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float x = q[0] / nq, y = q[1] / nq, z = q[2] / nq, w = q[3] / nq;
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - z * w); R[2] = 2.f * (x * z + y * w);
    R[3] = 2.f * (x * y + z * w); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - x * w);
    R[6] = 2.f * (x * z - y * w); R[7] = 2.f * (y * z + x * w); R[8] = 1.f - 2.f * (x * x + y * y);
}

__global__ void hpb_pose_update_kernel(const float *TCO, const float *K_crop, const float *outv, const float *tCR,
                                       int b, int variant, float *TCO_out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= b) return;
    float T[16];
    for (int k = 0; k < 16; ++k) T[k] = TCO[(size_t)n * 16 + k];
    const float fx = K_crop[(size_t)n * 9], fy = K_crop[(size_t)n * 9 + 4];
    float dR[9], vx, vy, vz;
    if (variant == HPB_POSE_COSYPOSE_QUAT) {
        const float *o = outv + (size_t)n * 7;
        const float q[4] = {o[0], o[1], o[2], o[3]};
        quat_to_R(q, dR);
        vx = o[4]; vy = o[5]; vz = o[6];
    } else {
        const float *o = outv + (size_t)n * 9;
        const float a[3] = {o[0], o[1], o[2]}, c[3] = {o[3], o[4], o[5]};
        ortho6d(a, c, dR);
        vx = o[6]; vy = o[7]; vz = o[8];
    }
    float ref[3];  // reference point in the camera frame
    if (variant == HPB_POSE_MEGAPOSE) {
        ref[0] = tCR[(size_t)n * 3]; ref[1] = tCR[(size_t)n * 3 + 1]; ref[2] = tCR[(size_t)n * 3 + 2];
    } else {
        ref[0] = T[3]; ref[1] = T[7]; ref[2] = T[11];
    }
    const float zsrc = ref[2];
    const float ztgt = vz * zsrc;
    const float rx = (vx / fx + ref[0] / zsrc) * ztgt;
    const float ry = (vy / fy + ref[1] / zsrc) * ztgt;
    float tn[3];
    if (variant == HPB_POSE_MEGAPOSE) {
        const float d0 = T[3] - ref[0], d1 = T[7] - ref[1], d2 = T[11] - ref[2];
        tn[0] = dR[0] * d0 + dR[1] * d1 + dR[2] * d2 + rx;
        tn[1] = dR[3] * d0 + dR[4] * d1 + dR[5] * d2 + ry;
        tn[2] = dR[6] * d0 + dR[7] * d1 + dR[8] * d2 + ztgt;
    } else {
        tn[0] = rx; tn[1] = ry; tn[2] = ztgt;
    }
    float *o = TCO_out + (size_t)n * 16;
    float Rn[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            Rn[i * 3 + j] = dR[i * 3] * T[j] + dR[i * 3 + 1] * T[4 + j] + dR[i * 3 + 2] * T[8 + j];
    o[0] = Rn[0]; o[1] = Rn[1]; o[2] = Rn[2]; o[3] = tn[0];
    o[4] = Rn[3]; o[5] = Rn[4]; o[6] = Rn[5]; o[7] = tn[1];
    o[8] = Rn[6]; o[9] = Rn[7]; o[10] = Rn[8]; o[11] = tn[2];
    o[12] = T[12]; o[13] = T[13]; o[14] = T[14]; o[15] = T[15];  // TCO.clone(): last row carried over
}

struct TcoInitParams {
    int variant;
    const float *boxes, *points;
    const int32_t *obj_ids;
    const float *K, *R;
    float z_mean;
    int n_pts, b;
    float *out;
};

__global__ void __launch_bounds__(256) hpb_tco_init_kernel(const TcoInitParams p) {
    const int n = blockIdx.x, tid = threadIdx.x;
    __shared__ float red[4][8];
    const float *K = p.K + (size_t)n * 9;
    const float *bx = p.boxes + (size_t)n * 4;
    const float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const float ucx = (bx[0] + bx[2]) / 2.0f, ucy = (bx[1] + bx[3]) / 2.0f;
    float R[9] = {0.f, 1.f, 0.f, 0.f, 0.f, -1.f, -1.f, 0.f, 0.f};
    if (p.variant == HPB_TCO_INIT_AUTODEPTH_WITH_R)
        for (int k = 0; k < 9; ++k) R[k] = p.R[(size_t)n * 9 + k];
    float *o = p.out + (size_t)n * 16;
    if (p.variant == HPB_TCO_INIT_FROM_BOXES) {
        if (tid == 0) {
            const float z = p.z_mean;
            o[0] = 1.f; o[1] = 0.f; o[2] = 0.f; o[3] = ((ucx - cx) * z) / fx;
            o[4] = 0.f; o[5] = 1.f; o[6] = 0.f; o[7] = ((ucy - cy) * z) / fy;
            o[8] = 0.f; o[9] = 0.f; o[10] = 1.f; o[11] = z;
            o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
        }
        return;
    }
    // z_guess = 1: translation (tx, ty, 1); only the x/y extents of the transformed points are needed
    const float tx = ((ucx - cx) * 1.0f) / fx, ty = ((ucy - cy) * 1.0f) / fy;
    const float *pts = p.points + (size_t)p.obj_ids[n] * p.n_pts * 3;
    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int i = tid; i < p.n_pts; i += blockDim.x) {
        const float x = __ldg(pts + 3 * i), y = __ldg(pts + 3 * i + 1), z = __ldg(pts + 3 * i + 2);
        const float X = R[0] * x + R[1] * y + R[2] * z + tx;
        const float Y = R[3] * x + R[4] * y + R[5] * z + ty;
        mnx = fminf(mnx, X); mxx = fmaxf(mxx, X);
        mny = fminf(mny, Y); mxy = fmaxf(mxy, Y);
    }
    for (int s = 16; s > 0; s >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, s));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, s));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, s));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, s));
    }
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = mnx; red[1][tid >> 5] = mny; red[2][tid >> 5] = mxx; red[3][tid >> 5] = mxy;
    }
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
            mnx = fminf(mnx, red[0][k]); mny = fminf(mny, red[1][k]);
            mxx = fmaxf(mxx, red[2][k]); mxy = fmaxf(mxy, red[3][k]);
        }
        const float dx3 = mxx - mnx, dy3 = mxy - mny;
        const float bdx = (bx[2] - bx[0]) + 1.0f, bdy = (bx[3] - bx[1]) + 1.0f;
        const float z = ((fy * dy3 / bdy) + (fx * dx3 / bdx)) / 2.0f;
        o[0] = R[0]; o[1] = R[1]; o[2] = R[2]; o[3] = ((ucx - cx) * z) / fx;
        o[4] = R[3]; o[5] = R[4]; o[6] = R[5]; o[7] = ((ucy - cy) * z) / fy;
        o[8] = R[6]; o[9] = R[7]; o[10] = R[8]; o[11] = z;
        o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
    }
}

__global__ void hpb_multiview_kernel(const float *TCO, const float *tCR, int b, int n_extra, int n_views,
                                     int keep_tco, const MvPositions mv, float *out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= b) return;
    const float *Tf = TCO + (size_t)n * 16;
    float *o = out + (size_t)n * n_views * 16;
    int v0 = 0;
    if (keep_tco) {
        for (int k = 0; k < 16; ++k) o[k] = Tf[k];
        v0 = 1;
    }
    if (n_extra == 0) return;
    MvFrame F;
    multiview_frame(Tf, tCR + (size_t)n * 3, F);
    for (int e = 0; e < n_extra; ++e) multiview_view(F, Tf, mv.p[3 * e], mv.p[3 * e + 1], mv.p[3 * e + 2], o + (size_t)(v0 + e) * 16);
}

__global__ void hpb_normalize_depth_kernel(float *depth, long long bstride, int c0, int c1, int c2, int c3, int c4,
                                           int c5, int c6, int c7, int n_planes, const float *tCR, int b, int hw,
                                           int kind) {
    const int chans[8] = {c0, c1, c2, c3, c4, c5, c6, c7};
    const long long total = (long long)b * n_planes * hw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int pix = (int)(i % hw);
        const long long r = i / hw;
        const int pl = (int)(r % n_planes), n = (int)(r / n_planes);
        float *d = depth + (size_t)n * bstride + (size_t)chans[pl] * hw + pix;
        const float z = tCR[(size_t)n * 3 + 2];
        float v = *d;
        if (kind == HPB_DEPTH_NORM_TCR_SCALE) v = v / z;
        else if (kind == HPB_DEPTH_NORM_TCR_SCALE_CLAMP_CENTER) v = fminf(fmaxf(v / z, 0.0f), 2.0f) - 1.0f;
        else if (kind == HPB_DEPTH_NORM_TCR_CENTER_CLAMP) v = fminf(fmaxf(v - z, -2.0f), 2.0f);
        *d = v;
    }
}

}  // namespace

int hpb_launch_normalize_T(hpb_ctx *ctx, const float *T, int b, float *out, cudaStream_t stream) {
    if (b == 0) return HPB_OK;
    hpb_normalize_T_kernel<<<(b + 127) / 128, 128, 0, stream>>>(T, b, out);
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}

int hpb_launch_pose_update(hpb_ctx *ctx, const float *TCO, const float *K_crop, const float *outv, const float *tCR,
                           int b, int variant, float *TCO_out, cudaStream_t stream) {
    if (b == 0) return HPB_OK;
    hpb_pose_update_kernel<<<(b + 127) / 128, 128, 0, stream>>>(TCO, K_crop, outv, tCR, b, variant, TCO_out);
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}

int hpb_launch_tco_init(hpb_ctx *ctx, int variant, const float *boxes, const float *points, int n_pts,
                        const int32_t *obj_ids, const float *K, const float *R, float z_mean, int b, float *out,
                        cudaStream_t stream) {
    if (b == 0) return HPB_OK;
    TcoInitParams p;
    p.variant = variant; p.boxes = boxes; p.points = points; p.obj_ids = obj_ids; p.K = K; p.R = R;
    p.z_mean = z_mean; p.n_pts = n_pts; p.b = b; p.out = out;
    hpb_tco_init_kernel<<<b, 256, 0, stream>>>(p);
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}

int hpb_launch_multiview(hpb_ctx *ctx, const float *TCO, const float *tCR, int b, const float *positions_host,
                         int n_extra, int n_views, int keep_tco, float *out, cudaStream_t stream) {
    if (b == 0) return HPB_OK;
    MvPositions mv;
    for (int i = 0; i < 26 * 3; ++i) mv.p[i] = i < 3 * n_extra ? positions_host[i] : 0.0f;
    hpb_multiview_kernel<<<(b + 63) / 64, 64, 0, stream>>>(TCO, tCR, b, n_extra, n_views, keep_tco, mv, out);
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}

int hpb_launch_normalize_depth(hpb_ctx *ctx, float *depth, int64_t bstride, const int32_t *chans, int n_planes,
                               const float *tCR, int b, int h, int w, int kind, cudaStream_t stream) {
    if (b == 0 || n_planes == 0 || kind == HPB_DEPTH_NORM_NONE) return HPB_OK;
    int c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n_planes; ++i) c[i] = chans[i];
    const long long total = (long long)b * n_planes * h * w;
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    hpb_normalize_depth_kernel<<<blocks, 256, 0, stream>>>(depth, bstride, c[0], c[1], c[2], c[3], c[4], c[5], c[6],
                                                           c[7], n_planes, tCR, b, h * w, kind);
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Network-input packing: float32 planar [b,C,h,w] -> bfloat16 pixel-interleaved [b,h,w,Cp] (torch channels_last), the
// layout cuDNN's tensor-core convolutions consume.  Channels C..Cp-1 are zero.  One thread per pixel: C coalesced plane
// reads, Cp*2 contiguous bytes written as 16-byte stores.  Rounding = round-to-nearest-even (== torch .to(bfloat16)).
// Replaces torch's x.to(bfloat16, channels_last) pass plus cuDNN's own channel-padding pass in front of the stem.
// ------------------------------------------------------------------------------------------------------------------
#include <cuda_bf16.h>

namespace {

template <int CP>
__global__ void __launch_bounds__(256) hpb_pack_input_kernel(const float *x, long long bstride, int C, long long npix, long long total,
                                                              uint4 *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long n = i / npix, px = i - n * npix;
    const float *src = x + n * bstride + px;
    __nv_bfloat16 v[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) v[c] = __float2bfloat16_rn(c < C ? __ldcs(src + (long long)c * npix) : 0.0f);
    uint4 *dst = out + i * (CP / 8);
#pragma unroll
    for (int q = 0; q < CP / 8; ++q) {
        uint4 w;
        w.x = (unsigned)__bfloat16_as_ushort(v[8 * q]) | ((unsigned)__bfloat16_as_ushort(v[8 * q + 1]) << 16);
        w.y = (unsigned)__bfloat16_as_ushort(v[8 * q + 2]) | ((unsigned)__bfloat16_as_ushort(v[8 * q + 3]) << 16);
        w.z = (unsigned)__bfloat16_as_ushort(v[8 * q + 4]) | ((unsigned)__bfloat16_as_ushort(v[8 * q + 5]) << 16);
        w.w = (unsigned)__bfloat16_as_ushort(v[8 * q + 6]) | ((unsigned)__bfloat16_as_ushort(v[8 * q + 7]) << 16);
        dst[q] = w;
    }
}

}  // namespace

int hpb_launch_pack_input(hpb_ctx *ctx, const float *x, int64_t bstride, int b, int C, int h, int w, void *out, int Cp,
                          cudaStream_t stream) {
    if (b == 0) return HPB_OK;
    const long long npix = (long long)h * w, total = npix * b;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    uint4 *o = reinterpret_cast<uint4 *>(out);
    switch (Cp) {
        case 8: hpb_pack_input_kernel<8><<<blocks, 256, 0, stream>>>(x, bstride, C, npix, total, o); break;
        case 16: hpb_pack_input_kernel<16><<<blocks, 256, 0, stream>>>(x, bstride, C, npix, total, o); break;
        case 24: hpb_pack_input_kernel<24><<<blocks, 256, 0, stream>>>(x, bstride, C, npix, total, o); break;
        case 32: hpb_pack_input_kernel<32><<<blocks, 256, 0, stream>>>(x, bstride, C, npix, total, o); break;
        case 40: hpb_pack_input_kernel<40><<<blocks, 256, 0, stream>>>(x, bstride, C, npix, total, o); break;
        default:
            hpb_set_error("hpb_pack_input_bf16: padded channel count %d not in {8,16,24,32,40}", Cp);
            return HPB_EINVAL;
    }
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// 3x3 / stride 2 / pad 1 max-pool on bfloat16 pixel-interleaved (NHWC) activations: the ResNet stem's nn.MaxPool2d(3, 2, 1)
// (torchvision_resnet.py:215).  Pure streaming op (read 4x what it writes); one thread = 8 channels (one 16-byte vector)
// of one output pixel, 9 vector loads, packed bf16x2 max (NaN-propagating like torch).  Max is exact, so the result is
// bit-identical to torch's kernel.
// ------------------------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
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

__global__ void __launch_bounds__(256) hpb_maxpool_kernel(const uint4 *in, int H, int W, int C8, int Ho, int Wo, long long total,
                                                           uint4 *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C8);
    long long t = i / C8;
    const int ow = (int)(t % Wo);
    t /= Wo;
    const int oh = (int)(t % Ho);
    const long long n = t / Ho;
    const uint4 *base = in + n * H * W * C8 + c;
    const unsigned NEG = 0xff80ff80u;  // bf16 -inf pair
    uint4 m = make_uint4(NEG, NEG, NEG, NEG);
    const int h0 = 2 * oh - 1, w0 = 2 * ow - 1;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int y = h0 + dy;
        if (y < 0 || y >= H) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int x = w0 + dx;
            if (x < 0 || x >= W) continue;
            m = bf16x8_max(m, __ldg(base + ((long long)y * W + x) * C8));
        }
    }
    out[i] = m;
}

}  // namespace

int hpb_launch_maxpool(hpb_ctx *ctx, const void *in, int b, int H, int W, int C, void *out, cudaStream_t stream) {
    if (b == 0) return HPB_OK;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1, C8 = C / 8;
    const long long total = (long long)b * Ho * Wo * C8;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    hpb_maxpool_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(in), H, W, C8, Ho, Wo, total,
                                                   reinterpret_cast<uint4 *>(out));
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Space-to-depth network-input packing for the ResNet stem.  A 7x7 / stride 2 / pad 3 convolution over C channels equals
// a 4x4 / stride 1 / no-pad convolution over the channels of z, where z[n, I, J, (r*2+s)*Cs + c] = xpad[n, c, 2I+r, 2J+s]
// (Cs = Cz / 4 channels reserved per sub-pixel, c < C used, the rest zero: every sub-pixel block starts 16-byte aligned)
// and xpad is x zero-padded by 3 pixels (w'[o,(r,s,c),a,b] = w[o,c,2a+r,2b+s], zero for tap index 7).  The stride-1 form
// has a 4x deeper reduction per tap and runs several times faster on the tensor cores than cuDNN's strided 9/27-channel
// stem.  This kernel builds z (bfloat16, pixel-interleaved, channels padded to Cz with zeros) straight from the float32
// planar network input the rasteriser / crop kernels wrote: one thread = 8 consecutive channels of one z pixel.
// ------------------------------------------------------------------------------------------------------------------
namespace {

constexpr int S2D_THREADS = 256;
constexpr int S2D_MARGIN = 4;  // zero columns left of x = 0 in the staged rows (needs >= 3; 4 keeps 8-byte stores aligned)

// One CTA = one z row (n, I): the two x rows it draws from (y = 2I-3, 2I-2) are staged in shared memory as bfloat16 for
// all C channels with coalesced 16-byte loads (each x element is read from HBM exactly once over the whole launch), with
// zero margins so that the padding of the convolution needs no bounds checks; then every thread emits 16-byte vectors of
// 8 consecutive z channels, consecutive threads writing consecutive addresses.
//   staged row index  (r * C + c), plus one all-zero row for the padded channels k >= 4C
//   staged column     xx + S2D_MARGIN for xx in [-S2D_MARGIN, W + 4)
__global__ void __launch_bounds__(S2D_THREADS) hpb_pack_s2d_kernel(const float *x, long long bstride, int C, int H, int W, int Hz, int Wz,
                                                                   int Cz8, int Wp, int Cs, unsigned cs_magic, uint4 *out) {
    extern __shared__ __align__(16) unsigned short s_rows[];  // [(2C + 1)][Wp]
    const int I = blockIdx.x;
    const long long n = blockIdx.y;
    const int tid = threadIdx.x;
    const float *src = x + n * bstride;
    const long long plane = (long long)H * W;
    const int n_rows = 2 * C;
    // ---- zero the margins and the zero row ----
    {
        const int right0 = W + S2D_MARGIN, nright = Wp - right0;
        const int per_row = S2D_MARGIN + nright;
        for (int i = tid; i < n_rows * per_row; i += S2D_THREADS) {
            const int row = i / per_row, j = i - row * per_row;
            s_rows[row * Wp + (j < S2D_MARGIN ? j : right0 + (j - S2D_MARGIN))] = 0;
        }
        for (int i = tid; i < Wp; i += S2D_THREADS) s_rows[n_rows * Wp + i] = 0;
    }
    // ---- stage the two source rows of every channel (float32 -> bfloat16, round to nearest even) ----
    const bool vec = (W % 4 == 0) && (bstride % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    if (vec) {
        const int W4 = W >> 2;
        for (int i = tid; i < n_rows * W4; i += S2D_THREADS) {
            const int row = i / W4, j = i - row * W4;
            const int r = row / C, c = row - r * C;
            const int y = 2 * I + r - 3;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < H) v = __ldcs(reinterpret_cast<const float4 *>(src + c * plane + (long long)y * W) + j);
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 w2;
            w2.x = *reinterpret_cast<const unsigned *>(&lo);
            w2.y = *reinterpret_cast<const unsigned *>(&hi);
            *reinterpret_cast<uint2 *>(s_rows + row * Wp + S2D_MARGIN + 4 * j) = w2;
        }
    } else {
        for (int i = tid; i < n_rows * W; i += S2D_THREADS) {
            const int row = i / W, j = i - row * W;
            const int r = row / C, c = row - r * C;
            const int y = 2 * I + r - 3;
            const float f = (y >= 0 && y < H) ? __ldcs(src + c * plane + (long long)y * W + j) : 0.0f;
            s_rows[row * Wp + S2D_MARGIN + j] = __bfloat16_as_ushort(__float2bfloat16_rn(f));
        }
    }
    __syncthreads();
    // ---- emit: thread -> fixed channel octet q, strided over J ----
    const int JT = S2D_THREADS / Cz8;  // z pixels per pass (launcher guarantees Cz8 <= S2D_THREADS)
    const int q = tid % Cz8, j0 = tid / Cz8;
    if (j0 >= JT) return;
    int off[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = 8 * q + e;
        const int rs = (int)(((unsigned)k * cs_magic) >> 16);  // k / Cs: sub-pixel (r,s) = rs, channel c within its block of Cs
        const int c = k - rs * Cs;
        off[e] = (rs < 4 && c < C) ? ((rs >> 1) * C + c) * Wp + (rs & 1) - 3 + S2D_MARGIN : n_rows * Wp;
    }
    uint4 *dst = out + ((n * Hz + I) * Wz) * Cz8 + q;
    for (int J = j0; J < Wz; J += JT) {
        unsigned short v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = s_rows[off[e] + 2 * J];  // 2J <= W + 4 < Wp also inside the zero row
        uint4 w;
        w.x = (unsigned)v[0] | ((unsigned)v[1] << 16);
        w.y = (unsigned)v[2] | ((unsigned)v[3] << 16);
        w.z = (unsigned)v[4] | ((unsigned)v[5] << 16);
        w.w = (unsigned)v[6] | ((unsigned)v[7] << 16);
        __stcs(dst + (long long)J * Cz8, w);
    }
}

}  // namespace

int hpb_launch_pack_s2d(hpb_ctx *ctx, const float *x, int64_t bstride, int b, int C, int H, int W, void *out, int Cz,
                        cudaStream_t stream) {
    if (b == 0) return HPB_OK;
    const int Hz = H / 2 + 3, Wz = W / 2 + 3, Cz8 = Cz / 8;
    if (Cz8 > S2D_THREADS || b > 65535) {
        hpb_set_error("hpb_pack_input_s2d_bf16: C_padded %d or batch %d too large", Cz, b);
        return HPB_EINVAL;
    }
    // staged row pitch (bf16 elements): W + margins, a multiple of 4 (8-byte stores) with an odd number of 8-byte units so that
    // consecutive channel rows start in different shared-memory banks
    int Wp = (W + S2D_MARGIN + 7 + 3) & ~3;
    if (((Wp / 4) & 1) == 0) Wp += 4;
    const size_t smem = (size_t)(2 * C + 1) * Wp * sizeof(unsigned short);
    if (smem > (size_t)ctx->max_smem_optin) {
        hpb_set_error("hpb_pack_input_s2d_bf16: %d channels x width %d do not fit in shared memory", C, W);
        return HPB_EINVAL;
    }
    HPB_CUDA_OK(cudaFuncSetAttribute(hpb_pack_s2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int Cs = Cz / 4;  // channels reserved per sub-pixel (r,s): z channel (r*2+s)*Cs + c
    const unsigned cs_magic = 65535u / (unsigned)Cs + 1u;
    dim3 grid(Hz, b);
    hpb_pack_s2d_kernel<<<grid, S2D_THREADS, smem, stream>>>(x, bstride, C, H, W, Hz, Wz, Cz8, Wp, Cs, cs_magic, reinterpret_cast<uint4 *>(out));
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}
