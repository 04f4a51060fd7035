// Batched triangle rasteriser for sm_100a: N single-object scenes -> RGB / normals / depth / mask in ONE launch.
//
// Replaces happypose/toolbox/renderer/panda3d_batch_renderer.py:194-286 (Panda3dBatchRenderer.render) and the
// Panda3D/OpenGL pipeline behind it (panda3d_scene_renderer.py:320-390, renderer/types.py:111-137,254-299,
// renderer/utils.py:46-79).  The semantics (pixel centres, near/far, two-sided, depth read-back rule, ambient-only
// colour, eye-normal 32-level wrap texture) are stated once in DESIGN.md; every float operation below is an
// explicit IEEE op / fmaf so results are bit-identical to the CPU oracle used by the tests (the library is built
// with -fmad=false).
//
// The workload is micro-polygon rendering: a 15 728-triangle mesh lands on ~19 000 pixels of a 320x240 view, i.e. a
// triangle covers 1-3 pixels and half of the triangles face away.  The kernel is therefore triangle-parallel with
// per-warp work compaction, and instruction-issue bound (not DRAM bound) by design; see DESIGN.md section 4.
//
// Structure: a persistent grid of thread-block CLUSTERS; each cluster (G = 1, 2, 4 or 8 CTAs, one CTA per SM) loops
// over scenes.  G > 1 is chosen for small batches so that one scene is spread over several SMs (the refiner renders
// only 4..64 scenes per launch); for b >= #SMs G = 1 and each SM renders whole scenes.
//   phase A  vertex stage (every CTA of the cluster, redundantly): object -> camera -> 24.8 fixed-point screen
//            position + 1/z, staged in SHARED memory (12 B per vertex; meshes that do not fit fall back to a per-CTA
//            global scratch slice).  The screen bounding box of the scene is reduced on the way.
//   phase B  triangle stage: the cluster's CTAs split the triangle list.
//            pass 1 (one lane per triangle, cheap): bounding box, signed area; triangles without a pixel centre,
//            degenerate ones and -- on closed meshes -- back faces are dropped; the survivors' ids are COMPACTED into
//            a per-warp shared-memory queue.
//            pass 2 (whenever 32 ids are queued): full set-up, exact integer edge functions with a top-left rule, a
//            per-triangle depth PLANE d(col,row); every lane walks its own bounding box in two warp-convergent loops
//            (coverage mask, then one covered pixel per iteration).  Depth test = 64-bit atomicMin of
//            (depth bits << 32 | triangle id << 1 | wide) on a per-cluster visibility buffer that stays resident in L2
//            (re-armed in phase C, never re-cleared).  Big triangles are walked by the whole warp.
//   phase C  resolve: the cluster's warps split the image into 32-pixel span units; one lane per pixel, coalesced
//            planar stores.  Spans that miss the scene's bounding box are zero-filled without reading the visibility
//            buffer; attributes are interpolated perspective-correctly from the un-normalised edge values (no area
//            division), texture is trilinearly filtered from an RGBA8 mip chain.  Background pixels are written as zeros
//            here, so the outputs need no separate clear pass.
//            Fused hand-off (hpb_render_s2d_bf16, template S2D = true): the same shading, but 4 lanes own one 2x2-pixel
//            cell of the ResNet stem's space-to-depth input and the warp writes that bf16 NHWC tensor directly (crop
//            channels read from the crop kernel's planes) -- the float32 network input and the packing pass disappear.
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "hpb_common.cuh"

namespace cg = cooperative_groups;

namespace {

#ifndef HPB_RASTER_THREADS
#define HPB_RASTER_THREADS 1024
#endif
constexpr int RASTER_THREADS = HPB_RASTER_THREADS;
constexpr int RASTER_WARPS = RASTER_THREADS / 32;
constexpr int SMALL_TRI_MAX = 32;  // bbox pixels a lane walks on its own; larger triangles are walked by the warp
constexpr int VTX_CLIPPED = (int)0x80000000;  // screen x/y of a vertex in front of the near plane

struct RasterParams {
    const HpbMeshDev *meshes;
    const int32_t *mesh_ids;
    const float *TCO;
    const float *K;
    const float *ambient;
    const float *lights;  // [b][n_lights][8]: (type 0 point / 1 directional, x, y, z in the OBJECT = world frame, r, g, b, -) or nullptr
    int n_lights;
    int b, h, w;
    float z_near;
    float inv_near, cd, a_f, b_f, eps_hi;
    uint32_t flags;
    float *rgb;
    float *nrm;
    float *depth;
    uint8_t *mask;
    long long rgb_bs, nrm_bs, depth_bs, mask_bs;
    int views;            // scene i writes at base + (i / views) * bstride + (i % views) * view_stride
    long long view_stride;
    unsigned long long *clipped_scenes;  // device counter: scenes in which the near plane cut the mesh (those triangles are dropped)
    unsigned long long *vis;      // [clusters][h*w]
    unsigned char *vert_scratch;  // [CTAs][max_nv * 12]: screen-space vertices of the meshes that do not fit in shared memory
    int max_nv;                   // largest (even-padded) vertex count of any uploaded mesh = scratch slice size / 12
    int smem_verts;               // vertices the dynamic shared memory holds; bigger meshes use the CTA's scratch slice
    int G;  // CTAs per cluster
    unsigned span_magic;  // floor(2^32 / ceil(w / 32)) + 1
    // space-to-depth output mode (hpb_render_s2d_bf16): the resolve writes the stem's bf16 NHWC input directly
    uint4 *s2d;           // [b][Hz][Wz][Cz8] uint4 (8 bf16 channels each) or nullptr
    const float *crops;   // [b][3][h][w] float32: the crop channels of the network input (crops_fmt 0)
    const uint2 *crops_h; // [b][h][w] (r,g,b,0) bfloat16, pixel-interleaved: hpb_crop_bf16x4's output (crops_fmt 1)
    long long crops_bs;   // batch stride in elements of the respective format
    int crops_fmt;
    int pad_prezeroed;    // the padding channels (>= 16 of every sub-pixel block) of the output are already zero (persistent buffer)
    int Hz, Wz, Cz8;      // Hz = h/2 + 3, Wz = w/2 + 3
    unsigned wz_magic;    // floor(2^32 / Wz) + 1
};

__device__ __forceinline__ int snap_fixed(float u) {
    float s = rintf(u * (float)HPB_SUBPIX);
    if (!(s > -HPB_GUARD)) s = -HPB_GUARD;  // also catches NaN
    if (s > HPB_GUARD) s = HPB_GUARD;
    return (int)s;
}

__device__ __forceinline__ int ceil_div_pix(int v) { return (v - 128 + 255) >> HPB_SUBPIX_BITS; }
__device__ __forceinline__ int floor_div_pix(int v) { return (v - 128) >> HPB_SUBPIX_BITS; }
__device__ __forceinline__ int edge_bias(int dx, int dy) { return (dy < 0 || (dy == 0 && dx > 0)) ? 0 : -1; }
__device__ __forceinline__ int min3i(int a, int b, int c) { return min(a, min(b, c)); }
__device__ __forceinline__ int max3i(int a, int b, int c) { return max(a, max(b, c)); }

__device__ __forceinline__ void put_fragment(float d, unsigned lo, unsigned long long *slot) {
    if (d <= 1.0f) {
        if (d < 0.0f) d = 0.0f;
        atomicMin(slot, ((unsigned long long)__float_as_uint(d) << 32) | lo);
    }
}

// Depth plane of a triangle oriented so that area2 > 0, anchored at the pixel (jx0, jy0) of its clamped bounding box:
//   d(col,row) = fmaf(Dx, col, fmaf(Dy, row, Dc)), the GL window depth (1/near - 1/z) / (1/near - 1/far).
// e?o are the (unbiased) edge values at the anchor pixel; (x?,y?) the oriented fixed-point vertices.
__device__ __forceinline__ void depth_plane(const RasterParams &p, float e0o, float e1o, float e2o, float area2f, int x0, int y0, int x1,
                                            int y1, int x2, int y2, float iz0, float iz1, float iz2, float &Dc, float &Dx,
                                            float &Dy) {
    const float inv = __frcp_rn(area2f);
    const float izc = fmaf(e2o * inv, iz2, fmaf(e1o * inv, iz1, (e0o * inv) * iz0));
    const float d1 = iz1 - iz0, d2 = iz2 - iz0;
    // per-pixel steps of the edge functions e1, e2: d/dx = -(edge dy) * 256, d/dy = (edge dx) * 256
    const float sx1 = (float)(y2 - y0) * 256.0f, sx2 = (float)(y0 - y1) * 256.0f;
    const float sy1 = (float)(x0 - x2) * 256.0f, sy2 = (float)(x1 - x0) * 256.0f;
    const float gx = fmaf(sx2, d2, sx1 * d1) * inv;
    const float gy = fmaf(sy2, d2, sy1 * d1) * inv;
    Dc = (p.inv_near - izc) * p.cd;
    Dx = -(gx * p.cd);
    Dy = -(gy * p.cd);
}

// Whole-warp walk of one (big) triangle's bounding box with 64-bit edge functions; all lanes hold the same triangle.
__device__ __forceinline__ void raster_tri_warp(const RasterParams &p, int x0, int y0, int x1, int y1, int x2, int y2,
                                                float Dc, float Dx, float Dy, int jx0, int jx1, int jy0, int jy1,
                                                unsigned lo, unsigned long long *vis, int lane) {
    const int b0 = edge_bias(x2 - x1, y2 - y1), b1 = edge_bias(x0 - x2, y0 - y2), b2 = edge_bias(x1 - x0, y1 - y0);
    const int roww = jx1 - jx0 + 1, total = roww * (jy1 - jy0 + 1);
    // lanes cover 32 consecutive pixels of the row-major bounding box at a time
    for (int idx = lane; idx < total; idx += 32) {
        const int row = idx / roww, col = idx - row * roww;
        const int px = jx0 + col, py = jy0 + row;
        const int fxp = px * HPB_SUBPIX + 128, fyp = py * HPB_SUBPIX + 128;
        const long long e0 = (long long)(x2 - x1) * (fyp - y1) - (long long)(y2 - y1) * (fxp - x1);
        const long long e1 = (long long)(x0 - x2) * (fyp - y2) - (long long)(y0 - y2) * (fxp - x2);
        const long long e2 = (long long)(x1 - x0) * (fyp - y0) - (long long)(y1 - y0) * (fxp - x0);
        if (((e0 + b0) | (e1 + b1) | (e2 + b2)) >= 0)
            put_fragment(fmaf(Dx, (float)col, fmaf(Dy, (float)row, Dc)), lo, vis + (long long)py * p.w + px);
    }
}

__device__ __forceinline__ float hp_log2(float x) {
    const unsigned u = __float_as_uint(x);
    const int e = (int)((u >> 23) & 0xff) - 127;
    const float m = __uint_as_float((u & 0x007fffffu) | 0x3f800000u) - 1.0f;
    float q = fmaf(m, 0.15922009f, -0.58208540f);
    q = fmaf(m, q, 1.42286531f);
    q = m * q;
    return (float)e + q;
}

template <bool POW2>
__device__ __forceinline__ uchar4 fetch_texel(const uchar4 *tex, int W, int H, int x, int y) {
    if (POW2) {
        x &= W - 1;
        y &= H - 1;
    } else {
        x %= W; if (x < 0) x += W;
        y %= H; if (y < 0) y += H;
    }
    return __ldg(tex + y * W + x);
}

__device__ __forceinline__ float bilerp(float fx, float fy, unsigned c00, unsigned c01, unsigned c10, unsigned c11) {
    const float a = (float)c00, b = (float)c01, c = (float)c10, d = (float)c11;
    const float top = fmaf(fx, b - a, a), bot = fmaf(fx, d - c, c);
    return fmaf(fy, bot - top, top);
}

template <bool POW2>
__device__ __forceinline__ float3 sample_bilinear(const uchar4 *tex, int Wi, int Hi, float u, float v) {
    const float x = fmaf(u, (float)Wi, -0.5f);
    const float y = fmaf(1.0f - v, (float)Hi, -0.5f);
    float xf = floorf(x), yf = floorf(y);
    const float fx = x - xf, fy = y - yf;
    if (!(xf > -1.0e9f)) xf = -1.0e9f;
    if (xf > 1.0e9f) xf = 1.0e9f;
    if (!(yf > -1.0e9f)) yf = -1.0e9f;
    if (yf > 1.0e9f) yf = 1.0e9f;
    const int x0 = (int)xf, y0 = (int)yf;
    uchar4 c00, c01, c10, c11;
    if (POW2) {  // wrap = bit mask; the two row offsets and the two column indices are shared by the four texels
        const unsigned xa = (unsigned)x0 & (unsigned)(Wi - 1), xb = (unsigned)(x0 + 1) & (unsigned)(Wi - 1);
        const unsigned ra = ((unsigned)y0 & (unsigned)(Hi - 1)) * (unsigned)Wi, rb = ((unsigned)(y0 + 1) & (unsigned)(Hi - 1)) * (unsigned)Wi;
        c00 = __ldg(tex + (ra + xa)); c01 = __ldg(tex + (ra + xb));
        c10 = __ldg(tex + (rb + xa)); c11 = __ldg(tex + (rb + xb));
    } else {
        c00 = fetch_texel<POW2>(tex, Wi, Hi, x0, y0); c01 = fetch_texel<POW2>(tex, Wi, Hi, x0 + 1, y0);
        c10 = fetch_texel<POW2>(tex, Wi, Hi, x0, y0 + 1); c11 = fetch_texel<POW2>(tex, Wi, Hi, x0 + 1, y0 + 1);
    }
    return make_float3(bilerp(fx, fy, c00.x, c01.x, c10.x, c11.x), bilerp(fx, fy, c00.y, c01.y, c10.y, c11.y),
                       bilerp(fx, fy, c00.z, c01.z, c10.z, c11.z));
}

template <bool POW2>
__device__ __forceinline__ float3 sample_trilinear(const HpbMeshDev &m, float u, float v, float lod) {
    const float maxl = (float)(m.tex_levels - 1);
    if (lod > maxl) lod = maxl;
    const float lf = floorf(lod);
    const float fl = lod - lf;
    const int li = (int)lf;
    const float3 ca = sample_bilinear<POW2>(m.tex + m.tex_off[li], m.tex_w[li], m.tex_h[li], u, v);
    if (fl > 0.0f && li + 1 < m.tex_levels) {
        const float3 cb = sample_bilinear<POW2>(m.tex + m.tex_off[li + 1], m.tex_w[li + 1], m.tex_h[li + 1], u, v);
        return make_float3(fmaf(fl, cb.x - ca.x, ca.x), fmaf(fl, cb.y - ca.y, ca.y), fmaf(fl, cb.z - ca.z, ca.z));
    }
    return ca;
}

// Panda3D's 32^3 "normal map" lookup (renderer/utils.py:63-79): texel k = floor(k*255/32), repeat wrap, linear.
// tab[k] = (T_k, T_{k+1} - T_k) with T_k = floor(k*255/32); lut[q] = (float)q / 255.0f (IEEE division, done once per
// CTA): the 8-bit framebuffer value returned as q/255.
__device__ __forceinline__ float encode_normal(float c, const float2 *tab, const float *lut) {
    const float s = c - floorf(c);
    const float t = fmaf(s, 32.0f, -0.5f);
    const float kf = floorf(t);
    const float f = t - kf;
    const float2 T = tab[((int)kf) & 31];
    const float val = fmaf(f, T.y, T.x);
    return lut[(int)(val + 0.5f)];  // val in [0, 247]: truncation == floor
}

__device__ __forceinline__ float quant8(float c, const float *lut) {
    float q = floorf(c + 0.5f);
    if (!(q > 0.0f)) q = 0.0f;
    if (q > 255.0f) q = 255.0f;
    return lut[(int)q];
}

// Pass 2 of the triangle stage: every active lane rasterises the (small) triangle `t` it was handed.
__device__ __forceinline__ void raster_small_tris(const RasterParams &p, const int4 *faces, const int2 *sxy, const float *siz,
                                                  bool active, int t, unsigned long long *vis) {
    int roww = 1, scnt = 0, jx0 = 0, jy0 = 0;
    unsigned E0 = 0, E1 = 0, E2 = 0, sx0 = 0, sx1 = 0, sx2 = 0, sy0 = 0, sy1 = 0, sy2 = 0;
    float Dc = 2.0f, Dx = 0.0f, Dy = 0.0f;
    if (active) {
        const int4 f = __ldg(faces + t);
        const int2 a = sxy[f.x];
        int2 b = sxy[f.y], c = sxy[f.z];
        const float iz0 = siz[f.x];
        float iz1 = siz[f.y], iz2 = siz[f.z];
        // small triangles: every product below is exact in 32 bits (|.| <= 2 * bbox_w * bbox_h < 2^30)
        int area2 = (b.x - a.x) * (c.y - a.y) - (c.x - a.x) * (b.y - a.y);
        if (area2 < 0) {  // orient so that area2 > 0
            const int2 tv = b; b = c; c = tv;
            const float tz = iz1; iz1 = iz2; iz2 = tz;
            area2 = -area2;
        }
        jx0 = max(ceil_div_pix(min3i(a.x, b.x, c.x)), 0);
        jy0 = max(ceil_div_pix(min3i(a.y, b.y, c.y)), 0);
        const int jx1 = min(floor_div_pix(max3i(a.x, b.x, c.x)), p.w - 1);
        const int jy1 = min(floor_div_pix(max3i(a.y, b.y, c.y)), p.h - 1);
        roww = jx1 - jx0 + 1;
        scnt = roww * (jy1 - jy0 + 1);
        const int px0 = jx0 * HPB_SUBPIX + 128, py0 = jy0 * HPB_SUBPIX + 128;
        const int dx0 = c.x - b.x, dy0 = c.y - b.y, dx1 = a.x - c.x, dy1 = a.y - c.y, dx2 = b.x - a.x, dy2 = b.y - a.y;
        const int e0o = dx0 * (py0 - b.y) - dy0 * (px0 - b.x);
        const int e1o = dx1 * (py0 - c.y) - dy1 * (px0 - c.x);
        const int e2o = dx2 * (py0 - a.y) - dy2 * (px0 - a.x);
        depth_plane(p, (float)e0o, (float)e1o, (float)e2o, (float)area2, a.x, a.y, b.x, b.y, c.x, c.y, iz0, iz1, iz2, Dc, Dx, Dy);
        // biased edge values: pixel covered <=> all three >= 0.  Unsigned arithmetic: the stepped values are exact modulo
        // 2^32 and the true values at the bounding-box pixels fit 31 bits; intermediate row jumps may wrap.
        E0 = (unsigned)(e0o + edge_bias(dx0, dy0));
        E1 = (unsigned)(e1o + edge_bias(dx1, dy1));
        E2 = (unsigned)(e2o + edge_bias(dx2, dy2));
        sx0 = (unsigned)(-dy0) * HPB_SUBPIX; sx1 = (unsigned)(-dy1) * HPB_SUBPIX; sx2 = (unsigned)(-dy2) * HPB_SUBPIX;
        sy0 = (unsigned)dx0 * HPB_SUBPIX; sy1 = (unsigned)dx1 * HPB_SUBPIX; sy2 = (unsigned)dx2 * HPB_SUBPIX;
    }
    // loop 1: coverage bit mask of the lane's bounding box (row-major), branch-free edge stepping.  Pixel k's bit is
    // shifted in from the right, so after mx iterations it sits at position mx-1-k.
    const int mx = __reduce_max_sync(0xffffffffu, scnt);
    unsigned cov = 0;
    {
        unsigned e0 = E0, e1 = E1, e2 = E2;
        const unsigned rj0 = sy0 - (unsigned)roww * sx0, rj1 = sy1 - (unsigned)roww * sx1, rj2 = sy2 - (unsigned)roww * sx2;
        int col = 0;
#pragma unroll 4
        for (int k = 0; k < mx; ++k) {
            cov = __funnelshift_l(~(e0 | e1 | e2), cov, 1);
            e0 += sx0; e1 += sx1; e2 += sx2;
            const bool wrap = ++col == roww;
            e0 += wrap ? rj0 : 0u; e1 += wrap ? rj1 : 0u; e2 += wrap ? rj2 : 0u;
            col = wrap ? 0 : col;
        }
    }
    // keep the bits of pixels k < scnt, then put pixel k at bit k
    cov = scnt > 0 ? (cov >> (mx - scnt)) : 0u;       // pixel k now at position scnt-1-k
    cov = scnt > 0 ? (__brev(cov) >> (32 - scnt)) : 0u;  // pixel k at bit k
    // loop 2: one covered pixel per iteration
    const int mxn = __reduce_max_sync(0xffffffffu, __popc(cov));
    const unsigned rcp = 65535u / (unsigned)roww + 1u;  // k / roww == (k * rcp) >> 16 for k, roww <= 32
    unsigned long long *org = vis + (long long)jy0 * p.w + jx0;
    const unsigned lo = (unsigned)t << 1;
    for (int j = 0; j < mxn; ++j) {
        if (cov) {
            const int k = __ffs(cov) - 1;
            cov &= cov - 1;
            const int row = (int)(((unsigned)k * rcp) >> 16), col = k - row * roww;
            put_fragment(fmaf(Dx, (float)col, fmaf(Dy, (float)row, Dc)), lo, org + row * p.w + col);
        }
    }
}

// Debug build only (-DHPB_PHASE_CLOCKS, scripts/raster_phases.py): SM clocks thread 0 of every CTA spends in each phase.
#ifdef HPB_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[8];
#define HPB_PHASE_MARK(i)                                                             \
    if (tid == 0) {                                                                   \
        const long long t_now = clock64();                                            \
        atomicAdd(&g_phase_clk[i], (unsigned long long)(t_now - t_prev));             \
        t_prev = t_now;                                                               \
    }
#else
#define HPB_PHASE_MARK(i)
#endif

// S2D = false: planar float32 outputs (hpb_render); true: bf16 space-to-depth network input (hpb_render_s2d_bf16).  Two
// instantiations so that the packed-output state of the second does not cost the first any registers.
// LIT: per-pixel Lambert shading by point / directional lights (the render_normals=False light rig of
// megapose/models/pose_rigid.py:105-141,421-422); a separate instantiation, so the hot path carries none of its registers.
template <bool S2D, bool LIT>
__global__ void __launch_bounds__(RASTER_THREADS, 1) hpb_raster_kernel(const RasterParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ HpbMeshDev sM;
    __shared__ float sT[16];
    __shared__ float sK[4];
    __shared__ float sAmb[3];
    __shared__ float sLight[HPB_MAX_LIGHTS][8];  // LIT: (type, position or direction in the CAMERA frame, colour)
    __shared__ int sFinite;
    __shared__ int sClipped;
    __shared__ int sBox[4];  // min x, min y, max x, max y of the snapped vertices (fixed point)
    __shared__ float sLut[256];
    __shared__ float2 sNrmTab[32];
    __shared__ int sQueue[RASTER_WARPS][64];  // per-warp queue of surviving triangle ids (pass 1 -> pass 2)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = p.G;
    const int rank = G > 1 ? (int)cg::this_cluster().block_rank() : 0;
    const int group = blockIdx.x / G, n_groups = gridDim.x / G;
    const int npix = p.h * p.w;
    unsigned long long *vis = p.vis + (size_t)group * npix;
    int *queue = sQueue[warp];

    if (tid < 256) sLut[tid] = (float)tid / 255.0f;
    if (tid < 32) {
        const int T0 = (tid * 255) >> 5, T1 = (((tid + 1) & 31) * 255) >> 5;
        sNrmTab[tid] = make_float2((float)T0, (float)T1 - (float)T0);
    }
#ifdef HPB_PHASE_CLOCKS
    long long t_prev = clock64();
#endif
    for (int hyp = group; hyp < p.b; hyp += n_groups) {
        {
            const int *src = reinterpret_cast<const int *>(p.meshes + p.mesh_ids[hyp]);
            int *dst = reinterpret_cast<int *>(&sM);
            for (int i = tid; i < (int)(sizeof(HpbMeshDev) / 4); i += RASTER_THREADS) dst[i] = src[i];
        }
        if (tid == 0) {
            sFinite = 1;
            sClipped = 0;
            sBox[0] = 0x7fffffff; sBox[1] = 0x7fffffff; sBox[2] = (int)0x80000000; sBox[3] = (int)0x80000000;
        }
        __syncthreads();
        if (tid < 16) {
            const float v = p.TCO[(size_t)hyp * 16 + tid];
            sT[tid] = v;
            if (!isfinite(v)) sFinite = 0;
        } else if (tid < 25) {
            const float v = p.K[(size_t)hyp * 9 + (tid - 16)];
            if (!isfinite(v)) sFinite = 0;
            if (tid == 16) sK[0] = v;       // fx
            else if (tid == 20) sK[1] = v;  // fy
            else if (tid == 18) sK[2] = v;  // cx
            else if (tid == 21) sK[3] = v;  // cy
        } else if (tid < 28) {
            float a = p.ambient ? p.ambient[(size_t)hyp * 3 + (tid - 25)] : 1.0f;
            if (!LIT && a > 1.0f) a = 1.0f;  // with lights the sum of all contributions is clamped per pixel instead
            if (!(a > 0.0f)) a = 0.0f;
            sAmb[tid - 25] = a;
        }
        __syncthreads();
        if (LIT && tid < p.n_lights) {  // lights are placed in the world (= object) frame: move them into the camera frame
            const float *L = p.lights + ((size_t)hyp * p.n_lights + tid) * 8;
            const float x = L[1], y = L[2], z = L[3];
            const bool dir = L[0] != 0.0f;
            float cx3 = fmaf(sT[2], z, fmaf(sT[1], y, sT[0] * x)), cy3 = fmaf(sT[6], z, fmaf(sT[5], y, sT[4] * x));
            float cz3 = fmaf(sT[10], z, fmaf(sT[9], y, sT[8] * x));
            if (!dir) { cx3 = cx3 + sT[3]; cy3 = cy3 + sT[7]; cz3 = cz3 + sT[11]; }
            sLight[tid][0] = L[0]; sLight[tid][1] = cx3; sLight[tid][2] = cy3; sLight[tid][3] = cz3;
            sLight[tid][4] = L[4]; sLight[tid][5] = L[5]; sLight[tid][6] = L[6]; sLight[tid][7] = 0.0f;
        }
        if (LIT) __syncthreads();
        HPB_PHASE_MARK(0)  // scene set-up
        const bool finite = sFinite != 0;
        const HpbMeshDev &m = sM;
        const int nv = m.nv, nf = m.nf;
        // screen-space vertex arrays of THIS scene's mesh: shared memory when it fits (decided per scene, so one huge
        // mesh in the database does not push the small ones out of shared memory), else the CTA's global scratch slice
        const int nvp = (nv + 1) & ~1;  // keeps the float array behind the int2 array 8-byte aligned
        unsigned char *vbase = nvp <= p.smem_verts ? smem_raw : p.vert_scratch + (size_t)blockIdx.x * p.max_nv * 12;
        int2 *sxy = reinterpret_cast<int2 *>(vbase);
        float *siz = reinterpret_cast<float *>(vbase + (size_t)nvp * 8);
        int bx0 = 1, bx1 = 0, by0 = 1, by1 = 0;  // pixel bounding box of the scene (empty by default)

        if (finite) {
            // ---------------- phase A: vertex stage ----------------
            const float fx = sK[0], fy = sK[1], cx = sK[2], cy = sK[3];
            int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = (int)0x80000000, mxy = (int)0x80000000;
            bool clipped = false;
            for (int i = tid; i < nv; i += RASTER_THREADS) {
                const float x = __ldg(m.pos + 3 * i), y = __ldg(m.pos + 3 * i + 1), z = __ldg(m.pos + 3 * i + 2);
                const float X = fmaf(sT[2], z, fmaf(sT[1], y, fmaf(sT[0], x, sT[3])));
                const float Y = fmaf(sT[6], z, fmaf(sT[5], y, fmaf(sT[4], x, sT[7])));
                const float Z = fmaf(sT[10], z, fmaf(sT[9], y, fmaf(sT[8], x, sT[11])));
                int2 o = make_int2(VTX_CLIPPED, VTX_CLIPPED);
                float iz = 0.0f;
                if (Z >= p.z_near) {
                    iz = 1.0f / Z;
                    o.x = snap_fixed(fmaf(fx, X * iz, cx));
                    o.y = snap_fixed(fmaf(fy, Y * iz, cy));
                    mnx = min(mnx, o.x); mxx = max(mxx, o.x);
                    mny = min(mny, o.y); mxy = max(mxy, o.y);
                } else {
                    clipped = true;
                }
                sxy[i] = o;
                siz[i] = iz;
            }
            mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
            mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
            if (lane == 0 && mnx <= mxx) {
                atomicMin(&sBox[0], mnx); atomicMin(&sBox[1], mny);
                atomicMax(&sBox[2], mxx); atomicMax(&sBox[3], mxy);
            }
            if (__any_sync(0xffffffffu, clipped) && lane == 0) sClipped = 1;
            __syncthreads();
            if (tid == 0 && rank == 0 && sClipped && p.clipped_scenes) atomicAdd(p.clipped_scenes, 1ull);
            HPB_PHASE_MARK(1)  // phase A
            if (sBox[0] <= sBox[2]) {
                bx0 = max(ceil_div_pix(sBox[0]), 0); bx1 = min(floor_div_pix(sBox[2]), p.w - 1);
                by0 = max(ceil_div_pix(sBox[1]), 0); by1 = min(floor_div_pix(sBox[3]), p.h - 1);
            }

            // ---------------- phase B: triangle stage ----------------
            if (bx0 <= bx1 && by0 <= by1) {
                // back faces of a closed surface are invisible unless the near plane has cut it open
                const int cull = (sClipped == 0 && fx > 0.0f && fy > 0.0f) ? m.cull_sign : 0;
                const int per = (nf + G - 1) / G;
                const int lo = rank * per, hi = min(nf, lo + per);
                int nq = 0;  // queued triangle ids of this warp (warp-uniform)
                for (int base = lo; base < hi; base += RASTER_THREADS) {  // trip count is warp-uniform
                    // ---- pass 1: classify ----
                    const int t = base + tid;
                    bool keep = false, big = false;
                    int2 a = make_int2(0, 0), b = a, c = a;
                    int4 f = make_int4(0, 0, 0, 0);
                    int jx0 = 0, jx1 = -1, jy0 = 0, jy1 = -1;
                    long long area2 = 0;
                    if (t < hi) {
                        f = __ldg(m.faces + t);
                        a = sxy[f.x]; b = sxy[f.y]; c = sxy[f.z];
                        const int tmnx = min3i(a.x, b.x, c.x), tmxx = max3i(a.x, b.x, c.x);
                        const int tmny = min3i(a.y, b.y, c.y), tmxy = max3i(a.y, b.y, c.y);
                        jx0 = max(ceil_div_pix(tmnx), 0); jx1 = min(floor_div_pix(tmxx), p.w - 1);
                        jy0 = max(ceil_div_pix(tmny), 0); jy1 = min(floor_div_pix(tmxy), p.h - 1);
                        area2 = (long long)(b.x - a.x) * (long long)(c.y - a.y) - (long long)(c.x - a.x) * (long long)(b.y - a.y);
                        keep = tmnx != VTX_CLIPPED && jx0 <= jx1 && jy0 <= jy1 && area2 != 0;
                        if (cull != 0) keep = keep && ((area2 < 0) == (cull < 0));
                        // 32-bit edge functions are exact when |e| <= 2 * bbox_w * bbox_h (fixed point) < 2^30
                        const unsigned bw = (unsigned)(tmxx - tmnx) + 2u * HPB_SUBPIX, bh = (unsigned)(tmxy - tmny) + 2u * HPB_SUBPIX;
                        big = keep && ((jx1 - jx0 + 1) * (jy1 - jy0 + 1) > SMALL_TRI_MAX || __umulhi(bw, bh) != 0u || bw * bh >= (1u << 29));
                    }
                    // ---- big triangles (rare): the whole warp walks one triangle at a time ----
                    unsigned bm = __ballot_sync(0xffffffffu, big);
                    if (bm) {
                        float Dc = 0.f, Dx = 0.f, Dy = 0.f;
                        if (big) {
                            float iz0 = siz[f.x], iz1 = siz[f.y], iz2 = siz[f.z];
                            if (area2 < 0) {
                                const int2 tv = b; b = c; c = tv;
                                const float tz = iz1; iz1 = iz2; iz2 = tz;
                                area2 = -area2;
                            }
                            const int px0 = jx0 * HPB_SUBPIX + 128, py0 = jy0 * HPB_SUBPIX + 128;
                            const long long e0o = (long long)(c.x - b.x) * (py0 - b.y) - (long long)(c.y - b.y) * (px0 - b.x);
                            const long long e1o = (long long)(a.x - c.x) * (py0 - c.y) - (long long)(a.y - c.y) * (px0 - c.x);
                            const long long e2o = (long long)(b.x - a.x) * (py0 - a.y) - (long long)(b.y - a.y) * (px0 - a.x);
                            depth_plane(p, (float)e0o, (float)e1o, (float)e2o, (float)area2, a.x, a.y, b.x, b.y, c.x, c.y, iz0, iz1, iz2,
                                        Dc, Dx, Dy);
                        }
                        const unsigned wide = area2 >= (1ll << 31) ? 1u : 0u;
                        while (bm) {
                            const int src = __ffs(bm) - 1;
                            bm &= bm - 1;
                            const int x0 = __shfl_sync(0xffffffffu, a.x, src), y0 = __shfl_sync(0xffffffffu, a.y, src);
                            const int x1 = __shfl_sync(0xffffffffu, b.x, src), y1 = __shfl_sync(0xffffffffu, b.y, src);
                            const int x2 = __shfl_sync(0xffffffffu, c.x, src), y2 = __shfl_sync(0xffffffffu, c.y, src);
                            const float qc = __shfl_sync(0xffffffffu, Dc, src), qx = __shfl_sync(0xffffffffu, Dx, src);
                            const float qy = __shfl_sync(0xffffffffu, Dy, src);
                            const int ax0 = __shfl_sync(0xffffffffu, jx0, src), ax1 = __shfl_sync(0xffffffffu, jx1, src);
                            const int ay0 = __shfl_sync(0xffffffffu, jy0, src), ay1 = __shfl_sync(0xffffffffu, jy1, src);
                            const unsigned wd = __shfl_sync(0xffffffffu, wide, src);
                            raster_tri_warp(p, x0, y0, x1, y1, x2, y2, qc, qx, qy, ax0, ax1, ay0, ay1,
                                            ((unsigned)(base + (tid & ~31) + src) << 1) | wd, vis, lane);
                        }
                    }
                    // ---- compaction: queue the small survivors; rasterise whenever 32 are waiting ----
                    const bool small = keep && !big;
                    const unsigned sm_mask = __ballot_sync(0xffffffffu, small);
                    if (small) queue[nq + __popc(sm_mask & ((1u << lane) - 1u))] = t;
                    nq += __popc(sm_mask);
                    __syncwarp();
                    if (nq >= 32) {
                        const int tq = queue[lane];
                        const int spill = queue[32 + lane];
                        __syncwarp();
                        nq -= 32;
                        if (lane < nq) queue[lane] = spill;
                        raster_small_tris(p, m.faces, sxy, siz, true, tq, vis);
                        __syncwarp();
                    }
                }
                if (nq > 0) raster_small_tris(p, m.faces, sxy, siz, lane < nq, lane < nq ? queue[lane] : 0, vis);
            }
            __threadfence();
        }
        HPB_PHASE_MARK(2)  // phase B, thread 0's own share
        if (G > 1) cg::this_cluster().sync();
        else __syncthreads();
        HPB_PHASE_MARK(3)  // wait for the slowest warp of phase B

        // ---------------- phase C: resolve ----------------
        const size_t oi = (size_t)(hyp / p.views), ov = (size_t)(hyp % p.views) * p.view_stride;
        float *rgb = (p.flags & HPB_RENDER_RGB) ? p.rgb + oi * p.rgb_bs + ov : nullptr;
        float *nrm = (p.flags & HPB_RENDER_NORMALS) ? p.nrm + oi * p.nrm_bs + ov : nullptr;
        float *dep = (p.flags & HPB_RENDER_DEPTH) ? p.depth + oi * p.depth_bs + ov : nullptr;
        uint8_t *msk = (p.flags & HPB_RENDER_MASK) ? p.mask + (size_t)hyp * p.mask_bs : nullptr;  // never view-interleaved
        const bool textured = m.tex != nullptr && m.uv != nullptr;
        const bool want_z = dep != nullptr || msk != nullptr;
        const bool do_rgb = (p.flags & HPB_RENDER_RGB) != 0, do_nrm = (p.flags & HPB_RENDER_NORMALS) != 0;
        // shading of ONE visible pixel (shared by the planar resolve and the space-to-depth resolve below)
        auto shade = [&](const unsigned long long key, const int px, const int py, float &r, float &g, float &bl, float &n0,
                         float &n1, float &n2, float &z) {
            if (key != HPB_VIS_EMPTY) {
                const unsigned klo = (unsigned)(key & 0xffffffffull);
                const int4 f = __ldg(m.faces + (klo >> 1));
                const int2 a = sxy[f.x], b = sxy[f.y], c = sxy[f.z];
                const float iz0 = siz[f.x], iz1 = siz[f.y], iz2 = siz[f.z];
                const int fxp = px * HPB_SUBPIX + 128, fyp = py * HPB_SUBPIX + 128;
                // Edge values at the pixel centre in the triangle's own winding (no re-orientation: the weights below
                // are ratios, so a common sign cancels).  For a covered pixel |e_i| <= |area2|, so 32-bit wrap-around
                // arithmetic is exact unless the triangle is flagged wide.
                const int dx0 = c.x - b.x, dy0 = c.y - b.y, dx1 = a.x - c.x, dy1 = a.y - c.y, dx2 = b.x - a.x, dy2 = b.y - a.y;
                float fe0, fe1, fe2;
                if (!(klo & 1u)) {
                    fe0 = (float)(int)((unsigned)dx0 * (unsigned)(fyp - b.y) - (unsigned)dy0 * (unsigned)(fxp - b.x));
                    fe1 = (float)(int)((unsigned)dx1 * (unsigned)(fyp - c.y) - (unsigned)dy1 * (unsigned)(fxp - c.x));
                    fe2 = (float)(int)((unsigned)dx2 * (unsigned)(fyp - a.y) - (unsigned)dy2 * (unsigned)(fxp - a.x));
                } else {
                    fe0 = (float)((long long)dx0 * (fyp - b.y) - (long long)dy0 * (fxp - b.x));
                    fe1 = (float)((long long)dx1 * (fyp - c.y) - (long long)dy1 * (fxp - c.x));
                    fe2 = (float)((long long)dx2 * (fyp - a.y) - (long long)dy2 * (fxp - a.x));
                }
                const float w0 = fe0 * iz0, w1 = fe1 * iz1, w2 = fe2 * iz2;
                const float s = __frcp_rn((w0 + w1) + w2);
                const float p0 = w0 * s, p1 = w1 * s, p2 = w2 * s;
                if (want_z) {
                    const float d = __uint_as_float((unsigned)(key >> 32));
                    z = p.a_f / (d - p.b_f);
                    if (d > p.eps_hi) z = 0.0f;
                }
                const float4 A0 = __ldg(m.nu + f.x), A1 = __ldg(m.nu + f.y), A2 = __ldg(m.nu + f.z);
                float lit0 = sAmb[0], lit1 = sAmb[1], lit2 = sAmb[2];
                if (do_nrm || LIT) {
                    // object-space normal interpolated over the triangle, rotated into the eye frame, normalised once
                    const float ox = fmaf(p2, A2.x, fmaf(p1, A1.x, p0 * A0.x));
                    const float oy = fmaf(p2, A2.y, fmaf(p1, A1.y, p0 * A0.y));
                    const float oz = fmaf(p2, A2.z, fmaf(p1, A1.z, p0 * A0.z));
                    float nx = fmaf(sT[2], oz, fmaf(sT[1], oy, sT[0] * ox));
                    float ny = fmaf(sT[6], oz, fmaf(sT[5], oy, sT[4] * ox));
                    float nz = fmaf(sT[10], oz, fmaf(sT[9], oy, sT[8] * ox));
                    const float len2 = fmaf(nz, nz, fmaf(ny, ny, nx * nx));
                    if (len2 > 0.0f) {
                        const float rl = __frcp_rn(__fsqrt_rn(len2));
                        nx *= rl; ny *= rl; nz *= rl;
                    }
                    if (do_nrm) {
                        n0 = encode_normal(nx, sNrmTab, sLut);
                        n1 = encode_normal(nz, sNrmTab, sLut);
                        n2 = encode_normal(-ny, sNrmTab, sLut);
                    }
                    if (LIT) {
                        // surface point in the camera frame from the pixel centre and the metric depth of the winning fragment
                        const float dk = __uint_as_float((unsigned)(key >> 32));
                        const float zc = p.a_f / (dk - p.b_f);
                        const float X = (((float)px + 0.5f) - sK[2]) / sK[0] * zc, Y = (((float)py + 0.5f) - sK[3]) / sK[1] * zc;
                        for (int li = 0; li < p.n_lights; ++li) {
                            float ndl;
                            if (sLight[li][0] != 0.0f) {  // directional: light travels along (x,y,z)
                                ndl = -fmaf(nz, sLight[li][3], fmaf(ny, sLight[li][2], nx * sLight[li][1]));
                            } else {
                                const float lx = sLight[li][1] - X, ly = sLight[li][2] - Y, lz = sLight[li][3] - zc;
                                const float l2 = fmaf(lz, lz, fmaf(ly, ly, lx * lx));
                                ndl = fmaf(nz, lz, fmaf(ny, ly, nx * lx));
                                if (l2 > 0.0f) ndl = ndl * __frcp_rn(__fsqrt_rn(l2));
                            }
                            if (ndl > 0.0f) {
                                lit0 = fmaf(ndl, sLight[li][4], lit0);
                                lit1 = fmaf(ndl, sLight[li][5], lit1);
                                lit2 = fmaf(ndl, sLight[li][6], lit2);
                            }
                        }
                        if (lit0 > 1.0f) lit0 = 1.0f;
                        if (lit1 > 1.0f) lit1 = 1.0f;
                        if (lit2 > 1.0f) lit2 = 1.0f;
                    }
                }
                if (do_rgb) {
                    float3 col = make_float3(255.0f, 255.0f, 255.0f);
                    if (textured) {
                        const float v0 = __ldg(m.tv + f.x), v1 = __ldg(m.tv + f.y), v2 = __ldg(m.tv + f.z);
                        const float u = fmaf(p2, A2.w, fmaf(p1, A1.w, p0 * A0.w));
                        const float v = fmaf(p2, v2, fmaf(p1, v1, p0 * v0));
                        // analytic screen-space derivatives of (u,v) for the mip level, from the per-pixel steps of the
                        // un-normalised perspective weights e_i / z_i
                        const float g0x = ((float)(-dy0) * 256.0f) * iz0, g1x = ((float)(-dy1) * 256.0f) * iz1, g2x = ((float)(-dy2) * 256.0f) * iz2;
                        const float g0y = ((float)dx0 * 256.0f) * iz0, g1y = ((float)dx1 * 256.0f) * iz1, g2y = ((float)dx2 * 256.0f) * iz2;
                        const float dDx = (g0x + g1x) + g2x, dDy = (g0y + g1y) + g2y;
                        const float dNux = fmaf(g2x, A2.w, fmaf(g1x, A1.w, g0x * A0.w));
                        const float dNuy = fmaf(g2y, A2.w, fmaf(g1y, A1.w, g0y * A0.w));
                        const float dNvx = fmaf(g2x, v2, fmaf(g1x, v1, g0x * v0));
                        const float dNvy = fmaf(g2y, v2, fmaf(g1y, v1, g0y * v0));
                        const float W0 = (float)m.tex_w[0], H0 = (float)m.tex_h[0];
                        const float ax = (dNux - u * dDx) * s * W0, bx = (dNvx - v * dDx) * s * H0;
                        const float ay = (dNuy - u * dDy) * s * W0, by = (dNvy - v * dDy) * s * H0;
                        const float r2x = fmaf(ax, ax, bx * bx), r2y = fmaf(ay, ay, by * by);
                        const float rho2 = r2x > r2y ? r2x : r2y;
                        float lod = 0.0f;
                        if (rho2 > 1.0f && rho2 < 1.0e30f) lod = 0.5f * hp_log2(rho2);
                        col = m.tex_pow2 ? sample_trilinear<true>(m, u, v, lod) : sample_trilinear<false>(m, u, v, lod);
                    } else if (m.vcol) {
                        const uchar4 c0 = __ldg(m.vcol + f.x), c1 = __ldg(m.vcol + f.y), c2 = __ldg(m.vcol + f.z);
                        col.x = fmaf(p2, (float)c2.x, fmaf(p1, (float)c1.x, p0 * (float)c0.x));
                        col.y = fmaf(p2, (float)c2.y, fmaf(p1, (float)c1.y, p0 * (float)c0.y));
                        col.z = fmaf(p2, (float)c2.z, fmaf(p1, (float)c1.z, p0 * (float)c0.z));
                    }
                    r = quant8(col.x * lit0, sLut);
                    g = quant8(col.y * lit1, sLut);
                    bl = quant8(col.z * lit2, sLut);
                }
            }
        };
        if constexpr (!S2D) {
        // Work unit = one 32-pixel span of one row, dealt to the warps of the cluster round-robin (fine-grained, so the
        // shaded and the empty spans spread evenly).  The row / span arithmetic is per warp, not per pixel; spans that
        // miss the scene's bounding box are zero-filled without touching the visibility buffer.
        const int nspan = (p.w + 31) >> 5;
        const int n_units = p.h * nspan;
        const int sp0 = bx0 >> 5, sp1 = bx1 >> 5;
        for (int u = rank * RASTER_WARPS + warp; u < n_units; u += G * RASTER_WARPS) {
            int py = (int)__umulhi((unsigned)u, p.span_magic);
            if (py * nspan > u) --py;
            const int sp = u - py * nspan;
            const int px = (sp << 5) + lane;
            const int pix = py * p.w + px;
            const bool in_row = px < p.w;
            float *rgb_row = rgb ? rgb + pix : nullptr;  // this lane's pixel in plane 0
            float *nrm_row = nrm ? nrm + pix : nullptr;
            if (py < by0 || py > by1 || sp < sp0 || sp > sp1) {  // warp-uniform: background span
                if (in_row) {
                    if (rgb) {
                        __stcs(rgb_row, 0.f);
                        __stcs(rgb_row + (unsigned)npix, 0.f);
                        __stcs(rgb_row + (unsigned)(2 * npix), 0.f);
                    }
                    if (nrm) {
                        __stcs(nrm_row, 0.f);
                        __stcs(nrm_row + (unsigned)npix, 0.f);
                        __stcs(nrm_row + (unsigned)(2 * npix), 0.f);
                    }
                    if (dep) __stcs(dep + pix, 0.f);
                    if (msk) msk[pix] = 0;
                }
                continue;
            }
            unsigned long long key = HPB_VIS_EMPTY;
            if (in_row && px >= bx0 && px <= bx1) {
                key = __ldcg(vis + pix);
                if (key != HPB_VIS_EMPTY) __stcg(vis + pix, HPB_VIS_EMPTY);  // re-arm for the next scene
            }
            float r = 0.f, g = 0.f, bl = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f, z = 0.f;
            if (key != HPB_VIS_EMPTY) shade(key, px, py, r, g, bl, n0, n1, n2, z);
            if (in_row) {
                if (rgb) {
                    __stcs(rgb_row, r);
                    __stcs(rgb_row + (unsigned)npix, g);
                    __stcs(rgb_row + (unsigned)(2 * npix), bl);
                }
                if (nrm) {
                    __stcs(nrm_row, n0);
                    __stcs(nrm_row + (unsigned)npix, n1);
                    __stcs(nrm_row + (unsigned)(2 * npix), n2);
                }
                if (dep) __stcs(dep + pix, z);
                if (msk) msk[pix] = z > 0.0f ? 1 : 0;
            }
        }
        } else {
            // ---- space-to-depth resolve: 4 lanes per CELL of the stem's input, one pixel per lane ----
            // z[n][I][J][(r*2+s)*Cs + c] = xpad[n][c][2I + r][2J + s]  (Cs = Cz / 4 channels reserved per sub-pixel, c < 9 used):
            // xpad = the 9-channel network input (3 crop channels from the crop kernel, 6 rendered channels shaded here)
            // zero-padded by 3 pixels.  Lane l of a warp shades sub-pixel (r,s) = l & 3 of cell l >> 2 of the warp's 8
            // consecutive cells, and a sub-pixel's channel block starts on a 16-byte boundary (Cs is a multiple of 8), so every
            // lane stores its own 9 bf16 values as two 16-byte vectors: no cross-lane traffic, no unaligned groups.  (Round 1
            // packed the 36 values densely, (r*2+s)*9 + c: the shuffles / funnel shifts / selects that re-aligned them cost 142
            // warp instructions per 8 cells -- a third of the whole kernel, more than the shading itself.)
            const int n_cells = p.Hz * p.Wz;
            const int sub = lane & 3;
            // this lane's 16-byte group inside cell 0 of the scene; a cell is addressed by a 32-bit offset from it
            uint4 *zlane = p.s2d + (size_t)hyp * n_cells * p.Cz8 + sub * (p.Cz8 >> 2);
            const uint2 *crop_h = p.crops_h ? p.crops_h + (size_t)hyp * p.crops_bs : nullptr;
            const float *crop_f = p.crops ? p.crops + (size_t)hyp * p.crops_bs : nullptr;
            // The crop pixel AND the visibility key of the NEXT work unit are loaded one iteration ahead (software pipelining
            // through registers).  ncu attributed 27 % of the kernel's stall samples to the first use of the crop load (it
            // streams from HBM) and, once that was hidden, 10 % to the key load (L2): the key heads a chain of dependent
            // loads (key -> face -> vertex attributes -> texels), so taking it off the chain shortens every shaded pixel.
            // The cell coordinates (I, J) advance incrementally, so the look-ahead costs no second division.
            const int q_stride = G * RASTER_WARPS * 8;
            const int dI = q_stride / p.Wz, dJ = q_stride - dI * p.Wz;
            int q = (rank * RASTER_WARPS + warp) * 8 + (lane >> 2);
            int I = (int)__umulhi((unsigned)q, p.wz_magic);
            if (I * p.Wz > q) --I;
            int J = q - I * p.Wz;
            const int oy = (sub >> 1) - 3, ox = (sub & 1) - 3;
            auto crop_fetch = [&](int qq, int II, int JJ) -> uint2 {
                const int yy = 2 * II + oy, xx = 2 * JJ + ox;
                if (crop_h && qq < n_cells && (unsigned)yy < (unsigned)p.h && (unsigned)xx < (unsigned)p.w)
                    return __ldcs(crop_h + (unsigned)(yy * p.w + xx));
                return make_uint2(0u, 0u);
            };
            // the scene's bounding box lies inside the image, so no separate image-bounds test
            auto key_fetch = [&](int qq, int II, int JJ) -> unsigned long long {
                const int yy = 2 * II + oy, xx = 2 * JJ + ox;
                if (qq < n_cells && yy >= by0 && yy <= by1 && xx >= bx0 && xx <= bx1) return __ldcg(vis + (unsigned)(yy * p.w + xx));
                return HPB_VIS_EMPTY;
            };
            uint2 cw = crop_fetch(q, I, J);
            unsigned long long key = key_fetch(q, I, J);
            while (q < n_cells) {
                int In = I + dI, Jn = J + dJ;
                if (Jn >= p.Wz) { Jn -= p.Wz; ++In; }
                const uint2 cw_next = crop_fetch(q + q_stride, In, Jn);
                const unsigned long long key_next = key_fetch(q + q_stride, In, Jn);
                const int py = 2 * I + oy, px = 2 * J + ox;
                float v[6];
#pragma unroll
                for (int c = 0; c < 6; ++c) v[c] = 0.f;
                if (key != HPB_VIS_EMPTY) {
                    __stcg(vis + (unsigned)(py * p.w + px), HPB_VIS_EMPTY);  // re-arm for the next scene
                    float zz = 0.f;
                    shade(key, px, py, v[0], v[1], v[2], v[3], v[4], v[5], zz);
                }
                // own values o0..o8 as bf16 pairs: e0 = (o0,o1) .. e3 = (o6,o7), e4 = (o8,0).  A pixel of the 3-pixel zero
                // border has an all-zero crop word and an empty key, so it needs no branch of its own.
                unsigned e0 = cw.x, c2 = cw.y & 0xffffu;  // crop channels 0,1 and 2 as bf16 bits
                if (!crop_h && (unsigned)py < (unsigned)p.h && (unsigned)px < (unsigned)p.w) {
                    const unsigned pix = (unsigned)(py * p.w + px);
                    const float c0 = __ldcs(crop_f + pix), c1 = __ldcs(crop_f + npix + pix), cb = __ldcs(crop_f + 2 * npix + pix);
                    const __nv_bfloat162 h01 = __floats2bfloat162_rn(c0, c1);
                    e0 = *reinterpret_cast<const unsigned *>(&h01);
                    c2 = (unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(cb));
                }
                const unsigned e1 = c2 | ((unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(v[0])) << 16);
                const __nv_bfloat162 h45 = __floats2bfloat162_rn(v[1], v[2]), h67 = __floats2bfloat162_rn(v[3], v[4]);
                const unsigned e2 = *reinterpret_cast<const unsigned *>(&h45);
                const unsigned e3 = *reinterpret_cast<const unsigned *>(&h67);
                const unsigned e4 = (unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(v[5]));
                uint4 *cell = zlane + (unsigned)(q * p.Cz8);
                __stcs(cell, make_uint4(e0, e1, e2, e3));
                __stcs(cell + 1, make_uint4(e4, 0u, 0u, 0u));
                // channels 16.. of the sub-pixel block are zero padding: written unless the caller keeps a persistent
                // pre-zeroed buffer that only this kernel writes
                if (!p.pad_prezeroed)
                    for (int x = 2; x < (p.Cz8 >> 2); ++x) __stcs(cell + x, make_uint4(0u, 0u, 0u, 0u));
                q += q_stride; I = In; J = Jn; cw = cw_next; key = key_next;
            }
        }
        HPB_PHASE_MARK(4)  // phase C, thread 0's own share
        // the next scene's phase A overwrites the vertex stage, and its phase B (from any CTA of the cluster) writes
        // the visibility buffer this CTA has just re-armed
        if (G > 1) cg::this_cluster().sync();
        else __syncthreads();
        HPB_PHASE_MARK(5)  // wait for the slowest warp of phase C
    }
}

__global__ void hpb_mip_kernel(const uchar4 *src, int sw, int sh, uchar4 *dst, int dw, int dh) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const int x0 = 2 * x, x1 = min(2 * x + 1, sw - 1), y0 = 2 * y, y1 = min(2 * y + 1, sh - 1);
    const uchar4 a = src[(size_t)y0 * sw + x0], b = src[(size_t)y0 * sw + x1];
    const uchar4 c = src[(size_t)y1 * sw + x0], d = src[(size_t)y1 * sw + x1];
    uchar4 o;
    o.x = (unsigned char)((a.x + b.x + c.x + d.x + 2) >> 2);
    o.y = (unsigned char)((a.y + b.y + c.y + d.y + 2) >> 2);
    o.z = (unsigned char)((a.z + b.z + c.z + d.z + 2) >> 2);
    o.w = (unsigned char)((a.w + b.w + c.w + d.w + 2) >> 2);
    dst[(size_t)y * dw + x] = o;
}

__global__ void hpb_tex_expand_kernel(const uint8_t *src, int n, int c, uchar4 *dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uchar4 o;
    o.x = src[(size_t)i * c];
    o.y = src[(size_t)i * c + 1];
    o.z = src[(size_t)i * c + 2];
    o.w = c == 4 ? src[(size_t)i * c + 3] : 255;
    dst[i] = o;
}

__global__ void hpb_fill_u64_kernel(unsigned long long *p, size_t n, unsigned long long v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

}  // namespace

#ifdef HPB_PHASE_CLOCKS
extern "C" int hpb_debug_phase_clocks(unsigned long long *out) {
    unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(out, g_phase_clk, sizeof(zero)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(g_phase_clk, zero, sizeof(zero)) != cudaSuccess) return -1;
    return 0;
}
#endif

int hpb_launch_mip(const uchar4 *src, int sw, int sh, uchar4 *dst, int dw, int dh, cudaStream_t stream) {
    dim3 blk(32, 8), grd((dw + 31) / 32, (dh + 7) / 8);
    hpb_mip_kernel<<<grd, blk, 0, stream>>>(src, sw, sh, dst, dw, dh);
    HPB_CUDA_OK(cudaGetLastError());
    return HPB_OK;
}

int hpb_launch_tex_expand(const uint8_t *src, int n, int c, uchar4 *dst, cudaStream_t stream) {
    hpb_tex_expand_kernel<<<(n + 255) / 256, 256, 0, stream>>>(src, n, c, dst);
    HPB_CUDA_OK(cudaGetLastError());
    return HPB_OK;
}

int hpb_launch_raster(hpb_ctx *ctx, const int32_t *mesh_ids, const float *TCO, const float *K, const float *ambient,
                      int b, int h, int w, float z_near, float z_far, uint32_t flags, float *rgb, int64_t rgb_bs,
                      float *nrm, int64_t nrm_bs, float *depth, int64_t depth_bs, uint8_t *mask, int64_t mask_bs,
                      int views, int64_t view_stride, cudaStream_t stream, const void *crops, int64_t crops_bs,
                      void *s2d_out, int Cz, int crops_fmt, int pad_prezeroed, const float *lights, int n_lights) {
    if (b == 0) return HPB_OK;
    const int npix = h * w;
    const int nv_pad = (ctx->max_nv + 1) & ~1;  // keeps the int2 array 8-byte aligned in every CTA's slice
    const size_t smem_need = (size_t)nv_pad * 12;
    static size_t static_smem = 0;  // the kernel's own __shared__ variables count against the per-block opt-in limit
    if (static_smem == 0) {
        cudaFuncAttributes fa;
        HPB_CUDA_OK(cudaFuncGetAttributes(&fa, hpb_raster_kernel<false, true>));
        static_smem = fa.sharedSizeBytes + 256;
    }
    const size_t smem_cap = (size_t)ctx->max_smem_optin > static_smem ? (size_t)ctx->max_smem_optin - static_smem : 0;
    const int verts_in_smem = smem_need <= smem_cap;
    // when some mesh does not fit, the others still use as much shared memory as there is (per-scene choice in the kernel)
    const size_t smem = verts_in_smem ? smem_need : (smem_cap / 24) * 24;
    const bool lit = lights != nullptr && n_lights > 0;
    void (*kern)(const RasterParams) = s2d_out ? hpb_raster_kernel<true, false> : (lit ? hpb_raster_kernel<false, true> : hpb_raster_kernel<false, false>);
    // all instantiations get the attribute: the cluster-occupancy query below is made on <false,false> whichever runs first
    HPB_CUDA_OK(cudaFuncSetAttribute(hpb_raster_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HPB_CUDA_OK(cudaFuncSetAttribute(hpb_raster_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HPB_CUDA_OK(cudaFuncSetAttribute(hpb_raster_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // clusters of 16 CTAs (non-portable size) for the smallest batches: a 4-view refiner launch then spreads over 64 SMs
    static bool np_ok = cudaFuncSetAttribute(hpb_raster_kernel<false, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                        cudaFuncSetAttribute(hpb_raster_kernel<false, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                        cudaFuncSetAttribute(hpb_raster_kernel<true, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    if (!np_ok) cudaGetLastError();

    // how many clusters of 1/2/4/8 CTAs can be co-resident (one CTA per SM); queried once per shared-memory size
    if (ctx->max_clusters_smem != smem) {
        for (int k = 0; k < 5; ++k) {
            const int G = 1 << k;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(G * 8, 1, 1);
            cfg.blockDim = dim3(RASTER_THREADS, 1, 1);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int n = 0;
            if (G == 1) n = ctx->sm_count;
            else if (G == 16 && !np_ok) n = 0;
            else if (cudaOccupancyMaxActiveClusters(&n, hpb_raster_kernel<false, true>, &cfg) != cudaSuccess) { n = 0; cudaGetLastError(); }
            ctx->max_clusters[k] = n;
        }
        ctx->max_clusters_smem = smem;
    }
    // persistent grid: the largest cluster size that still gives every scene its own cluster
    int k = 0;
    while (k < 4 && ctx->max_clusters[k + 1] >= b) ++k;
    const int G = 1 << k;
    const int n_groups = b < ctx->max_clusters[k] ? b : ctx->max_clusters[k];
    const int n_ctas = n_groups * G;

    // workspace (grown on demand, never freed; the visibility buffer is armed once and re-armed by the kernel itself)
    {
        const int rc = hpb_raster_reserve(ctx, h, w, stream);
        if (rc != HPB_OK) return rc;
    }
    if (!verts_in_smem) {
        const size_t need = (size_t)ctx->sm_count * nv_pad * 12;
        void *vs = ctx->vert_scratch;
        const int rc = hpb_ws_grow(ctx, &vs, &ctx->vert_scratch_bytes, need, stream, "rasteriser vertex-scratch");
        if (rc != HPB_OK) return rc;
        ctx->vert_scratch = (unsigned char *)vs;
    }
    {
        const int rc = hpb_stream_enter(ctx, stream);
        if (rc != HPB_OK) return rc;
    }

    RasterParams p;
    p.meshes = ctx->meshes_dev;
    p.mesh_ids = mesh_ids;
    p.TCO = TCO;
    p.K = K;
    p.ambient = ambient;
    p.lights = lit ? lights : nullptr;
    p.n_lights = lit ? n_lights : 0;
    p.b = b; p.h = h; p.w = w;
    p.z_near = z_near;
    p.inv_near = 1.0f / z_near;
    const float inv_far = 1.0f / z_far;
    p.cd = 1.0f / (p.inv_near - inv_far);
    const double a_d = 1.0 / (1.0 / (double)z_far - 1.0 / (double)z_near);  // renderer/utils.py:56-57
    p.a_f = (float)a_d;
    p.b_f = (float)(-a_d / (double)z_near);
    p.eps_hi = (float)(1.0 - 0.001);
    p.flags = flags;
    p.rgb = rgb; p.nrm = nrm; p.depth = depth; p.mask = mask;
    p.rgb_bs = rgb_bs; p.nrm_bs = nrm_bs; p.depth_bs = depth_bs; p.mask_bs = mask_bs;
    p.views = views; p.view_stride = view_stride;
    p.vis = ctx->vis;
    p.clipped_scenes = ctx->clipped_scenes;
    p.vert_scratch = ctx->vert_scratch;
    p.max_nv = nv_pad;
    p.smem_verts = (int)(smem / 12);
    p.G = G;
    p.span_magic = (unsigned)((1ull << 32) / (unsigned)((w + 31) / 32)) + 1u;
    p.s2d = reinterpret_cast<uint4 *>(s2d_out);
    p.crops = crops_fmt == 0 ? reinterpret_cast<const float *>(crops) : nullptr;
    p.crops_h = crops_fmt == 1 ? reinterpret_cast<const uint2 *>(crops) : nullptr;
    p.crops_bs = crops_bs;
    p.crops_fmt = crops_fmt;
    p.pad_prezeroed = pad_prezeroed;
    p.Hz = h / 2 + 3;
    p.Wz = w / 2 + 3;
    p.Cz8 = Cz / 8;
    p.wz_magic = (unsigned)((1ull << 32) / (unsigned)p.Wz) + 1u;

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_ctas, 1, 1);
    cfg.blockDim = dim3(RASTER_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = G > 1 ? 1 : 0;
    HPB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
    ctx->launches++;
    hpb_stream_leave(ctx, stream);
    return HPB_OK;
}

// Visibility buffer for (h x w) renders: one slice per SM, armed (all-empty) once; the kernel re-arms what it reads.
int hpb_raster_reserve(hpb_ctx *ctx, int h, int w, cudaStream_t stream) {
    const size_t vis_need = (size_t)ctx->sm_count * h * w;
    if (ctx->vis_elems >= vis_need && ctx->vis) return HPB_OK;
    void *v = ctx->vis;
    size_t bytes = ctx->vis_elems * sizeof(unsigned long long);
    const int rc = hpb_ws_grow(ctx, &v, &bytes, vis_need * sizeof(unsigned long long), stream, "rasteriser visibility");
    if (rc != HPB_OK) return rc;
    ctx->vis = (unsigned long long *)v;
    ctx->vis_elems = vis_need;
    hpb_fill_u64_kernel<<<ctx->sm_count * 4, 256, 0, stream>>>(ctx->vis, vis_need, HPB_VIS_EMPTY);
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}
