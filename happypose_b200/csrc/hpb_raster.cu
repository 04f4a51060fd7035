// Batched triangle rasteriser for sm_100a: N single-object scenes -> RGB / normals / depth / mask in ONE launch.
//
// Replaces happypose/toolbox/renderer/panda3d_batch_renderer.py:194-286 (Panda3dBatchRenderer.render) and the
// Panda3D/OpenGL pipeline behind it (panda3d_scene_renderer.py:320-390, renderer/types.py:111-137,254-299,
// renderer/utils.py:46-79).  The semantics (pixel centres, near/far, two-sided, depth read-back rule, ambient-only
// colour, eye-normal 32-level wrap texture) are stated once in DESIGN.md; every float operation below is an
// explicit IEEE op / fmaf so results are bit-identical to the CPU oracle used by the tests (the library is built
// with -fmad=false).
//
// Structure (one persistent CTA per SM, looping over scenes):
//   phase A  vertex stage: object -> camera -> 24.8 fixed-point screen position + 1/z, staged in SHARED memory
//            (12 B per vertex; meshes that do not fit fall back to a per-CTA global scratch slice).
//   phase B  triangle stage: one thread per triangle, exact integer edge functions with a top-left rule over the
//            triangle's pixel bounding box; depth test = 64-bit atomicMin of (depth bits << 32 | triangle id) on a
//            per-CTA visibility buffer that stays resident in L2 (it is re-armed in phase C, never re-cleared).
//   phase C  resolve: one thread per pixel, coalesced planar stores; the winning triangle is re-set-up, attributes
//            are interpolated perspective-correctly, texture is trilinearly filtered from an RGBA8 mip chain.
//            Background pixels are written as zeros here, so the outputs need no separate clear pass.
#include "hpb_common.cuh"

namespace {

struct RasterParams {
    const HpbMeshDev *meshes;
    const int32_t *mesh_ids;
    const float *TCO;
    const float *K;
    const float *ambient;
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
    unsigned long long *vis;   // [gridDim.x][h*w]
    HpbSVert *vert_scratch;    // [gridDim.x][max_nv] when vertices do not fit in shared memory
    int max_nv;
    int verts_in_smem;
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

struct TriSetup {
    int x0, y0, x1, y1, x2, y2;
    int i0, i1, i2;
    long long area2;
};

// Orients the triangle so area2 > 0.  Returns false for near-clipped or degenerate triangles.
__device__ __forceinline__ bool setup_tri(const HpbSVert *sv, int4 f, TriSetup &t, float &iz0, float &iz1, float &iz2) {
    const HpbSVert a = sv[f.x], b = sv[f.y], c = sv[f.z];
    if (a.iz == 0.0f || b.iz == 0.0f || c.iz == 0.0f) return false;
    long long area2 = (long long)(b.x - a.x) * (long long)(c.y - a.y) - (long long)(c.x - a.x) * (long long)(b.y - a.y);
    if (area2 == 0) return false;
    t.x0 = a.x; t.y0 = a.y; t.i0 = f.x; iz0 = a.iz;
    if (area2 < 0) {
        t.x1 = c.x; t.y1 = c.y; t.i1 = f.z; iz1 = c.iz;
        t.x2 = b.x; t.y2 = b.y; t.i2 = f.y; iz2 = b.iz;
        area2 = -area2;
    } else {
        t.x1 = b.x; t.y1 = b.y; t.i1 = f.y; iz1 = b.iz;
        t.x2 = c.x; t.y2 = c.y; t.i2 = f.z; iz2 = c.iz;
    }
    t.area2 = area2;
    return true;
}

__device__ __forceinline__ float depth_from_iz(const RasterParams &p, float iz) { return (p.inv_near - iz) * p.cd; }

// Phase B for one triangle; I = int (all intermediates provably fit 32 bits) or long long.
template <typename I>
__device__ __forceinline__ void raster_tri(const RasterParams &p, const TriSetup &t, float iz0, float iz1, float iz2,
                                           int jx0, int jx1, int jy0, int jy1, unsigned tri_id,
                                           unsigned long long *vis) {
    const int b0 = edge_bias(t.x2 - t.x1, t.y2 - t.y1);
    const int b1 = edge_bias(t.x0 - t.x2, t.y0 - t.y2);
    const int b2 = edge_bias(t.x1 - t.x0, t.y1 - t.y0);
    const int px0 = jx0 * HPB_SUBPIX + 128, py0 = jy0 * HPB_SUBPIX + 128;
    // edge functions at the first pixel centre and their per-pixel steps
    I e0r = (I)(t.x2 - t.x1) * (I)(py0 - t.y1) - (I)(t.y2 - t.y1) * (I)(px0 - t.x1);
    I e1r = (I)(t.x0 - t.x2) * (I)(py0 - t.y2) - (I)(t.y0 - t.y2) * (I)(px0 - t.x2);
    I e2r = (I)(t.x1 - t.x0) * (I)(py0 - t.y0) - (I)(t.y1 - t.y0) * (I)(px0 - t.x0);
    const I sx0 = -(I)(t.y2 - t.y1) * HPB_SUBPIX, sy0 = (I)(t.x2 - t.x1) * HPB_SUBPIX;
    const I sx1 = -(I)(t.y0 - t.y2) * HPB_SUBPIX, sy1 = (I)(t.x0 - t.x2) * HPB_SUBPIX;
    const I sx2 = -(I)(t.y1 - t.y0) * HPB_SUBPIX, sy2 = (I)(t.x1 - t.x0) * HPB_SUBPIX;
    const float inv = 1.0f / (float)t.area2;
    for (int py = jy0; py <= jy1; ++py) {
        I e0 = e0r, e1 = e1r, e2 = e2r;
        unsigned long long *row = vis + (long long)py * p.w;
        for (int px = jx0; px <= jx1; ++px) {
            if (((e0 + b0) | (e1 + b1) | (e2 + b2)) >= 0) {
                const float l0 = (float)e0 * inv, l1 = (float)e1 * inv, l2 = (float)e2 * inv;
                const float iz = fmaf(l2, iz2, fmaf(l1, iz1, l0 * iz0));
                float d = depth_from_iz(p, iz);
                if (d <= 1.0f) {
                    if (d < 0.0f) d = 0.0f;
                    const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | tri_id;
                    atomicMin(row + px, key);
                }
            }
            e0 += sx0; e1 += sx1; e2 += sx2;
        }
        e0r += sy0; e1r += sy1; e2r += sy2;
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

__device__ __forceinline__ float3 fetch_texel(const HpbMeshDev &m, int lvl, int x, int y) {
    const int W = m.tex_w[lvl], H = m.tex_h[lvl];
    x %= W; if (x < 0) x += W;
    y %= H; if (y < 0) y += H;
    const uchar4 c = __ldg(m.tex + m.tex_off[lvl] + (long long)y * W + x);
    return make_float3((float)c.x, (float)c.y, (float)c.z);
}

__device__ __forceinline__ float3 sample_bilinear(const HpbMeshDev &m, int lvl, float u, float v) {
    const float W = (float)m.tex_w[lvl], H = (float)m.tex_h[lvl];
    const float x = fmaf(u, W, -0.5f);
    const float y = fmaf(1.0f - v, H, -0.5f);
    float xf = floorf(x), yf = floorf(y);
    const float fx = x - xf, fy = y - yf;
    if (!(xf > -1.0e9f)) xf = -1.0e9f;
    if (xf > 1.0e9f) xf = 1.0e9f;
    if (!(yf > -1.0e9f)) yf = -1.0e9f;
    if (yf > 1.0e9f) yf = 1.0e9f;
    const int x0 = (int)xf, y0 = (int)yf;
    const float3 c00 = fetch_texel(m, lvl, x0, y0), c01 = fetch_texel(m, lvl, x0 + 1, y0);
    const float3 c10 = fetch_texel(m, lvl, x0, y0 + 1), c11 = fetch_texel(m, lvl, x0 + 1, y0 + 1);
    float3 o;
    {
        const float top = fmaf(fx, c01.x - c00.x, c00.x), bot = fmaf(fx, c11.x - c10.x, c10.x);
        o.x = fmaf(fy, bot - top, top);
    }
    {
        const float top = fmaf(fx, c01.y - c00.y, c00.y), bot = fmaf(fx, c11.y - c10.y, c10.y);
        o.y = fmaf(fy, bot - top, top);
    }
    {
        const float top = fmaf(fx, c01.z - c00.z, c00.z), bot = fmaf(fx, c11.z - c10.z, c10.z);
        o.z = fmaf(fy, bot - top, top);
    }
    return o;
}

// Panda3D's 32^3 "normal map" lookup (renderer/utils.py:63-79): texel k = floor(k*255/32), repeat wrap, linear.
__device__ __forceinline__ float encode_normal(float c) {
    const float s = c - floorf(c);
    const float t = fmaf(s, 32.0f, -0.5f);
    const float kf = floorf(t);
    const float f = t - kf;
    const int k0 = ((int)kf) & 31;
    const int k1 = (k0 + 1) & 31;
    const float T0 = (float)((k0 * 255) >> 5);
    const float T1 = (float)((k1 * 255) >> 5);
    const float val = fmaf(f, T1 - T0, T0);
    return floorf(val + 0.5f) / 255.0f;
}

__device__ __forceinline__ float quant8(float c) {
    float q = floorf(c + 0.5f);
    if (!(q > 0.0f)) q = 0.0f;
    if (q > 255.0f) q = 255.0f;
    return q / 255.0f;
}

__device__ __forceinline__ float3 eye_normal(const float *T, const float *n) {
    const float nx = __ldg(n), ny = __ldg(n + 1), nz = __ldg(n + 2);
    float ex = fmaf(T[2], nz, fmaf(T[1], ny, T[0] * nx));
    float ey = fmaf(T[6], nz, fmaf(T[5], ny, T[4] * nx));
    float ez = fmaf(T[10], nz, fmaf(T[9], ny, T[8] * nx));
    const float l2 = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
    if (l2 > 0.0f) {
        const float r = 1.0f / sqrtf(l2);
        ex *= r; ey *= r; ez *= r;
    }
    return make_float3(ex, ey, ez);
}

__global__ void __launch_bounds__(1024, 1) hpb_raster_kernel(const RasterParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float sT[16];
    __shared__ float sK[4];
    __shared__ float sAmb[3];
    __shared__ int sFinite;

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int npix = p.h * p.w;
    unsigned long long *vis = p.vis + (size_t)blockIdx.x * npix;
    HpbSVert *sv = p.verts_in_smem ? reinterpret_cast<HpbSVert *>(smem_raw)
                                   : p.vert_scratch + (size_t)blockIdx.x * p.max_nv;

    for (int hyp = blockIdx.x; hyp < p.b; hyp += gridDim.x) {
        const HpbMeshDev &m = p.meshes[p.mesh_ids[hyp]];
        if (tid == 0) sFinite = 1;
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
            if (a > 1.0f) a = 1.0f;
            if (!(a > 0.0f)) a = 0.0f;
            sAmb[tid - 25] = a;
        }
        __syncthreads();
        const bool finite = sFinite != 0;
        const int nv = m.nv, nf = m.nf;

        if (finite) {
            // ---------------- phase A: vertex stage ----------------
            const float fx = sK[0], fy = sK[1], cx = sK[2], cy = sK[3];
            for (int i = tid; i < nv; i += nthr) {
                const float x = __ldg(m.pos + 3 * i), y = __ldg(m.pos + 3 * i + 1), z = __ldg(m.pos + 3 * i + 2);
                const float X = fmaf(sT[2], z, fmaf(sT[1], y, fmaf(sT[0], x, sT[3])));
                const float Y = fmaf(sT[6], z, fmaf(sT[5], y, fmaf(sT[4], x, sT[7])));
                const float Z = fmaf(sT[10], z, fmaf(sT[9], y, fmaf(sT[8], x, sT[11])));
                HpbSVert o;
                if (!(Z >= p.z_near)) {
                    o.x = 0; o.y = 0; o.iz = 0.0f;
                } else {
                    const float iz = 1.0f / Z;
                    o.iz = iz;
                    o.x = snap_fixed(fmaf(fx, X * iz, cx));
                    o.y = snap_fixed(fmaf(fy, Y * iz, cy));
                }
                sv[i] = o;
            }
            __syncthreads();

            // ---------------- phase B: triangle stage ----------------
            for (int t = tid; t < nf; t += nthr) {
                const int4 f = __ldg(m.faces + t);
                TriSetup ts;
                float iz0, iz1, iz2;
                if (!setup_tri(sv, f, ts, iz0, iz1, iz2)) continue;
                const int mnx = min(ts.x0, min(ts.x1, ts.x2)), mxx = max(ts.x0, max(ts.x1, ts.x2));
                const int mny = min(ts.y0, min(ts.y1, ts.y2)), mxy = max(ts.y0, max(ts.y1, ts.y2));
                const int jx0 = max(ceil_div_pix(mnx), 0), jx1 = min(floor_div_pix(mxx), p.w - 1);
                const int jy0 = max(ceil_div_pix(mny), 0), jy1 = min(floor_div_pix(mxy), p.h - 1);
                if (jx0 > jx1 || jy0 > jy1) continue;
                // 32-bit path: |edge function| <= 2 * bbox_w * bbox_h (fixed point) over the (clamped) bbox
                const long long bw = (long long)mxx - mnx + 2 * HPB_SUBPIX, bh = (long long)mxy - mny + 2 * HPB_SUBPIX;
                if (bw * bh < (1ll << 29))
                    raster_tri<int>(p, ts, iz0, iz1, iz2, jx0, jx1, jy0, jy1, (unsigned)t, vis);
                else
                    raster_tri<long long>(p, ts, iz0, iz1, iz2, jx0, jx1, jy0, jy1, (unsigned)t, vis);
            }
            __threadfence();
            __syncthreads();
        }

        // ---------------- phase C: resolve ----------------
        const size_t oi = (size_t)(hyp / p.views), ov = (size_t)(hyp % p.views) * p.view_stride;
        float *rgb = (p.flags & HPB_RENDER_RGB) ? p.rgb + oi * p.rgb_bs + ov : nullptr;
        float *nrm = (p.flags & HPB_RENDER_NORMALS) ? p.nrm + oi * p.nrm_bs + ov : nullptr;
        float *dep = (p.flags & HPB_RENDER_DEPTH) ? p.depth + oi * p.depth_bs + ov : nullptr;
        uint8_t *msk = (p.flags & HPB_RENDER_MASK) ? p.mask + (size_t)hyp * p.mask_bs : nullptr;  // never view-interleaved
        for (int pix = tid; pix < npix; pix += nthr) {
            unsigned long long key = HPB_VIS_EMPTY;
            if (finite) {
                key = __ldcg(vis + pix);
                if (key != HPB_VIS_EMPTY) __stcg(vis + pix, HPB_VIS_EMPTY);  // re-arm for the next scene
            }
            float r = 0.f, g = 0.f, bl = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f, z = 0.f;
            if (key != HPB_VIS_EMPTY) {
                const int py = pix / p.w, px = pix - py * p.w;
                const unsigned t = (unsigned)(key & 0xffffffffull);
                const int4 f = __ldg(m.faces + t);
                TriSetup ts;
                float iz0, iz1, iz2;
                setup_tri(sv, f, ts, iz0, iz1, iz2);
                const int fxp = px * HPB_SUBPIX + 128, fyp = py * HPB_SUBPIX + 128;
                const long long e0 = (long long)(ts.x2 - ts.x1) * (fyp - ts.y1) - (long long)(ts.y2 - ts.y1) * (fxp - ts.x1);
                const long long e1 = (long long)(ts.x0 - ts.x2) * (fyp - ts.y2) - (long long)(ts.y0 - ts.y2) * (fxp - ts.x2);
                const long long e2 = (long long)(ts.x1 - ts.x0) * (fyp - ts.y0) - (long long)(ts.y1 - ts.y0) * (fxp - ts.x0);
                const float inv = 1.0f / (float)ts.area2;
                const float l0 = (float)e0 * inv, l1 = (float)e1 * inv, l2 = (float)e2 * inv;
                const float w0 = l0 * iz0, w1 = l1 * iz1, w2 = l2 * iz2;
                const float iz = fmaf(l2, iz2, fmaf(l1, iz1, w0));
                const float d = __uint_as_float((unsigned)(key >> 32));
                z = p.a_f / (d - p.b_f);
                if (d > p.eps_hi) z = 0.0f;
                const float s = 1.0f / iz;
                const float p0 = w0 * s, p1 = w1 * s, p2 = w2 * s;
                if (nrm) {
                    const float3 a = eye_normal(sT, m.nrm + 3 * ts.i0);
                    const float3 b = eye_normal(sT, m.nrm + 3 * ts.i1);
                    const float3 c = eye_normal(sT, m.nrm + 3 * ts.i2);
                    float nx = fmaf(p2, c.x, fmaf(p1, b.x, p0 * a.x));
                    float ny = fmaf(p2, c.y, fmaf(p1, b.y, p0 * a.y));
                    float nz = fmaf(p2, c.z, fmaf(p1, b.z, p0 * a.z));
                    const float len2 = fmaf(nz, nz, fmaf(ny, ny, nx * nx));
                    if (len2 > 0.0f) {
                        const float rl = 1.0f / sqrtf(len2);
                        nx *= rl; ny *= rl; nz *= rl;
                    }
                    n0 = encode_normal(nx);
                    n1 = encode_normal(nz);
                    n2 = encode_normal(-ny);
                }
                if (rgb) {
                    float3 col = make_float3(255.0f, 255.0f, 255.0f);
                    if (m.tex && m.uv) {
                        const float u0 = __ldg(m.uv + 2 * ts.i0), v0 = __ldg(m.uv + 2 * ts.i0 + 1);
                        const float u1 = __ldg(m.uv + 2 * ts.i1), v1 = __ldg(m.uv + 2 * ts.i1 + 1);
                        const float u2 = __ldg(m.uv + 2 * ts.i2), v2 = __ldg(m.uv + 2 * ts.i2 + 1);
                        const float u = fmaf(p2, u2, fmaf(p1, u1, p0 * u0));
                        const float v = fmaf(p2, v2, fmaf(p1, v1, p0 * v0));
                        const float sc = (float)HPB_SUBPIX * inv;
                        const float dl0x = (float)(-(ts.y2 - ts.y1)) * sc, dl0y = (float)(ts.x2 - ts.x1) * sc;
                        const float dl1x = (float)(-(ts.y0 - ts.y2)) * sc, dl1y = (float)(ts.x0 - ts.x2) * sc;
                        const float dl2x = (float)(-(ts.y1 - ts.y0)) * sc, dl2y = (float)(ts.x1 - ts.x0) * sc;
                        const float g0x = dl0x * iz0, g1x = dl1x * iz1, g2x = dl2x * iz2;
                        const float g0y = dl0y * iz0, g1y = dl1y * iz1, g2y = dl2y * iz2;
                        const float dDx = g0x + g1x + g2x, dDy = g0y + g1y + g2y;
                        const float dNux = fmaf(g2x, u2, fmaf(g1x, u1, g0x * u0));
                        const float dNuy = fmaf(g2y, u2, fmaf(g1y, u1, g0y * u0));
                        const float dNvx = fmaf(g2x, v2, fmaf(g1x, v1, g0x * v0));
                        const float dNvy = fmaf(g2y, v2, fmaf(g1y, v1, g0y * v0));
                        const float W0 = (float)m.tex_w[0], H0 = (float)m.tex_h[0];
                        const float ax = (dNux - u * dDx) * s * W0, bx = (dNvx - v * dDx) * s * H0;
                        const float ay = (dNuy - u * dDy) * s * W0, by = (dNvy - v * dDy) * s * H0;
                        const float r2x = fmaf(ax, ax, bx * bx), r2y = fmaf(ay, ay, by * by);
                        const float rho2 = r2x > r2y ? r2x : r2y;
                        float lod = 0.0f;
                        if (rho2 > 1.0f && rho2 < 1.0e30f) lod = 0.5f * hp_log2(rho2);
                        const float maxl = (float)(m.tex_levels - 1);
                        if (lod > maxl) lod = maxl;
                        const float lf = floorf(lod);
                        const float fl = lod - lf;
                        const int li = (int)lf;
                        const float3 ca = sample_bilinear(m, li, u, v);
                        if (fl > 0.0f && li + 1 < m.tex_levels) {
                            const float3 cb = sample_bilinear(m, li + 1, u, v);
                            col.x = fmaf(fl, cb.x - ca.x, ca.x);
                            col.y = fmaf(fl, cb.y - ca.y, ca.y);
                            col.z = fmaf(fl, cb.z - ca.z, ca.z);
                        } else {
                            col = ca;
                        }
                    } else if (m.vcol) {
                        const uchar4 c0 = __ldg(m.vcol + ts.i0), c1 = __ldg(m.vcol + ts.i1), c2 = __ldg(m.vcol + ts.i2);
                        col.x = fmaf(p2, (float)c2.x, fmaf(p1, (float)c1.x, p0 * (float)c0.x));
                        col.y = fmaf(p2, (float)c2.y, fmaf(p1, (float)c1.y, p0 * (float)c0.y));
                        col.z = fmaf(p2, (float)c2.z, fmaf(p1, (float)c1.z, p0 * (float)c0.z));
                    }
                    r = quant8(col.x * sAmb[0]);
                    g = quant8(col.y * sAmb[1]);
                    bl = quant8(col.z * sAmb[2]);
                }
            }
            if (rgb) {
                __stcs(rgb + pix, r);
                __stcs(rgb + npix + pix, g);
                __stcs(rgb + 2 * (size_t)npix + pix, bl);
            }
            if (nrm) {
                __stcs(nrm + pix, n0);
                __stcs(nrm + npix + pix, n1);
                __stcs(nrm + 2 * (size_t)npix + pix, n2);
            }
            if (dep) __stcs(dep + pix, z);
            if (msk) msk[pix] = z > 0.0f ? 1 : 0;
        }
        __syncthreads();
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
                      int views, int64_t view_stride, cudaStream_t stream) {
    if (b == 0) return HPB_OK;
    const int npix = h * w;
    // persistent grid: one CTA per SM (1024 threads, vertices in shared memory), never more CTAs than scenes
    const int grid = b < ctx->sm_count ? b : ctx->sm_count;

    const size_t smem_need = (size_t)ctx->max_nv * sizeof(HpbSVert);
    const size_t smem_cap = (size_t)ctx->max_smem_optin > 4096 ? (size_t)ctx->max_smem_optin - 2048 : 0;
    const int verts_in_smem = smem_need <= smem_cap;

    // workspace (grown on demand; the visibility buffer is armed once and re-armed by the kernel itself)
    const size_t vis_need = (size_t)ctx->sm_count * npix;
    if (ctx->vis_elems < vis_need) {
        if (ctx->vis) HPB_CUDA_OK(cudaFree(ctx->vis));
        ctx->vis = nullptr;
        ctx->vis_elems = 0;
        HPB_CUDA_OK(cudaMalloc(&ctx->vis, vis_need * sizeof(unsigned long long)));
        ctx->vis_elems = vis_need;
        hpb_fill_u64_kernel<<<ctx->sm_count * 4, 256, 0, stream>>>(ctx->vis, vis_need, HPB_VIS_EMPTY);
        HPB_CUDA_OK(cudaGetLastError());
        ctx->launches++;
    }
    if (!verts_in_smem) {
        const size_t need = (size_t)ctx->sm_count * ctx->max_nv;
        if (ctx->vert_scratch_elems < need) {
            if (ctx->vert_scratch) HPB_CUDA_OK(cudaFree(ctx->vert_scratch));
            ctx->vert_scratch = nullptr;
            ctx->vert_scratch_elems = 0;
            HPB_CUDA_OK(cudaMalloc(&ctx->vert_scratch, need * sizeof(HpbSVert)));
            ctx->vert_scratch_elems = need;
        }
    }

    RasterParams p;
    p.meshes = ctx->meshes_dev;
    p.mesh_ids = mesh_ids;
    p.TCO = TCO;
    p.K = K;
    p.ambient = ambient;
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
    p.vert_scratch = ctx->vert_scratch;
    p.max_nv = ctx->max_nv;
    p.verts_in_smem = verts_in_smem;

    const size_t smem = verts_in_smem ? smem_need : 0;
    HPB_CUDA_OK(cudaFuncSetAttribute(hpb_raster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hpb_raster_kernel<<<grid, 1024, smem, stream>>>(p);
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}
