/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the rasterisation semantics of
 * happypose's Panda3D renderer.  Nothing under happypose_b200/ may include, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * use it, as the checker / host-core baseline.
 *
 * PARITY: Panda3D 1.10.13 + OpenGL cannot be installed here, so the renderer itself never runs next to this
 * oracle.  It is pinned to the only real Panda3D pixels the reference ships, the golden figures
 * tests/data/panda3d_obj_{batch,scene}_render.png (written by tests/test_batch_renderer_panda3d.py:148-163):
 * tests/test_oracle_figure_pin.py renders the same scene with this oracle, resamples it the way matplotlib
 * built the figure (oracle/figure_pin.py) and compares: silhouette IoU 0.9995 / centroid within 0.01 px /
 * identical extents, depth panel within 1 grey level (corr 0.98), normal channels R,G corr 0.96 and B corr
 * +0.77 (the opposite B sign convention gives -0.77), rgb channel means within 1 level (corr 0.9).  That
 * pins projection, flip, depth read-back, normal colour convention and texture orientation at FIGURE
 * resolution (one figure pixel ~ 8 image pixels); per-pixel texture filtering, MSAA and near-plane clipping
 * remain unpinned (below).  The reference's structural test assertions (:71-242) are re-asserted on this
 * oracle in tests/test_oracle_raster.py.  The semantics restated:
 *
 *   projection / pixel centres  toolbox/renderer/types.py:111-137 (set_lens_parameters) and
 *                               :254-293 (flipud read-back): pixel (i,j) samples the ray through
 *                               (u,v) = (j+0.5, i+0.5), u = fx*X/Z + cx, v = fy*Y/Z + cy (OpenCV
 *                               camera axes, skew ignored, only K[0,0],K[1,1],K[0,2],K[1,2] read).
 *   near / far                  types.py:96-97 (z_near = 0.1, z_far = 10).
 *   culling                     none, two sided (panda3d_scene_renderer.py:102).  Implemented as: back faces are
 *                               skipped only on meshes proven closed and consistently oriented (cull_sign != 0,
 *                               computed by oracle/raster.py:closed_surface_sign) in scenes the near plane does not
 *                               cut -- there every pixel a back face covers is also covered by a nearer front face,
 *                               so the two-sided image is unchanged (tests/test_oracle_raster.py checks it).
 *   depth                       GL depth d in [0,1], z = a/(d-b), a = 1/(1/far-1/near), b = -a/near,
 *                               d > 1-0.001 -> 0 (toolbox/renderer/utils.py:46-60).
 *   mask                        depth > 0 (panda3d_scene_renderer.py:360-367).
 *   rgb                         black background (:72); ambient-only lighting on the hot path
 *                               (megapose/models/pose_rigid.py:415-420): colour = texture (mip-mapped,
 *                               :68) * min(sum(ambient),1); 8-bit framebuffer, returned as k/255
 *                               (panda3d_batch_renderer.py:249).
 *   normals                     eye-space unit normal (object normal interpolated over the triangle, rotated by
 *                               R_CO, normalised per pixel) looked up in a 32^3 RGB texture whose texel
 *                               (x,y,z) = floor((x,y,z)*255/32), repeat wrap, linear filter
 *                               (toolbox/renderer/utils.py:63-79, panda3d_scene_renderer.py:221-230),
 *                               in Panda camera axes (x right, y forward, z up).
 *   non-finite pose / K         all-zero images (panda3d_batch_renderer.py:81-111).
 *
 * Stated deviations from a GL pipeline (also in DESIGN.md): one sample per pixel (the reference
 * asks for 4x MSAA, :70-71), isotropic trilinear filtering (reference: anisotropic 16, :69),
 * triangles with a vertex in front of the near plane are dropped instead of clipped, vertices are
 * snapped to a 1/256-pixel grid and coverage uses exact integer edge functions with a top-left
 * rule; depth ties go to the lower triangle index (draw order under GL_LESS); window depth is evaluated
 * from a per-triangle plane anchored at the first pixel of the triangle's clamped bounding box.
 *
 * All float arithmetic is written with explicit fmaf()/IEEE ops and this file must be built with
 * -ffp-contract=off so the CUDA kernels (compiled with -fmad=false) can follow it bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HPO_FLAG_RGB 1u
#define HPO_FLAG_NORMALS 2u
#define HPO_FLAG_DEPTH 4u
#define HPO_FLAG_MASK 8u

#define SUBPIX_BITS 8
#define SUBPIX 256
#define GUARD 4194304 /* 2^22 fixed-point units = 16384 px */

typedef struct {
    int64_t n_verts;
    int64_t n_faces;
    const float *pos;     /* [nv,3] metres */
    const float *nrm;     /* [nv,3] unit, object frame */
    const float *uv;      /* [nv,2] or NULL */
    const uint8_t *vcol;  /* [nv,4] RGBA or NULL */
    const int32_t *faces; /* [nf,3] */
    const uint8_t *tex;   /* RGBA8 mip chain or NULL */
    int32_t tex_levels;
    const int32_t *tex_w; /* [levels] */
    const int32_t *tex_h;
    const int64_t *tex_off; /* [levels] offset in texels */
    int32_t cull_sign;      /* sign of the screen-space area2 of front faces of a closed surface, 0 = two-sided */
} hpo_mesh;

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* log2 of a positive finite float, identical on CPU and GPU: exponent + cubic in the mantissa. */
static inline float hp_log2(float x) {
    uint32_t u = f2u(x);
    int e = (int)((u >> 23) & 0xff) - 127;
    float m = u2f((u & 0x007fffffu) | 0x3f800000u) - 1.0f; /* [0,1) */
    /* minimax-ish cubic for log2(1+m), p(1) = 1, max err 8.8e-4 */
    float p = fmaf(m, 0.15922009f, -0.58208540f);
    p = fmaf(m, p, 1.42286531f);
    p = m * p;
    return (float)e + p;
}

static inline int snap(float u) {
    float s = rintf(u * (float)SUBPIX);
    if (!(s > -(float)GUARD)) s = -(float)GUARD; /* also catches NaN */
    if (s > (float)GUARD) s = (float)GUARD;
    return (int)s;
}

static inline int ceil_div_pix(int v) { /* smallest j with j*256+128 >= v */
    int t = v - 128;
    return (t + 255) >> SUBPIX_BITS; /* arithmetic shift: floor((t+255)/256) = ceil(t/256) */
}
static inline int floor_div_pix(int v) { /* largest j with j*256+128 <= v */
    return (v - 128) >> SUBPIX_BITS;
}

typedef struct {
    int x0, y0, x1, y1, x2, y2; /* oriented so that area2 > 0 */
    float iz0, iz1, iz2;
    int64_t area2;
    int orig_sign; /* sign of area2 in the face's own winding */
    int b0, b1, b2; /* tie-rule bias per edge function */
} tri_setup;

static inline int edge_bias(int dx, int dy) { return (dy < 0 || (dy == 0 && dx > 0)) ? 0 : -1; }

static int setup_tri(const int *sx, const int *sy, const float *viz, const uint8_t *vbad, const int32_t *f, tri_setup *t) {
    int i0 = f[0], i1 = f[1], i2 = f[2];
    if (vbad[i0] | vbad[i1] | vbad[i2]) return 0;
    int x0 = sx[i0], y0 = sy[i0], x1 = sx[i1], y1 = sy[i1], x2 = sx[i2], y2 = sy[i2];
    int64_t area2 = (int64_t)(x1 - x0) * (int64_t)(y2 - y0) - (int64_t)(x2 - x0) * (int64_t)(y1 - y0);
    if (area2 == 0) return 0;
    t->orig_sign = area2 < 0 ? -1 : 1;
    if (area2 < 0) {
        int tx = x1, ty = y1, ti = i1;
        x1 = x2; y1 = y2; i1 = i2;
        x2 = tx; y2 = ty; i2 = ti;
        area2 = -area2;
    }
    t->x0 = x0; t->y0 = y0; t->x1 = x1; t->y1 = y1; t->x2 = x2; t->y2 = y2;
    t->iz0 = viz[i0]; t->iz1 = viz[i1]; t->iz2 = viz[i2];
    t->area2 = area2;
    t->b0 = edge_bias(x2 - x1, y2 - y1);
    t->b1 = edge_bias(x0 - x2, y0 - y2);
    t->b2 = edge_bias(x1 - x0, y1 - y0);
    return 1;
}

static inline void edges_at(const tri_setup *t, int px, int py, int64_t *e0, int64_t *e1, int64_t *e2) {
    *e0 = (int64_t)(t->x2 - t->x1) * (int64_t)(py - t->y1) - (int64_t)(t->y2 - t->y1) * (int64_t)(px - t->x1);
    *e1 = (int64_t)(t->x0 - t->x2) * (int64_t)(py - t->y2) - (int64_t)(t->y0 - t->y2) * (int64_t)(px - t->x2);
    *e2 = (int64_t)(t->x1 - t->x0) * (int64_t)(py - t->y0) - (int64_t)(t->y1 - t->y0) * (int64_t)(px - t->x0);
}

static inline void fetch_texel(const hpo_mesh *m, int lvl, int x, int y, float *c) {
    int W = m->tex_w[lvl], H = m->tex_h[lvl];
    x %= W; if (x < 0) x += W;
    y %= H; if (y < 0) y += H;
    const uint8_t *p = m->tex + 4 * (m->tex_off[lvl] + (int64_t)y * W + x);
    c[0] = (float)p[0]; c[1] = (float)p[1]; c[2] = (float)p[2];
}

static void sample_bilinear(const hpo_mesh *m, int lvl, float u, float v, float *out) {
    float W = (float)m->tex_w[lvl], H = (float)m->tex_h[lvl];
    float x = fmaf(u, W, -0.5f);
    float y = fmaf(1.0f - v, H, -0.5f);
    float xf = floorf(x), yf = floorf(y);
    float fx = x - xf, fy = y - yf;
    /* keep the int conversion defined for wild uv */
    if (!(xf > -1.0e9f)) xf = -1.0e9f;
    if (xf > 1.0e9f) xf = 1.0e9f;
    if (!(yf > -1.0e9f)) yf = -1.0e9f;
    if (yf > 1.0e9f) yf = 1.0e9f;
    int x0 = (int)xf, y0 = (int)yf;
    float c00[3], c01[3], c10[3], c11[3];
    fetch_texel(m, lvl, x0, y0, c00);
    fetch_texel(m, lvl, x0 + 1, y0, c01);
    fetch_texel(m, lvl, x0, y0 + 1, c10);
    fetch_texel(m, lvl, x0 + 1, y0 + 1, c11);
    for (int k = 0; k < 3; ++k) {
        float top = fmaf(fx, c01[k] - c00[k], c00[k]);
        float bot = fmaf(fx, c11[k] - c10[k], c10[k]);
        out[k] = fmaf(fy, bot - top, top);
    }
}

static inline float encode_normal(float c) {
    float s = c - floorf(c);
    float t = fmaf(s, 32.0f, -0.5f);
    float kf = floorf(t);
    float f = t - kf;
    int k0 = ((int)kf) & 31;
    int k1 = (k0 + 1) & 31;
    float T0 = (float)((k0 * 255) >> 5);
    float T1 = (float)((k1 * 255) >> 5);
    float val = fmaf(f, T1 - T0, T0);
    return floorf(val + 0.5f) / 255.0f;
}

static inline float quant8(float c) {
    float q = floorf(c + 0.5f);
    if (!(q > 0.0f)) q = 0.0f;
    if (q > 255.0f) q = 255.0f;
    return q / 255.0f;
}

/* Render one hypothesis.  Outputs are CHW float32 planes (any may be NULL); mask is uint8 0/1. */
static int render_one(const hpo_mesh *m, const float *T, const float *K, int h, int w, float znear, float zfar,
                      const float *ambient, const float *lights, int n_lights, uint32_t flags, float *rgb, float *nrm_out, float *depth,
                      uint8_t *mask) {
    const int64_t npix = (int64_t)h * w;
    if (rgb) memset(rgb, 0, sizeof(float) * 3 * npix);
    if (nrm_out) memset(nrm_out, 0, sizeof(float) * 3 * npix);
    if (depth) memset(depth, 0, sizeof(float) * npix);
    if (mask) memset(mask, 0, npix);
    for (int i = 0; i < 16; ++i) if (!isfinite(T[i])) return 0;
    for (int i = 0; i < 9; ++i) if (!isfinite(K[i])) return 0;

    const int64_t nv = m->n_verts, nf = m->n_faces;
    int *sx = (int *)malloc(sizeof(int) * nv), *sy = (int *)malloc(sizeof(int) * nv);
    float *viz = (float *)malloc(sizeof(float) * nv);
    uint8_t *vbad = (uint8_t *)malloc(nv);
    uint64_t *zb = (uint64_t *)malloc(sizeof(uint64_t) * npix);
    if (!sx || !sy || !viz || !vbad || !zb) return -1;
    memset(zb, 0xff, sizeof(uint64_t) * npix);

    const float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const float inv_near = 1.0f / znear, inv_far = 1.0f / zfar;
    const float cd = 1.0f / (inv_near - inv_far);
    const double a_d = 1.0 / (1.0 / (double)zfar - 1.0 / (double)znear);
    const float a_f = (float)a_d, b_f = (float)(-a_d / (double)znear);
    const float eps_hi = (float)(1.0 - 0.001);

    int any_clipped = 0;
    for (int64_t i = 0; i < nv; ++i) {
        const float x = m->pos[3 * i], y = m->pos[3 * i + 1], z = m->pos[3 * i + 2];
        const float X = fmaf(T[2], z, fmaf(T[1], y, fmaf(T[0], x, T[3])));
        const float Y = fmaf(T[6], z, fmaf(T[5], y, fmaf(T[4], x, T[7])));
        const float Z = fmaf(T[10], z, fmaf(T[9], y, fmaf(T[8], x, T[11])));
        if (!(Z >= znear)) { vbad[i] = 1; sx[i] = sy[i] = 0; viz[i] = 0.0f; any_clipped = 1; continue; }
        vbad[i] = 0;
        const float iz = 1.0f / Z;
        viz[i] = iz;
        sx[i] = snap(fmaf(fx, X * iz, cx));
        sy[i] = snap(fmaf(fy, Y * iz, cy));
    }

    /* pass 1: visibility (depth bits << 32 | triangle id << 1 | wide), min wins.  Back faces of a closed surface are
     * skipped unless the near plane has cut the surface open. */
    const int cull = (!any_clipped && fx > 0.0f && fy > 0.0f) ? m->cull_sign : 0;
    for (int64_t t = 0; t < nf; ++t) {
        tri_setup ts;
        if (!setup_tri(sx, sy, viz, vbad, m->faces + 3 * t, &ts)) continue;
        if (cull != 0 && ts.orig_sign != cull) continue;
        int mnx = ts.x0 < ts.x1 ? ts.x0 : ts.x1; if (ts.x2 < mnx) mnx = ts.x2;
        int mxx = ts.x0 > ts.x1 ? ts.x0 : ts.x1; if (ts.x2 > mxx) mxx = ts.x2;
        int mny = ts.y0 < ts.y1 ? ts.y0 : ts.y1; if (ts.y2 < mny) mny = ts.y2;
        int mxy = ts.y0 > ts.y1 ? ts.y0 : ts.y1; if (ts.y2 > mxy) mxy = ts.y2;
        int jx0 = ceil_div_pix(mnx), jx1 = floor_div_pix(mxx);
        int jy0 = ceil_div_pix(mny), jy1 = floor_div_pix(mxy);
        if (jx0 < 0) jx0 = 0;
        if (jy0 < 0) jy0 = 0;
        if (jx1 > w - 1) jx1 = w - 1;
        if (jy1 > h - 1) jy1 = h - 1;
        if (jx0 > jx1 || jy0 > jy1) continue;
        /* depth plane anchored at the first pixel of the clamped bounding box */
        float Dc, Dx, Dy;
        {
            int64_t e0o, e1o, e2o;
            edges_at(&ts, jx0 * SUBPIX + 128, jy0 * SUBPIX + 128, &e0o, &e1o, &e2o);
            const float inv = 1.0f / (float)ts.area2;
            const float izc = fmaf((float)e2o * inv, ts.iz2, fmaf((float)e1o * inv, ts.iz1, ((float)e0o * inv) * ts.iz0));
            const float d1 = ts.iz1 - ts.iz0, d2 = ts.iz2 - ts.iz0;
            const float sx1 = (float)(ts.y2 - ts.y0) * 256.0f, sx2 = (float)(ts.y0 - ts.y1) * 256.0f;
            const float sy1 = (float)(ts.x0 - ts.x2) * 256.0f, sy2 = (float)(ts.x1 - ts.x0) * 256.0f;
            const float gx = fmaf(sx2, d2, sx1 * d1) * inv;
            const float gy = fmaf(sy2, d2, sy1 * d1) * inv;
            Dc = (inv_near - izc) * cd;
            Dx = -(gx * cd);
            Dy = -(gy * cd);
        }
        const uint64_t lo = ((uint64_t)(uint32_t)t << 1) | (ts.area2 >= ((int64_t)1 << 31) ? 1u : 0u);
        for (int py = jy0; py <= jy1; ++py) {
            for (int px = jx0; px <= jx1; ++px) {
                int64_t e0, e1, e2;
                edges_at(&ts, px * SUBPIX + 128, py * SUBPIX + 128, &e0, &e1, &e2);
                if ((e0 + ts.b0) < 0 || (e1 + ts.b1) < 0 || (e2 + ts.b2) < 0) continue;
                float d = fmaf(Dx, (float)(px - jx0), fmaf(Dy, (float)(py - jy0), Dc));
                if (!(d <= 1.0f)) continue;
                if (d < 0.0f) d = 0.0f;
                const uint64_t key = ((uint64_t)f2u(d) << 32) | lo;
                uint64_t *slot = zb + (int64_t)py * w + px;
                if (key < *slot) *slot = key;
            }
        }
    }

    /* pass 2: shade the winning triangle of every covered pixel */
    float amb[3] = {1.0f, 1.0f, 1.0f};
    const int lit = lights != NULL && n_lights > 0;
    if (ambient) for (int k = 0; k < 3; ++k) { amb[k] = ambient[k]; if (!lit && amb[k] > 1.0f) amb[k] = 1.0f; if (!(amb[k] > 0.0f)) amb[k] = 0.0f; }
    /* point / directional lights (render_normals=False light rig, pose_rigid.py:105-141,421-422): given in the world =
     * object frame, moved into the camera frame once per scene; per-pixel Lambert, no attenuation, sum clamped at 1 */
    float Lc[8][8];
    for (int i = 0; lit && i < n_lights; ++i) {
        const float *L = lights + 8 * i;
        const float x = L[1], y = L[2], z = L[3];
        float cx3 = fmaf(T[2], z, fmaf(T[1], y, T[0] * x)), cy3 = fmaf(T[6], z, fmaf(T[5], y, T[4] * x));
        float cz3 = fmaf(T[10], z, fmaf(T[9], y, T[8] * x));
        if (L[0] == 0.0f) { cx3 = cx3 + T[3]; cy3 = cy3 + T[7]; cz3 = cz3 + T[11]; }
        Lc[i][0] = L[0]; Lc[i][1] = cx3; Lc[i][2] = cy3; Lc[i][3] = cz3; Lc[i][4] = L[4]; Lc[i][5] = L[5]; Lc[i][6] = L[6];
    }
    for (int py = 0; py < h; ++py) {
        for (int px = 0; px < w; ++px) {
            const int64_t pi = (int64_t)py * w + px;
            const uint64_t key = zb[pi];
            if (key == ~(uint64_t)0) continue;
            const int64_t t = (int64_t)((key & 0xffffffffu) >> 1);
            /* edge values at the pixel centre in the face's own winding: the weights are ratios, so the common sign
             * of e0, e1, e2 (= the sign of area2) cancels and no area division is needed */
            const int32_t *f = m->faces + 3 * t;
            const int i0 = f[0], i1 = f[1], i2 = f[2];
            const int fxp = px * SUBPIX + 128, fyp = py * SUBPIX + 128;
            const int dx0 = sx[i2] - sx[i1], dy0 = sy[i2] - sy[i1];
            const int dx1 = sx[i0] - sx[i2], dy1 = sy[i0] - sy[i2];
            const int dx2 = sx[i1] - sx[i0], dy2 = sy[i1] - sy[i0];
            const float fe0 = (float)((int64_t)dx0 * (int64_t)(fyp - sy[i1]) - (int64_t)dy0 * (int64_t)(fxp - sx[i1]));
            const float fe1 = (float)((int64_t)dx1 * (int64_t)(fyp - sy[i2]) - (int64_t)dy1 * (int64_t)(fxp - sx[i2]));
            const float fe2 = (float)((int64_t)dx2 * (int64_t)(fyp - sy[i0]) - (int64_t)dy2 * (int64_t)(fxp - sx[i0]));
            const float iz0 = viz[i0], iz1 = viz[i1], iz2 = viz[i2];
            const float w0 = fe0 * iz0, w1 = fe1 * iz1, w2 = fe2 * iz2;
            const float s = 1.0f / ((w0 + w1) + w2);
            const float p0 = w0 * s, p1 = w1 * s, p2 = w2 * s;
            if (depth || mask) {
                const float d = u2f((uint32_t)(key >> 32));
                float z = a_f / (d - b_f);
                if (d > eps_hi) z = 0.0f;
                if (depth) depth[pi] = z;
                if (mask) mask[pi] = z > 0.0f;
            }
            float lit3[3] = {amb[0], amb[1], amb[2]};
            if (nrm_out || lit) {
                /* object-space normal interpolated perspective-correctly, then rotated into the eye frame and
                 * normalised once per pixel (for an orthonormal R identical to interpolating per-vertex eye normals) */
                float ox = 0.0f, oy = 0.0f, oz = 0.0f;
                if (m->nrm) {
                    const float *n0 = m->nrm + 3 * i0, *n1 = m->nrm + 3 * i1, *n2 = m->nrm + 3 * i2;
                    ox = fmaf(p2, n2[0], fmaf(p1, n1[0], p0 * n0[0]));
                    oy = fmaf(p2, n2[1], fmaf(p1, n1[1], p0 * n0[1]));
                    oz = fmaf(p2, n2[2], fmaf(p1, n1[2], p0 * n0[2]));
                }
                float nx = fmaf(T[2], oz, fmaf(T[1], oy, T[0] * ox));
                float ny = fmaf(T[6], oz, fmaf(T[5], oy, T[4] * ox));
                float nz = fmaf(T[10], oz, fmaf(T[9], oy, T[8] * ox));
                const float len2 = fmaf(nz, nz, fmaf(ny, ny, nx * nx));
                if (len2 > 0.0f) { const float r = 1.0f / sqrtf(len2); nx *= r; ny *= r; nz *= r; }
                if (nrm_out) {
                    nrm_out[pi] = encode_normal(nx);
                    nrm_out[npix + pi] = encode_normal(nz);
                    nrm_out[2 * npix + pi] = encode_normal(-ny);
                }
                if (lit) {
                    const float dk = u2f((uint32_t)(key >> 32));
                    const float zc = a_f / (dk - b_f);
                    const float X = (((float)px + 0.5f) - K[2]) / K[0] * zc, Y = (((float)py + 0.5f) - K[5]) / K[4] * zc;
                    for (int i = 0; i < n_lights; ++i) {
                        float ndl;
                        if (Lc[i][0] != 0.0f) {
                            ndl = -fmaf(nz, Lc[i][3], fmaf(ny, Lc[i][2], nx * Lc[i][1]));
                        } else {
                            const float lx = Lc[i][1] - X, ly = Lc[i][2] - Y, lz = Lc[i][3] - zc;
                            const float l2 = fmaf(lz, lz, fmaf(ly, ly, lx * lx));
                            ndl = fmaf(nz, lz, fmaf(ny, ly, nx * lx));
                            if (l2 > 0.0f) ndl = ndl * (1.0f / sqrtf(l2));
                        }
                        if (ndl > 0.0f)
                            for (int k = 0; k < 3; ++k) lit3[k] = fmaf(ndl, Lc[i][4 + k], lit3[k]);
                    }
                    for (int k = 0; k < 3; ++k) if (lit3[k] > 1.0f) lit3[k] = 1.0f;
                }
            }
            if (rgb) {
                float col[3] = {255.0f, 255.0f, 255.0f};
                if (m->tex && m->uv) {
                    const float *t0 = m->uv + 2 * i0, *t1 = m->uv + 2 * i1, *t2 = m->uv + 2 * i2;
                    const float u = fmaf(p2, t2[0], fmaf(p1, t1[0], p0 * t0[0]));
                    const float v = fmaf(p2, t2[1], fmaf(p1, t1[1], p0 * t0[1]));
                    /* analytic screen-space derivatives of (u,v) for the mip level, from the per-pixel steps of the
                     * un-normalised perspective weights e_i / z_i */
                    const float g0x = ((float)(-dy0) * 256.0f) * iz0, g1x = ((float)(-dy1) * 256.0f) * iz1, g2x = ((float)(-dy2) * 256.0f) * iz2;
                    const float g0y = ((float)dx0 * 256.0f) * iz0, g1y = ((float)dx1 * 256.0f) * iz1, g2y = ((float)dx2 * 256.0f) * iz2;
                    const float dDx = (g0x + g1x) + g2x, dDy = (g0y + g1y) + g2y;
                    const float dNux = fmaf(g2x, t2[0], fmaf(g1x, t1[0], g0x * t0[0]));
                    const float dNuy = fmaf(g2y, t2[0], fmaf(g1y, t1[0], g0y * t0[0]));
                    const float dNvx = fmaf(g2x, t2[1], fmaf(g1x, t1[1], g0x * t0[1]));
                    const float dNvy = fmaf(g2y, t2[1], fmaf(g1y, t1[1], g0y * t0[1]));
                    const float W0 = (float)m->tex_w[0], H0 = (float)m->tex_h[0];
                    const float ax = (dNux - u * dDx) * s * W0, bx = (dNvx - v * dDx) * s * H0;
                    const float ay = (dNuy - u * dDy) * s * W0, by = (dNvy - v * dDy) * s * H0;
                    const float r2x = fmaf(ax, ax, bx * bx), r2y = fmaf(ay, ay, by * by);
                    const float rho2 = r2x > r2y ? r2x : r2y;
                    float lod = 0.0f;
                    if (rho2 > 1.0f && rho2 < 1.0e30f) lod = 0.5f * hp_log2(rho2);
                    const float maxl = (float)(m->tex_levels - 1);
                    if (lod > maxl) lod = maxl;
                    const float lf = floorf(lod);
                    const float fl = lod - lf;
                    const int li = (int)lf;
                    float ca[3], cb[3];
                    sample_bilinear(m, li, u, v, ca);
                    if (fl > 0.0f && li + 1 < m->tex_levels) {
                        sample_bilinear(m, li + 1, u, v, cb);
                        for (int k = 0; k < 3; ++k) col[k] = fmaf(fl, cb[k] - ca[k], ca[k]);
                    } else {
                        for (int k = 0; k < 3; ++k) col[k] = ca[k];
                    }
                } else if (m->vcol) {
                    const uint8_t *c0 = m->vcol + 4 * i0, *c1 = m->vcol + 4 * i1, *c2 = m->vcol + 4 * i2;
                    for (int k = 0; k < 3; ++k)
                        col[k] = fmaf(p2, (float)c2[k], fmaf(p1, (float)c1[k], p0 * (float)c0[k]));
                }
                for (int k = 0; k < 3; ++k) rgb[k * npix + pi] = quant8(col[k] * lit3[k]);
            }
        }
    }
    free(sx); free(sy); free(viz); free(vbad); free(zb);
    (void)flags;
    return 0;
}

/* Batched entry point.  mesh_ids[b] indexes `meshes`; outputs are [b,3,h,w] / [b,1,h,w].
 * Renders hypotheses n0 <= n < n1 serially; the Python wrapper runs disjoint ranges on host threads
 * (ctypes releases the GIL), mirroring the reference's n_workers pool (panda3d_batch_renderer.py:288-330). */
int hpo_render_batch(const hpo_mesh *meshes, const int32_t *mesh_ids, const float *TCO, const float *K,
                     const float *ambient, const float *lights, int n_lights, int n0, int n1, int h, int w, float znear, float zfar,
                     uint32_t flags, float *rgb, float *normals, float *depth, uint8_t *mask) {
    const int64_t npix = (int64_t)h * w;
    int err = 0;
    for (int n = n0; n < n1; ++n) {
        int r = render_one(meshes + mesh_ids[n], TCO + 16 * n, K + 9 * n, h, w, znear, zfar,
                           ambient ? ambient + 3 * n : NULL, lights ? lights + (size_t)8 * n_lights * n : NULL, n_lights, flags,
                           (flags & HPO_FLAG_RGB) && rgb ? rgb + 3 * npix * n : NULL,
                           (flags & HPO_FLAG_NORMALS) && normals ? normals + 3 * npix * n : NULL,
                           (flags & HPO_FLAG_DEPTH) && depth ? depth + npix * n : NULL,
                           (flags & HPO_FLAG_MASK) && mask ? mask + npix * n : NULL);
        if (r != 0) err = r;
    }
    return err;
}

/* Box-filter mip chain of an RGBA8 image, the same rule the product uses at mesh upload:
 * level l+1 has size max(1, floor(size/2)); each texel is the rounded mean of the 2x2 block
 * (clamped at the edge for odd sizes). */
void hpo_mip_downsample(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh) {
    for (int y = 0; y < dh; ++y) {
        for (int x = 0; x < dw; ++x) {
            int x0 = 2 * x, x1 = 2 * x + 1 < sw ? 2 * x + 1 : sw - 1;
            int y0 = 2 * y, y1 = 2 * y + 1 < sh ? 2 * y + 1 : sh - 1;
            for (int k = 0; k < 4; ++k) {
                int sum = src[4 * (y0 * sw + x0) + k] + src[4 * (y0 * sw + x1) + k] +
                          src[4 * (y1 * sw + x0) + k] + src[4 * (y1 * sw + x1) + k];
                dst[4 * (y * dw + x) + k] = (uint8_t)((sum + 2) >> 2);
            }
        }
    }
}
