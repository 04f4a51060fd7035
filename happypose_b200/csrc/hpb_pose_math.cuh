// Device functions shared by the per-row pose kernels (hpb_pose.cu), the crop-box kernel (hpb_crop.cu) and the fused
// refiner prologue (hpb_prologue.cu).  One definition each, so the fused kernel is bit-identical to the chain it replaces.
#pragma once
#include "hpb_common.cuh"

namespace hpbm {

// columns (x, y, z) of R from the 6-D representation (a = first column, b = second column)
__device__ __forceinline__ void ortho6d(const float a[3], const float bb[3], float R[9]) {
    const float na = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    const float x0 = a[0] / na, x1 = a[1] / na, x2 = a[2] / na;
    float z0 = x1 * bb[2] - x2 * bb[1], z1 = x2 * bb[0] - x0 * bb[2], z2 = x0 * bb[1] - x1 * bb[0];
    const float nz = sqrtf(z0 * z0 + z1 * z1 + z2 * z2);
    z0 /= nz; z1 /= nz; z2 /= nz;
    const float y0 = z1 * x2 - z2 * x1, y1 = z2 * x0 - z0 * x2, y2 = z0 * x1 - z1 * x0;
    R[0] = x0; R[1] = y0; R[2] = z0;
    R[3] = x1; R[4] = y1; R[5] = z1;
    R[6] = x2; R[7] = y2; R[8] = z2;
}


// normalize_T (toolbox/lib3d/transform_ops.py:107-120): Gram-Schmidt on columns 0,1 of R, translation kept, last row 0 0 0 1
__device__ __forceinline__ void normalize_T_row(const float *t, float *o) {
    const float a[3] = {t[0], t[4], t[8]}, c[3] = {t[1], t[5], t[9]};
    const float tr[3] = {t[3], t[7], t[11]};
    float R[9];
    ortho6d(a, c, R);
    o[0] = R[0]; o[1] = R[1]; o[2] = R[2]; o[3] = tr[0];
    o[4] = R[3]; o[5] = R[4]; o[6] = R[5]; o[7] = tr[1];
    o[8] = R[6]; o[9] = R[7]; o[10] = R[8]; o[11] = tr[2];
    o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
}

// ---- multiview camera placement (float64 like the reference's numpy path, cast to float32 at the end) ----
struct M4 { double m[16]; };

__device__ __forceinline__ M4 m4_mul(const M4 &a, const M4 &b) {
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += a.m[i * 4 + k] * b.m[k * 4 + j];
            r.m[i * 4 + j] = s;
        }
    return r;
}
__device__ __forceinline__ M4 m4_rigid_inv(const M4 &a) {  // (R, t) -> (R^T, -R^T t)
    M4 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i * 4 + j] = a.m[j * 4 + i];
    for (int i = 0; i < 3; ++i) r.m[i * 4 + 3] = -(r.m[i * 4] * a.m[3] + r.m[i * 4 + 1] * a.m[7] + r.m[i * 4 + 2] * a.m[11]);
    r.m[12] = r.m[13] = r.m[14] = 0;
    r.m[15] = 1;
    return r;
}
// Panda3D look_at(): +Y exactly at the target, +Z as close to `up` as possible, +X = Y x Z; node axes as columns.
__device__ __forceinline__ void look_at(const double pos[3], const double tgt[3], const double up[3], double R[9]) {
    double f[3] = {tgt[0] - pos[0], tgt[1] - pos[1], tgt[2] - pos[2]};
    const double nf = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    f[0] /= nf; f[1] /= nf; f[2] /= nf;
    double r[3] = {f[1] * up[2] - f[2] * up[1], f[2] * up[0] - f[0] * up[2], f[0] * up[1] - f[1] * up[0]};
    const double nr = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    r[0] /= nr; r[1] /= nr; r[2] /= nr;
    const double u[3] = {r[1] * f[2] - r[2] * f[1], r[2] * f[0] - r[0] * f[2], r[0] * f[1] - r[1] * f[0]};
    R[0] = r[0]; R[1] = f[0]; R[2] = u[0];
    R[3] = r[1]; R[4] = f[1]; R[5] = u[1];
    R[6] = r[2]; R[7] = f[2]; R[8] = u[2];
}

// camera positions of the extra views travel BY VALUE in the kernel parameters: nothing is read from host memory after the
// launch call returns, so the launch can be captured into a CUDA graph
struct MvPositions {
    float p[26 * 3];
};


// Frame shared by all extra views of one row (toolbox/lib3d/multiview.py:28-66): camera C0 in the world (= object) frame,
// the reference point, C0's up vector, and the rotation of a camera at C0 looking at the reference point.
struct MvFrame {
    M4 C0W;
    double ref[3], up[3], c0[3], Rp[9], radius;
};

__device__ __forceinline__ void multiview_frame(const float *Tf, const float *tCR3, MvFrame &F) {
    M4 T;
    for (int k = 0; k < 16; ++k) T.m[k] = (double)Tf[k];
    double c[3] = {(double)tCR3[0], (double)tCR3[1], (double)tCR3[2]};
    T.m[12] = T.m[13] = T.m[14] = 0; T.m[15] = 1;
    M4 TOC = m4_rigid_inv(T);
    bool fin = true;
    for (int k = 0; k < 12; ++k) fin = fin && isfinite(TOC.m[k]);
    if (!fin) {  // multiview.py:44-46
        for (int k = 0; k < 16; ++k) TOC.m[k] = (k % 5 == 0) ? 1.0 : 0.0;
        c[0] = c[1] = c[2] = 0;
    }
    const M4 CCGL = {{1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, 1}};
    const M4 WC0 = m4_mul(TOC, CCGL);
    F.ref[0] = TOC.m[0] * c[0] + TOC.m[1] * c[1] + TOC.m[2] * c[2] + TOC.m[3];
    F.ref[1] = TOC.m[4] * c[0] + TOC.m[5] * c[1] + TOC.m[6] * c[2] + TOC.m[7];
    F.ref[2] = TOC.m[8] * c[0] + TOC.m[9] * c[1] + TOC.m[10] * c[2] + TOC.m[11];
    F.radius = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    F.up[0] = WC0.m[2]; F.up[1] = WC0.m[6]; F.up[2] = WC0.m[10];
    F.c0[0] = WC0.m[3]; F.c0[1] = WC0.m[7]; F.c0[2] = WC0.m[11];
    look_at(F.c0, F.ref, F.up, F.Rp);
    F.C0W = m4_rigid_inv(WC0);
}

// One extra view: camera at c0 + Rp * (q * radius) looking at the reference point; writes TCV_O = inv(TC0_CV) @ TCO (float32)
__device__ __forceinline__ void multiview_view(const MvFrame &F, const float *Tf, float qx, float qy, float qz, float *ov) {
    const M4 CCGL = {{1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, 1}};
    const M4 CCGLi = {{1, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0, 0, 0, 0, 0, 1}};
    const double q[3] = {qx * F.radius, qy * F.radius, qz * F.radius};
    const double pos[3] = {F.c0[0] + F.Rp[0] * q[0] + F.Rp[1] * q[1] + F.Rp[2] * q[2],
                           F.c0[1] + F.Rp[3] * q[0] + F.Rp[4] * q[1] + F.Rp[5] * q[2],
                           F.c0[2] + F.Rp[6] * q[0] + F.Rp[7] * q[1] + F.Rp[8] * q[2]};
    double Rn[9];
    look_at(pos, F.ref, F.up, Rn);
    M4 WN;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) WN.m[i * 4 + j] = Rn[i * 3 + j];
        WN.m[i * 4 + 3] = pos[i];
    }
    WN.m[12] = WN.m[13] = WN.m[14] = 0; WN.m[15] = 1;
    const M4 TC0_CV = m4_mul(m4_mul(CCGL, m4_mul(F.C0W, WN)), CCGLi);
    // cast to float32, invert as a rigid transform and compose with TCO in float32
    // (invert_transform_matrices(TC0_CV) @ TCO, multiview.py:236)
    float A[16];
    for (int k = 0; k < 16; ++k) A[k] = (float)TC0_CV.m[k];
    float Ai[12];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Ai[i * 4 + j] = A[j * 4 + i];
    for (int i = 0; i < 3; ++i) Ai[i * 4 + 3] = -(Ai[i * 4] * A[3] + Ai[i * 4 + 1] * A[7] + Ai[i * 4 + 2] * A[11]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j)
            ov[i * 4 + j] = Ai[i * 4] * Tf[j] + Ai[i * 4 + 1] * Tf[4 + j] + Ai[i * 4 + 2] * Tf[8 + j] + Ai[i * 4 + 3] * Tf[12 + j];
    for (int j = 0; j < 4; ++j)
        ov[12 + j] = A[12] * Tf[j] + A[13] * Tf[4 + j] + A[14] * Tf[8 + j] + A[15] * Tf[12 + j];
}

// Crop geometry of one row, computed by ALL threads of a CTA (blockDim.x a multiple of 32, at most 256): projects the
// point set with P = K @ T[:3], reduces min / max, then thread 0 builds boxes_rend, boxes_crop (deepim_boxes,
// toolbox/lib3d/cropping.py:27-75; obs box == rend box on this path) and K_crop (get_K_crop_resize,
// camera_geometry.py:70-122).  `sP` [12] and `red` [4][8] are shared-memory scratch; the call contains __syncthreads().
__device__ __forceinline__ void crop_boxes_cta(const float *K, const float *T, const float *tCR3, const float *pts, int n_pts, int H, int W,
                                               int h, int w, float lamb, float *K_crop_out, float *boxes_rend_out, float *boxes_crop_out,
                                               float *sP, float (*red)[8]) {
    const int tid = threadIdx.x;
    __syncthreads();  // scratch may still be read by the previous call
    if (tid < 12) {
        const int i = tid / 4, j = tid % 4;  // P = K @ TCO[:3]
        sP[tid] = fmaf(K[i * 3 + 2], T[8 + j], fmaf(K[i * 3 + 1], T[4 + j], K[i * 3] * T[j]));
    }
    __syncthreads();
    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int i = tid; i < n_pts; i += blockDim.x) {
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
        const float *c = tCR3;
        const float su = fmaf(K[2], c[2], fmaf(K[1], c[1], K[0] * c[0]));
        const float sv = fmaf(K[5], c[2], fmaf(K[4], c[1], K[3] * c[0]));
        float sz = fmaf(K[8], c[2], fmaf(K[7], c[1], K[6] * c[0]));
        sz = fmaxf(0.1f, sz);
        const float xc = su / sz, yc = sv / sz;
        const float r = (float)max(H, W) / (float)min(H, W);
        const float xdist = fmaxf(fabsf(mnx - xc), fabsf(mxx - xc));
        const float ydist = fmaxf(fabsf(mny - yc), fabsf(mxy - yc));
        const float width = fmaxf(xdist, ydist * r) * 2.0f * lamb;
        const float height = fmaxf(xdist / r, ydist) * 2.0f * lamb;
        const float x1 = xc - width / 2.0f, y1 = yc - height / 2.0f, x2 = xc + width / 2.0f, y2 = yc + height / 2.0f;
        if (boxes_rend_out) { boxes_rend_out[0] = mnx; boxes_rend_out[1] = mny; boxes_rend_out[2] = mxx; boxes_rend_out[3] = mxy; }
        if (boxes_crop_out) { boxes_crop_out[0] = x1; boxes_crop_out[1] = y1; boxes_crop_out[2] = x2; boxes_crop_out[3] = y2; }
        const float final_w = (float)max(h, w), final_h = (float)min(h, w);
        const float cw = x2 - x1, ch = y2 - y1;
        const float ccj = (x1 + x2) / 2.0f, cci = (y1 + y2) / 2.0f;
        const float cx = K[2] + (cw - 1.0f) / 2.0f - ccj;
        const float cy = K[5] + (ch - 1.0f) / 2.0f - cci;
        const float dcx = cx - (cw - 1.0f) / 2.0f, dcy = cy - (ch - 1.0f) / 2.0f;
        const float sx = final_w / cw, sy = final_h / ch;
        float ko[9];
        for (int k = 0; k < 9; ++k) ko[k] = K[k];
        ko[0] = sx * K[0];
        ko[4] = sy * K[4];
        ko[2] = (final_w - 1.0f) / 2.0f + sx * dcx;
        ko[5] = (final_h - 1.0f) / 2.0f + sy * dcy;
        for (int k = 0; k < 9; ++k) K_crop_out[k] = ko[k];
    }
}

}  // namespace hpbm
