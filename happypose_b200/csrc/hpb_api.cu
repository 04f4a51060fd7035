// C-ABI layer of libhpb200.so (see include/hpb200.h).  Plain pointers and sizes only; no torch types.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <new>
#include <utility>

#include "hpb_common.cuh"

int hpb_launch_crop_boxes(hpb_ctx *ctx, int H, int W, const float *points, int n_pts, const int32_t *obj_ids,
                          const float *K, const float *TCO, const float *tCR, int b, int h, int w, float lamb,
                          float *K_crop, float *boxes_rend, float *boxes_crop, cudaStream_t stream);
int hpb_launch_crop_pixels(hpb_ctx *ctx, const float *images, int n_im, int C, int H, int W, const int32_t *im_ids,
                           const float *boxes, int b, int h, int w, float *crops, int64_t crops_bs, int tap_bits,
                           cudaStream_t stream, void *crops_bf16x4 = nullptr);
int hpb_launch_normalize_T(hpb_ctx *ctx, const float *T, int b, float *out, cudaStream_t stream);
int hpb_launch_pose_update(hpb_ctx *ctx, const float *TCO, const float *K_crop, const float *outv, const float *tCR,
                           int b, int variant, float *TCO_out, cudaStream_t stream);
int hpb_launch_tco_init(hpb_ctx *ctx, int variant, const float *boxes, const float *points, int n_pts,
                        const int32_t *obj_ids, const float *K, const float *R, float z_mean, int b, float *out,
                        cudaStream_t stream);
int hpb_launch_multiview(hpb_ctx *ctx, const float *TCO, const float *tCR, int b, const float *positions_host,
                         int n_extra, int n_views, int keep_tco, float *out, cudaStream_t stream);
int hpb_launch_pack_input(hpb_ctx *ctx, const float *x, int64_t bstride, int b, int C, int h, int w, void *out, int Cp,
                          cudaStream_t stream);
int hpb_launch_maxpool(hpb_ctx *ctx, const void *in, int b, int H, int W, int C, void *out, cudaStream_t stream);
int hpb_launch_pack_s2d(hpb_ctx *ctx, const float *x, int64_t bstride, int b, int C, int H, int W, void *out, int Cz,
                        cudaStream_t stream);
int hpb_launch_normalize_depth(hpb_ctx *ctx, float *depth, int64_t bstride, const int32_t *chans, int n_planes,
                               const float *tCR, int b, int h, int w, int kind, cudaStream_t stream);
int hpb_launch_topk(hpb_ctx *ctx, const float *scores, const int32_t *groups, int n, int n_groups, int K,
                    int64_t *out_idx, int32_t *out_count, cudaStream_t stream);

int hpb_launch_icp_points(hpb_ctx *ctx, const float *depth_measured, int n_im, const float *depth_rendered, const uint8_t *masks,
                          const int32_t *im_ids, const float *K, int N, int H, int W, float delta, int64_t cap, float *pts_tgt,
                          float *pts_src, int32_t *counts, uint8_t *mask_out, int32_t *idx_tgt, int32_t *idx_src, cudaStream_t stream);

int hpb_launch_refiner_prologue(hpb_ctx *ctx, const float *TCO_in, const float *K, const int32_t *obj_ids, const float *pts_crop,
                                int n_crop, const float *pts_mv, int n_mv, int b, int H, int W, int h, int w, float lamb,
                                const float *positions_host, int n_extra, int n_views, int keep_tco, float *T_norm, float *tCR,
                                float *TCV_O, float *K_crop, float *boxes_rend, float *boxes_crop, float *KV_crop,
                                cudaStream_t stream);

int hpb_launch_maxpool_tma(hpb_ctx *ctx, const void *in, int b, int H, int W, int C, void *out, cudaStream_t stream);
int hpb_launch_conv3x3_tc(hpb_ctx *ctx, const void *x, int b, int H, int W, int C, const void *w, const float *bias, int O, const void *res,
                          void *out, cudaStream_t stream);
int hpb_launch_stem_tc(hpb_ctx *ctx, const void *z, int b, int Hz, int Wz, int C, const void *w, const float *bias, int O, void *out,
                       unsigned long long kmask, cudaStream_t stream);

static thread_local char g_err[512] = "";

int hpb_ws_grow(hpb_ctx *ctx, void **ptr, size_t *cur_bytes, size_t need, cudaStream_t stream, const char *what) {
    if (*cur_bytes >= need && *ptr) return HPB_OK;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) cudaGetLastError();
    if (st != cudaStreamCaptureStatusNone) {
        hpb_set_error("%s workspace must grow to %zu bytes while the stream is being captured into a CUDA graph; "
                      "run the call once eagerly or call hpb_reserve() before capturing", what, need);
        return HPB_EINVAL;
    }
    void *fresh = nullptr;
    HPB_CUDA_OK(cudaMalloc(&fresh, need));
    if (*ptr) ctx->retired.push_back(*ptr);  // captured graphs may still reference it
    *ptr = fresh;
    *cur_bytes = need;
    ctx->workspace_epoch++;
    return HPB_OK;
}

int hpb_stream_enter(hpb_ctx *ctx, cudaStream_t stream) {
    if (ctx->done_ev_valid && ctx->last_stream != stream) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) cudaGetLastError();
        if (st == cudaStreamCaptureStatusNone) HPB_CUDA_OK(cudaStreamWaitEvent(stream, ctx->done_ev, 0));
    }
    return HPB_OK;
}

void hpb_stream_leave(hpb_ctx *ctx, cudaStream_t stream) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) cudaGetLastError();
    if (st != cudaStreamCaptureStatusNone) return;  // inside a graph: ordering is the graph's own
    if (!ctx->done_ev && cudaEventCreateWithFlags(&ctx->done_ev, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        ctx->done_ev = nullptr;
        return;
    }
    if (cudaEventRecord(ctx->done_ev, stream) == cudaSuccess) {
        ctx->done_ev_valid = true;
        ctx->last_stream = stream;
    } else {
        cudaGetLastError();
    }
}

void hpb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}


// Sign of the screen-space area2 of FRONT faces if the mesh is a closed, consistently oriented surface, else 0.
// Vertices are welded by exact position (texture seams duplicate vertices); every undirected edge of the welded mesh
// must be used as often forwards as backwards (a closed 2-chain), and every connected component must enclose a volume
// of the same sign.  Then, for every pixel, #front-face hits == #back-face hits along the viewing ray, so a covered
// pixel is always covered by a front face: skipping back faces cannot open holes.  oracle/raster.py restates this.
static int hpb_closed_surface_sign(const float *verts, int64_t nv, const int32_t *faces, int64_t nf) {
    std::vector<int32_t> order((size_t)nv), wid((size_t)nv);
    for (int64_t i = 0; i < nv; ++i) order[(size_t)i] = (int32_t)i;
    auto key = [&](int32_t i, int k) { float v = verts[3 * (size_t)i + k]; return v == 0.0f ? 0.0f : v; };
    for (int64_t i = 0; i < 3 * nv; ++i)
        if (!std::isfinite(verts[i])) return 0;
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
        for (int k = 0; k < 3; ++k) {
            const float x = key(a, k), y = key(b, k);
            if (x != y) return x < y;
        }
        return a < b;
    });
    int32_t n_w = 0;
    for (int64_t i = 0; i < nv; ++i) {
        const int32_t cur = order[(size_t)i];
        if (i > 0) {
            const int32_t prev = order[(size_t)i - 1];
            if (!(key(cur, 0) == key(prev, 0) && key(cur, 1) == key(prev, 1) && key(cur, 2) == key(prev, 2))) ++n_w;
        }
        wid[(size_t)cur] = n_w;
    }
    ++n_w;
    std::vector<int32_t> parent((size_t)n_w);
    for (int32_t i = 0; i < n_w; ++i) parent[(size_t)i] = i;
    auto find = [&](int32_t x) {
        while (parent[(size_t)x] != x) { parent[(size_t)x] = parent[(size_t)parent[(size_t)x]]; x = parent[(size_t)x]; }
        return x;
    };
    std::vector<std::pair<uint64_t, int32_t>> edges;
    edges.reserve((size_t)nf * 3);
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = 0; i < nv; ++i)
        for (int k = 0; k < 3; ++k) { const double v = verts[3 * i + k]; mn[k] = v < mn[k] ? v : mn[k]; mx[k] = v > mx[k] ? v : mx[k]; }
    for (int64_t t = 0; t < nf; ++t) {
        const int32_t w[3] = {wid[(size_t)faces[3 * t]], wid[(size_t)faces[3 * t + 1]], wid[(size_t)faces[3 * t + 2]]};
        if (w[0] == w[1] || w[1] == w[2] || w[0] == w[2]) continue;  // zero-area face: never rasterised
        for (int k = 0; k < 3; ++k) {
            const int32_t u = w[k], v = w[(k + 1) % 3];
            const uint64_t lo = (uint64_t)(u < v ? u : v), hi = (uint64_t)(u < v ? v : u);
            edges.emplace_back((lo << 32) | hi, u < v ? 1 : -1);
        }
        const int32_t ra = find(w[0]), rb = find(w[1]), rc = find(w[2]);
        parent[(size_t)rb] = ra;
        parent[(size_t)find(rc)] = ra;
    }
    if (edges.empty()) return 0;
    std::sort(edges.begin(), edges.end());
    for (size_t i = 0; i < edges.size();) {
        size_t j = i;
        int sum = 0;
        while (j < edges.size() && edges[j].first == edges[i].first) sum += edges[j++].second;
        if (sum != 0) return 0;
        i = j;
    }
    std::vector<double> vol((size_t)n_w, 0.0);
    std::vector<char> used((size_t)n_w, 0);
    for (int64_t t = 0; t < nf; ++t) {
        const int32_t i0 = faces[3 * t], i1 = faces[3 * t + 1], i2 = faces[3 * t + 2];
        const int32_t w0 = wid[(size_t)i0], w1 = wid[(size_t)i1], w2 = wid[(size_t)i2];
        if (w0 == w1 || w1 == w2 || w0 == w2) continue;
        const double a[3] = {verts[3 * (size_t)i0], verts[3 * (size_t)i0 + 1], verts[3 * (size_t)i0 + 2]};
        const double b[3] = {verts[3 * (size_t)i1], verts[3 * (size_t)i1 + 1], verts[3 * (size_t)i1 + 2]};
        const double c[3] = {verts[3 * (size_t)i2], verts[3 * (size_t)i2 + 1], verts[3 * (size_t)i2 + 2]};
        const double det = a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
        const int32_t r = find(w0);
        vol[(size_t)r] += det / 6.0;
        used[(size_t)r] = 1;
    }
    const double diag2 = (mx[0] - mn[0]) * (mx[0] - mn[0]) + (mx[1] - mn[1]) * (mx[1] - mn[1]) + (mx[2] - mn[2]) * (mx[2] - mn[2]);
    const double tol = 1e-9 * diag2 * std::sqrt(diag2);
    int sign = 0;
    for (int32_t r = 0; r < n_w; ++r) {
        if (!used[(size_t)r]) continue;
        if (!(std::fabs(vol[(size_t)r]) > tol)) return 0;
        const int sg = vol[(size_t)r] > 0 ? 1 : -1;
        if (sign != 0 && sg != sign) return 0;
        sign = sg;
    }
    // outward winding (positive volume): front faces project with area2 < 0 in (x right, y down) pixel coordinates
    return sign > 0 ? -1 : (sign < 0 ? 1 : 0);
}

extern "C" {

int hpb_version(void) { return HPB_VERSION; }
const char *hpb_last_error(void) { return g_err; }

int hpb_create(int device, hpb_ctx **out) {
    HPB_REQUIRE(out != nullptr, "out is NULL");
    int n_dev = 0;
    HPB_CUDA_OK(cudaGetDeviceCount(&n_dev));
    HPB_REQUIRE(device >= 0 && device < n_dev, "no such CUDA device");
    HpbDeviceGuard guard(device);
    hpb_ctx *ctx = new (std::nothrow) hpb_ctx();
    if (!ctx) {
        hpb_set_error("hpb_create: out of host memory");
        return HPB_ENOMEM;
    }
    ctx->device = device;
    cudaDeviceProp prop;
    HPB_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (prop.major < 10) {
        hpb_set_error("hpb_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                      prop.major, prop.minor);
        delete ctx;
        return HPB_EINVAL;
    }
    if (cudaMalloc(&ctx->clipped_scenes, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(ctx->clipped_scenes, 0, sizeof(unsigned long long)) != cudaSuccess) {
        hpb_set_error("hpb_create: device allocation failed");
        delete ctx;
        return HPB_ENOMEM;
    }
    *out = ctx;
    return HPB_OK;
}

int hpb_destroy(hpb_ctx *ctx) {
    if (!ctx) return HPB_OK;
    HpbDeviceGuard guard(ctx->device);
    for (auto &m : ctx->meshes) {
        cudaFree(m.pos); cudaFree(m.nrm); cudaFree(m.uv); cudaFree(m.vcol); cudaFree(m.faces); cudaFree(m.tex); cudaFree(m.nu); cudaFree(m.tv);
    }
    for (void *p : ctx->retired) cudaFree(p);
    if (ctx->done_ev) cudaEventDestroy(ctx->done_ev);
    cudaFree(ctx->clipped_scenes);
    cudaFree(ctx->meshes_dev);
    cudaFree(ctx->vis);
    cudaFree(ctx->vert_scratch);
    cudaFree(ctx->frame_pack);
    cudaFree(ctx->topk_ws);
    delete ctx;
    return HPB_OK;
}

int64_t hpb_launch_count(const hpb_ctx *ctx) { return ctx ? ctx->launches : 0; }
int64_t hpb_workspace_epoch(const hpb_ctx *ctx) { return ctx ? ctx->workspace_epoch : 0; }

int hpb_raster_clipped_scenes(hpb_ctx *ctx, int64_t *count, int reset) {
    HPB_REQUIRE(ctx && count, "NULL argument");
    HpbDeviceGuard guard(ctx->device);
    unsigned long long v = 0;
    HPB_CUDA_OK(cudaMemcpy(&v, ctx->clipped_scenes, sizeof(v), cudaMemcpyDeviceToHost));  // synchronises the device
    if (reset) HPB_CUDA_OK(cudaMemset(ctx->clipped_scenes, 0, sizeof(v)));
    *count = (int64_t)v;
    return HPB_OK;
}

int hpb_reserve(hpb_ctx *ctx, int render_h, int render_w, int64_t frame_pixels, int64_t topk_rows, int64_t topk_groups) {
    HPB_REQUIRE(ctx, "NULL ctx");
    HPB_REQUIRE(render_h >= 0 && render_w >= 0 && frame_pixels >= 0 && topk_rows >= 0 && topk_groups >= 0, "negative size");
    HpbDeviceGuard guard(ctx->device);
    int rc = HPB_OK;
    if (render_h > 0 && render_w > 0) rc = hpb_raster_reserve(ctx, render_h, render_w, nullptr);
    if (rc == HPB_OK && frame_pixels > 0) rc = hpb_crop_reserve(ctx, frame_pixels, nullptr);
    if (rc == HPB_OK && topk_rows > 0) rc = hpb_topk_reserve(ctx, topk_rows, topk_groups, nullptr);
    if (rc == HPB_OK) HPB_CUDA_OK(cudaStreamSynchronize(nullptr));
    return rc;
}
int hpb_mesh_count(const hpb_ctx *ctx) { return ctx ? (int)ctx->meshes.size() : 0; }

int hpb_mesh_upload(hpb_ctx *ctx, const float *verts_xyz, const float *normals, const float *uv,
                    const uint8_t *vcolor, int64_t n_verts, const int32_t *faces, int64_t n_faces,
                    const uint8_t *tex, int tex_h, int tex_w, int tex_c, int32_t *mesh_id) {
    HPB_REQUIRE(ctx && verts_xyz && faces && mesh_id, "NULL argument");
    HPB_REQUIRE(n_verts > 0 && n_verts < (1ll << 31) && n_faces > 0 && n_faces < (1ll << 31), "bad mesh size");
    HPB_REQUIRE(!tex || (tex_h > 0 && tex_w > 0 && (tex_c == 3 || tex_c == 4)), "bad texture shape");
    for (int64_t i = 0; i < 3 * n_faces; ++i) HPB_REQUIRE(faces[i] >= 0 && faces[i] < n_verts, "face index out of range");
    HpbDeviceGuard guard(ctx->device);
    HpbMeshHost m;
    memset(&m.dev, 0, sizeof(m.dev));
    const size_t nv = (size_t)n_verts, nf = (size_t)n_faces;

    HPB_CUDA_OK(cudaMalloc(&m.pos, nv * 3 * sizeof(float)));
    HPB_CUDA_OK(cudaMemcpy(m.pos, verts_xyz, nv * 3 * sizeof(float), cudaMemcpyHostToDevice));

    std::vector<float> gen;
    if (!normals) {  // area-weighted smooth normals in float64 (a mesh file without normals)
        std::vector<double> acc(nv * 3, 0.0);
        for (size_t t = 0; t < nf; ++t) {
            const int32_t *f = faces + 3 * t;
            double a[3], b[3];
            for (int k = 0; k < 3; ++k) {
                a[k] = (double)verts_xyz[3 * f[1] + k] - (double)verts_xyz[3 * f[0] + k];
                b[k] = (double)verts_xyz[3 * f[2] + k] - (double)verts_xyz[3 * f[0] + k];
            }
            const double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
            for (int v = 0; v < 3; ++v)
                for (int k = 0; k < 3; ++k) acc[3 * (size_t)f[v] + k] += c[k];
        }
        gen.resize(nv * 3);
        for (size_t i = 0; i < nv; ++i) {
            const double l = std::sqrt(acc[3 * i] * acc[3 * i] + acc[3 * i + 1] * acc[3 * i + 1] + acc[3 * i + 2] * acc[3 * i + 2]);
            for (int k = 0; k < 3; ++k) gen[3 * i + k] = l > 0 ? (float)(acc[3 * i + k] / l) : 0.0f;
        }
        normals = gen.data();
    }
    HPB_CUDA_OK(cudaMalloc(&m.nrm, nv * 3 * sizeof(float)));
    HPB_CUDA_OK(cudaMemcpy(m.nrm, normals, nv * 3 * sizeof(float), cudaMemcpyHostToDevice));
    if (uv) {
        HPB_CUDA_OK(cudaMalloc(&m.uv, nv * 2 * sizeof(float)));
        HPB_CUDA_OK(cudaMemcpy(m.uv, uv, nv * 2 * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (vcolor) {
        HPB_CUDA_OK(cudaMalloc(&m.vcol, nv * 4));
        HPB_CUDA_OK(cudaMemcpy(m.vcol, vcolor, nv * 4, cudaMemcpyHostToDevice));
    }
    {
        std::vector<int32_t> f4(nf * 4);
        for (size_t t = 0; t < nf; ++t) {
            f4[4 * t] = faces[3 * t]; f4[4 * t + 1] = faces[3 * t + 1]; f4[4 * t + 2] = faces[3 * t + 2]; f4[4 * t + 3] = 0;
        }
        HPB_CUDA_OK(cudaMalloc(&m.faces, nf * 16));
        HPB_CUDA_OK(cudaMemcpy(m.faces, f4.data(), nf * 16, cudaMemcpyHostToDevice));
    }
    {
        std::vector<float> nu(nv * 4), tv(nv, 0.0f);
        for (size_t i = 0; i < nv; ++i) {
            nu[4 * i] = normals[3 * i]; nu[4 * i + 1] = normals[3 * i + 1]; nu[4 * i + 2] = normals[3 * i + 2];
            nu[4 * i + 3] = uv ? uv[2 * i] : 0.0f;
            if (uv) tv[i] = uv[2 * i + 1];
        }
        HPB_CUDA_OK(cudaMalloc(&m.nu, nv * 16));
        HPB_CUDA_OK(cudaMemcpy(m.nu, nu.data(), nv * 16, cudaMemcpyHostToDevice));
        HPB_CUDA_OK(cudaMalloc(&m.tv, nv * 4));
        HPB_CUDA_OK(cudaMemcpy(m.tv, tv.data(), nv * 4, cudaMemcpyHostToDevice));
    }
    m.dev.nu = (const float4 *)m.nu;
    m.dev.tv = (const float *)m.tv;
    m.closed_sign = hpb_closed_surface_sign(verts_xyz, n_verts, faces, n_faces);
    m.dev.cull_sign = m.closed_sign;
    m.dev.pos = (const float *)m.pos;
    m.dev.nrm = (const float *)m.nrm;
    m.dev.uv = (const float *)m.uv;
    m.dev.vcol = (const uchar4 *)m.vcol;
    m.dev.faces = (const int4 *)m.faces;
    m.dev.nv = (int)n_verts;
    m.dev.nf = (int)n_faces;
    if (tex && uv) {
        // RGBA8 mip chain: level l+1 = max(1, size/2), rounded 2x2 box filter, built on the device
        int w = tex_w, h = tex_h, levels = 0;
        long long total = 0;
        while (true) {
            HPB_REQUIRE(levels < HPB_MAX_MIPS, "texture too large");
            m.dev.tex_w[levels] = w; m.dev.tex_h[levels] = h; m.dev.tex_off[levels] = total;
            total += (long long)w * h;
            ++levels;
            if (w == 1 && h == 1) break;
            w = w > 1 ? w / 2 : 1;
            h = h > 1 ? h / 2 : 1;
        }
        m.dev.tex_levels = levels;
        m.dev.tex_pow2 = ((tex_w & (tex_w - 1)) == 0 && (tex_h & (tex_h - 1)) == 0) ? 1 : 0;
        HPB_CUDA_OK(cudaMalloc(&m.tex, (size_t)total * 4));
        uint8_t *staging = nullptr;
        const size_t raw = (size_t)tex_w * tex_h * tex_c;
        HPB_CUDA_OK(cudaMalloc(&staging, raw));
        HPB_CUDA_OK(cudaMemcpy(staging, tex, raw, cudaMemcpyHostToDevice));
        int rc = hpb_launch_tex_expand(staging, tex_w * tex_h, tex_c, (uchar4 *)m.tex, 0);
        for (int l = 1; l < levels && rc == HPB_OK; ++l)
            rc = hpb_launch_mip((const uchar4 *)m.tex + m.dev.tex_off[l - 1], m.dev.tex_w[l - 1], m.dev.tex_h[l - 1],
                                (uchar4 *)m.tex + m.dev.tex_off[l], m.dev.tex_w[l], m.dev.tex_h[l], 0);
        cudaError_t e = cudaDeviceSynchronize();
        cudaFree(staging);
        if (rc != HPB_OK) return rc;
        HPB_CUDA_OK(e);
        m.dev.tex = (const uchar4 *)m.tex;
    }
    ctx->meshes.push_back(m);
    if ((int)n_verts > ctx->max_nv) ctx->max_nv = (int)n_verts;
    // refresh the device-side table
    const int n = (int)ctx->meshes.size();
    if (n > ctx->meshes_dev_cap) {
        const int cap = n * 2 + 8;
        HpbMeshDev *nd = nullptr;
        HPB_CUDA_OK(cudaMalloc(&nd, sizeof(HpbMeshDev) * cap));
        HPB_CUDA_OK(cudaDeviceSynchronize());
        if (ctx->meshes_dev) {
            // the old table stays valid for the mesh ids it holds (captured graphs reference it): retire, do not free
            ctx->retired.push_back(ctx->meshes_dev);
            ctx->workspace_epoch++;
        }
        ctx->meshes_dev = nd;
        ctx->meshes_dev_cap = cap;
        std::vector<HpbMeshDev> all(n);
        for (int i = 0; i < n; ++i) all[i] = ctx->meshes[i].dev;
        HPB_CUDA_OK(cudaMemcpy(ctx->meshes_dev, all.data(), sizeof(HpbMeshDev) * n, cudaMemcpyHostToDevice));
    } else {
        HPB_CUDA_OK(cudaMemcpy(ctx->meshes_dev + (n - 1), &ctx->meshes[n - 1].dev, sizeof(HpbMeshDev), cudaMemcpyHostToDevice));
    }
    *mesh_id = n - 1;
    return HPB_OK;
}

int hpb_mesh_get_mip(hpb_ctx *ctx, int32_t mesh_id, int level, uint8_t *out_host, int *w, int *h, int *levels) {
    HPB_REQUIRE(ctx, "NULL ctx");
    if (mesh_id < 0 || mesh_id >= (int)ctx->meshes.size()) {
        hpb_set_error("hpb_mesh_get_mip: unknown mesh id %d", mesh_id);
        return HPB_ENOTFOUND;
    }
    const HpbMeshDev &d = ctx->meshes[mesh_id].dev;
    if (levels) *levels = d.tex_levels;
    if (d.tex_levels == 0) return HPB_OK;
    HPB_REQUIRE(level >= 0 && level < d.tex_levels, "no such mip level");
    if (w) *w = d.tex_w[level];
    if (h) *h = d.tex_h[level];
    if (out_host) {
        HpbDeviceGuard guard(ctx->device);
        HPB_CUDA_OK(cudaMemcpy(out_host, d.tex + d.tex_off[level], (size_t)d.tex_w[level] * d.tex_h[level] * 4, cudaMemcpyDeviceToHost));
    }
    return HPB_OK;
}

int hpb_mesh_closed_sign(hpb_ctx *ctx, int32_t mesh_id, int *sign) {
    HPB_REQUIRE(ctx && sign, "NULL argument");
    if (mesh_id < 0 || mesh_id >= (int)ctx->meshes.size()) {
        hpb_set_error("hpb_mesh_closed_sign: unknown mesh id %d", mesh_id);
        return HPB_ENOTFOUND;
    }
    *sign = ctx->meshes[mesh_id].closed_sign;
    return HPB_OK;
}

int hpb_mesh_set_cull(hpb_ctx *ctx, int32_t mesh_id, int enable) {
    HPB_REQUIRE(ctx, "NULL ctx");
    if (mesh_id < 0 || mesh_id >= (int)ctx->meshes.size()) {
        hpb_set_error("hpb_mesh_set_cull: unknown mesh id %d", mesh_id);
        return HPB_ENOTFOUND;
    }
    HpbDeviceGuard guard(ctx->device);
    HpbMeshHost &m = ctx->meshes[mesh_id];
    m.dev.cull_sign = enable ? m.closed_sign : 0;
    HPB_CUDA_OK(cudaDeviceSynchronize());
    HPB_CUDA_OK(cudaMemcpy(ctx->meshes_dev + mesh_id, &m.dev, sizeof(HpbMeshDev), cudaMemcpyHostToDevice));
    return HPB_OK;
}

int hpb_render(hpb_ctx *ctx, const int32_t *mesh_ids_dev, const float *TCO_dev, const float *K_dev,
               const float *ambient_dev, int b, int h, int w, float z_near, float z_far, uint32_t flags,
               float *rgb_dev, int64_t rgb_bstride, float *normals_dev, int64_t normals_bstride, float *depth_dev,
               int64_t depth_bstride, uint8_t *mask_dev, int64_t mask_bstride, int views, int64_t view_stride,
               const float *lights_dev, int n_lights, void *stream) {
    HPB_REQUIRE(ctx, "NULL ctx");
    HPB_REQUIRE(b >= 0 && h > 0 && w > 0 && (long long)h * w < (1ll << 30), "bad batch / resolution");
    HPB_REQUIRE(views >= 1 && b % views == 0, "b must be a multiple of views");
    if (b == 0) return HPB_OK;
    HPB_REQUIRE(mesh_ids_dev && TCO_dev && K_dev, "NULL input");
    HPB_REQUIRE(z_near > 0.f && z_far > z_near, "bad near/far");
    HPB_REQUIRE(!ctx->meshes.empty(), "no mesh uploaded");
    HPB_REQUIRE(!(flags & HPB_RENDER_RGB) || rgb_dev, "rgb requested but NULL");
    HPB_REQUIRE(!(flags & HPB_RENDER_NORMALS) || normals_dev, "normals requested but NULL");
    HPB_REQUIRE(!(flags & HPB_RENDER_DEPTH) || depth_dev, "depth requested but NULL");
    HPB_REQUIRE(!(flags & HPB_RENDER_MASK) || mask_dev, "mask requested but NULL");
    HPB_REQUIRE(flags != 0, "nothing to render");
    HPB_REQUIRE(n_lights >= 0 && n_lights <= HPB_MAX_LIGHTS, "at most 8 point / directional lights per scene");
    HPB_REQUIRE(n_lights == 0 || lights_dev, "lights_dev is NULL");
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_raster(ctx, mesh_ids_dev, TCO_dev, K_dev, ambient_dev, b, h, w, z_near, z_far, flags, rgb_dev,
                             rgb_bstride, normals_dev, normals_bstride, depth_dev, depth_bstride, mask_dev,
                             mask_bstride, views, view_stride, (cudaStream_t)stream, nullptr, 0, nullptr, 0, 0, 0, lights_dev,
                             n_lights);
}

int hpb_render_s2d_bf16(hpb_ctx *ctx, const int32_t *mesh_ids_dev, const float *TCO_dev, const float *K_dev,
                        const float *ambient_dev, int b, int h, int w, float z_near, float z_far, const void *crops_dev,
                        int64_t crops_bstride, int crops_format, void *out_dev, int C_padded, int pad_prezeroed, void *stream) {
    HPB_REQUIRE(ctx, "NULL ctx");
    HPB_REQUIRE(b >= 0 && h > 0 && w > 0 && (long long)h * w < (1ll << 30), "bad batch / resolution");
    HPB_REQUIRE(h % 2 == 0 && w % 2 == 0, "space-to-depth output needs even height and width");
    HPB_REQUIRE(C_padded >= 64 && C_padded % 32 == 0, "C_padded must be a multiple of 32 with C_padded / 4 >= 9 channels per sub-pixel");
    if (b == 0) return HPB_OK;
    HPB_REQUIRE(mesh_ids_dev && TCO_dev && K_dev && crops_dev && out_dev, "NULL input");
    HPB_REQUIRE((reinterpret_cast<uintptr_t>(out_dev) & 15u) == 0, "out_dev must be 16-byte aligned");
    HPB_REQUIRE(crops_format == HPB_CROPS_F32_PLANAR || crops_format == HPB_CROPS_BF16X4, "unknown crops_format");
    HPB_REQUIRE(crops_bstride >= (crops_format == HPB_CROPS_F32_PLANAR ? 3ll : 1ll) * h * w, "crops_bstride smaller than one crop");
    HPB_REQUIRE((reinterpret_cast<uintptr_t>(crops_dev) & (crops_format == HPB_CROPS_BF16X4 ? 7u : 3u)) == 0, "crops_dev misaligned");
    HPB_REQUIRE(z_near > 0.f && z_far > z_near, "bad near/far");
    HPB_REQUIRE(!ctx->meshes.empty(), "no mesh uploaded");
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_raster(ctx, mesh_ids_dev, TCO_dev, K_dev, ambient_dev, b, h, w, z_near, z_far,
                             HPB_RENDER_RGB | HPB_RENDER_NORMALS, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0, 1, 0,
                             (cudaStream_t)stream, crops_dev, crops_bstride, out_dev, C_padded, crops_format, pad_prezeroed ? 1 : 0);
}

int hpb_crop_boxes(hpb_ctx *ctx, int H, int W, const float *points_dev, int n_obj, int n_pts,
                   const int32_t *obj_ids_dev, const float *K_dev, const float *TCO_dev, const float *tCR_dev, int b,
                   int h, int w, float lamb, float *K_crop_dev, float *boxes_rend_dev, float *boxes_crop_dev,
                   void *stream) {
    HPB_REQUIRE(ctx, "NULL ctx");
    HPB_REQUIRE(b >= 0 && H > 0 && W > 0 && h > 0 && w > 0 && n_obj > 0 && n_pts > 0, "bad sizes");
    if (b == 0) return HPB_OK;
    HPB_REQUIRE(points_dev && obj_ids_dev && K_dev && TCO_dev && tCR_dev, "NULL input");
    HPB_REQUIRE(K_crop_dev && boxes_rend_dev && boxes_crop_dev, "NULL output");
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_crop_boxes(ctx, H, W, points_dev, n_pts, obj_ids_dev, K_dev, TCO_dev, tCR_dev, b, h, w, lamb,
                                 K_crop_dev, boxes_rend_dev, boxes_crop_dev, (cudaStream_t)stream);
}

int hpb_crop(hpb_ctx *ctx, const float *images_dev, int n_im, int C, int H, int W, const int32_t *im_ids_dev,
             const float *points_dev, int n_obj, int n_pts, const int32_t *obj_ids_dev, const float *K_dev,
             const float *TCO_dev, const float *tCR_dev, int b, int h, int w, float lamb, float *crops_dev,
             int64_t crops_bstride, float *K_crop_dev, float *boxes_rend_dev, float *boxes_crop_dev, int tap_bits,
             void *stream) {
    HPB_REQUIRE(ctx, "NULL ctx");
    HPB_REQUIRE(C == 3 || C == 4, "images must have 3 or 4 channels");  // cropping.py:161
    HPB_REQUIRE(tap_bits == 0 || tap_bits == 16 || tap_bits == 32, "tap_bits must be 0 (context default), 16 or 32");
    if (tap_bits == 0) tap_bits = ctx->crop_tap_bits;
    HPB_REQUIRE(n_im > 0 && images_dev && im_ids_dev && crops_dev, "NULL image input/output");
    int rc = hpb_crop_boxes(ctx, H, W, points_dev, n_obj, n_pts, obj_ids_dev, K_dev, TCO_dev, tCR_dev, b, h, w, lamb,
                            K_crop_dev, boxes_rend_dev, boxes_crop_dev, stream);
    if (rc != HPB_OK || b == 0) return rc;
    HpbDeviceGuard guard(ctx->device);
    for (int s = 0; s < b; s += 32768) {  // grid.y limit
        const int nb = b - s < 32768 ? b - s : 32768;
        rc = hpb_launch_crop_pixels(ctx, images_dev, n_im, C, H, W, im_ids_dev + s, boxes_crop_dev + (size_t)s * 4, nb,
                                    h, w, crops_dev + (size_t)s * crops_bstride, crops_bstride, tap_bits, (cudaStream_t)stream);
        if (rc != HPB_OK) return rc;
    }
    return HPB_OK;
}

int hpb_crop_bf16x4(hpb_ctx *ctx, const float *images_dev, int n_im, int H, int W, const int32_t *im_ids_dev,
                    const float *points_dev, int n_obj, int n_pts, const int32_t *obj_ids_dev, const float *K_dev,
                    const float *TCO_dev, const float *tCR_dev, int b, int h, int w, float lamb, void *crops_dev,
                    int64_t crops_bstride, float *K_crop_dev, float *boxes_rend_dev, float *boxes_crop_dev, int tap_bits,
                    void *stream) {
    HPB_REQUIRE(ctx, "NULL ctx");
    HPB_REQUIRE(n_im > 0 && images_dev && im_ids_dev && crops_dev, "NULL image input/output");
    HPB_REQUIRE((reinterpret_cast<uintptr_t>(crops_dev) & 7u) == 0, "crops_dev must be 8-byte aligned");
    HPB_REQUIRE(crops_bstride >= (int64_t)h * w, "crops_bstride smaller than one crop");
    HPB_REQUIRE(tap_bits == 0 || tap_bits == 16 || tap_bits == 32, "tap_bits must be 0 (context default), 16 or 32");
    if (tap_bits == 0) tap_bits = ctx->crop_tap_bits;
    int rc = hpb_crop_boxes(ctx, H, W, points_dev, n_obj, n_pts, obj_ids_dev, K_dev, TCO_dev, tCR_dev, b, h, w, lamb,
                            K_crop_dev, boxes_rend_dev, boxes_crop_dev, stream);
    if (rc != HPB_OK || b == 0) return rc;
    HpbDeviceGuard guard(ctx->device);
    for (int s = 0; s < b; s += 32768) {  // grid.y limit
        const int nb = b - s < 32768 ? b - s : 32768;
        rc = hpb_launch_crop_pixels(ctx, images_dev, n_im, 3, H, W, im_ids_dev + s, boxes_crop_dev + (size_t)s * 4, nb, h, w,
                                    nullptr, crops_bstride, tap_bits, (cudaStream_t)stream,
                                    reinterpret_cast<unsigned char *>(crops_dev) + (size_t)s * crops_bstride * 8);
        if (rc != HPB_OK) return rc;
    }
    return HPB_OK;
}

int hpb_set_crop_tma(hpb_ctx *ctx, int enable) {
    HPB_REQUIRE(ctx, "NULL ctx");
    ctx->crop_tma = enable ? 1 : 0;
    return HPB_OK;
}

int hpb_set_stem_tc_halo(hpb_ctx *ctx, int enable) {
    HPB_REQUIRE(ctx, "NULL context");
    ctx->stem_tc_halo = enable < 0 || enable > 2 ? 1 : enable;
    return HPB_OK;
}

int hpb_set_maxpool_tma(hpb_ctx *ctx, int enable) {
    HPB_REQUIRE(ctx, "NULL ctx");
    ctx->maxpool_tma = enable ? 1 : 0;
    return HPB_OK;
}

int hpb_set_crop_tap_precision(hpb_ctx *ctx, int bits) {
    HPB_REQUIRE(ctx, "NULL ctx");
    HPB_REQUIRE(bits == 32 || bits == 16, "bits must be 32 or 16");
    ctx->crop_tap_bits = bits;
    return HPB_OK;
}

int hpb_normalize_T(hpb_ctx *ctx, const float *T_dev, int b, float *T_out_dev, void *stream) {
    HPB_REQUIRE(ctx && b >= 0, "bad argument");
    if (b == 0) return HPB_OK;
    HPB_REQUIRE(T_dev && T_out_dev, "NULL pointer");
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_normalize_T(ctx, T_dev, b, T_out_dev, (cudaStream_t)stream);
}

int hpb_pose_update(hpb_ctx *ctx, const float *TCO_dev, const float *K_crop_dev, const float *out_dev,
                    const float *tCR_dev, int b, int variant, float *TCO_out_dev, void *stream) {
    HPB_REQUIRE(ctx && b >= 0, "bad argument");
    HPB_REQUIRE(variant >= 0 && variant <= 2, "unknown variant");
    if (b == 0) return HPB_OK;
    HPB_REQUIRE(TCO_dev && K_crop_dev && out_dev && TCO_out_dev, "NULL pointer");
    HPB_REQUIRE(variant != HPB_POSE_MEGAPOSE || tCR_dev, "tCR required by the MegaPose variant");
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_pose_update(ctx, TCO_dev, K_crop_dev, out_dev, tCR_dev, b, variant, TCO_out_dev, (cudaStream_t)stream);
}

int hpb_tco_init(hpb_ctx *ctx, int variant, const float *boxes_dev, const float *points_dev, int n_obj, int n_pts,
                 const int32_t *obj_ids_dev, const float *K_dev, const float *R_dev, float z_mean, int b,
                 float *TCO_out_dev, void *stream) {
    HPB_REQUIRE(ctx && b >= 0, "bad argument");
    HPB_REQUIRE(variant >= 0 && variant <= 2, "unknown variant");
    if (b == 0) return HPB_OK;
    HPB_REQUIRE(boxes_dev && K_dev && TCO_out_dev, "NULL pointer");
    if (variant != HPB_TCO_INIT_FROM_BOXES)
        HPB_REQUIRE(points_dev && obj_ids_dev && n_obj > 0 && n_pts > 0, "points required");
    HPB_REQUIRE(variant != HPB_TCO_INIT_AUTODEPTH_WITH_R || R_dev, "R required");
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_tco_init(ctx, variant, boxes_dev, points_dev, n_pts, obj_ids_dev, K_dev, R_dev, z_mean, b,
                               TCO_out_dev, (cudaStream_t)stream);
}

// camera positions (in units of |tCR|, in the frame of the look-at camera at C0) of a multiview type
// (toolbox/lib3d/multiview.py:95-163); returns HPB_OK and fills pos / n_extra / keep
static int hpb_mv_positions(int mv_type, int n_views, int remove_tco_rendering, float *pos, int *n_extra_out, int *keep_out) {
    int n_extra = 0, keep = 1;
    if (n_views == 1) {  // multiview.py:190-197: identity view only
        n_extra = 0;
        keep = 1;
    } else {
        keep = remove_tco_rendering ? 0 : 1;
        if (mv_type == HPB_MV_TCO_FRONT_1VIEW) {
            const float q[3] = {0, 0, 0};
            memcpy(pos, q, sizeof(q));
            n_extra = 1;
        } else if (mv_type == HPB_MV_TCO_FRONT_3VIEWS) {
            const float q[9] = {0, 0, 0, 1, 0, 0, -1, 0, 0};
            memcpy(pos, q, sizeof(q));
            n_extra = 3;
        } else if (mv_type == HPB_MV_SPHERE_26VIEWS) {
            const int ys[3] = {0, 1, 2}, xs[3] = {0, -1, 1}, zs[3] = {0, 1, -1};
            for (int yi = 0; yi < 3; ++yi)
                for (int xi = 0; xi < 3; ++xi)
                    for (int zi = 0; zi < 3; ++zi) {
                        if (xs[xi] == 0 && ys[yi] == 1 && zs[zi] == 0) continue;
                        pos[3 * n_extra] = (float)xs[xi]; pos[3 * n_extra + 1] = (float)ys[yi]; pos[3 * n_extra + 2] = (float)zs[zi];
                        ++n_extra;
                    }
        } else {
            hpb_set_error("unknown multiview type %d", mv_type);
            return HPB_EINVAL;
        }
        HPB_REQUIRE(n_views == n_extra + keep, "n_views does not match the multiview type");
    }
    *n_extra_out = n_extra;
    *keep_out = keep;
    return HPB_OK;
}

int hpb_multiview(hpb_ctx *ctx, const float *TCO_dev, const float *tCR_dev, int b, int mv_type, int n_views,
                  int remove_tco_rendering, float *TCV_O_dev, void *stream) {
    HPB_REQUIRE(ctx && b >= 0 && n_views >= 1, "bad argument");
    if (b == 0) return HPB_OK;
    HPB_REQUIRE(TCO_dev && tCR_dev && TCV_O_dev, "NULL pointer");
    float pos[26 * 3];
    int n_extra = 0, keep = 1;
    const int rc = hpb_mv_positions(mv_type, n_views, remove_tco_rendering, pos, &n_extra, &keep);
    if (rc != HPB_OK) return rc;
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_multiview(ctx, TCO_dev, tCR_dev, b, pos, n_extra, n_views, keep, TCV_O_dev, (cudaStream_t)stream);
}

int hpb_refiner_prologue(hpb_ctx *ctx, const float *TCO_dev, const float *K_dev, const int32_t *obj_ids_dev,
                         const float *points_crop_dev, int n_obj, int n_pts_crop, const float *points_mv_dev, int n_pts_mv, int b,
                         int H, int W, int h, int w, float lamb, int mv_type, int n_views, int remove_tco_rendering,
                         float *T_norm_dev, float *tCR_dev, float *TCV_O_dev, float *K_crop_dev, float *boxes_rend_dev,
                         float *boxes_crop_dev, float *KV_crop_dev, void *stream) {
    HPB_REQUIRE(ctx && b >= 0 && n_views >= 1 && n_obj > 0 && n_pts_crop > 0, "bad argument");
    HPB_REQUIRE(H > 0 && W > 0 && h > 0 && w > 0, "bad sizes");
    if (b == 0) return HPB_OK;
    HPB_REQUIRE(TCO_dev && K_dev && obj_ids_dev && points_crop_dev, "NULL input");
    HPB_REQUIRE(T_norm_dev && tCR_dev && TCV_O_dev && K_crop_dev && boxes_rend_dev && boxes_crop_dev, "NULL output");
    HPB_REQUIRE(!KV_crop_dev || (points_mv_dev && n_pts_mv > 0), "KV_crop needs the multiview point set");
    float pos[26 * 3];
    int n_extra = 0, keep = 1;
    const int rc = hpb_mv_positions(mv_type, n_views, remove_tco_rendering, pos, &n_extra, &keep);
    if (rc != HPB_OK) return rc;
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_refiner_prologue(ctx, TCO_dev, K_dev, obj_ids_dev, points_crop_dev, n_pts_crop, points_mv_dev, n_pts_mv, b, H, W,
                                       h, w, lamb, pos, n_extra, n_views, keep, T_norm_dev, tCR_dev, TCV_O_dev, K_crop_dev,
                                       boxes_rend_dev, boxes_crop_dev, KV_crop_dev, (cudaStream_t)stream);
}

int hpb_crop_pixels(hpb_ctx *ctx, const float *images_dev, int n_im, int C, int H, int W, const int32_t *im_ids_dev,
                    const float *boxes_crop_dev, int b, int h, int w, float *crops_dev, int64_t crops_bstride, int tap_bits,
                    void *stream) {
    HPB_REQUIRE(ctx && b >= 0 && H > 0 && W > 0 && h > 0 && w > 0, "bad sizes");
    HPB_REQUIRE(C == 3 || C == 4, "images must have 3 or 4 channels");
    HPB_REQUIRE(tap_bits == 0 || tap_bits == 16 || tap_bits == 32, "tap_bits must be 0 (context default), 16 or 32");
    if (b == 0) return HPB_OK;
    HPB_REQUIRE(n_im > 0 && images_dev && im_ids_dev && boxes_crop_dev && crops_dev, "NULL pointer");
    if (tap_bits == 0) tap_bits = ctx->crop_tap_bits;
    HpbDeviceGuard guard(ctx->device);
    for (int s0 = 0; s0 < b; s0 += 32768) {  // grid.y limit
        const int nb = b - s0 < 32768 ? b - s0 : 32768;
        const int rc = hpb_launch_crop_pixels(ctx, images_dev, n_im, C, H, W, im_ids_dev + s0, boxes_crop_dev + (size_t)s0 * 4, nb, h, w,
                                              crops_dev + (size_t)s0 * crops_bstride, crops_bstride, tap_bits, (cudaStream_t)stream);
        if (rc != HPB_OK) return rc;
    }
    return HPB_OK;
}

int hpb_normalize_depth(hpb_ctx *ctx, float *depth_dev, int64_t bstride, const int32_t *plane_channels_host,
                        int n_planes, const float *tCR_dev, int b, int h, int w, int kind, void *stream) {
    HPB_REQUIRE(ctx && b >= 0 && n_planes >= 0 && n_planes <= 8, "bad argument (at most 8 depth planes)");
    HPB_REQUIRE(kind >= 0 && kind <= 3, "unknown depth normalisation");
    if (b == 0 || n_planes == 0) return HPB_OK;
    HPB_REQUIRE(depth_dev && plane_channels_host && tCR_dev, "NULL pointer");
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_normalize_depth(ctx, depth_dev, bstride, plane_channels_host, n_planes, tCR_dev, b, h, w, kind,
                                      (cudaStream_t)stream);
}

int hpb_topk_segmented(hpb_ctx *ctx, const float *scores_dev, const int32_t *group_ids_dev, int n, int n_groups,
                       int K, int64_t *out_idx_dev, int32_t *out_count_dev, void *stream) {
    HPB_REQUIRE(ctx && n >= 0 && n_groups >= 0, "bad argument");
    HPB_REQUIRE(out_count_dev, "NULL out_count");
    HPB_REQUIRE(n == 0 || (scores_dev && group_ids_dev && out_idx_dev), "NULL pointer");
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_topk(ctx, scores_dev, group_ids_dev, n, n_groups, K, out_idx_dev, out_count_dev, (cudaStream_t)stream);
}

int hpb_icp_points(hpb_ctx *ctx, const float *depth_measured_dev, int n_im, const float *depth_rendered_dev,
                   const uint8_t *masks_dev, const int32_t *im_ids_dev, const float *K_dev, int N, int H, int W,
                   float depth_delta_thresh, int64_t capacity, float *points_tgt_dev, float *points_src_dev,
                   int32_t *counts_dev, uint8_t *mask_out_dev, int32_t *index_tgt_dev, int32_t *index_src_dev, void *stream) {
    HPB_REQUIRE(ctx && N >= 0 && n_im > 0 && H > 0 && W > 0 && capacity > 0, "bad argument");
    if (N == 0) return HPB_OK;
    HPB_REQUIRE(depth_measured_dev && depth_rendered_dev && im_ids_dev && K_dev, "NULL input");
    HPB_REQUIRE(points_tgt_dev && points_src_dev && counts_dev, "NULL output");
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_icp_points(ctx, depth_measured_dev, n_im, depth_rendered_dev, masks_dev, im_ids_dev, K_dev, N, H, W,
                                 depth_delta_thresh, capacity, points_tgt_dev, points_src_dev, counts_dev, mask_out_dev,
                                 index_tgt_dev, index_src_dev, (cudaStream_t)stream);
}

int hpb_pack_input_bf16(hpb_ctx *ctx, const float *x_dev, int64_t x_bstride, int b, int C, int h, int w, void *out_dev,
                        int C_padded, void *stream) {
    HPB_REQUIRE(ctx && b >= 0 && C > 0 && h > 0 && w > 0, "bad argument");
    HPB_REQUIRE(C_padded >= C && C_padded % 8 == 0, "C_padded must be a multiple of 8 and >= C");
    if (b == 0) return HPB_OK;
    HPB_REQUIRE(x_dev && out_dev, "NULL pointer");
    HPB_REQUIRE(((uintptr_t)out_dev & 15) == 0, "output must be 16-byte aligned");
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_pack_input(ctx, x_dev, x_bstride, b, C, h, w, out_dev, C_padded, (cudaStream_t)stream);
}

int hpb_maxpool3x3s2_bf16_nhwc(hpb_ctx *ctx, const void *in_dev, int b, int H, int W, int C, void *out_dev, void *stream) {
    HPB_REQUIRE(ctx && b >= 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad argument (C must be a multiple of 8)");
    if (b == 0) return HPB_OK;
    HPB_REQUIRE(in_dev && out_dev, "NULL pointer");
    HPB_REQUIRE((((uintptr_t)in_dev | (uintptr_t)out_dev) & 15) == 0, "buffers must be 16-byte aligned");
    HpbDeviceGuard guard(ctx->device);
    if (ctx->maxpool_tma) {  // TMA-staged tiles for the shapes it serves (C = 64, the ResNet stem); otherwise the plain kernel
        const int rc = hpb_launch_maxpool_tma(ctx, in_dev, b, H, W, C, out_dev, (cudaStream_t)stream);
        if (rc != HPB_ENOTFOUND) return rc;
    }
    return hpb_launch_maxpool(ctx, in_dev, b, H, W, C, out_dev, (cudaStream_t)stream);
}

int hpb_conv3x3_bias_relu_bf16_nhwc(hpb_ctx *ctx, const void *x_dev, int b, int H, int W, int C, const void *w_dev, const float *bias_dev, int O,
                                    const void *residual_dev, void *out_dev, void *stream) {
    HPB_REQUIRE(ctx && b >= 0 && H > 0 && W > 0 && C > 0 && O > 0, "bad argument");
    if (b == 0) return HPB_OK;
    HPB_REQUIRE(x_dev && w_dev && bias_dev && out_dev, "NULL pointer");
    HPB_REQUIRE(out_dev != x_dev, "the convolution cannot run in place");
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_conv3x3_tc(ctx, x_dev, b, H, W, C, w_dev, bias_dev, O, residual_dev, out_dev, (cudaStream_t)stream);
}

int hpb_stem_conv4x4_relu_bf16_nhwc(hpb_ctx *ctx, const void *z_dev, int b, int Hz, int Wz, int C, const void *w_dev, const float *bias_dev,
                                    int O, uint64_t k_slice_mask, void *out_dev, void *stream) {
    HPB_REQUIRE(ctx && b >= 0 && Hz > 3 && Wz > 3 && C > 0 && O > 0, "bad argument");
    if (b == 0) return HPB_OK;
    HPB_REQUIRE(z_dev && w_dev && bias_dev && out_dev, "NULL pointer");
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_stem_tc(ctx, z_dev, b, Hz, Wz, C, w_dev, bias_dev, O, out_dev, (unsigned long long)k_slice_mask, (cudaStream_t)stream);
}

int hpb_pack_input_s2d_bf16(hpb_ctx *ctx, const float *x_dev, int64_t x_bstride, int b, int C, int H, int W, void *out_dev,
                            int C_padded, void *stream) {
    HPB_REQUIRE(ctx && b >= 0 && C > 0 && H > 0 && W > 0, "bad argument");
    HPB_REQUIRE(H % 2 == 0 && W % 2 == 0, "space-to-depth needs even H and W");
    HPB_REQUIRE(C_padded >= 4 * C && C_padded % 32 == 0 && C <= 64, "C_padded must be a multiple of 32 and >= 4*C (C <= 64)");
    if (b == 0) return HPB_OK;
    HPB_REQUIRE(x_dev && out_dev, "NULL pointer");
    HPB_REQUIRE(((uintptr_t)out_dev & 15) == 0, "output must be 16-byte aligned");
    HpbDeviceGuard guard(ctx->device);
    return hpb_launch_pack_s2d(ctx, x_dev, x_bstride, b, C, H, W, out_dev, C_padded, (cudaStream_t)stream);
}

}  // extern "C"
