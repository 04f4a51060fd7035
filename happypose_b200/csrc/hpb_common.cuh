// Internal definitions shared by the kernels and the C-ABI layer of libhpb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/hpb200.h"

#define HPB_MAX_MIPS 16
#define HPB_MAX_LIGHTS 8
#define HPB_SUBPIX_BITS 8
#define HPB_SUBPIX 256
#define HPB_GUARD 4194304.0f  // 2^22 fixed-point units (16384 px) guard band for snapped vertices
#define HPB_VIS_EMPTY 0xffffffffffffffffull

// Device-side mesh record (array indexed by mesh id lives in device memory).
struct HpbMeshDev {
    const float *pos;    // [nv,3] metres
    const float *nrm;    // [nv,3] unit, object frame
    const float *uv;     // [nv,2] or nullptr
    const uchar4 *vcol;  // [nv] RGBA or nullptr
    const int4 *faces;   // [nf] (i0,i1,i2,0): padded to 16 B so one thread loads a triangle in one LDG.128
    const uchar4 *tex;   // RGBA8 mip chain or nullptr
    const float4 *nu;    // [nv] (nx, ny, nz, u): the resolve stage fetches a vertex's normal and u with ONE LDG.128
    const float *tv;     // [nv] v (0 when the mesh has no uv)
    int nv, nf;
    // Back-face culling is semantically invisible on a closed, consistently oriented surface (every pixel a back
    // face covers is also covered by a nearer front face), so the rasteriser skips back faces of such meshes.
    // cull_sign = sign of the screen-space area2 of FRONT faces (-1: outward CCW winding, +1: inward), 0 = not
    // provably closed -> two-sided rendering of every triangle (panda3d_scene_renderer.py:102).
    int cull_sign;
    int tex_levels;
    int tex_pow2;  // 1 when level-0 width and height are powers of two (then every level is): wrap = bit mask
    int tex_w[HPB_MAX_MIPS];
    int tex_h[HPB_MAX_MIPS];
    long long tex_off[HPB_MAX_MIPS];  // texel offset of each level
};

// Screen-space vertices produced by phase A of the rasteriser are kept as two arrays (12 B per vertex):
//   int2 xy  24.8 fixed-point screen position      float iz  1 / Z_cam (0 marks a near-clipped vertex)

struct HpbMeshHost {
    void *pos = nullptr, *nrm = nullptr, *uv = nullptr, *vcol = nullptr, *faces = nullptr, *tex = nullptr, *nu = nullptr, *tv = nullptr;
    HpbMeshDev dev;
    int closed_sign = 0;  // result of the closed-surface analysis (dev.cull_sign is this, or 0 when culling is disabled)
};

struct hpb_ctx {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    std::vector<HpbMeshHost> meshes;
    HpbMeshDev *meshes_dev = nullptr;  // device copy of all HpbMeshDev records
    int meshes_dev_cap = 0;
    int max_nv = 0;
    // rasteriser workspace: one visibility buffer (+ vertex scratch for big meshes) per resident CTA
    unsigned long long *vis = nullptr;
    size_t vis_elems = 0;
    unsigned char *vert_scratch = nullptr;  // [resident CTAs][max_nv * 12 B] when a mesh does not fit in shared memory
    size_t vert_scratch_bytes = 0;
    int max_clusters[5] = {0, 0, 0, 0, 0};  // co-resident clusters of size 1,2,4,8,16 (queried once per shared-memory size)
    size_t max_clusters_smem = (size_t)-1;
    // crop workspace: pixel-interleaved copy of the observed frames
    void *frame_pack = nullptr;
    size_t frame_pack_bytes = 0;
    int crop_tap_bits = 32;  // 16: the interleaved frame copy sampled by the crop kernel holds fp16 (hpb_set_crop_tap_precision)
    // top-k workspace
    void *topk_ws = nullptr;
    size_t topk_ws_bytes = 0;
    int64_t launches = 0;
    int crop_tma = 0;     // hpb_crop*: TMA-fed shared-memory ring for fp16 taps of RGB frames (hpb_set_crop_tma); measured: no gain
    int stem_tc_halo = 1;  // hpb_stem_conv4x4_relu_bf16_nhwc: one halo box per tile (1) or one box per tap (0); hpb_set_stem_tc_halo
    int maxpool_tma = 1;  // hpb_maxpool3x3s2_bf16_nhwc: TMA-staged tile kernel for C = 64 (hpb_set_maxpool_tma)
    unsigned long long *clipped_scenes = nullptr;  // device counter (hpb_raster_clipped_scenes)
    // Workspaces only ever GROW, and a buffer that is replaced is retired (kept allocated until hpb_destroy), never freed:
    // kernel parameters baked into captured CUDA graphs keep pointing at valid, correctly initialised memory.
    // workspace_epoch counts replacements so that a caller can re-capture its graphs onto the new buffers.
    std::vector<void *> retired;
    int64_t workspace_epoch = 0;
    // single-stream use of the shared scratch is enforced by an event hand-over when the launching stream changes
    cudaStream_t last_stream = nullptr;
    cudaEvent_t done_ev = nullptr;
    bool done_ev_valid = false;
};

// Grow *ptr to at least `need` bytes (no-op when large enough).  Never frees: the old buffer is retired.  Fails with a clear
// message when the stream is being captured into a CUDA graph (allocation is illegal there): call hpb_reserve() first.
int hpb_ws_grow(hpb_ctx *ctx, void **ptr, size_t *cur_bytes, size_t need, cudaStream_t stream, const char *what);
// Orders work on `stream` behind the last launch that used the context's shared scratch on another stream.
int hpb_stream_enter(hpb_ctx *ctx, cudaStream_t stream);
void hpb_stream_leave(hpb_ctx *ctx, cudaStream_t stream);
int hpb_raster_reserve(hpb_ctx *ctx, int h, int w, cudaStream_t stream);
int hpb_topk_reserve(hpb_ctx *ctx, int64_t n, int64_t n_groups, cudaStream_t stream);
int hpb_crop_reserve(hpb_ctx *ctx, int64_t frame_pixels, cudaStream_t stream);

void hpb_set_error(const char *fmt, ...);

#define HPB_CUDA_OK(expr)                                                                      \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            hpb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return HPB_ECUDA;                                                                   \
        }                                                                                       \
    } while (0)

#define HPB_REQUIRE(cond, msg)                                        \
    do {                                                              \
        if (!(cond)) {                                                \
            hpb_set_error("%s: %s (%s)", __func__, msg, #cond);       \
            return HPB_EINVAL;                                        \
        }                                                             \
    } while (0)

struct HpbDeviceGuard {
    int prev = -1;
    explicit HpbDeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~HpbDeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// kernel launchers implemented in the .cu files
int hpb_launch_raster(hpb_ctx *ctx, const int32_t *mesh_ids, const float *TCO, const float *K, const float *ambient,
                      int b, int h, int w, float z_near, float z_far, uint32_t flags, float *rgb, int64_t rgb_bs,
                      float *nrm, int64_t nrm_bs, float *depth, int64_t depth_bs, uint8_t *mask, int64_t mask_bs,
                      int views, int64_t view_stride, cudaStream_t stream, const void *crops = nullptr,
                      int64_t crops_bs = 0, void *s2d_out = nullptr, int Cz = 0, int crops_fmt = 0, int pad_prezeroed = 0,
                      const float *lights = nullptr, int n_lights = 0);
int hpb_launch_mip(const uchar4 *src, int sw, int sh, uchar4 *dst, int dw, int dh, cudaStream_t stream);
int hpb_launch_tex_expand(const uint8_t *src, int n, int c, uchar4 *dst, cudaStream_t stream);
