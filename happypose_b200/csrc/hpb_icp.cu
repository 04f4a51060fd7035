// Input stage of the ICP depth refiner on the device: for every pose estimate, the object mask and the two
// point clouds the reference builds on the host, one object at a time, from full-resolution depth maps.
//
// Replaces, in happypose/pose_estimators/megapose/inference/icp_refiner.py:
//   refine_poses :271-289      per-object .cpu().numpy() of the measured and the rendered depth map + compute_masks
//                              (refiner_utils.py, mask_type="threshold")
//   icp_refinement :138-176    getXYZ (:106-135) of both maps and the boolean-mask selections
//       target  = measured points with 0.2 < d < 5 inside the mask
//       source  = rendered points at the same pixels where something was rendered (d_rendered > 0)
// The rendered depth maps come from hpb_render(render_depth) at frame resolution and never leave the device.  What stays on
// the host is the registration itself (OpenCV's ppf_match_3d ICP) and its normal estimation (cv2.inpaint + Gaussian filter).
//
// Arithmetic follows getXYZ exactly: the pixel offsets are stored in an int16 table there, so (x - cx) and (y - cy) are
// truncated towards zero; x = (float)u * d / fx in float32 (the library is built with -fmad=false).  Points are emitted in
// row-major pixel order, like numpy's boolean-mask indexing: one CTA per object walks the image in chunks of 1024 pixels
// with a block-wide exclusive scan and a running offset (ordered stream compaction).
#include "hpb_common.cuh"

namespace {

constexpr int ICP_THREADS = 1024;

struct IcpParams {
    const float *depth_measured;  // [n_im, H, W]
    const float *depth_rendered;  // [N, H, W]
    const uint8_t *masks;         // [n_im, H, W] or nullptr -> threshold mask
    const int32_t *im_ids;        // [N]
    const float *K;               // [N, 9]
    int N, H, W;
    float delta;
    long long cap;                // points per object the outputs can hold
    float *pts_tgt, *pts_src;     // [N, cap, 3]
    int32_t *counts;              // [N, 2] (n_tgt, n_src): the TRUE counts, even beyond cap
    uint8_t *mask_out;            // [N, H, W] or nullptr: the mask used (rendered == measured mask, refiner_utils.py)
    int32_t *idx_tgt, *idx_src;   // [N, cap] or nullptr: linear pixel index of every emitted point (to gather host-side normals)
};

__global__ void __launch_bounds__(ICP_THREADS) hpb_icp_points_kernel(const IcpParams p) {
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int sWarpT[32], sWarpS[32];
    __shared__ int sBaseT, sBaseS;
    const long long npix = (long long)p.H * p.W;
    const int im = p.im_ids[n];
    const float *dm = p.depth_measured + (size_t)im * npix;
    const float *dr = p.depth_rendered + (size_t)n * npix;
    const uint8_t *mk = p.masks ? p.masks + (size_t)im * npix : nullptr;
    const float fx = p.K[(size_t)n * 9], fy = p.K[(size_t)n * 9 + 4];
    // np.int16 table of (index - c): float64 subtraction, then truncation towards zero (icp_refiner.py:107-110)
    const double cx = (double)p.K[(size_t)n * 9 + 2], cy = (double)p.K[(size_t)n * 9 + 5];
    float *ot = p.pts_tgt + (size_t)n * p.cap * 3, *os = p.pts_src + (size_t)n * p.cap * 3;
    if (tid == 0) { sBaseT = 0; sBaseS = 0; }
    __syncthreads();
    for (long long base = 0; base < npix; base += ICP_THREADS) {
        const long long i = base + tid;
        bool vt = false, vs = false;
        float m = 0.f, r = 0.f;
        int x = 0, y = 0;
        if (i < npix) {
            m = __ldg(dm + i);
            r = __ldg(dr + i);
            y = (int)(i / p.W);
            x = (int)(i - (long long)y * p.W);
            bool mask;
            if (mk) {
                mask = mk[i] != 0;
            } else {  // compute_masks("threshold"): NaN differences are not "> delta", like numpy
                mask = m > 0.f && r > 0.f && !(fabsf(m - r) > p.delta);
            }
            if (p.mask_out) p.mask_out[(size_t)n * npix + i] = mask ? 1 : 0;
            vt = m > 0.2f && m < 5.f && mask;
            vs = vt && r > 0.f;
        }
        const unsigned bt = __ballot_sync(0xffffffffu, vt), bs = __ballot_sync(0xffffffffu, vs);
        if (lane == 0) { sWarpT[warp] = __popc(bt); sWarpS[warp] = __popc(bs); }
        __syncthreads();
        int offT = sBaseT, offS = sBaseS;  // running totals + the warps in front of this one
        for (int w2 = 0; w2 < warp; ++w2) { offT += sWarpT[w2]; offS += sWarpS[w2]; }
        offT += __popc(bt & ((1u << lane) - 1u));
        offS += __popc(bs & ((1u << lane) - 1u));
        if (vt || vs) {
            const float u = (float)(short)(int)((double)x - cx), v = (float)(short)(int)((double)y - cy);
            if (vt && offT < p.cap) {
                float *o = ot + (size_t)offT * 3;
                o[0] = u * m * 1.0f / fx; o[1] = v * m * 1.0f / fy; o[2] = m;
                if (p.idx_tgt) p.idx_tgt[(size_t)n * p.cap + offT] = (int32_t)i;
            }
            if (vs && offS < p.cap) {
                float *o = os + (size_t)offS * 3;
                o[0] = u * r * 1.0f / fx; o[1] = v * r * 1.0f / fy; o[2] = r;
                if (p.idx_src) p.idx_src[(size_t)n * p.cap + offS] = (int32_t)i;
            }
        }
        __syncthreads();
        if (tid == 0) {
            int tT = 0, tS = 0;
            for (int w2 = 0; w2 < ICP_THREADS / 32; ++w2) { tT += sWarpT[w2]; tS += sWarpS[w2]; }
            sBaseT += tT; sBaseS += tS;
        }
        __syncthreads();
    }
    if (tid == 0) { p.counts[2 * n] = sBaseT; p.counts[2 * n + 1] = sBaseS; }
}

}  // namespace

int hpb_launch_icp_points(hpb_ctx *ctx, const float *depth_measured, int n_im, const float *depth_rendered, const uint8_t *masks,
                          const int32_t *im_ids, const float *K, int N, int H, int W, float delta, int64_t cap, float *pts_tgt,
                          float *pts_src, int32_t *counts, uint8_t *mask_out, int32_t *idx_tgt, int32_t *idx_src, cudaStream_t stream) {
    if (N == 0) return HPB_OK;
    (void)n_im;
    IcpParams p;
    p.depth_measured = depth_measured; p.depth_rendered = depth_rendered; p.masks = masks; p.im_ids = im_ids; p.K = K;
    p.N = N; p.H = H; p.W = W; p.delta = delta; p.cap = cap; p.pts_tgt = pts_tgt; p.pts_src = pts_src; p.counts = counts;
    p.mask_out = mask_out; p.idx_tgt = idx_tgt; p.idx_src = idx_src;
    hpb_icp_points_kernel<<<N, ICP_THREADS, 0, stream>>>(p);
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}
