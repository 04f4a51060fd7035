// Fused prologue of one refiner iteration (PosePredictor.forward, megapose/models/pose_rigid.py:570-612): everything
// between the previous iteration's pose and the first image kernel, in ONE launch instead of ~12 tiny ones
// (normalize_T, tCR slice + copy, make_TCO_multiview, tCV_R slice + copy, the row's crop boxes / K_crop, the per-view
// ids / K expansion, compute_crops_multiview, KV_crop[:, 0] = K_crop).  The refiner runs 5 dependent iterations on a
// handful of rows, so these launches are pure latency: ~25 us of a 400 us iteration.
//
// One CTA per row.  Thread 0 normalises the pose; one thread per extra view places its camera (float64 look-at, the same
// device functions as hpb_multiview); then the whole CTA projects the row's 2000-point set (crop box, K_crop) and the
// 200-point set once per view (KV_crop).  Built from the functions in hpb_pose_math.cuh, so every output is bit-identical
// to hpb_normalize_T -> hpb_multiview -> hpb_crop_boxes (x2).
#include "hpb_pose_math.cuh"

namespace {

using namespace hpbm;

constexpr int PROLOGUE_MAX_VIEWS = 27;  // sphere_26views + TCO

struct PrologueParams {
    const float *TCO_in, *K;
    const int32_t *obj_ids;
    const float *pts_crop, *pts_mv;  // [n_obj, n_crop, 3], [n_obj, n_mv, 3]
    int n_crop, n_mv, b;
    int H, W, h, w;
    float lamb;
    int n_extra, n_views, keep_tco;
    MvPositions mv;
    float *T_norm, *tCR, *TCV_O, *K_crop, *boxes_rend, *boxes_crop, *KV_crop;
};

__global__ void __launch_bounds__(256) hpb_refiner_prologue_kernel(const PrologueParams p) {
    const int n = blockIdx.x, tid = threadIdx.x;
    __shared__ float sT[16];
    __shared__ float sTV[PROLOGUE_MAX_VIEWS][16];
    __shared__ float sKc[9];
    __shared__ float sP[12];
    __shared__ float red[4][8];
    __shared__ MvFrame sF;
    const float *K = p.K + (size_t)n * 9;
    if (tid == 0) {
        float o[16];
        normalize_T_row(p.TCO_in + (size_t)n * 16, o);
        for (int k = 0; k < 16; ++k) { sT[k] = o[k]; p.T_norm[(size_t)n * 16 + k] = o[k]; }
        p.tCR[(size_t)n * 3] = o[3]; p.tCR[(size_t)n * 3 + 1] = o[7]; p.tCR[(size_t)n * 3 + 2] = o[11];
        if (p.n_extra > 0) {
            const float c3[3] = {o[3], o[7], o[11]};  // reference point = object origin (tOR = 0, pose_rigid.py:574-576)
            multiview_frame(o, c3, sF);
        }
    }
    __syncthreads();
    const int v0 = p.keep_tco ? 1 : 0;
    if (tid < p.n_views) {
        float ov[16];
        if (tid < v0) {
            for (int k = 0; k < 16; ++k) ov[k] = sT[k];
        } else {
            const int e = tid - v0;
            multiview_view(sF, sT, p.mv.p[3 * e], p.mv.p[3 * e + 1], p.mv.p[3 * e + 2], ov);
        }
        for (int k = 0; k < 16; ++k) { sTV[tid][k] = ov[k]; p.TCV_O[((size_t)n * p.n_views + tid) * 16 + k] = ov[k]; }
    }
    __syncthreads();
    const int obj = p.obj_ids[n];
    const float tcr[3] = {sT[3], sT[7], sT[11]};
    crop_boxes_cta(K, sT, tcr, p.pts_crop + (size_t)obj * p.n_crop * 3, p.n_crop, p.H, p.W, p.h, p.w, p.lamb, sKc,
                   p.boxes_rend + (size_t)n * 4, p.boxes_crop + (size_t)n * 4, sP, red);
    __syncthreads();
    if (tid < 9) p.K_crop[(size_t)n * 9 + tid] = sKc[tid];
    if (p.KV_crop) {
        for (int v = 0; v < p.n_views; ++v) {
            float *kv = p.KV_crop + ((size_t)n * p.n_views + v) * 9;
            if (v < v0) {  // KV_crop[:, 0] = K_crop (pose_rigid.py:610-611)
                if (tid < 9) kv[tid] = sKc[tid];
                continue;
            }
            const float tv[3] = {sTV[v][3], sTV[v][7], sTV[v][11]};
            crop_boxes_cta(K, sTV[v], tv, p.pts_mv + (size_t)obj * p.n_mv * 3, p.n_mv, p.H, p.W, p.h, p.w, p.lamb, kv, nullptr, nullptr,
                           sP, red);
        }
    }
}

}  // namespace

int hpb_launch_refiner_prologue(hpb_ctx *ctx, const float *TCO_in, const float *K, const int32_t *obj_ids, const float *pts_crop,
                                int n_crop, const float *pts_mv, int n_mv, int b, int H, int W, int h, int w, float lamb,
                                const float *positions_host, int n_extra, int n_views, int keep_tco, float *T_norm, float *tCR,
                                float *TCV_O, float *K_crop, float *boxes_rend, float *boxes_crop, float *KV_crop,
                                cudaStream_t stream) {
    if (b == 0) return HPB_OK;
    if (n_views > PROLOGUE_MAX_VIEWS) {
        hpb_set_error("hpb_refiner_prologue: at most %d views", PROLOGUE_MAX_VIEWS);
        return HPB_EINVAL;
    }
    PrologueParams p;
    p.TCO_in = TCO_in; p.K = K; p.obj_ids = obj_ids; p.pts_crop = pts_crop; p.pts_mv = pts_mv; p.n_crop = n_crop; p.n_mv = n_mv;
    p.b = b; p.H = H; p.W = W; p.h = h; p.w = w; p.lamb = lamb; p.n_extra = n_extra; p.n_views = n_views; p.keep_tco = keep_tco;
    for (int i = 0; i < 26 * 3; ++i) p.mv.p[i] = i < 3 * n_extra ? positions_host[i] : 0.0f;
    p.T_norm = T_norm; p.tCR = tCR; p.TCV_O = TCV_O; p.K_crop = K_crop; p.boxes_rend = boxes_rend; p.boxes_crop = boxes_crop;
    p.KV_crop = KV_crop;
    hpb_refiner_prologue_kernel<<<b, 256, 0, stream>>>(p);
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches++;
    return HPB_OK;
}
