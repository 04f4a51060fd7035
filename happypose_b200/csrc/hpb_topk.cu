// Segmented top-K for sm_100a: replaces the pandas sort_values().groupby().head(K) of
// happypose/toolbox/utils/tensor_collection.py:201-230 (filter_top_pose_estimates), which forces a
// device->host copy of every coarse logit (megapose/inference/pose_estimator.py:462-463).
//
// Output contract (bit-exact indices given identical scores): surviving row indices in GLOBAL descending
// score order, at most K per group; ties broken by the lower row index; NaN scores last.
//
// Everything is exact integer work on 64-bit keys  key(i) = (order(score_i) << 32) | i  where order() maps a
// float to an unsigned that ascends as the score descends.  Kernels:
//   1. keys + per-group histogram            (one thread per row, atomicAdd on n_groups counters)
//   2. exclusive scan of the histogram       (single CTA; n_groups is small) + output count
//   3. bucket rows by group                  (atomic cursor; order inside a bucket is irrelevant)
//   4. rank inside the group by counting     (one CTA per group, keys tiled through shared memory) and scatter
//      the K best of each group into that group's sorted slot list
//   5. global position of every survivor     = sum over groups of lower_bound(sorted slots of the group, key)
//      (binary searches; no global sort needed) -> out_idx[pos] = row
#include "hpb_common.cuh"

namespace {

__device__ __forceinline__ unsigned order_desc(float s) {
    if (s != s) return 0xffffffffu;  // NaN last
    unsigned u = __float_as_uint(s);
    // ascending-orderable transform, then invert for descending
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    unsigned inv = ~u;
    if (inv == 0xffffffffu) inv = 0xfffffffeu;  // keep NaN strictly last
    return inv;
}

__global__ void topk_keys_kernel(const float *scores, const int32_t *groups, int n, int n_groups,
                                 unsigned long long *keys, int *counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = scores[i];
    if (s == 0.0f) s = 0.0f;  // -0.0 == +0.0 must tie (pandas compares values)
    keys[i] = ((unsigned long long)order_desc(s) << 32) | (unsigned)i;
    const int g = groups[i];
    if (g >= 0 && g < n_groups) atomicAdd(counts + g, 1);
}

__global__ void topk_scan_kernel(const int *counts, int n_groups, int K, int *offsets, int *kept_offsets,
                                 int *cursors, int32_t *out_count) {
    // single thread: n_groups is at most a few thousand detections
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int acc = 0, kacc = 0;
        for (int g = 0; g < n_groups; ++g) {
            offsets[g] = acc;
            kept_offsets[g] = kacc;
            cursors[g] = 0;
            acc += counts[g];
            kacc += min(counts[g], K);
        }
        offsets[n_groups] = acc;
        kept_offsets[n_groups] = kacc;
        *out_count = kacc;
    }
}

__global__ void topk_bucket_kernel(const int32_t *groups, const unsigned long long *keys, int n, int n_groups,
                                   const int *offsets, int *cursors, unsigned long long *bucketed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = groups[i];
    if (g < 0 || g >= n_groups) return;
    const int pos = atomicAdd(cursors + g, 1);
    bucketed[offsets[g] + pos] = keys[i];
}

constexpr int RANK_TILE = 1024;

__global__ void __launch_bounds__(256) topk_rank_kernel(const unsigned long long *bucketed, const int *offsets,
                                                        const int *kept_offsets, int K, unsigned long long *kept) {
    const int g = blockIdx.x;
    const int beg = offsets[g], cnt = offsets[g + 1] - beg;
    const int kbeg = kept_offsets[g];
    __shared__ unsigned long long tile[RANK_TILE];
    for (int base = 0; base < cnt; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const unsigned long long mine = i < cnt ? bucketed[beg + i] : 0ull;
        int rank = 0;
        for (int t0 = 0; t0 < cnt; t0 += RANK_TILE) {
            const int tn = min(RANK_TILE, cnt - t0);
            __syncthreads();
            for (int k = threadIdx.x; k < tn; k += blockDim.x) tile[k] = bucketed[beg + t0 + k];
            __syncthreads();
            if (i < cnt)
                for (int k = 0; k < tn; ++k) rank += tile[k] < mine;
        }
        if (i < cnt && rank < K) kept[kbeg + rank] = mine;  // keys are unique -> ranks are a permutation
    }
}

__global__ void topk_place_kernel(const unsigned long long *kept, const int *kept_offsets, int n_groups,
                                  int64_t *out_idx) {
    const int total = kept_offsets[n_groups];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const unsigned long long mine = kept[i];
    int pos = 0;
    for (int g = 0; g < n_groups; ++g) {
        int lo = kept_offsets[g], hi = kept_offsets[g + 1];
        const int b0 = lo;
        while (lo < hi) {  // lower_bound: number of this group's survivors with a smaller key
            const int mid = (lo + hi) >> 1;
            if (kept[mid] < mine) lo = mid + 1;
            else hi = mid;
        }
        pos += lo - b0;
    }
    out_idx[pos] = (int64_t)(mine & 0xffffffffull);
}

}  // namespace

int hpb_launch_topk(hpb_ctx *ctx, const float *scores, const int32_t *groups, int n, int n_groups, int K,
                    int64_t *out_idx, int32_t *out_count, cudaStream_t stream) {
    if (n == 0 || n_groups == 0 || K <= 0) {
        HPB_CUDA_OK(cudaMemsetAsync(out_count, 0, sizeof(int32_t), stream));
        return HPB_OK;
    }
    // workspace layout: keys[n] | bucketed[n] | kept[n] | counts[G] | offsets[G+1] | kept_offsets[G+1] | cursors[G]
    {
        const int rc = hpb_topk_reserve(ctx, n, n_groups, stream);
        if (rc != HPB_OK) return rc;
        const int rc2 = hpb_stream_enter(ctx, stream);
        if (rc2 != HPB_OK) return rc2;
    }
    unsigned long long *keys = (unsigned long long *)ctx->topk_ws;
    unsigned long long *bucketed = keys + n;
    unsigned long long *kept = bucketed + n;
    int *counts = (int *)(kept + n);
    int *offsets = counts + n_groups;
    int *kept_offsets = offsets + n_groups + 1;
    int *cursors = kept_offsets + n_groups + 1;
    HPB_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int) * n_groups, stream));
    const int tb = 256, nb = (n + tb - 1) / tb;
    topk_keys_kernel<<<nb, tb, 0, stream>>>(scores, groups, n, n_groups, keys, counts);
    topk_scan_kernel<<<1, 32, 0, stream>>>(counts, n_groups, K, offsets, kept_offsets, cursors, out_count);
    topk_bucket_kernel<<<nb, tb, 0, stream>>>(groups, keys, n, n_groups, offsets, cursors, bucketed);
    topk_rank_kernel<<<n_groups, 256, 0, stream>>>(bucketed, offsets, kept_offsets, K, kept);
    const long long max_kept = (long long)n_groups * K < n ? (long long)n_groups * K : n;
    topk_place_kernel<<<(int)((max_kept + tb - 1) / tb), tb, 0, stream>>>(kept, kept_offsets, n_groups, out_idx);
    HPB_CUDA_OK(cudaGetLastError());
    ctx->launches += 5;
    hpb_stream_leave(ctx, stream);
    return HPB_OK;
}

int hpb_topk_reserve(hpb_ctx *ctx, int64_t n, int64_t n_groups, cudaStream_t stream) {
    const size_t need = sizeof(unsigned long long) * 3 * (size_t)n + sizeof(int) * (4 * (size_t)n_groups + 2);
    return hpb_ws_grow(ctx, &ctx->topk_ws, &ctx->topk_ws_bytes, need, stream, "top-K");
}
