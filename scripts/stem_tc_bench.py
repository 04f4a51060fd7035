"""A/B of the stem: cuDNN fused conv+bias+ReLU vs the tcgen05 implicit GEMM (hpb_stem_tc.cu), 576 rows (run on the GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from happypose_b200 import ops
from happypose_b200._capi import Context

torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
ctx = Context.get(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 576


def timeit(fn, warm=5, it=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


z = torch.randn(B, 64, 123, 163, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
from happypose_b200.megapose.fast_resnet import s2d_weight
w = s2d_weight(torch.randn(64, 9, 7, 7, device=dev) * 0.05, 64).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
mask = ops.stem_k_slice_mask(w)
print(json.dumps({"k_slices_non_zero": bin(mask).count("1")}))
bias = torch.randn(64, device=dev)
bias_h = bias.to(torch.bfloat16)
ms_cudnn = timeit(lambda: torch.cudnn_convolution_relu(z, w, bias_h, (1, 1), (0, 0), (1, 1), 1))
ref = torch.cudnn_convolution_relu(z, w, bias_h, (1, 1), (0, 0), (1, 1), 1)
flop = 2.0 * B * 120 * 160 * 64 * 1024
names = {1: "halo box per tile, staged TMA-store epilogue (shipped)", 2: "halo box per tile, register-store epilogue", 0: "box per tap"}
res = {m: {"masked": [], "dense": []} for m in names}
outs = {}
for rnd in range(4):  # interleaved rounds: the first configuration timed after an idle gap runs at lower clocks
    for mode in (1, 2, 0):
        ctx.lib.hpb_set_stem_tc_halo(ctx.handle, mode)
        outs[mode] = ops.stem_conv4x4_relu_bf16(ctx, z, w, bias, mask)
        res[mode]["masked"].append(timeit(lambda: ops.stem_conv4x4_relu_bf16(ctx, z, w, bias, mask)))
        res[mode]["dense"].append(timeit(lambda: ops.stem_conv4x4_relu_bf16(ctx, z, w, bias)))
    res.setdefault("cudnn", []).append(timeit(lambda: torch.cudnn_convolution_relu(z, w, bias_h, (1, 1), (0, 0), (1, 1), 1)))
ctx.lib.hpb_set_stem_tc_halo(ctx.handle, 1)
ms_cudnn = min(res["cudnn"])
for mode, name in names.items():
    ms_tc, ms_dense = min(res[mode]["masked"]), min(res[mode]["dense"])
    print(json.dumps({"rows": B, "scheme": name, "cudnn_ms": round(ms_cudnn, 4), "tcgen05_ms": round(ms_tc, 4), "tcgen05_all_slices_ms": round(ms_dense, 4),
                      "rounds_ms": [round(x, 4) for x in res[mode]["masked"]], "issued_tflops": round(flop * 49 / 64 * (128 / 120) / ms_tc / 1e9, 1),
                      "cudnn_tflops_on_all_slices": round(flop / ms_cudnn / 1e9, 1),
                      "max_abs_diff_vs_cudnn": float((outs[mode].float() - ref.float()).abs().max()), "frac_equal": float((outs[mode] == ref).float().mean())}))
