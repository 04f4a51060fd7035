"""A/B of a ResNet-34 layer1 convolution at the coarse batch (576 x 64 x 60 x 80, bf16 NHWC): cuDNN fused conv+bias+ReLU and
conv+bias+add+ReLU vs the tcgen05 kernel (hpb_conv3x3_tc.cu).  Interleaved rounds (run on the GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from happypose_b200 import ops
from happypose_b200._capi import Context

torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
ctx = Context.get(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 576


def timeit(fn, warm=3, it=10):
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


x = torch.randn(B, 64, 60, 80, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
z = torch.randn(B, 64, 60, 80, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
w = (torch.randn(64, 64, 3, 3, device=dev) * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
bias = torch.randn(64, device=dev)
bias_h = bias.to(torch.bfloat16)
res = {"cudnn_relu": [], "cudnn_add_relu": [], "tc_relu": [], "tc_add_relu": []}
for rnd in range(4):
    res["cudnn_relu"].append(timeit(lambda: torch.cudnn_convolution_relu(x, w, bias_h, (1, 1), (1, 1), (1, 1), 1)))
    res["tc_relu"].append(timeit(lambda: ops.conv3x3_bias_relu_bf16(ctx, x, w, bias)))
    res["cudnn_add_relu"].append(timeit(lambda: torch.cudnn_convolution_add_relu(x, w, z, 1.0, bias_h, (1, 1), (1, 1), (1, 1), 1)))
    res["tc_add_relu"].append(timeit(lambda: ops.conv3x3_bias_relu_bf16(ctx, x, w, bias, z)))
flop = 2.0 * B * 60 * 80 * 64 * 576
a = ops.conv3x3_bias_relu_bf16(ctx, x, w, bias, z).float()
r = torch.cudnn_convolution_add_relu(x, w, z, 1.0, bias_h, (1, 1), (1, 1), (1, 1), 1).float()
print(json.dumps({"rows": B, **{k: [round(v, 4) for v in vs] for k, vs in res.items()},
                  "tflops": {k: round(flop / min(vs) / 1e9, 1) for k, vs in res.items()},
                  "max_abs_diff_vs_cudnn": float((a - r).abs().max()), "frac_equal": float((a == r).float().mean())}))
