"""Which stem shape does cuDNN run fastest?  (run on the GPU box)  The 7x7/s2 stem of the 9-channel network input as a 4x4/s1
convolution of the space-to-depth input with C = 4 * Cs channels (Cs channels reserved per sub-pixel, 9 used), against the
direct 7x7/s2 convolution of a channel-padded NHWC input.  bf16, channels_last, fused bias + ReLU, 576 rows."""
import json, sys
import torch

torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
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


for C in (64, 48, 40, 32):
    x = torch.randn(B, C, 123, 163, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(64, C, 4, 4, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = torch.randn(64, device=dev, dtype=torch.bfloat16)
    try:
        ms = timeit(lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1), (0, 0), (1, 1), 1))
        print(json.dumps({"stem": f"s2d 4x4/s1, C={C}", "rows": B, "ms": round(ms, 4), "K": 16 * C}))
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"stem": f"s2d 4x4/s1, C={C}", "error": str(e)[:200]}))
for C in (16, 12, 9):
    x = torch.randn(B, C, 240, 320, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(64, C, 7, 7, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = torch.randn(64, device=dev, dtype=torch.bfloat16)
    try:
        ms = timeit(lambda: torch.cudnn_convolution_relu(x, w, b, (2, 2), (3, 3), (1, 1), 1))
        print(json.dumps({"stem": f"direct 7x7/s2, C={C}", "rows": B, "ms": round(ms, 4), "K": 49 * C}))
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"stem": f"direct 7x7/s2, C={C}", "error": str(e)[:200]}))
# 2x space-to-depth twice (4x4 block -> 144 channels, padded to Cs per sub-pixel): 2x2/s1 ... not equivalent to 7x7/s2; skipped.
# the 3x3/s2 max-pool that follows, for scale
y = torch.randn(B, 64, 120, 160, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
ms = timeit(lambda: torch.nn.functional.max_pool2d(y, 3, 2, 1))
print(json.dumps({"op": "torch max_pool2d 3x3/s2 of the stem output", "ms": round(ms, 4)}))
