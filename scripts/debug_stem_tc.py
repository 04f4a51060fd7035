"""Decode which input element the tcgen05 stem kernel actually reads for every (tap, output pixel, channel): identity weights on
ONE tap make out[o, y, x] = z[o, y + kh, x + kw]; z encodes its own x, y or channel index (exact in bf16).
This is how the un-swizzle rule of a descriptor whose start address is not 1 KB aligned was found (hpb_stem_tc.cu header):
usage: python scripts/debug_stem_tc.py [variant 0|1|2]   (run on the GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from happypose_b200 import ops
from happypose_b200._capi import Context

dev = torch.device("cuda:0")
ctx = Context.get(dev)
halo = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx.lib.hpb_set_stem_tc_halo(ctx.handle, halo)
Hz, Wz = 19, 19  # one image, 16 x 16 outputs: two tiles of the halo scheme, two of the box-per-tap scheme
yy, xx, cc = torch.meshgrid(torch.arange(Hz), torch.arange(Wz), torch.arange(64), indexing="ij")  # [Hz, Wz, 64]
enc = {"x": xx, "y": yy, "c": cc}
bias = torch.zeros(64, device=dev)
for tap in range(16):
    kh, kw = tap >> 2, tap & 3
    w = torch.zeros(64, 64, 4, 4)
    w[torch.arange(64), torch.arange(64), kh, kw] = 1.0
    w = w.to(torch.bfloat16).to(dev).contiguous(memory_format=torch.channels_last)
    dec = {}
    for name, e in enc.items():
        z = e.permute(2, 0, 1)[None].float().to(torch.bfloat16).to(dev).contiguous(memory_format=torch.channels_last)
        out = ops.stem_conv4x4_relu_bf16(ctx, z, w, bias)
        if out is None:
            print("declined"); sys.exit(0)
        dec[name] = out[0].float().permute(1, 2, 0).round().long().cpu()  # [Hc, Wc, 64]
    Hc, Wc = dec["x"].shape[:2]
    oy, ox, oc = torch.meshgrid(torch.arange(Hc), torch.arange(Wc), torch.arange(64), indexing="ij")
    dx, dy, dc = dec["x"] - (ox + kw), dec["y"] - (oy + kh), dec["c"] - oc
    bad = (dx != 0) | (dy != 0) | (dc != 0)
    print(f"tap kh={kh} kw={kw}: wrong {int(bad.sum())} of {bad.numel()}")
    if bad.any():
        # summarise by (output x % 8, channel chunk): the displacement read
        seen = {}
        for y, x, c in bad.nonzero()[:4000].tolist():
            key = (x % 8, c // 8)
            val = (int(dy[y, x, c]), int(dx[y, x, c]), int(dc[y, x, c]) // 8 if dc[y, x, c] % 8 == 0 else int(dc[y, x, c]))
            seen.setdefault(key, set()).add(val)
        for key in sorted(seen)[:24]:
            print("   (x%8, chunk)", key, "-> (dy, dx, dchunk)", sorted(seen[key])[:4])
