"""One small TMA-ring crop launch (run under compute-sanitizer when the kernel misbehaves)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from happypose_b200 import ops, _capi
from happypose_b200._capi import Context
from happypose_b200.utils import transform_utils

dev = torch.device("cuda:0")
ctx = Context.get(dev)
d = np.load(B.MESH)
pos = (d["verts"].astype(np.float64) * 0.001).astype(np.float32)
grid = transform_utils.load_SO3_grid(576).to(dev)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pts_all = torch.as_tensor(pos[None]).to(dev)
pts = torch.as_tensor(pos[np.random.RandomState(0).choice(len(pos), 2000, replace=False)][None]).to(dev)
img = torch.rand(1, 3, 480, 640, device=dev)
zero = torch.zeros(b, dtype=torch.int32, device=dev)
K = torch.as_tensor(B.K_BBQ).to(dev).expand(b, 3, 3).contiguous()
boxes = torch.as_tensor(B.BBOX_BBQ).to(dev).expand(b, 4).contiguous()
TCO = ops.tco_init(ctx, _capi.TCO_INIT_AUTODEPTH_WITH_R, boxes, K, pts_all, zero, grid[:b])
tCR = TCO[:, :3, 3].contiguous()
outs = []
for tma in (0, 1):
    ctx.check(ctx.lib.hpb_set_crop_tma(ctx.handle, tma), "set")
    c = ops.crop_bf16x4(ctx, img, zero, pts, zero, K, TCO, tCR, (240, 320), tap_bits=16)[0]
    torch.cuda.synchronize()
    outs.append(c)
    print("tma", tma, "ok", float(c.float().abs().sum()))
print("equal:", torch.equal(outs[0].view(torch.int16), outs[1].view(torch.int16)))
