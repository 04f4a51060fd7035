"""Per-phase SM-clock breakdown of hpb_raster_kernel (debug build with -DHPB_PHASE_CLOCKS, never the shipped library).
usage (GPU box): python scripts/raster_phases.py [b]"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from happypose_b200 import _build
dbg = os.path.join(ROOT, "build", "libhpb200_dbg.so")
if "--build" in sys.argv or not os.path.exists(dbg):
    os.makedirs(os.path.dirname(dbg), exist_ok=True)
    cmd = [_build._nvcc()] + _build.NVCC_FLAGS + ["-DHPB_PHASE_CLOCKS", "-o", dbg] + [os.path.join(_build.CSRC, s) for s in _build.SOURCES] + ["-lcudart"]
    subprocess.run(cmd, check=True)
    if "--build" in sys.argv:
        sys.exit(0)
_build.LIB_PATH = dbg
_build.is_stale = lambda: False
import numpy as np, torch
import bench as B
from happypose_b200 import ops, _capi
from happypose_b200._capi import Context
from happypose_b200.utils import transform_utils

dev = torch.device("cuda:0")
ctx = Context.get(dev)
lib = ctypes.CDLL(dbg)
d = np.load(B.MESH)
pos = (d["verts"].astype(np.float64) * 0.001).astype(np.float32)
mid = ops.mesh_upload(ctx, pos, d["faces"], d["normals"], d["uv"], texture=d["texture"])
grid = transform_utils.load_SO3_grid(576).to(dev)
pts_all = torch.as_tensor(pos[None]).to(dev)
pts = torch.as_tensor(pos[np.random.RandomState(0).choice(len(pos), 2000, replace=False)][None]).to(dev)
img = torch.rand(1, 3, 480, 640, device=dev)
b = int([a for a in sys.argv[1:] if a.isdigit()][0]) if any(a.isdigit() for a in sys.argv[1:]) else 576
R = grid[torch.arange(b, device=dev) % 576]
zero = torch.zeros(b, dtype=torch.int32, device=dev)
K = torch.as_tensor(B.K_BBQ).to(dev).expand(b, 3, 3).contiguous()
boxes = torch.as_tensor(B.BBOX_BBQ).to(dev).expand(b, 4).contiguous()
TCO = ops.tco_init(ctx, _capi.TCO_INIT_AUTODEPTH_WITH_R, boxes, K, pts_all, zero, R)
ids = torch.full((b,), mid, dtype=torch.int32, device=dev)
x = torch.empty((b, 9, 240, 320), device=dev)
_, K_crop, _, _ = ops.crop(ctx, img, zero, pts, zero, K, TCO, TCO[:, :3, 3].contiguous(), (240, 320), out=x)
out = (ctypes.c_ulonglong * 8)()
for _ in range(3):
    ops.render(ctx, ids, TCO, K_crop, (240, 320), render_normals=True, out=x, out_channel_offset=3)
torch.cuda.synchronize()
lib.hpb_debug_phase_clocks(out)
n = 10
for _ in range(n):
    ops.render(ctx, ids, TCO, K_crop, (240, 320), render_normals=True, out=x, out_channel_offset=3)
torch.cuda.synchronize()
lib.hpb_debug_phase_clocks(out)
names = ["setup", "A vertex (+barrier)", "B triangles (thread 0)", "B barrier wait", "C resolve (thread 0)", "C barrier wait"]
tot = sum(out[i] for i in range(6))
print(f"b={b}: clocks per scene (thread 0 of the CTA), total {tot / n / b:.0f}")
for i, nm in enumerate(names):
    print(f"  {nm:26s} {out[i] / n / b:9.0f}  {100 * out[i] / tot:5.1f}%")
