"""BASELINE config #5 class on ONE GPU: GSO-scale meshes (icosphere level 6 + bumps: 40 962 vertices, 81 920 triangles,
vertex colours; their 12 B / vertex screen-space arrays exceed shared memory, so the vertex stage runs in the CTA's
global scratch slice), n_mesh distinct meshes, b hypotheses per launch cycling through them.  Prints render / crop /
fused hand-off throughput.  usage (GPU box): python scripts/config5_bench.py [n_mesh] [b]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from happypose_b200 import ops, _capi
from happypose_b200._capi import Context
from tests.scenes import icosphere, random_rotations

def timeit(fn, warm=2, it=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

n_mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 16
b = int(sys.argv[2]) if len(sys.argv) > 2 else 576
dev = torch.device("cuda:0")
ctx = Context(dev)  # own context: its scratch is sized by these meshes
rs = np.random.RandomState(5)
v0, f, n0 = icosphere(6, 0.08)
mids, pts = [], []
for m in range(n_mesh):
    k1, k2 = rs.randint(3, 12, 2)
    bump = 1.0 + 0.05 * np.sin(k1 * n0[:, :1]) * np.cos(k2 * n0[:, 1:2]) + 0.004 * rs.randn(len(v0), 1)
    v = (v0 * bump).astype(np.float32)
    col = (rs.rand(len(v), 3) * 255).astype(np.uint8)
    mids.append(ops.mesh_upload(ctx, v, f, None, vcolor=col) if False else ops.mesh_upload(ctx, v, f, n0, vcolor=col))
    pts.append(v[rs.choice(len(v), 2000, replace=False)])
pts = torch.as_tensor(np.stack(pts)).to(dev)
obj = torch.as_tensor(np.arange(b) % n_mesh, dtype=torch.int32).to(dev)
ids = torch.as_tensor(np.array(mids, np.int32)[np.arange(b) % n_mesh]).to(dev)
T = np.tile(np.eye(4, dtype=np.float32), (b, 1, 1))
T[:, :3, :3] = random_rotations(rs, b)
T[:, :3, 3] = np.stack([rs.uniform(-0.1, 0.1, b), rs.uniform(-0.08, 0.08, b), rs.uniform(0.5, 1.5, b)], 1)
T = torch.as_tensor(T).to(dev)
K = torch.as_tensor(B.K_BBQ).to(dev).expand(b, 3, 3).contiguous()
img = torch.rand(1, 3, 480, 640, device=dev)
zero = torch.zeros(b, dtype=torch.int32, device=dev)
x = torch.empty((b, 9, 240, 320), device=dev)
tCR = T[:, :3, 3].contiguous()
_, K_crop, _, _ = ops.crop(ctx, img, zero, pts, obj, K, T, tCR, (240, 320), out=x)
peak = B.measured_peak_gbs()[0]
ms_r = timeit(lambda: ops.render(ctx, ids, T, K_crop, (240, 320), render_normals=True, out=x, out_channel_offset=3))
cov = float((x[:, 3:9].abs().sum(1) > 0).float().mean())
ms_c = timeit(lambda: ops.crop(ctx, img, zero, pts, obj, K, T, tCR, (240, 320), out=x))
ms_f = timeit(lambda: ops.render_s2d_bf16(ctx, ids, T, K_crop, x[:, :3], 64))
print(json.dumps({"config": "#5 class, one GPU", "meshes": n_mesh, "triangles_per_mesh": int(len(f)), "vertices_per_mesh": int(len(v0)), "b": b,
                  "coverage": round(cov, 3), "render_ms": round(ms_r, 3), "render_hyps_per_s": round(b / ms_r * 1e3),
                  "render_triangles_per_s": round(b * len(f) / ms_r * 1e3), "render_GBps": round(b * 6 * 240 * 320 * 4 / 1e9 / ms_r * 1e3, 1),
                  "render_frac": round(b * 6 * 240 * 320 * 4 / 1e9 / ms_r * 1e3 / peak, 4), "crop_ms": round(ms_c, 3),
                  "fused_render_ms": round(ms_f, 3), "render_plus_crop_hyps_per_s": round(b / (ms_f + ms_c) * 1e3)}))
