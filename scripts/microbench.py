"""Quick kernel timing (CUDA events) for the rasteriser and crop kernels on BASELINE config #1/#3 shapes."""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from happypose_b200 import ops
from happypose_b200._capi import Context
from tests.scenes import random_crop_scene


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


def main():
    ctx = Context.get("cuda:0")
    d = np.load("tests/golden/obj_000001.npz")
    pos = (d["verts"].astype(np.float64) * 0.001).astype(np.float32)
    mid = ops.mesh_upload(ctx, pos, d["faces"], d["normals"], d["uv"], texture=d["texture"])
    rs = np.random.RandomState(0)
    bs = [int(a) for a in sys.argv[1:]] or [4, 32, 148, 576, 2304]
    for b in bs:
        T, K = random_crop_scene(rs, b)
        T, K = torch.as_tensor(T).cuda(), torch.as_tensor(K).cuda()
        ids = torch.full((b,), mid, dtype=torch.int32, device="cuda")
        x = torch.empty((b, 9, 240, 320), device="cuda")
        ms = timeit(lambda: ops.render(ctx, ids, T, K, (240, 320), render_normals=True, out=x, out_channel_offset=3))
        byts = b * 6 * 240 * 320 * 4
        print(json.dumps({"kernel": "raster rgb+normals", "b": b, "ms": ms, "hyps_per_s": b / ms * 1e3, "GBps": byts / ms / 1e6, "Mtris_per_s": b * 15728 / ms / 1e3}))
        ms = timeit(lambda: ops.render(ctx, ids, T, K, (240, 320), render_rgb=False, render_depth=True))
        print(json.dumps({"kernel": "raster depth only", "b": b, "ms": ms, "hyps_per_s": b / ms * 1e3}))
        img = torch.rand(1, 3, 480, 640, device="cuda")
        pts = torch.as_tensor(pos[np.random.RandomState(0).choice(len(pos), 2000, replace=False)][None]).cuda()
        Kim = torch.tensor([[605.95, 0, 319.03], [0, 605.01, 249.68], [0, 0, 1]]).cuda().expand(b, 3, 3).contiguous()
        Tim = T.clone(); Tim[:, 2, 3] = 0.6
        zero = torch.zeros(b, dtype=torch.int32, device="cuda")
        ms = timeit(lambda: ops.crop(ctx, img, zero, pts, zero, Kim, Tim, Tim[:, :3, 3], (240, 320), out=x))
        print(json.dumps({"kernel": "crop", "b": b, "ms": ms, "hyps_per_s": b / ms * 1e3, "GBps": b * 3 * 240 * 320 * 4 / ms / 1e6}))


if __name__ == "__main__":
    main()
