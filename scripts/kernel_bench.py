"""Kernel timing (CUDA events) for the rasteriser and crop kernels on the bench's real-use poses (SO(3) grid + bbox init).
usage: kernel_bench.py [b ...]   (b = scenes per launch)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from happypose_b200 import ops, _capi, _build
if os.environ.get("HPB200_LIB"):  # A/B of a library variant built elsewhere (e.g. -DHPB_RASTER_THREADS=768)
    _build.LIB_PATH = os.environ["HPB200_LIB"]
from happypose_b200._capi import Context
from happypose_b200.utils import transform_utils

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
    dev = torch.device("cuda:0")
    ctx = Context.get(dev)
    d = np.load(B.MESH)
    pos = (d["verts"].astype(np.float64) * 0.001).astype(np.float32)
    mid = ops.mesh_upload(ctx, pos, d["faces"], d["normals"], d["uv"], texture=d["texture"])
    grid = transform_utils.load_SO3_grid(576).to(dev)
    pts_all = torch.as_tensor(pos[None]).to(dev)
    pts = torch.as_tensor(pos[np.random.RandomState(0).choice(len(pos), 2000, replace=False)][None]).to(dev)
    img = torch.rand(1, 3, 480, 640, device=dev)
    bs = [int(a) for a in sys.argv[1:]] or [1, 4, 64, 576, 2304]
    peak = B.measured_peak_gbs()[0]
    for b in bs:
        R = grid[torch.arange(b, device=dev) % 576]
        zero = torch.zeros(b, dtype=torch.int32, device=dev)
        K = torch.as_tensor(B.K_BBQ).to(dev).expand(b, 3, 3).contiguous()
        boxes = torch.as_tensor(B.BBOX_BBQ).to(dev).expand(b, 4).contiguous()
        TCO = ops.tco_init(ctx, _capi.TCO_INIT_AUTODEPTH_WITH_R, boxes, K, pts_all, zero, R)
        ids = torch.full((b,), mid, dtype=torch.int32, device=dev)
        x = torch.empty((b, 9, 240, 320), device=dev)
        tCR = TCO[:, :3, 3].contiguous()
        _, K_crop, _, _ = ops.crop(ctx, img, zero, pts, zero, K, TCO, tCR, (240, 320), out=x)
        ms = timeit(lambda: ops.render(ctx, ids, TCO, K_crop, (240, 320), render_normals=True, out=x, out_channel_offset=3))
        gb = b * 6 * 240 * 320 * 4 / 1e9
        cov = float((x[:, 3:6].sum(1) > 0).float().mean())
        print(json.dumps({"kernel": "raster rgb+normals", "b": b, "ms": round(ms, 4), "hyps_per_s": round(b / ms * 1e3), "GBps": round(gb / ms * 1e3, 1), "frac": round(gb / ms * 1e3 / peak, 4), "coverage": round(cov, 3)}))
        crops = x[:, :3]
        ms = timeit(lambda: ops.render_s2d_bf16(ctx, ids, TCO, K_crop, crops, 64))
        gb_moved = b * (123 * 163 * 64 * 2 + 3 * 240 * 320 * 4) / 1e9
        print(json.dumps({"kernel": "raster rgb+normals -> bf16 s2d stem input (fused hand-off)", "b": b, "ms": round(ms, 4), "hyps_per_s": round(b / ms * 1e3),
                          "GBps": round(gb_moved / ms * 1e3, 1), "frac": round(gb_moved / ms * 1e3 / peak, 4), "fp32_equivalent_GBps": round(gb / ms * 1e3, 1)}))
        crops_h, _, _, _ = ops.crop_bf16x4(ctx, img, zero, pts, zero, K, TCO, tCR, (240, 320), tap_bits=16)
        zbuf = torch.empty((b, 64, 123, 163), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last).zero_()
        ms = timeit(lambda: ops.render_s2d_bf16(ctx, ids, TCO, K_crop, crops_h, 64, out=zbuf, pad_prezeroed=True))
        gb_moved = b * (123 * 163 * 128 + 240 * 320 * 8) / 1e9
        print(json.dumps({"kernel": "fused hand-off, bf16x4 crop + aligned 128 B cells (shipped)", "b": b, "ms": round(ms, 4), "hyps_per_s": round(b / ms * 1e3),
                          "moved_GBps": round(gb_moved / ms * 1e3, 1), "GBps_8d": round(gb / ms * 1e3, 1), "frac_8d": round(gb / ms * 1e3 / peak, 4),
                          "checksum": int(zbuf.view(torch.int16).to(torch.int64).sum())}))
        ms = timeit(lambda: ops.crop_bf16x4(ctx, img, zero, pts, zero, K, TCO, tCR, (240, 320), tap_bits=16))
        gbc = b * 3 * 240 * 320 * 4 / 1e9
        print(json.dumps({"kernel": "crop_bf16x4 (boxes+pixels), fp16 taps (shipped)", "b": b, "ms": round(ms, 4), "GBps_8d": round(gbc / ms * 1e3, 1), "frac_8d": round(gbc / ms * 1e3 / peak, 4)}))
        ms = timeit(lambda: ops.crop(ctx, img, zero, pts, zero, K, TCO, tCR, (240, 320), out=x))
        gb = b * 3 * 240 * 320 * 4 / 1e9
        print(json.dumps({"kernel": "crop (boxes+pixels)", "b": b, "ms": round(ms, 4), "hyps_per_s": round(b / ms * 1e3), "GBps": round(gb / ms * 1e3, 1), "frac": round(gb / ms * 1e3 / peak, 4)}))
        ms = timeit(lambda: ops.crop(ctx, img, zero, pts, zero, K, TCO, tCR, (240, 320), out=x, tap_bits=16))
        print(json.dumps({"kernel": "crop (boxes+pixels), fp16 taps", "b": b, "ms": round(ms, 4), "hyps_per_s": round(b / ms * 1e3), "GBps": round(gb / ms * 1e3, 1), "frac": round(gb / ms * 1e3 / peak, 4)}))

def crop_bench():
    """fp16-tap crop at the coarse batch: per-lane gather kernel vs the TMA ring kernel (same process, same box)."""
    dev = torch.device("cuda:0")
    ctx = Context.get(dev)
    d = np.load(B.MESH)
    pos = (d["verts"].astype(np.float64) * 0.001).astype(np.float32)
    grid = transform_utils.load_SO3_grid(576).to(dev)
    pts_all = torch.as_tensor(pos[None]).to(dev)
    pts = torch.as_tensor(pos[np.random.RandomState(0).choice(len(pos), 2000, replace=False)][None]).to(dev)
    img = torch.rand(1, 3, 480, 640, device=dev)
    peak = B.measured_peak_gbs()[0]
    for b in (576, 2304):
        R = grid[torch.arange(b, device=dev) % 576]
        zero = torch.zeros(b, dtype=torch.int32, device=dev)
        K = torch.as_tensor(B.K_BBQ).to(dev).expand(b, 3, 3).contiguous()
        boxes = torch.as_tensor(B.BBOX_BBQ).to(dev).expand(b, 4).contiguous()
        TCO = ops.tco_init(ctx, _capi.TCO_INIT_AUTODEPTH_WITH_R, boxes, K, pts_all, zero, R)
        tCR = TCO[:, :3, 3].contiguous()
        gb = b * 3 * 240 * 320 * 4 / 1e9
        for tma in (0, 1, 0, 1):
            ctx.check(ctx.lib.hpb_set_crop_tma(ctx.handle, tma), "hpb_set_crop_tma")
            ms = timeit(lambda: ops.crop_bf16x4(ctx, img, zero, pts, zero, K, TCO, tCR, (240, 320), tap_bits=16), it=20)
            ms32 = timeit(lambda: ops.crop(ctx, img, zero, pts, zero, K, TCO, tCR, (240, 320), tap_bits=16), it=20)
            print(json.dumps({"kernel": "crop fp16 taps (boxes + pixels)", "b": b, "tma": tma, "bf16x4_ms": round(ms, 4), "planar_f32_ms": round(ms32, 4),
                              "bf16x4_GBps_8d": round(gb / ms * 1e3, 1), "bf16x4_frac_8d": round(gb / ms * 1e3 / peak, 4)}))
    ctx.check(ctx.lib.hpb_set_crop_tma(ctx.handle, 0), "hpb_set_crop_tma")


def maxpool_bench():
    """The stem's max-pool at the coarse batch: plain kernel vs the TMA-staged tile kernel (same process, same box)."""
    dev = torch.device("cuda:0")
    ctx = Context.get(dev)
    peak = B.measured_peak_gbs()[0]
    x = torch.randn(576, 64, 120, 160, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gb = (x.numel() * 2 + x.numel() // 4 * 2) / 1e9
    for tma in (0, 1, 0, 1):
        ctx.check(ctx.lib.hpb_set_maxpool_tma(ctx.handle, tma), "hpb_set_maxpool_tma")
        ms = timeit(lambda: ops.maxpool3x3s2_bf16(ctx, x), it=20)
        print(json.dumps({"kernel": "maxpool3x3s2 bf16 NHWC [576,64,120,160]", "tma": tma, "ms": round(ms, 4), "GBps": round(gb / ms * 1e3, 1),
                          "frac": round(gb / ms * 1e3 / peak, 4)}))
    ctx.check(ctx.lib.hpb_set_maxpool_tma(ctx.handle, 1), "hpb_set_maxpool_tma")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "maxpool":
        maxpool_bench()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "crop":
        crop_bench()
        sys.exit(0)
    main()
