"""Pure-write HBM bandwidth on this box for the crop kernel's output footprint (what a store-only kernel achieves)."""
import torch
dev = torch.device("cuda:0")
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for b in (576, 2304):
    x = torch.empty((b, 9, 240, 320), device=dev)
    flat = torch.empty(b * 3 * 240 * 320, device=dev)
    src = torch.rand(b * 3 * 240 * 320, device=dev)
    gb = b * 3 * 240 * 320 * 4 / 1e9
    for name, fn, bytes_ in (("fill contiguous", lambda: flat.zero_(), gb), ("fill x[:, :3] (strided chunks)", lambda: x[:, :3].zero_(), gb),
                             ("fill all 9 channels", lambda: x.zero_(), 3 * gb), ("copy contiguous (read+write)", lambda: flat.copy_(src), 2 * gb)):
        ms = t(fn)
        print(f"b={b:5d} {name:34s} {ms:.4f} ms  {bytes_ / ms * 1e3:7.1f} GB/s")
# driver memset over the same strided footprint (pitch = one hypothesis of x, width = its 3 crop channels)
import ctypes
rt = None
for name in ("libcudart.so", "libcudart.so.12"):
    try:
        rt = ctypes.CDLL(name); break
    except OSError:
        pass
if rt is None:
    import glob, os
    c = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
    rt = ctypes.CDLL(c[0])
rt.cudaMemset2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]
for b in (576, 2304):
    x = torch.empty((b, 9, 240, 320), device=dev)
    gb = b * 3 * 240 * 320 * 4 / 1e9
    st = torch.cuda.current_stream().cuda_stream
    ms = t(lambda: rt.cudaMemset2DAsync(x.data_ptr(), 9 * 76800 * 4, 0, 3 * 76800 * 4, b, st))
    print(f"b={b:5d} cudaMemset2D width 3/9 of the pitch    {ms:.4f} ms  {gb / ms * 1e3:7.1f} GB/s")
    ms = t(lambda: rt.cudaMemset2DAsync(x.data_ptr() + 3 * 76800 * 4, 9 * 76800 * 4, 0, 6 * 76800 * 4, b, st))
    print(f"b={b:5d} cudaMemset2D width 6/9 of the pitch    {ms:.4f} ms  {2 * gb / ms * 1e3:7.1f} GB/s")
    y = torch.empty((b, 3, 240, 320), device=dev)
    ms = t(lambda: y.zero_())
    print(f"b={b:5d} fill separate [b,3,h,w] tensor         {ms:.4f} ms  {gb / ms * 1e3:7.1f} GB/s")
