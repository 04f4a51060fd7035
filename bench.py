#!/usr/bin/env python
"""bench.py -- MegaPose poses/sec on BASELINE config #1 (1 object per GPU, 640x480 synthetic frame, 576 coarse hypotheses
+ top-1 + 5 refiner iterations x 4 views + 1 scoring pass, random-init ResNet-34 coarse/refiner in bf16).

    python bench.py --gpus N --steps K --warmup W                 this repo (CUDA kernels behind libhpb200.so)
    python bench.py --impl reference --gpus N --steps K --warmup W  the reference's CPU path (oracle port) on host cores

One "step" = one PoseEstimator.run_inference_pipeline call = D poses per GPU (D = --dets, default 1).  With N > 1 (one
process per GPU, torchrun) the N*D detections of the frame form one hypothesis table whose rows are sharded across the
ranks; the only collective is the all-gather of the coarse / scoring logits (happypose_b200/distributed.py).
Per-GPU work is fixed as N grows ("weak").  Rank 0 prints ONE JSON line.

Reported (see DESIGN.md "Measurement"):
  value      poses/s, frame and detections already resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public API from pinned HOST buffers (H2D of frame/K/boxes and D2H of the final poses
             inside the timed region)
  roofline   hpb_raster_kernel (dominant kernel of this library): algorithmic bytes / CUDA-event duration measured
             live around every launch of the timed region, against MEASURED_PEAKS.json hbm_gbs
  kernels    the same live figures for every bracketed kernel (rasteriser, crop)
  hyps       rendered hyps/s (render + crop of 576 hypotheses into the 9-channel network input, no network)
  cpu_baseline  the CPU oracle pipeline (C rasteriser port + torch-CPU ResNet + torchvision roi_align) on a bounded
             sample, extrapolated to one pose, on all host cores (N=1, rank 0 only)
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

K_BBQ = np.array([[605.95, 0, 319.03], [0, 605.01, 249.68], [0, 0, 1]], np.float32)  # docs/book/megapose/inference.md:33
BBOX_BBQ = np.array([384, 234, 522, 455], np.float32)                               # docs/book/megapose/inference.md:37
MESH = os.path.join(ROOT, "tests", "golden", "obj_000001.npz")
LABEL = "obj_000001"
M_GRID, N_REFINER_ITERS, N_VIEWS = 576, 5, 4
H_IM, W_IM, H_R, W_R = 480, 640, 240, 320
METRIC, UNIT = "megapose_poses_per_sec", "poses/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dets", type=int, default=1, help="detections (poses) per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="do not poll nvidia-smi during the timed regions (debugging)")
    ap.add_argument("--clock-sampler", default="smi", choices=["nvml", "smi"],
                    help="nvml: in-process NVML polling thread; smi: the recipe's `nvidia-smi -lms` child started before the warm-up")
    ap.add_argument("--profile-range", action="store_true",
                    help="bracket the kernel-timed region with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    ap.add_argument("--config", default="megapose1", choices=["megapose1", "cosypose21", "sweep", "tless240", "gso1000"],
                    help="megapose1 = BASELINE configs[0] (the headline metric); the others are configs[1..4] (bench_configs.py)")
    ap.add_argument("--meshes", type=int, default=0, help="gso1000: number of meshes (default 1000)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="megapose1 at N > 1: weak = --dets detections PER GPU (default); strong = --dets detections in total, their "
                         "576 rows each split across the ranks (the latency case: 72 rows per GPU at N = 8)")
    args = ap.parse_args()
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = 10
    return args


def config_dict(args, world):
    return {
        "workload": "BASELINE configs[0]: MegaPose barbecue-sauce-style example, 1 object per GPU, 640x480 synthetic RGB frame, "
                    "576 coarse hypotheses + top-1 + 5 refiner iterations x 4 views + 1 scoring pass, random-init ResNet-34 (bf16)",
        "mesh": "tests/golden/obj_000001.npz (reference tests/data/obj_000001.ply: 9951 verts, 15728 faces, 512^2 texture)",
        "detections_per_gpu": args.dets if getattr(args, "scaling", "weak") == "weak" else args.dets / world, "coarse_hypotheses": M_GRID, "refiner_iterations": N_REFINER_ITERS, "refiner_views": N_VIEWS,
        "render_size": [H_R, W_R], "frame": [H_IM, W_IM], "bsz_images": 576, "bsz_objects": 16,
        "parallelism": f"hypothesis-sharded x{world}" + (" (strong: the detections' rows are split across the ranks)" if getattr(args, "scaling", "weak") == "strong" else ""),
        "l2": "no explicit flush: every step writes/reads 1.6 GB of network input per pose (> 126 MB L2)",
        "batching": "bsz_images=576, bsz_objects=16 (both reference defaults of run_inference_pipeline callers; SURVEY C1 quotes 128 / 8)",
        "frames": "`value`: the pipeline's result frames stay deferred (no pandas DataFrame is built inside the timed region); "
                  "`e2e`: the final frame is materialised every step (pose_score to numpy + poses to host); the 576-row coarse "
                  "frame the reference always builds is deferred in both",
        "host": "gc.collect() + gc.freeze() after the warm-up (a serving process's setting; profiles/r2_e2e_gc_pause.txt)",
    }


def detections_arrays(n_det):
    """Seeded synthetic detections of frame 0: the barbecue-sauce bbox, jittered for rows > 0."""
    rs = np.random.RandomState(1)
    boxes = np.tile(BBOX_BBQ, (n_det, 1)) + rs.uniform(-25, 25, (n_det, 4)).astype(np.float32)
    boxes[0] = BBOX_BBQ
    return boxes.astype(np.float32)


# ----------------------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md): sampled during the timed region
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle-reason samples taken DURING the timed regions.

    mode "smi" (default): the profiling recipe's `nvidia-smi --query-gpu=... -lms` child process, started BEFORE the
    warm-up (its start-up stalls kernel launches for milliseconds, so it must not begin inside a timed region) and left
    running; its time-stamped lines are filtered to the timed regions afterwards.  The sampling happens in another
    process, so it does not touch this process's host thread.
    mode "nvml": an in-process NVML polling thread (nvidia_ml_py).  Round 1 used it; at N >= 2 it cost the end-to-end
    region 3.4 ms per step (15.6 vs 12.2 ms/step, profiles/r2_e2e_n2_clock_sampler.txt): rank 0's host thread competes
    with the poller while every step ends in a host synchronisation and every collective waits for the slowest rank.
    That, not the pipeline, was the N >= 2 e2e efficiency of 0.77 in SCALE_r01.json.  Kept for debugging only."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_s=0.02, mode="nvml"):
        self.samples = []  # (sm MHz, power W, reasons bitmask)
        self.sm_max = None
        self.proc = None
        self.thread = None
        self._stop = threading.Event()
        self._armed = threading.Event()
        self.windows = []  # [t0, t1] wall-clock intervals of the timed regions (smi mode keeps samples inside them)
        if mode == "smi":
            self._start_smi(index)
            return
        try:
            import pynvml as N

            N.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            h = N.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            self.N, self.h = N, h
            N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)  # fail here, not in the thread
            self.period = period_s
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            self._start_smi(index)

    def _loop(self):
        N, h = self.N, self.h
        while not self._stop.is_set():
            if self._armed.is_set():
                try:
                    sm = float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
                    try:
                        pw = N.nvmlDeviceGetPowerUsage(h) / 1e3
                    except Exception:
                        pw = 0.0
                    try:
                        rs = int(N.nvmlDeviceGetCurrentClocksEventReasons(h))
                    except Exception:
                        rs = int(N.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    self.samples.append((sm, pw, rs))
                except Exception:
                    pass
            self._stop.wait(self.period)

    def _start_smi(self, index):
        self.path = f"/tmp/hpb_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(index)],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def arm(self, on=True):
        """Samples are kept only while armed (= while a timed region is running)."""
        (self._armed.set if on else self._armed.clear)()
        if on:
            self.windows.append([time.time(), None])
        elif self.windows and self.windows[-1][1] is None:
            self.windows[-1][1] = time.time()

    def stop(self):
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no samples"]}
            N = self.N
            bits = 0
            for _, _, r in self.samples:
                bits |= r
            names = [("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", 0x8), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", 0x20), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", 0x4)]
            reasons = [n for n, attr, dflt in names if bits & int(getattr(N, attr, dflt))]
            return {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_min_mhz": min(s[0] for s in self.samples), "sm_max_mhz": self.sm_max,
                    "power_w_max": max(s[1] for s in self.samples), "samples": len(self.samples), "reasons": sorted(reasons), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        import datetime

        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if self.windows and not any(w0 - 0.05 <= ts <= (w1 or ts) + 0.05 for w0, w1 in self.windows):
                    continue  # outside the timed regions
                sm.append(float(p[1])); mx.append(float(p[2])); power.append(float(p[3]))
            except ValueError:
                continue
            for nme, v in zip(names, p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power), "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi"}


def measured_peak_tflops():
    """Dense bf16 tensor throughput for the tensor-core roofline: MEASURED_PEAKS.json (sustained when present), else the
    profiling recipe's nominal 2250 TFLOP/s."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        for k in ("bf16_tflops_sustained", "bf16_tflops"):
            if isinstance(d.get(k), (int, float)) and d[k] > 0:
                return float(d[k]), f"measured (MEASURED_PEAKS.json {k})"
    except (OSError, ValueError):
        pass
    return 2250.0, "fallback (B200_PROFILING.md nominal dense bf16)"


def measured_peak_gbs():
    """HBM peak for the roofline: the driver-written MEASURED_PEAKS.json (the sustained figure when the file distinguishes
    burst / sustained -- the kernels are timed inside a long step), else the profiling recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        flat = {}

        def walk(prefix, obj):
            if isinstance(obj, dict):
                for k, v in obj.items():
                    walk(f"{prefix}.{k}" if prefix else str(k), v)
            elif isinstance(obj, (int, float)):
                flat[prefix.lower()] = float(obj)

        walk("", d)
        cands = {k: v for k, v in flat.items() if "hbm" in k and any(t in k for t in ("gb", "tb", "bw", "bandwidth"))}
        for pick in (lambda k: "sustain" in k, lambda k: k == "hbm_gbs", lambda k: "burst" not in k, lambda k: True):
            for k, v in cands.items():
                if pick(k) and v > 0:
                    if v < 100:  # TB/s
                        v *= 1e3
                    return v, f"measured (MEASURED_PEAKS.json {k})"
        raise KeyError("no hbm bandwidth key")
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------------------------
# CPU side: the oracle pipeline = the reference's algorithm on host cores
# ----------------------------------------------------------------------------------------------------------------
def _torchvision_roi_align():
    """The reference crops with torchvision.ops.roi_align (toolbox/lib3d/cropping.py:167); use the real (C++, CPU) op in
    the baseline when it is importable so the baseline is not slowed down by the numpy restatement."""
    try:
        from torchvision.ops import roi_align as tv_roi_align
    except Exception:
        return None

    def roi_align(images, rois, output_size, sampling_ratio=4):
        out = tv_roi_align(torch.as_tensor(np.asarray(images, np.float32)), torch.as_tensor(np.asarray(rois, np.float32)),
                           output_size=tuple(output_size), spatial_scale=1.0, sampling_ratio=int(sampling_ratio))
        return out.numpy()

    return roi_align


class CpuPipeline:
    """Bounded sample of the pose workload on the CPU: n_coarse coarse hypotheses (crop + raster + ResNet-34 9ch) and one
    refiner iteration of one hypothesis (crop + 4 rasters + ResNet-34 27ch); pose time = 577/n_coarse * t_c + 5 * t_r."""

    def __init__(self, n_coarse=8, cores=None):
        from oracle import np_oracle as O
        from oracle import pipeline_oracle as P
        from happypose_b200.megapose.backbones import make_backbone
        from types import SimpleNamespace

        self.O, self.P = O, P
        tv = _torchvision_roi_align()
        self.roi_kind = "torchvision.ops.roi_align (CPU)" if tv is not None else "numpy restatement"
        if tv is not None:
            O.roi_align = tv
        self.cores = int(cores) if cores else len(os.sched_getaffinity(0))
        torch.set_num_threads(self.cores)
        d = np.load(MESH)
        self.scene = P.make_scene([{k: d[k] for k in d.files}], [0.001])
        torch.manual_seed(0)

        def cpu_model(n_in, n_views, head, out_dim):
            bb = make_backbone("vanilla_resnet34", n_in).float().eval()
            return SimpleNamespace(backbone=bb, heads={head: torch.nn.Linear(bb.n_features, out_dim).eval()}, render_size=(H_R, W_R),
                                   input_depth=False, render_normals=True, render_depth=False, n_rendered_views=n_views,
                                   multiview_type="TCO+front_3views", remove_TCO_rendering=False, depth_normalization_type="none", pose_dim=9)

        self.coarse = cpu_model(9, 1, "renderings_logits", 1)
        self.refiner = cpu_model(27, 4, "pose", 9)
        with torch.no_grad():
            self.refiner.heads["pose"].weight.mul_(1e-2)
            self.refiner.heads["pose"].bias.copy_(torch.tensor([1.0, 0, 0, 0, 1, 0, 0, 0, 1]))
        self.n_coarse = n_coarse
        rs = np.random.RandomState(0)
        self.image = rs.rand(1, 3, H_IM, W_IM).astype(np.float32)
        quats = np.load(os.path.join(ROOT, "happypose_b200", "data", "so3_grid_576.npy"))
        R = O.unitquat_to_rotmat(quats[:: M_GRID // n_coarse][:n_coarse]).astype(np.float32)
        self.K_rows = np.tile(K_BBQ, (n_coarse, 1, 1))
        boxes = np.tile(BBOX_BBQ, (n_coarse, 1))
        self.zeros = np.zeros(n_coarse, int)
        self.TCO0 = O.TCO_init_from_boxes_autodepth_with_R(boxes, self.scene.points[self.zeros], self.K_rows, R)

    def sample(self):
        """-> seconds per pose extrapolated from one bounded sample."""
        P, n = self.P, self.n_coarse
        t0 = time.perf_counter()
        P.forward_coarse(self.coarse, self.scene, self.image, self.K_rows, self.zeros, self.zeros, self.TCO0, n_threads=self.cores)
        t1 = time.perf_counter()
        P.forward_refiner(self.refiner, self.scene, self.image, self.K_rows[:1], self.zeros[:1], self.zeros[:1], self.TCO0[:1], 1, n_threads=self.cores)
        t2 = time.perf_counter()
        return (t1 - t0) / n * (M_GRID + 1) + (t2 - t1) * N_REFINER_ITERS

    def single_thread(self):
        """SURVEY 8(d) also asks for the reference's own setting, OMP_NUM_THREADS = 1 (megapose/__init__.py:20-39): the same
        sample on ONE thread (a smaller one: 2 coarse hypotheses + 1 refiner iteration), all-cores setting restored after."""
        keep = self.cores, self.n_coarse
        try:
            self.cores, self.n_coarse = 1, 2
            torch.set_num_threads(1)
            saved = self.K_rows, self.zeros, self.TCO0
            self.K_rows, self.zeros, self.TCO0 = self.K_rows[:2], self.zeros[:2], self.TCO0[:2]
            self.sample()
            t = self.sample()
            self.K_rows, self.zeros, self.TCO0 = saved
        finally:
            self.cores, self.n_coarse = keep
            torch.set_num_threads(self.cores)
        return {"value": 1.0 / t, "unit": UNIT, "cores": 1,
                "sample": "2 of 577 coarse/scoring hypotheses + 1 of 5 refiner iterations, one thread (the reference package's own OMP_NUM_THREADS=1)"}

    def describe(self):
        return (f"{self.n_coarse} of 577 coarse/scoring hypotheses + 1 of 5 refiner iterations (1 hypothesis x 4 views) per sample, "
                f"extrapolated linearly to one pose; C rasteriser port (oracle/raster_oracle.c) + {self.roi_kind} + torch-CPU fp32 ResNet-34")


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    pipe = CpuPipeline(n_coarse=8)
    for _ in range(max(args.warmup, 1)):
        pipe.sample()
    times = [pipe.sample() for _ in range(args.steps)]
    t_pose = sum(times) / len(times)
    value = args.dets / (t_pose * args.dets)  # a CPU box processes the detections one after another
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_pose * args.dets * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": pipe.cores, "kind": "port", "sample": pipe.describe(),
                         "single_thread": pipe.single_thread()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "Panda3D/OpenGL cannot be installed offline, so the reference arm is the CPU oracle port of the same path "
                "(it omits the reference's worker-process IPC and GL read-back, i.e. it is a faster baseline than the real one)",
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------------------
# this repo
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import pandas as pd
    from happypose_b200 import distributed as hdist, ops
    from happypose_b200.datasets.object_dataset import RigidObject, RigidObjectDataset
    from happypose_b200.inference.types import ObservationTensor
    from happypose_b200.megapose.pose_estimator import PoseEstimator
    from happypose_b200.megapose.pose_models_cfg import make_pose_models
    from happypose_b200.utils.tensor_collection import PandasTensorCollection

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    rank, local_rank, world = hdist.init_distributed_mode()
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    dev = torch.device("cuda", local_rank)
    torch.backends.cudnn.benchmark = True

    ds = RigidObjectDataset([RigidObject(label=LABEL, mesh_path=MESH, mesh_units="mm")])
    coarse, refiner, mesh_db = make_pose_models(ds, device=dev, seed=0)
    with torch.no_grad():  # identity-biased pose head: random weights must not throw the object out of the frame
        refiner.pose_fc.weight.mul_(1e-2)
        refiner.pose_fc.bias.copy_(torch.tensor([1.0, 0, 0, 0, 1, 0, 0, 0, 1]))
    est = PoseEstimator(refiner_model=refiner, coarse_model=coarse, bsz_objects=16, bsz_images=576, SO3_grid_size=M_GRID,
                        shard_across_ranks=True)  # every rank passes the same frame + detections; rows are split
    est.use_cuda_graphs = not args.no_graphs
    ctx = coarse._ctx()

    n_det = args.dets * (world if args.scaling == "weak" else 1)
    boxes_np = detections_arrays(n_det)
    rs = np.random.RandomState(0)
    image_host = torch.as_tensor(rs.rand(1, 3, H_IM, W_IM).astype(np.float32)).pin_memory()
    K_host = torch.as_tensor(K_BBQ[None]).pin_memory()
    boxes_host = torch.as_tensor(boxes_np).pin_memory()

    def make_detections(boxes_dev):
        infos = pd.DataFrame({"label": [LABEL] * n_det, "batch_im_id": [0] * n_det, "score": [1.0] * n_det})
        return PandasTensorCollection(infos=infos, bboxes=boxes_dev)

    obs_dev = ObservationTensor(image_host.to(dev), K_host.to(dev))
    boxes_dev = boxes_host.to(dev)

    def step_resident():
        final, _ = est.run_inference_pipeline(obs_dev, detections=make_detections(boxes_dev), n_refiner_iterations=N_REFINER_ITERS, n_pose_hypotheses=1)
        return final

    h2d = image_host.numel() * 4 + K_host.numel() * 4 + boxes_host.numel() * 4

    def step_e2e():
        obs = ObservationTensor(image_host.to(dev, non_blocking=True), K_host.to(dev, non_blocking=True))
        det = make_detections(boxes_host.to(dev, non_blocking=True))
        final, _ = est.run_inference_pipeline(obs, detections=det, n_refiner_iterations=N_REFINER_ITERS, n_pose_hypotheses=1)
        scores = final.infos["pose_score"].to_numpy()  # materialises the deferred result frames (D2H of scores / survivor ids)
        poses = final.poses.cpu()
        return poses, scores

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        walls = []
        for _ in range(steps):
            t0 = time.perf_counter()
            fn()
            walls.append(time.perf_counter() - t0)
        e1.record()
        sync_all()
        if os.environ.get("HPB_BENCH_STEP_TIMES"):  # debugging aid: host wall time of every step of this region
            print(f"[rank {rank}] {getattr(fn, '__name__', 'step')} wall ms/step: " + " ".join(f"{1e3 * w:.2f}" for w in walls), file=sys.stderr)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # the clock sampler starts BEFORE the warm-up (its start-up cost stays outside every timed region) and keeps samples
    # only while a timed region is running
    sampler = ClockSampler(local_rank, mode=args.clock_sampler) if (rank == 0 and not args.no_clocks) else None
    for _ in range(max(args.warmup, 3)):
        step_resident()
    # Serving-process hygiene, after the warm-up: collect once, then move everything alive (modules, meshes, pandas / torch
    # internals: millions of objects) to the permanent generation.  Without it a full (generation-2) collection -- 77 ms in
    # this process, measured with HPB_BENCH_STEP_TIMES=1 -- lands every few dozen steps, and whether one falls into a 20-step
    # timed region is luck: it put one 88 ms step into the e2e region of every N >= 2 run of one build and none into another's.
    # The collector stays enabled; later full collections only scan what the steps themselves allocate.
    gc.collect()
    gc.freeze()
    if os.environ.get("HPB_BENCH_STEP_TIMES"):
        _gc_t = {}

        def _gc_cb(phase, info):
            if phase == "start":
                _gc_t["t"] = time.perf_counter()
            elif info.get("generation", 0) >= 1:
                print(f"[rank {rank}] gc generation {info['generation']}: {1e3 * (time.perf_counter() - _gc_t.get('t', 0)):.2f} ms", file=sys.stderr)

        gc.callbacks.append(_gc_cb)

    # ---- timed region 1: resident inputs (value); the public API as a user runs it (CUDA graphs on), clocks sampled -----
    l0 = ctx.launch_count()
    if sampler is not None:
        sampler.arm(True)
    ms_total = timed(step_resident, args.steps)
    if sampler is not None:
        sampler.arm(False)
    launches = ctx.launch_count() - l0
    # ---- timed region 1b: the same K steps with every rasteriser / crop launch bracketed by CUDA events on the launching
    # stream (live roofline).  Event brackets cannot be recorded into a CUDA graph, so graph replay is off in this pass:
    # `launches` counts the kernels enqueued eagerly here when region 1 replayed graphs (graph replays bypass the C ABI's
    # launch counter).
    timer = ops.KernelTimer()
    ops.set_kernel_timer(timer)
    l0 = ctx.launch_count()
    if args.profile_range:
        torch.cuda.profiler.start()
    ms_bracketed = timed(step_resident, args.steps)
    if args.profile_range:
        torch.cuda.profiler.stop()
    launches = max(launches, ctx.launch_count() - l0)
    ops.set_kernel_timer(None)
    ksum = timer.summary()
    fsum = timer.flop_summary()
    # ---- timed region 2: end to end from pinned host buffers -----------------------------------------------------
    for _ in range(max(args.warmup, 3)):  # the e2e variant gets its own warm-up right before its timed region
        step_e2e()
    if sampler is not None:
        sampler.arm(True)
    ms_e2e = timed(step_e2e, args.steps)
    if sampler is not None:
        sampler.arm(False)
    clocks = sampler.stop() if sampler is not None else None
    poses, scores = step_e2e()
    assert poses.shape == (n_det, 4, 4) and torch.isfinite(poses).all(), "pipeline produced non-finite poses"

    # ---- render+crop throughput (BASELINE's second metric), no network -------------------------------------------
    b = M_GRID
    grid = est._SO3_grid.to(dev)
    from happypose_b200 import _capi
    obj0 = torch.zeros(b, dtype=torch.int32, device=dev)
    mesh_ids = coarse.renderer.mesh_ids([LABEL] * b)
    K_rows = obs_dev.K.expand(b, 3, 3).contiguous()
    TCO = ops.tco_init(ctx, _capi.TCO_INIT_AUTODEPTH_WITH_R, boxes_dev[:1].expand(b, 4).contiguous(), K_rows, mesh_db.points, obj0, grid)
    x = torch.empty((b, 9, H_R, W_R), device=dev)
    pts = mesh_db.points_subset(2000)

    def render_crop():
        _, K_crop, _, _ = ops.crop(ctx, obs_dev.images, obj0, pts, obj0, K_rows, TCO, TCO[:, :3, 3].contiguous(), (H_R, W_R), out=x)
        ops.render(ctx, mesh_ids, TCO, K_crop, (H_R, W_R), render_normals=True, out=x, out_channel_offset=3)

    z_buf = torch.empty((b, 64, H_R // 2 + 3, W_R // 2 + 3), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last).zero_()

    def render_crop_fused():  # what the pipeline's coarse stage does: bf16 crop pixels, then the rasteriser writes the stem input
        crops_h, K_crop, _, _ = ops.crop_bf16x4(ctx, obs_dev.images, obj0, pts, obj0, K_rows, TCO, TCO[:, :3, 3].contiguous(), (H_R, W_R),
                                                tap_bits=coarse.crop_tap_bits)
        ops.render_s2d_bf16(ctx, mesh_ids, TCO, K_crop, crops_h, 64, out=z_buf, pad_prezeroed=True)

    for _ in range(3):
        render_crop()
        render_crop_fused()
    hyp_iters = 20
    ms_hyp_planar = timed(render_crop, hyp_iters)
    ms_hyp = timed(render_crop_fused, hyp_iters)
    hyps_per_s = world * b * hyp_iters / (ms_hyp / 1e3)

    if rank != 0:
        return
    poses_per_step = n_det
    value = poses_per_step * args.steps / (ms_total / 1e3)
    e2e_value = poses_per_step * args.steps / (ms_e2e / 1e3)
    peak, peak_src = measured_peak_gbs()
    rk = ksum.get("hpb_raster_kernel", {"gbps": 0.0, "ms_avg": 0.0, "launches": 0, "bytes_avg": 0, "ms_total": 0.0})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "bf16 (networks) / f32 (rasteriser, crop, pose kernels)", "data": "synthetic", "config": config_dict(args, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(poses.numel() * 4 + scores.size * 8),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        # roofline of the dominant kernel of this library, on SURVEY 8(d)'s ALGORITHMIC bytes: 1 843 200 B per rendered view
        # (6 float32 planes of 240x320: the reference's layout of the result), whatever layout the launch actually wrote.
        # The bytes the launches really move (bf16 stem-input cells + the crop read for the fused 576-scene launch) are the
        # secondary `moved_*` keys.
        "roofline": {"bound": "hbm", "kernel": "hpb_raster_kernel", "achieved": rk.get("fp32_equivalent_gbps", 0.0), "peak": peak, "unit": "GB/s",
                     "frac": rk.get("fp32_equivalent_gbps", 0.0) / peak, "peak_source": peak_src, "traffic": None,
                     "algorithmic_bytes_per_view": 6 * H_R * W_R * 4,
                     "launches_timed": rk["launches"], "avg_launch_ms": rk["ms_avg"],
                     "moved_bytes_per_launch": rk["bytes_avg"], "moved_achieved": rk["gbps"], "moved_frac": rk["gbps"] / peak,
                     "share_of_step": rk["ms_total"] / ms_bracketed if ms_bracketed > 0 else None,
                     "bracketed_ms_per_step": ms_bracketed / args.steps,
                     # the same figures for the largest launch class alone (the 576-scene coarse launch; the average above
                     # also contains the 4..16-scene refiner / scoring launches, which cannot fill 148 SMs)
                     "largest_launch": ({"launches": rk["largest"]["launches"], "avg_launch_ms": rk["largest"]["ms_avg"],
                                         "achieved": rk["largest"]["fp32_equivalent_gbps"], "frac": rk["largest"]["fp32_equivalent_gbps"] / peak,
                                         "moved_bytes": rk["largest"]["bytes"], "moved_achieved": rk["largest"]["gbps"],
                                         "moved_frac": rk["largest"]["gbps"] / peak} if "largest" in rk else None)},
        "kernels": {k: {"gbps": v["gbps"], "frac": v["gbps"] / peak, "avg_launch_ms": v["ms_avg"], "launches": v["launches"],
                        "share_of_step": v["ms_total"] / ms_bracketed,
                        "largest_launch_gbps": v["largest"]["gbps"], "largest_launch_frac": v["largest"]["gbps"] / peak,
                        "largest_launch_ms": v["largest"]["ms_avg"]} for k, v in ksum.items()},
        "hyps": {"metric": "rendered_hyps_per_sec", "value": hyps_per_s, "unit": "hyps/s", "b": b, "ms_per_launch_pair": ms_hyp / hyp_iters,
                 "path": "hpb_crop_bf16x4 (fp16 frame taps) -> hpb_render_s2d_bf16 (the pipeline's coarse hand-off: bf16 space-to-depth stem input, 128 B per cell)",
                 # 2 764 800 algorithmic bytes per coarse hypothesis (the 9-channel float32 network input)
                 "gbps": b * 9 * H_R * W_R * 4 / 1e9 / (ms_hyp / hyp_iters / 1e3),
                 "frac": b * 9 * H_R * W_R * 4 / 1e9 / (ms_hyp / hyp_iters / 1e3) / peak,
                 # bytes really moved per hypothesis: bf16 crop pixels written + read back, stem-input cells written
                 "moved_gbps": b * (2 * H_R * W_R * 8 + (H_R // 2 + 3) * (W_R // 2 + 3) * 128) / 1e9 / (ms_hyp / hyp_iters / 1e3),
                 "planar_fp32": {"value": world * b * hyp_iters / (ms_hyp_planar / 1e3), "ms_per_launch_pair": ms_hyp_planar / hyp_iters,
                                 "gbps": b * 9 * H_R * W_R * 4 / 1e9 / (ms_hyp_planar / hyp_iters / 1e3),
                                 "frac": b * 9 * H_R * W_R * 4 / 1e9 / (ms_hyp_planar / hyp_iters / 1e3) / peak}},
    }
    if "hpb_stem_tc_kernel" in fsum:
        # the one tensor-core kernel of the library (tcgen05 implicit GEMM of the stem): issued MMA flops against the measured
        # dense bf16 throughput (the sustained figure: the kernel runs inside a long step)
        tf_peak, tf_src = measured_peak_tflops()
        st = fsum["hpb_stem_tc_kernel"]
        line["tensor_roofline"] = {"bound": "tensor", "kernel": "hpb_stem_tc_kernel", "achieved": st["largest"]["tflops"], "peak": tf_peak,
                                   "unit": "TFLOP/s", "frac": st["largest"]["tflops"] / tf_peak, "peak_source": tf_src,
                                   "launch": "the 576-row coarse batch (largest launch class)", "avg_launch_ms": st["largest"]["ms_avg"],
                                   "launches_timed": st["launches"], "share_of_step": st["ms_total"] / ms_bracketed if ms_bracketed > 0 else None,
                                   "note": "issued flops = tiles x non-zero weight slices x M128 N64 K16; the 64-channel space-to-depth "
                                           "cell carries 9 real channels per 16-channel slice, so the reference conv1's own flops are 9/16 of these"}
    traffic_file = os.path.join(ROOT, "profiles", "raster_traffic.json")
    if os.path.exists(traffic_file):
        try:
            tr = json.load(open(traffic_file))
            # dram bytes of ONE ncu --set full capture of the 576-scene launch, scaled to this run's average launch by
            # the measured traffic / algorithmic ratio (1.07: mesh + texture reads on top of the image writes)
            line["roofline"]["traffic"] = tr["traffic_over_algorithmic"] * rk["bytes_avg"]  # ratio measured against MOVED bytes
            line["roofline"]["traffic_capture"] = {"launch": tr.get("launch", "576 scenes"), "dram_bytes": tr["traffic_bytes_per_launch"],
                                                   "algorithmic_bytes": tr["algorithmic_bytes_of_that_launch"]}
        except Exception:
            pass
    if world == 1 and not args.no_cpu_baseline:
        pipe = CpuPipeline(n_coarse=8)
        pipe.sample()
        t_pose = min(pipe.sample() for _ in range(2))
        line["cpu_baseline"] = {"value": 1.0 / t_pose, "unit": UNIT, "cores": pipe.cores, "kind": "port", "sample": pipe.describe(),
                                "single_thread": pipe.single_thread()}
    emit(line)


_JSON_FD = None


def _reserve_stdout():
    """stdout must carry exactly ONE JSON line: anything a library prints to fd 1 (NCCL's version banner, cuDNN notices)
    is sent to stderr instead, and emit() writes the line to the original stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    args = parse_args()
    _reserve_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.config != "megapose1":
            import bench_configs

            if not torch.cuda.is_available():
                raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
            bench_configs.CONFIGS[args.config](args, sys.modules[__name__])
        else:
            run_ours(args)
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
