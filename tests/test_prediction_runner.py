"""Evaluation-harness caller (happypose_b200/evaluation/prediction_runner.py; reference
megapose/evaluation/prediction_runner.py:44-291): per-frame loop, result keys, timing columns, scene sharding and the
file-free gather across ranks (CPU, gloo, world size 2).  A GPU test with the real PoseEstimator is in test_gpu_pipeline.py."""
import os
import socket

import numpy as np
import pandas as pd
import torch
import torch.multiprocessing as mp

from happypose_b200.inference.types import InferenceConfig
from happypose_b200.utils.tensor_collection import PandasTensorCollection


class FakeEstimator:
    """Returns what PoseEstimator.run_inference_pipeline returns (pose_estimator.py:650-666), derived from the detections."""

    def __init__(self):
        self.calls = []

    def run_inference_pipeline(self, obs, detections=None, run_detector=None, coarse_estimates=None, n_refiner_iterations=5,
                               n_pose_hypotheses=1, run_depth_refiner=False, bsz_images=None, bsz_objects=None):
        self.calls.append((tuple(obs.images.shape), len(detections), n_refiner_iterations, n_pose_hypotheses, bsz_images, bsz_objects))
        n = len(detections)
        infos = detections.infos.copy()
        infos["pose_score"] = np.linspace(0.9, 0.5, n)
        poses = torch.eye(4).repeat(n, 1, 1) * float(obs.K[0, 0, 0])
        final = PandasTensorCollection(infos, poses=poses)
        coarse = PandasTensorCollection(pd.concat([infos] * 3).reset_index(drop=True), poses=poses.repeat(3, 1, 1))
        extra = {"coarse": {"preds": coarse, "data": {"time": 0.25, "TCO": torch.zeros(1)}},
                 "refiner": {"preds": final, "data": {"time": 0.5}}, "scoring": {"preds": final, "data": {"time": 0.125}},
                 "refiner_all_hypotheses": {}, "coarse_filter": {}, "timing_str": "", "time": 1.0}
        return final, extra


def make_frames(n_frames):
    frames = []
    for i in range(n_frames):
        n_det = 1 + i % 3
        det = PandasTensorCollection(pd.DataFrame({"label": [f"obj_{k}" for k in range(n_det)], "batch_im_id": [0] * n_det,
                                                   "instance_id": list(range(n_det)), "visib_fract": [0.5] * n_det}),
                                     bboxes=torch.rand(n_det, 4))
        frames.append({"rgb": np.zeros((48, 64, 3), np.uint8), "K": np.eye(3, dtype=np.float32) * (i + 1), "detections": det,
                       "im_info": {"scene_id": 7, "view_id": i}})
    return frames


def test_runner_keys_columns_and_timing():
    from happypose_b200.evaluation.prediction_runner import PredictionRunner, compute_pose_est_total_time

    cfg = InferenceConfig(detection_type="gt", n_refiner_iterations=3, n_pose_hypotheses=2, bsz_images=64, bsz_objects=4)
    est = FakeEstimator()
    frames = make_frames(4)
    preds = PredictionRunner(frames, cfg).get_predictions(est)
    assert sorted(preds.keys()) == ["coarse", "final", "refiner/final", "refiner/iteration=3"]
    n_rows = sum(len(f["detections"]) for f in frames)
    assert len(preds["final"]) == n_rows and len(preds["coarse"]) == 3 * n_rows
    assert len(est.calls) == 5 and est.calls[0] == est.calls[1]       # the first frame is run twice (warm-up, :233-240)
    assert est.calls[0][2:] == (3, 2, 64, 4)                           # the inference config reaches the estimator
    f = preds["final"].infos
    assert list(f["view_id"].unique()) == [0, 1, 2, 3] and (f["scene_id"] == 7).all()
    assert (f["time"] == 0.75).all() and (preds["coarse"].infos["time"] == 0.25).all()   # :265-291
    assert preds["final"].poses.shape == (n_rows, 4, 4)
    data = {"coarse": {"time": 1.0}, "refiner": {"time": 2.0}, "depth_refiner": {"time": 4.0}}
    assert [compute_pose_est_total_time(data, k) for k in ("coarse", "refiner/final", "depth_refiner", "final")] == [1.0, 3.0, 7.0, 7.0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    from happypose_b200 import distributed as hdist
    from happypose_b200.evaluation.prediction_runner import PredictionRunner

    hdist.init_distributed_mode(backend="gloo")
    cfg = InferenceConfig(detection_type="gt", n_refiner_iterations=5)
    frames = make_frames(5)
    runner = PredictionRunner(frames, cfg, sharding="scenes")
    assert list(runner.sampler) == ([0, 2, 4] if rank == 0 else [1, 3])   # DistributedSceneSampler: whole frames per rank
    preds = runner.get_predictions(FakeEstimator())
    f = preds["final"].infos
    assert sorted(f["view_id"].unique().tolist()) == [0, 1, 2, 3, 4]       # gathered on every rank, no tmp files
    assert len(preds["final"]) == sum(len(fr["detections"]) for fr in frames)
    assert preds["final"].poses.shape[0] == len(f)
    # rank order: rank 0's frames first
    assert f["view_id"].tolist()[:3] == [0, 2, 2]
    # hypothesis sharding: every rank walks every frame, nothing to gather
    runner2 = PredictionRunner(frames, cfg, sharding="hypotheses")
    assert list(runner2.sampler) == [0, 1, 2, 3, 4]
    hdist.barrier()
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([len(f)]))
    dist.destroy_process_group()


def test_scene_sharding_and_gather_world_size_2_gloo(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok_0.npy") and os.path.exists(tmp_path / "ok_1.npy")
