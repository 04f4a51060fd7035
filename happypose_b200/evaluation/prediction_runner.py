"""Evaluation-harness caller of the hot path (mirror of
happypose/pose_estimators/megapose/evaluation/prediction_runner.py:44-291): iterates the frames of a scene dataset, runs
PoseEstimator.run_inference_pipeline on each and collects the predictions of every stage with their timings.

Differences in HOW:
  * results are gathered across ranks with ONE all_gather_object over the process group (NCCL / gloo) instead of pickle
    files in a tmp dir + barriers (toolbox/utils/tensor_collection.py:166-187);
  * two ways to use several GPUs.  sharding="scenes" is the reference's: a DistributedSceneSampler deals whole frames to
    the ranks (prediction_runner.py:65-67) -- throughput.  sharding="hypotheses": every rank sees every frame and the
    PoseEstimator (built with shard_across_ranks=True) splits each frame's hypothesis rows across the GPUs of the box --
    the per-image LATENCY case; the harness evaluates one image at a time (evaluation.py:195 asserts batch_size == 1).

`scene_ds` is any sequence of frames: dicts with "rgb" [H,W,3] uint8, "K" [3,3], optional "depth" [H,W] (metres),
"im_info" {scene_id, view_id}, and -- for detection_type "gt" / "exte" -- "detections" (a DetectionsType with bboxes) and
optionally "initial_data" (coarse estimates for coarse_estimation_type="external").  The reference's SceneDataset /
SceneObservation classes (data loading) are out of scope; SceneObservation.collate_fn's batch of one is this dict.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import distributed as hdist
from ..inference.types import DetectionsType, InferenceConfig, ObservationTensor, PoseEstimatesType
from ..inference.utils import add_instance_id
from ..utils import tensor_collection as tc


class DistributedSceneSampler:
    """toolbox/datasets/samplers.py: frame indices rank, rank + world, ... (no padding: every frame is predicted once)."""

    def __init__(self, n_frames: int, num_replicas: int, rank: int):
        self.indices = list(range(rank, n_frames, num_replicas))

    def __iter__(self):
        return iter(self.indices)

    def __len__(self):
        return len(self.indices)


def compute_pose_est_total_time(all_preds_data: dict, pred_name: str) -> float:
    """prediction_runner.py:265-291."""
    dt_coarse = all_preds_data["coarse"]["time"]
    dt_coarse_refiner = dt_coarse + all_preds_data["refiner"]["time"]
    if "depth_refiner" in all_preds_data:
        dt_coarse_refiner_depth = dt_coarse_refiner + all_preds_data["depth_refiner"]["time"]
    if pred_name.startswith("coarse"):
        return dt_coarse
    if pred_name.startswith("refiner"):
        return dt_coarse_refiner
    if pred_name == "depth_refiner":
        return dt_coarse_refiner_depth
    if pred_name == "final":
        return dt_coarse_refiner_depth if "depth_refiner" in all_preds_data else dt_coarse_refiner
    raise ValueError(f"{pred_name} extra data not in {all_preds_data.keys()}")


class PredictionRunner:
    def __init__(self, scene_ds: Sequence[dict], inference_cfg: InferenceConfig, batch_size: int = 1, n_workers: int = 4,
                 sharding: str = "scenes") -> None:
        assert batch_size == 1, "the evaluation harness predicts one image at a time (evaluation.py:195)"
        assert sharding in ("scenes", "hypotheses")
        self.inference_cfg = inference_cfg
        self.rank = hdist.get_rank()
        self.world_size = hdist.get_world_size()
        self.sharding = sharding
        self.scene_ds = scene_ds
        self.batch_size = batch_size
        if sharding == "scenes":
            self.sampler = DistributedSceneSampler(len(scene_ds), num_replicas=self.world_size, rank=self.rank)
        else:
            self.sampler = DistributedSceneSampler(len(scene_ds), num_replicas=1, rank=0)
        self.load_depth = any(frame.get("depth") is not None for frame in scene_ds) if len(scene_ds) else False

    # ------------------------------------------------------------------------------------------------------------
    def run_inference_pipeline(self, pose_estimator, obs_tensor: ObservationTensor, detections: Optional[DetectionsType],
                               initial_estimates: Optional[PoseEstimatesType] = None) -> Tuple[Dict[str, PoseEstimatesType], dict]:
        """prediction_runner.py:79-165 -> (all_preds, all_preds_data) with keys 'final', 'refiner/iteration=N',
        'refiner/final', 'coarse' [, 'depth_refiner']."""
        cfg = self.inference_cfg
        if cfg.detection_type in ("gt", "exte"):
            run_detector = False
        elif cfg.detection_type == "detector":
            detections = None
            run_detector = True
        else:
            raise ValueError(f"Unknown detection type {cfg.detection_type}")
        coarse_estimates = None
        if cfg.coarse_estimation_type == "external":
            coarse_estimates = add_instance_id(initial_estimates)
            coarse_estimates.infos["instance_id"] = 0
            run_detector = False
        preds, extra_data = pose_estimator.run_inference_pipeline(
            obs_tensor, detections=detections, run_detector=run_detector, coarse_estimates=coarse_estimates,
            n_refiner_iterations=cfg.n_refiner_iterations, n_pose_hypotheses=cfg.n_pose_hypotheses,
            run_depth_refiner=cfg.run_depth_refiner, bsz_images=cfg.bsz_images, bsz_objects=cfg.bsz_objects)
        ref_it_str = f"refiner/iteration={cfg.n_refiner_iterations}"
        data_TCO_refiner = extra_data["refiner"]["preds"]
        all_preds = {"final": preds, ref_it_str: data_TCO_refiner, "refiner/final": data_TCO_refiner, "coarse": extra_data["coarse"]["preds"]}
        coarse_data = dict(extra_data["coarse"]["data"] or {})
        coarse_data.pop("TCO", None)
        all_preds_data = {"coarse": coarse_data, "refiner": extra_data["refiner"]["data"], "scoring": extra_data["scoring"]}
        if cfg.run_depth_refiner:
            all_preds["depth_refiner"] = extra_data["depth_refiner"]["preds"]
            all_preds_data["depth_refiner"] = extra_data["depth_refiner"].get("data", {"time": 0.0})
        for v in all_preds.values():
            if "mask" in v.tensors:
                v.delete_tensor("mask")
        return all_preds, all_preds_data

    # ------------------------------------------------------------------------------------------------------------
    def get_predictions(self, pose_estimator, gather: bool = True) -> Dict[str, PoseEstimatesType]:
        """prediction_runner.py:167-262: every stage's predictions over this rank's frames, with time / scene_id / view_id
        columns; with `gather` the ranks' results are concatenated (rank order) on every rank."""
        predictions_list = defaultdict(list)
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        first = True
        for idx in self.sampler:
            data = self.scene_ds[idx]
            im_info = data.get("im_info", {"scene_id": 0, "view_id": idx})
            rgb = torch.as_tensor(np.asarray(data["rgb"])).unsqueeze(0).permute(0, 3, 1, 2).contiguous()
            depth = None if data.get("depth") is None else torch.as_tensor(np.asarray(data["depth"], np.float32)).unsqueeze(0)
            K = torch.as_tensor(np.asarray(data["K"], np.float32)).unsqueeze(0)
            detections = data.get("detections")
            dt_det_exte = 0.0
            if detections is not None:
                detections = detections.to(dev)
                if self.inference_cfg.detection_type == "gt" and "visib_fract" in detections.infos:
                    detections = detections[np.where(detections.infos["visib_fract"].to_numpy() > 0.05)[0].tolist()]  # :211-219
                if self.inference_cfg.detection_type == "exte" and "time" in detections.infos and len(detections) > 0:
                    dt_det_exte += float(detections.infos["time"].iloc[0])
            initial = data.get("initial_data")
            if initial is not None:
                initial = initial.to(dev)
            obs_tensor = ObservationTensor.from_torch_batched(rgb, depth, K).to(dev)
            with torch.no_grad():
                if first:  # the reference also runs the first frame twice (warm-up, :233-240)
                    self.run_inference_pipeline(pose_estimator, obs_tensor, detections, initial_estimates=initial)
                    first = False
                all_preds, all_preds_data = self.run_inference_pipeline(pose_estimator, obs_tensor, detections, initial_estimates=initial)
            for pred_name, pred in all_preds.items():
                infos = pred.infos.copy()
                infos["time"] = dt_det_exte + compute_pose_est_total_time(all_preds_data, pred_name)
                infos["scene_id"] = im_info["scene_id"]
                infos["view_id"] = im_info["view_id"]
                predictions_list[pred_name].append(tc.PandasTensorCollection(infos, **{k: v for k, v in pred.tensors.items()}))
        predictions = {k: tc.concatenate(v) for k, v in predictions_list.items()}
        if gather and self.sharding == "scenes" and hdist.is_distributed():
            names = sorted(predictions.keys())
            # every rank must take part in every collective, also a rank that got no frame
            all_names = [None] * self.world_size
            torch.distributed.all_gather_object(all_names, names)
            for name in sorted({n for ns in all_names for n in ns}):
                local = predictions.get(name, tc.PandasTensorCollection(infos=__import__("pandas").DataFrame()))
                predictions[name] = local.gather_distributed()
        return predictions
