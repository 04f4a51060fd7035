"""CosyPose PoseEstimator re-hosted on the B200 kernels (mirror of
happypose/pose_estimators/cosypose/cosypose/integrated/pose_estimator.py:31-470): TCO init from the detections,
coarse model for n_coarse_iterations, refiner for n_refiner_iterations; no SO(3) grid and no top-K."""
from __future__ import annotations

import time
from typing import Any, List, Optional, Tuple

import numpy as np
import torch

from ..inference.types import DetectionsType, ObservationTensor, PoseEstimatesType
from ..inference.utils import filter_detections
from ..lib3d.cosypose_ops import TCO_init_from_boxes, TCO_init_from_boxes_zup_autodepth
from ..utils.tensor_collection import PandasTensorCollection
from ..utils.timer import CudaTimer, SimpleTimer


class PoseEstimator(torch.nn.Module):
    def __init__(self, refiner_model=None, coarse_model=None, detector_model=None, bsz_objects: int = 8, bsz_images: int = 256) -> None:
        super().__init__()
        self.coarse_model = coarse_model
        self.refiner_model = refiner_model
        self.detector_model = detector_model
        self.bsz_objects = bsz_objects
        self.bsz_images = bsz_images
        model = refiner_model if refiner_model is not None else coarse_model
        if model is None:
            raise ValueError("At least one of refiner_model or  coarse_model must be specified.")
        self.cfg = getattr(model, "cfg", None)
        self.mesh_db = model.mesh_db
        self.eval()
        self.keep_all_outputs = False
        self.keep_all_coarse_outputs = False
        self.refiner_outputs = None
        self.coarse_outputs = None
        self.debug_dict: dict = {}

    def make_TCO_init(self, detections, K):
        """pose_estimator.py:125-135."""
        dev = detections.bboxes.device
        im_ids = torch.as_tensor(detections.infos["batch_im_id"].to_numpy()).to(K.device)
        K = K[im_ids]
        boxes = detections.bboxes.float()
        init_method = getattr(getattr(self.coarse_model, "cfg", None), "init_method", "v0")
        if init_method == "z-up+auto-depth":
            labels = detections.infos["label"].tolist()
            db = self.coarse_model.mesh_db
            TCO_init = TCO_init_from_boxes_zup_autodepth(boxes, db.points_subset(2000), K, obj_ids=db.label_ids(labels, dev))
        else:
            TCO_init = TCO_init_from_boxes(z_range=(1.0, 1.0), boxes=boxes, K=K)
        return PandasTensorCollection(infos=detections.infos, poses=TCO_init)

    @torch.no_grad()
    def _run_model(self, model, prefix, observation, data_TCO_input, n_iterations, keep_all_outputs, cuda_timer):
        start_time = time.time()
        B = data_TCO_input.poses.shape[0]
        df = data_TCO_input.infos
        dev = observation.images.device
        im_ids_all = torch.as_tensor(df["batch_im_id"].to_numpy().astype(np.int32)).to(dev)
        keys = ("poses", "poses_input", "K_crop", "K", "boxes_rend", "boxes_crop")
        chunks = {n: {k: [] for k in keys} for n in range(1, n_iterations + 1)}
        all_outputs = []
        model_time = 0.0
        bidx, iidx = np.zeros(B, np.int64), np.zeros(B, np.int64)
        labels_all = df["label"].tolist()
        for batch_idx, s in enumerate(range(0, B, self.bsz_objects)):
            e = min(s + self.bsz_objects, B)
            bidx[s:e], iidx[s:e] = batch_idx, np.arange(e - s)
            timer_ = CudaTimer(enabled=cuda_timer) if torch.cuda.is_available() else SimpleTimer()
            timer_.start()
            outputs_ = model(images=observation.images, K=observation.K[im_ids_all[s:e].long()], TCO=data_TCO_input.poses[s:e],
                             n_iterations=n_iterations, labels=labels_all[s:e], im_ids=im_ids_all[s:e])
            timer_.stop()
            model_time += timer_.elapsed()
            if keep_all_outputs:
                all_outputs.append(outputs_)
            for n in range(1, n_iterations + 1):
                o = outputs_[f"iteration={n}"]
                for k, v in zip(keys, (o.TCO_output, o.TCO_input, o.K_crop, o.K, o.boxes_rend, o.boxes_crop)):
                    chunks[n][k].append(v)
        infos = df.copy()
        infos[f"{prefix}_batch_idx"] = bidx
        infos[f"{prefix}_instance_idx"] = iidx
        preds = {f"iteration={n}": PandasTensorCollection(infos, **{k: torch.cat(chunks[n][k]) for k in keys}) for n in range(1, n_iterations + 1)}
        extra_data = {"n_iterations": n_iterations, "outputs": all_outputs, "model_time": model_time, "time": time.time() - start_time}
        return preds, extra_data

    def forward_coarse_model(self, observation, data_TCO_input, n_iterations=5, keep_all_outputs=False, cuda_timer=False):
        return self._run_model(self.coarse_model, "coarse", observation, data_TCO_input, n_iterations, keep_all_outputs, cuda_timer)

    def forward_refiner(self, observation, data_TCO_input, n_iterations=5, keep_all_outputs=False, cuda_timer=False):
        return self._run_model(self.refiner_model, "refiner", observation, data_TCO_input, n_iterations, keep_all_outputs, cuda_timer)

    def forward_detection_model(self, observation, detection_th: float = 0.7, mask_th: float = 0.8, *args: Any, **kwargs: Any):
        return self.detector_model.get_detections(
            observation=observation, one_instance_per_class=False, detection_th=detection_th, output_masks=False, mask_th=mask_th)

    @torch.no_grad()
    def run_inference_pipeline(
        self,
        observation: ObservationTensor,
        detections: Optional[DetectionsType] = None,
        data_TCO_init: Optional[PandasTensorCollection] = None,
        run_detector: Optional[bool] = None,
        n_refiner_iterations: int = 1,
        n_coarse_iterations: int = 1,
        bsz_images: Optional[int] = None,
        bsz_objects: Optional[int] = None,
        coarse_estimates: Optional[PoseEstimatesType] = None,
        detection_th: float = 0.7,
        mask_th: float = 0.8,
        labels_to_keep: Optional[List[str]] = None,
    ) -> Tuple[PoseEstimatesType, dict]:
        """pose_estimator.py:137-247."""
        timing_str = ""
        timer = SimpleTimer()
        timer.start()
        if bsz_images is not None:
            self.bsz_images = bsz_images
        if bsz_objects is not None:
            self.bsz_objects = bsz_objects
        if coarse_estimates is None:
            assert detections is not None or run_detector, "You must either pass in `detections` or set run_detector=True"
            if detections is None and run_detector:
                start_time = time.time()
                detections = self.forward_detection_model(observation, detection_th, mask_th).to(observation.images.device)
                timing_str += f"detection={time.time() - start_time:.2f}, "
        preds = {}
        coarse_extra_data = refiner_extra_data = None
        if data_TCO_init is None:
            assert detections is not None
            assert self.coarse_model is not None
            assert n_coarse_iterations > 0
            if labels_to_keep is not None:
                detections = filter_detections(detections, labels_to_keep)
            data_TCO_init = self.make_TCO_init(detections, observation.K)
            coarse_preds, coarse_extra_data = self.forward_coarse_model(observation, data_TCO_init, n_iterations=n_coarse_iterations)
            for n in range(1, n_coarse_iterations + 1):
                preds[f"coarse/iteration={n}"] = coarse_preds[f"iteration={n}"]
            data_TCO_coarse = coarse_preds[f"iteration={n_coarse_iterations}"]
        else:
            assert n_coarse_iterations == 0
            preds["external_coarse"] = data_TCO_init
            data_TCO_coarse = data_TCO_init
        data_TCO = data_TCO_coarse
        if n_refiner_iterations >= 1:
            assert self.refiner_model is not None
            refiner_preds, refiner_extra_data = self.forward_refiner(observation, data_TCO_coarse, n_iterations=n_refiner_iterations)
            for n in range(1, n_refiner_iterations + 1):
                preds[f"refiner/iteration={n}"] = refiner_preds[f"iteration={n}"]
            data_TCO = refiner_preds[f"iteration={n_refiner_iterations}"]
        timer.stop()
        extra_data: dict = {}
        extra_data["coarse"] = {"preds": data_TCO_coarse, "data": coarse_extra_data}
        extra_data["refiner_all_hypotheses"] = {"preds": preds, "data": refiner_extra_data}
        extra_data["refiner"] = {"preds": data_TCO, "data": refiner_extra_data}
        extra_data["timing_str"] = f"total={timer.elapsed():.2f}, {timing_str}"
        extra_data["time"] = timer.elapsed()
        return data_TCO, extra_data
