"""MegaPose PoseEstimator re-hosted on the B200 kernels.

Mirror of happypose/pose_estimators/megapose/inference/pose_estimator.py:55-668: same constructor, same methods
(run_inference_pipeline, forward_coarse_model, forward_refiner, forward_scoring_model, forward_detection_model,
run_depth_refiner, load_SO3_grid), same return types and `extra_data` keys.  Differences in HOW:

  * the B*M hypothesis table is built with vectorised pandas (no `pd.DataFrame([row] * M)` loop, :351-362);
  * per-row ids (mesh id, mesh-db id, frame id, bbox id, grid id) are device tensors computed once; batches are
    slices of them, so the hot loops contain no pandas / label lookups (:379-395);
  * the frame is never expanded per hypothesis (`observation.images[batch_im_ids_]`, :390);
  * coarse logits stay on the device for the top-K (hpb_topk_segmented); the DataFrame columns are filled with ONE
    device->host copy at the end of each stage;
  * with torch.distributed initialised (one process per GPU) hypotheses are sharded across ranks and the logits are
    all-gathered (happypose_b200/distributed.py) -- the reference shards whole scenes instead.
"""
from __future__ import annotations

import time
from collections import defaultdict
from typing import Any, List, Optional, Tuple

import numpy as np
import pandas as pd
import torch

from .. import _capi, distributed as hdist, ops
from ..inference.types import DetectionsType, ObservationTensor, PoseEstimatesType, assert_detections_valid
from ..inference.utils import add_instance_id, filter_detections
from ..utils import tensor_collection as tc
from ..utils import transform_utils
from ..utils.tensor_collection import PandasTensorCollection, filter_top_pose_estimates
from ..utils.timer import CudaTimer, SimpleTimer, Timer

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


class PoseEstimator(torch.nn.Module):
    """Performs inference for pose estimation."""

    def __init__(
        self,
        refiner_model: Optional[torch.nn.Module] = None,
        coarse_model: Optional[torch.nn.Module] = None,
        detector_model: Optional[torch.nn.Module] = None,
        depth_refiner: Optional[Any] = None,
        bsz_objects: int = 8,
        bsz_images: int = 256,
        SO3_grid_size: int = 576,
        shard_across_ranks: bool = False,
        use_cuda_graphs: bool = False,
    ) -> None:
        super().__init__()
        self.coarse_model = coarse_model
        self.refiner_model = refiner_model
        self.detector_model = detector_model
        self.depth_refiner = depth_refiner
        self.bsz_objects = bsz_objects
        self.bsz_images = bsz_images
        # Opt-in (the reference's distributed use is the opposite: every rank gets DIFFERENT scenes and the results are
        # gathered afterwards, evaluation/prediction_runner.py:65-67).  With shard_across_ranks=True every rank must call
        # the pipeline with the SAME observation and detections; the rows of the hypothesis table are then split across
        # the ranks.  The first call for every input shape verifies that the ranks agree (see _check_sharded_inputs).
        self.shard_across_ranks = shard_across_ranks
        self._sharding_checked: set = set()
        if SO3_grid_size is not None:
            self.load_SO3_grid(SO3_grid_size)
        if self.refiner_model is not None:
            self.cfg = getattr(self.refiner_model, "cfg", None)
            self.mesh_db = self.refiner_model.mesh_db
        elif self.coarse_model is not None:
            self.cfg = getattr(self.coarse_model, "cfg", None)
            self.mesh_db = self.coarse_model.mesh_db
        else:
            raise ValueError("At least one of refiner_model or  coarse_model must be specified.")
        self.eval()
        self.use_cuda_graphs = use_cuda_graphs
        self.keep_all_outputs = False
        self.keep_all_coarse_outputs = False
        self.refiner_outputs = None
        self.coarse_outputs = None
        self.debug_dict: dict = {}

    @property
    def use_cuda_graphs(self) -> bool:
        """Replay the launch-bound stages (refiner iterations, small scoring batches) as CUDA graphs."""
        return self._use_cuda_graphs

    @use_cuda_graphs.setter
    def use_cuda_graphs(self, value: bool) -> None:
        self._use_cuda_graphs = bool(value)
        for m in (self.coarse_model, self.refiner_model):
            if m is not None and hasattr(m, "use_cuda_graphs"):
                m.use_cuda_graphs = bool(value)

    def load_SO3_grid(self, grid_size: int) -> None:
        self._SO3_grid = transform_utils.load_SO3_grid(grid_size).to(device)

    # ------------------------------------------------------------------------------------------
    def _row_ids(self, model, data, dev) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(mesh-db ids, renderer mesh ids, frame ids) of every row, as int32 device tensors.  `data` is a
        PandasTensorCollection (its device-resident row tensors are used when a previous stage attached them) or a
        DataFrame.  Labels are resolved once per distinct label (KeyError on unknown labels, like mesh_db.select / the
        renderer)."""
        if isinstance(data, PandasTensorCollection):
            rt = data.row_tensors
            if all(k in rt for k in ("obj_ids", "mesh_ids", "im_ids")):
                return rt["obj_ids"].to(dev), rt["mesh_ids"].to(dev), rt["im_ids"].to(dev)
            df = data.infos
        else:
            df = data
        labels = df["label"].to_numpy()
        uniq, inv = np.unique(labels, return_inverse=True)
        db = np.array([model.mesh_db.label_to_id[u] for u in uniq], np.int32)[inv]
        rd = np.array([model.renderer._label_to_mesh_id[u] for u in uniq], np.int32)[inv]
        im = df["batch_im_id"].to_numpy().astype(np.int32)
        stacked = torch.as_tensor(np.stack([db, rd, im])).to(dev)
        return stacked[0].contiguous(), stacked[1].contiguous(), stacked[2].contiguous()

    @staticmethod
    def _my_rows(n: int, shard: bool) -> Tuple[int, int]:
        return hdist.shard_range(n) if (shard and hdist.is_distributed()) else (0, n)

    def _check_sharded_inputs(self, observation: ObservationTensor, detections) -> None:
        """Row sharding assumes identical inputs on every rank; mismatched inputs would otherwise hang or silently mix
        scenes.  Checked (one all-reduce + host read) the first time a given input shape is seen."""
        if not (self.shard_across_ranks and hdist.is_distributed()):
            return
        key = (len(detections), tuple(observation.images.shape))
        if key in self._sharding_checked:
            return
        import zlib

        labels = "|".join(str(x) for x in detections.infos["label"].tolist())
        ims = ",".join(str(int(x)) for x in detections.infos["batch_im_id"].tolist())
        sig = [float(len(detections)), float(zlib.crc32((labels + "#" + ims).encode()) % (1 << 24)),
               float(detections.bboxes.double().sum().item()), float(observation.K.double().sum().item()),
               float(observation.images[..., ::16, ::16].double().sum().item())]
        if not hdist.all_ranks_equal(sig, observation.images.device):
            raise ValueError("PoseEstimator(shard_across_ranks=True) needs the same observation and detections on every rank "
                             "(rows of one hypothesis table are split across the ranks); for one-scene-per-rank evaluation "
                             "construct the estimator with shard_across_ranks=False")
        self._sharding_checked.add(key)

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_refiner(
        self,
        observation: ObservationTensor,
        data_TCO_input: PoseEstimatesType,
        n_iterations: int = 5,
        keep_all_outputs: bool = False,
        cuda_timer: bool = False,
        **refiner_kwargs,
    ) -> Tuple[dict, dict]:
        """Runs the refiner for n_iterations on batches of bsz_objects rows (pose_estimator.py:105-220).
        preds: {'iteration=n': PoseEstimatesType}; extra_data: n_iterations, outputs, model_time, time."""
        timer = Timer()
        timer.start()
        start_time = time.time()
        assert self.refiner_model is not None
        model = self.refiner_model
        dev = observation.images.device
        B = data_TCO_input.poses.shape[0]
        obj_ids, mesh_ids, im_ids = self._row_ids(model, data_TCO_input, dev)
        K_rows = observation.K[im_ids.long()]
        lo, hi = self._my_rows(B, self.shard_across_ranks)

        keys = ("poses", "poses_input", "K_crop", "K", "boxes_rend", "boxes_crop")
        chunks = {n: {k: [] for k in keys} for n in range(1, n_iterations + 1)}
        all_outputs = []
        model_time = 0.0
        # refiner_batch_idx / refiner_instance_idx of ALL rows (every rank's shard), so `infos` is the same on every rank
        batch_idx_col = np.zeros(B, np.int64)
        inst_idx_col = np.zeros(B, np.int64)
        sharded = self.shard_across_ranks and hdist.is_distributed()
        for r in range(hdist.get_world_size() if sharded else 1):
            lo_r, hi_r = hdist.shard_range(B, r, hdist.get_world_size()) if sharded else (0, B)
            rel = np.arange(hi_r - lo_r)
            batch_idx_col[lo_r:hi_r] = rel // self.bsz_objects
            inst_idx_col[lo_r:hi_r] = rel % self.bsz_objects
        for batch_idx, s in enumerate(range(lo, hi, self.bsz_objects)):
            e = min(s + self.bsz_objects, hi)
            timer_ = CudaTimer(enabled=cuda_timer) if torch.cuda.is_available() else SimpleTimer()
            timer_.start()
            outputs_ = model.forward_ids(
                observation.images, K_rows[s:e], obj_ids[s:e], mesh_ids[s:e], im_ids[s:e], data_TCO_input.poses[s:e],
                n_iterations=n_iterations, labels=data_TCO_input.infos["label"].iloc[s:e].tolist() if keep_all_outputs else None,
                **refiner_kwargs)
            timer_.stop()
            model_time += timer_.elapsed()
            if keep_all_outputs:
                all_outputs.append(outputs_)
            for n in range(1, n_iterations + 1):
                o = outputs_[f"iteration={n}"]
                for k, v in zip(keys, (o.TCO_output, o.TCO_input, o.K_crop, o.K, o.boxes_rend, o.boxes_crop)):
                    chunks[n][k].append(v)

        def make_infos(cache=[]):  # one column set shared by all iterations, built when first looked at
            if not cache:
                cache.append(tc.cols_assign(data_TCO_input._cols(), refiner_batch_idx=batch_idx_col, refiner_instance_idx=inst_idx_col))
            return cache[0]

        row_tensors = {"obj_ids": obj_ids, "mesh_ids": mesh_ids, "im_ids": im_ids}
        if "group_ids" in data_TCO_input.row_tensors:
            row_tensors["group_ids"] = data_TCO_input.row_tensors["group_ids"]
        shape_tail = {"poses": (4, 4), "poses_input": (4, 4), "K_crop": (3, 3), "K": (3, 3), "boxes_rend": (4,), "boxes_crop": (4,)}
        local = {(n, k): (torch.cat(chunks[n][k], dim=0) if chunks[n][k] else torch.zeros((0,) + shape_tail[k], device=dev))
                 for n in range(1, n_iterations + 1) for k in keys}
        if hdist.is_distributed() and self.shard_across_ranks:
            # ONE all-gather for the whole stage: every (iteration, tensor) of this rank's rows packed side by side
            # ([rows, n_iterations * 58] floats) instead of one collective per tensor per iteration
            local = hdist.all_gather_rows_packed(local, B)
        preds = {}
        for n in range(1, n_iterations + 1):
            tensors = {k: local[(n, k)] for k in keys}
            preds[f"iteration={n}"] = PandasTensorCollection(make_infos, n_rows=B, row_tensors=row_tensors, **tensors)
            preds[f"iteration={n}"].meta.update(data_TCO_input.meta)

        elapsed = time.time() - start_time
        extra_data = {"n_iterations": n_iterations, "outputs": all_outputs, "model_time": model_time, "time": elapsed}
        timer.stop()
        return preds, extra_data

    # ------------------------------------------------------------------------------------------
    def _score_rows(self, observation, ids, TCO, cuda_timer, return_debug_data):
        """Shared by the coarse and scoring stages: forward_coarse over batches of bsz_images rows of this rank's
        slice, then all-gather.  Returns (logits [n], scores [n], render_time, model_time, n_batches, debug)."""
        model = self.coarse_model
        dev = observation.images.device
        n = TCO.shape[0]
        obj_ids, mesh_ids, im_ids = ids
        K_rows = observation.K[im_ids.long()]
        lo, hi = self._my_rows(n, self.shard_across_ranks)
        logits_l, scores_l, crops_l, renders_l = [], [], [], []
        render_time = model_time = 0.0
        n_batches = 0
        for s in range(lo, hi, self.bsz_images):
            e = min(s + self.bsz_images, hi)
            out_ = model.forward_coarse_ids(
                observation.images, K_rows[s:e], obj_ids[s:e], mesh_ids[s:e], im_ids[s:e], TCO[s:e],
                cuda_timer=cuda_timer, return_debug_data=return_debug_data)
            render_time += out_["render_time"]
            model_time += out_["model_time"]
            logits_l.append(out_["logits"])
            scores_l.append(out_["scores"])
            if return_debug_data:
                crops_l.append(out_["images_crop"])
                renders_l.append(out_["renders"])
            n_batches += 1
        n_out = model.n_rendered_views
        logits = torch.cat(logits_l) if logits_l else torch.zeros((0, n_out), device=dev)
        scores = torch.cat(scores_l) if scores_l else torch.zeros((0, n_out), device=dev)
        if hdist.is_distributed() and self.shard_across_ranks:
            # THE collective of the path: all-gather of the fp32 coarse scores before the replicated top-K
            both = hdist.all_gather_rows(torch.cat([logits, scores], dim=1).contiguous(), n)
            logits, scores = both[:, :n_out].contiguous(), both[:, n_out:].contiguous()
        debug = {}
        if return_debug_data:
            debug = {"images_crop": torch.cat(crops_l), "renders": torch.cat(renders_l)}
        return logits, scores, render_time, model_time, n_batches, debug

    @torch.no_grad()
    def forward_scoring_model(
        self,
        observation: ObservationTensor,
        data_TCO: PoseEstimatesType,
        cuda_timer: bool = False,
        return_debug_data: bool = False,
    ) -> Tuple[PoseEstimatesType, dict]:
        """Scores the estimates with the coarse model; adds pose_logit / pose_score (pose_estimator.py:223-325)."""
        start_time = time.time()
        assert self.coarse_model is not None
        n = len(data_TCO)
        ids = self._row_ids(self.coarse_model, data_TCO, observation.images.device)
        logits, scores, render_time, model_time, n_batches, debug_data = self._score_rows(
            observation, ids, data_TCO.poses, cuda_timer, return_debug_data)
        # the stage's single D2H copy: enqueued here, awaited when the frame is looked at
        both = tc.HostCopy(torch.stack([logits.reshape(n, -1)[:, 0], scores.reshape(n, -1)[:, 0]]))

        def add_scores(cols, both=both):
            host = both.numpy()
            return tc.cols_assign(cols, pose_logit=host[0], pose_score=host[1])

        data_TCO.map_cols(add_scores)  # in place, like the reference (df["pose_logit"] = ...; data_TCO.infos = df)

        elapsed = time.time() - start_time
        timing_str = f"time: {elapsed:.2f}, model_time: {model_time:.2f}, render_time: {render_time:.2f}"
        extra_data = {
            "render_time": render_time, "model_time": model_time, "time": elapsed, "logits": logits, "scores": scores,
            "debug": debug_data, "n_batches": n_batches, "timing_str": timing_str,
        }
        return data_TCO, extra_data

    @torch.no_grad()
    def forward_coarse_model(
        self,
        observation: ObservationTensor,
        detections: DetectionsType,
        cuda_timer: bool = False,
        return_debug_data: bool = False,
        _instance_ids_after_enqueue: bool = False,
    ) -> Tuple[PoseEstimatesType, dict]:
        """SO(3)-grid hypotheses for every detection, scored by the coarse model (pose_estimator.py:328-485).

        `_instance_ids_after_enqueue` (run_inference_pipeline): `instance_id` is added to the detections AFTER the coarse
        stage has been enqueued -- the kernels only need labels, frame ids and boxes, so the pandas bookkeeping (instance
        ids, top-K group ids) runs while the GPU is already busy instead of in front of an idle GPU."""
        start_time = time.time()
        if not _instance_ids_after_enqueue:
            assert_detections_valid(detections)
        coarse_model = self.coarse_model
        dev = observation.images.device
        SO3_grid = self._SO3_grid.to(dev)
        B = len(detections)
        M = SO3_grid.shape[0]

        # every detection row repeated M times, with hypothesis_id / bbox_id columns (:351-362).  The row ids the kernels
        # need are built on the device from B-sized host arrays; the B*M-row DataFrame itself is deferred (make_infos).
        df = detections.infos
        obj_ids_d, mesh_ids_d, im_ids_d = self._row_ids(coarse_model, detections, dev)
        rep = torch.arange(B, device=dev).repeat_interleave(M)            # row -> detection position
        m_idx = torch.arange(M, device=dev).repeat(B)                     # row -> grid rotation
        obj_ids, mesh_ids, im_ids = obj_ids_d[rep].contiguous(), mesh_ids_d[rep].contiguous(), im_ids_d[rep].contiguous()
        bboxes = detections.bboxes.to(dev)[rep].float()
        K_rows = observation.K[im_ids.long()]
        # initial poses of ALL rows on every rank (cheap; lets the top-K be replicated without exchanging poses)
        TCO = ops.tco_init(
            coarse_model._ctx(), _capi.TCO_INIT_AUTODEPTH_WITH_R, bboxes, K_rows, coarse_model.mesh_db.points, obj_ids,
            SO3_grid[m_idx])

        logits, scores, render_time, model_time, n_batches, dbg = self._score_rows(
            observation, (obj_ids, mesh_ids, im_ids), TCO, cuda_timer, return_debug_data)
        # ---- host bookkeeping, behind the enqueued coarse stage ----
        if _instance_ids_after_enqueue:
            detections = add_instance_id(detections)
            assert_detections_valid(detections)
            df = detections.infos
        # groups of the later top-K = (batch_im_id, label, instance_id); computed on the B detection rows
        det_groups_np = tc.group_ids_from_columns(df, ["batch_im_id", "label", "instance_id"])
        n_groups = int(det_groups_np.max()) + 1
        # pinned + non_blocking: a pageable H2D copy would block the host until the coarse stage enqueued above has run
        det_groups = torch.as_tensor(det_groups_np)
        det_groups = (det_groups.pin_memory().to(dev, non_blocking=True) if dev.type == "cuda" else det_groups.to(dev))
        row_tensors = {"obj_ids": obj_ids, "mesh_ids": mesh_ids, "im_ids": im_ids, "group_ids": det_groups[rep].contiguous()}
        logits = logits.reshape([B, M])
        scores = scores.reshape([B, M])
        debug_data = {}
        if return_debug_data:
            H, W = dbg["images_crop"].shape[2:]
            debug_data = {"images_crop": dbg["images_crop"].reshape([B, M, -1, H, W]), "renders": dbg["renders"].reshape([B, M, -1, H, W])}

        # the stage's single D2H copy: enqueued here, awaited when the frame is looked at
        both = tc.HostCopy(torch.stack([logits.flatten(), scores.flatten()]))

        def make_infos(df=df, both=both):
            pos = np.repeat(np.arange(B), M)
            cols = tc.cols_take(tc.cols_of(df), pos)
            host = both.numpy()
            return tc.cols_assign(cols, hypothesis_id=np.tile(np.arange(M), B), bbox_id=np.repeat(df.index.to_numpy(), M),
                                  coarse_logit=host[0], coarse_score=host[1])

        elapsed = time.time() - start_time
        timing_str = f"time: {elapsed:.2f}, model_time: {model_time:.2f}, render_time: {render_time:.2f}"
        extra_data = {
            "render_time": render_time, "model_time": model_time, "time": elapsed, "logits": logits, "scores": scores,
            "TCO": TCO.reshape([B, M, 4, 4]), "debug": debug_data, "n_batches": n_batches, "timing_str": timing_str,
        }
        data_TCO = PandasTensorCollection(make_infos, n_rows=B * M, row_tensors=row_tensors, poses=TCO, bboxes=bboxes)
        data_TCO.meta.update({"group_cols": ["batch_im_id", "label", "instance_id"], "n_groups": n_groups, "min_group_size": M})
        return data_TCO, extra_data

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_detection_model(self, observation: ObservationTensor, *args: Any, **kwargs: Any) -> DetectionsType:
        return self.detector_model.get_detections(observation, *args, **kwargs)

    def run_depth_refiner(self, observation: ObservationTensor, predictions: PoseEstimatesType) -> Tuple[PoseEstimatesType, dict]:
        assert self.depth_refiner is not None, "You must specify a depth refiner"
        return self.depth_refiner.refine_poses(predictions, depth=observation.depth, K=observation.K)

    @torch.no_grad()
    def run_inference_pipeline(
        self,
        observation: ObservationTensor,
        detections: Optional[DetectionsType] = None,
        run_detector: Optional[bool] = None,
        n_refiner_iterations: int = 5,
        n_pose_hypotheses: int = 1,
        keep_all_refiner_outputs: bool = False,
        run_depth_refiner: bool = False,
        bsz_images: Optional[int] = None,
        bsz_objects: Optional[int] = None,
        cuda_timer: Optional[bool] = False,
        coarse_estimates: Optional[PoseEstimatesType] = None,
        labels_to_keep: Optional[List[str]] = None,
    ) -> Tuple[PoseEstimatesType, dict]:
        """detector (optional) -> coarse -> top-K -> refiner -> scoring -> top-1 [-> depth refiner]
        (pose_estimator.py:516-668)."""
        timing_str = ""
        timer = SimpleTimer()
        timer.start()
        if bsz_images is not None:
            self.bsz_images = bsz_images
        if bsz_objects is not None:
            self.bsz_objects = bsz_objects
        group_cols = ["batch_im_id", "label", "instance_id"]

        if coarse_estimates is None:
            assert detections is not None or run_detector, "You must either pass in `detections` or set run_detector=True"
            if detections is None and run_detector:
                start_time = time.time()
                detections = self.forward_detection_model(observation)
                detections = detections.to(observation.images.device)
                timing_str += f"detection={time.time() - start_time:.2f}, "
            assert detections is not None
            if labels_to_keep is not None:
                detections = filter_detections(detections, labels_to_keep)
            assert len(detections) > 0, "TOFIX: currently, dealing with absence of detections is not supported"
            self._check_sharded_inputs(observation, detections)
            # detections = add_instance_id(detections) (pose_estimator.py:578) happens inside, behind the coarse enqueue
            data_TCO_coarse, coarse_extra_data = self.forward_coarse_model(
                observation=observation, detections=detections, cuda_timer=cuda_timer, _instance_ids_after_enqueue=True)
            timing_str += f"coarse={coarse_extra_data['time']:.2f}, "
            data_TCO_filtered = filter_top_pose_estimates(
                data_TCO_coarse, top_K=n_pose_hypotheses, group_cols=group_cols, filter_field="coarse_logit",
                scores_device=coarse_extra_data["logits"])
        else:
            data_TCO_coarse = coarse_estimates
            coarse_extra_data = None
            data_TCO_filtered = coarse_estimates

        preds, refiner_extra_data = self.forward_refiner(
            observation, data_TCO_filtered, n_iterations=n_refiner_iterations,
            keep_all_outputs=keep_all_refiner_outputs, cuda_timer=cuda_timer)
        data_TCO_refined = preds[f"iteration={n_refiner_iterations}"]
        timing_str += f"refiner={refiner_extra_data['time']:.2f}, "

        data_TCO_scored, scoring_extra_data = self.forward_scoring_model(observation, data_TCO_refined, cuda_timer=cuda_timer)
        timing_str += f"scoring={scoring_extra_data['time']:.2f}, "

        data_TCO_final_scored = filter_top_pose_estimates(
            data_TCO_scored, top_K=1, group_cols=group_cols, filter_field="pose_logit",
            scores_device=scoring_extra_data["logits"])

        if run_depth_refiner:
            depth_refiner_start = time.time()
            data_TCO_depth_refiner, _ = self.run_depth_refiner(observation, data_TCO_final_scored)
            data_TCO_final = data_TCO_depth_refiner
            timing_str += f"depth refiner={time.time() - depth_refiner_start:.2f}"
        else:
            data_TCO_depth_refiner = None
            data_TCO_final = data_TCO_final_scored

        timer.stop()
        timing_str = f"total={timer.elapsed():.2f}, {timing_str}"
        extra_data: dict = {}
        extra_data["coarse"] = {"preds": data_TCO_coarse, "data": coarse_extra_data}
        extra_data["coarse_filter"] = {"preds": data_TCO_filtered}
        extra_data["refiner_all_hypotheses"] = {"preds": preds, "data": refiner_extra_data}
        extra_data["scoring"] = {"preds": data_TCO_scored, "data": scoring_extra_data}
        extra_data["refiner"] = {"preds": data_TCO_final_scored, "data": refiner_extra_data}
        extra_data["timing_str"] = timing_str
        extra_data["time"] = timer.elapsed()
        if run_depth_refiner:
            extra_data["depth_refiner"] = {"preds": data_TCO_depth_refiner}
        return data_TCO_final, extra_data
