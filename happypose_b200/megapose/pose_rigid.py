"""MegaPose PosePredictor re-hosted on the B200 kernels.

Mirror of happypose/pose_estimators/megapose/models/pose_rigid.py:96-788: same constructor, same methods
(forward, forward_coarse, forward_coarse_tensor, crop_inputs, compute_crops_multiview, render_images_multiview,
normalize_images, normalize_depth, update_pose, net_forward), same output dataclasses and dict keys.  What changed
is the data flow per iteration:

  reference                                                  here
  ---------------------------------------------------------  ----------------------------------------------------
  normalize_T: ~12 eager ops                                 hpb_normalize_T (1 launch)
  make_TCO_multiview: per-sample CPU loop over Panda3D nodes hpb_multiview (1 launch, poses stay on the device)
  crop_inputs: 2 projections, ~40 eager ops, roi_align       hpb_crop (2 launches) writing x[:, :C] in place
  on images[batch_im_ids] (b copies of the frame)            frame indexed by im_id, never expanded
  compute_crops_multiview                                    hpb_crop_boxes (1 launch)
  renderer.render: b Panda3D frames in CPU workers + H2D     hpb_render (1 launch) writing x[:, C:] in place
  normalize_images: 2 full-tensor clones (+ depth ops)       hpb_normalize_depth in place, only for RGB-D models
  torch.cat((images_crop, renders))                          none: crop and renders already share x
  ResNet in fp32                                             same torch module, bf16 + channels_last
  update_pose: ~30 eager ops                                 hpb_pose_update (1 launch)

The only extension to the signatures is the optional `im_ids` argument: when given, `images` / `K` hold the
distinct frames ([n_im, C, H, W] / [n_im, 3, 3]) and im_ids[b] selects the frame of each row.
"""
from __future__ import annotations

import time
from collections import defaultdict
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import numpy as np
import torch
from torch import nn

from .. import _capi, ops
from . import fast_resnet
from .._capi import Context
from ..lib3d.rigid_mesh_database import BatchedMeshes
from ..renderer.panda3d_batch_renderer import Panda3dBatchRenderer
from ..renderer.types import Panda3dLightData, Resolution
from ..utils.cuda_graphs import GraphCache
from ..utils.timer import CudaTimer, SimpleTimer


@dataclass
class PosePredictorOutputCosypose:
    TCO_output: torch.Tensor
    TCO_input: torch.Tensor
    renders: torch.Tensor
    images_crop: torch.Tensor
    labels: List[str]
    K: torch.Tensor
    K_crop: torch.Tensor
    boxes_rend: torch.Tensor
    boxes_crop: torch.Tensor
    model_outputs: torch.Tensor


@dataclass
class PosePredictorOutput:
    TCO_output: torch.Tensor
    TCO_input: torch.Tensor
    renders: torch.Tensor
    images_crop: torch.Tensor
    TCV_O_input: torch.Tensor
    KV_crop: torch.Tensor
    tCR: torch.Tensor
    labels: List[str]
    K: torch.Tensor
    K_crop: torch.Tensor
    network_outputs: Dict[str, torch.Tensor]
    boxes_rend: torch.Tensor
    boxes_crop: torch.Tensor
    renderings_logits: torch.Tensor
    timing_dict: Dict[str, float]


@dataclass
class PosePredictorDebugData:
    output: Optional[PosePredictorOutput] = None
    images: Optional[torch.Tensor] = None
    origin_uv: Optional[torch.Tensor] = None
    ref_point_uv: Optional[torch.Tensor] = None
    origin_uv_crop: Optional[torch.Tensor] = None
    pose_predictor_outputs: Optional[torch.Tensor] = None


class PosePredictor(nn.Module):
    def __init__(
        self,
        backbone: torch.nn.Module,
        renderer: Panda3dBatchRenderer,
        mesh_db: BatchedMeshes,
        render_size: Resolution = (240, 320),
        multiview_type: str = "front_3views",
        views_inplane_rotations: bool = False,
        remove_TCO_rendering: bool = False,
        predict_pose_update: bool = True,
        predict_rendered_views_logits: bool = False,
        render_normals: bool = True,
        n_rendered_views: int = 1,
        input_depth: bool = False,
        render_depth: bool = False,
        depth_normalization_type: Optional[str] = None,
        compute_dtype: torch.dtype = torch.bfloat16,
    ):
        super().__init__()
        self.backbone = backbone
        self.renderer = renderer
        self.render_size = render_size
        self.n_rendered_views = n_rendered_views
        self.input_depth = input_depth
        self.multiview_type = multiview_type
        # Stored, and -- exactly like the reference -- not used by forward(): pose_rigid.py:578-584 calls make_TCO_multiview
        # without views_inplane_rotations; the only caller that turns it on is the training loss
        # (megapose/training/megapose_forward_loss.py:116-123), through lib3d.multiview.make_TCO_multiview, which supports it.
        self.views_inplane_rotations = views_inplane_rotations
        self.render_normals = render_normals
        self.render_depth = render_depth
        self.depth_normalization_type = depth_normalization_type
        self.remove_TCO_rendering = remove_TCO_rendering
        self.mesh_db = mesh_db
        self.compute_dtype = compute_dtype

        n_features = backbone.n_features
        assert isinstance(n_features, int)
        self.heads: Dict[str, Union[torch.nn.Linear, Callable]] = {}
        self.predict_pose_update = predict_pose_update
        if self.predict_pose_update:
            self._pose_dim = 9
            self.pose_fc = nn.Linear(n_features, self._pose_dim, bias=True)
            self.heads["pose"] = self.pose_fc
        self.predict_rendered_views_logits = predict_rendered_views_logits
        if self.predict_rendered_views_logits:
            self.views_logits_head = nn.Linear(n_features, self.n_rendered_views, bias=True)
            self.heads["renderings_logits"] = self.views_logits_head

        # channel bookkeeping (pose_rigid.py:151-180)
        self._input_rgb_dims = [0, 1, 2]
        self._input_depth_dims = [3] if self.input_depth else []
        self._render_rgb_dims = [0, 1, 2]
        self._render_normal_dims = [3, 4, 5] if self.render_normals else []
        self._render_depth_dims = [3 + len(self._render_normal_dims)] if self.render_depth else []
        self._n_single_render_channels = 3 + len(self._render_normal_dims) + len(self._render_depth_dims)

        self.debug = False
        self.timing_dict: Dict[str, float] = defaultdict(float)
        self.debug_data = PosePredictorDebugData()
        self._net_ready = False
        self._folded = None
        self._crop_tap_bits: Optional[int] = None  # None = follow compute_dtype (see crop_tap_bits)
        self.use_direct_s2d = True  # forward_coarse: rasteriser writes the stem's bf16 input directly (see _direct_s2d_ok)
        # replay launch-bound batches as CUDA graphs (utils/cuda_graphs.py); off by default, PoseEstimator turns it on
        self.use_cuda_graphs = False
        self.graph_max_batch = 64
        self._graphs = GraphCache(lambda: self.renderer._ctx.workspace_epoch())

    # ---- properties of the reference -----------------------------------------------------------
    @property
    def input_rgb_dims(self) -> List[int]:
        return self._input_rgb_dims

    @property
    def input_depth_dims(self) -> List[int]:
        return self._input_depth_dims

    @property
    def render_rgb_dims(self) -> List[int]:
        return self._render_rgb_dims

    @property
    def render_depth_dims(self) -> List[int]:
        return self._render_depth_dims

    @property
    def n_input_channels(self) -> int:
        return 3 + len(self._input_depth_dims)

    # ---- helpers -------------------------------------------------------------------------------
    def _ctx(self) -> Context:
        return self.renderer._ctx

    @property
    def crop_tap_bits(self) -> int:
        """Precision of the frame samples of hpb_crop: float32 (the reference's arithmetic) when the network computes in
        float32; fp16 taps of RGB frames (|error| <= 2.5e-4 on [0,1], BASELINE bar 1e-3) when the network input is
        rounded to bf16 / fp16 anyway -- the crop kernel is bound by L1 sector look-ups, 8-byte taps halve them."""
        if self._crop_tap_bits is not None:
            return self._crop_tap_bits
        return 16 if self.compute_dtype in (torch.bfloat16, torch.float16) else 32

    @crop_tap_bits.setter
    def crop_tap_bits(self, bits: Optional[int]) -> None:
        assert bits in (None, 16, 32)
        self._crop_tap_bits = bits

    def _ids(self, labels: List[str], im_ids, bsz: int, device):
        obj_ids = self.mesh_db.label_ids(labels, device)
        mesh_ids = self.renderer.mesh_ids(labels)
        if im_ids is None:
            im_ids = torch.arange(bsz, dtype=torch.int32, device=device)
        return obj_ids, mesh_ids, torch.as_tensor(im_ids).to(device=device, dtype=torch.int32)

    @staticmethod
    def _row_K(K: torch.Tensor, im_ids: Optional[torch.Tensor], bsz: int) -> torch.Tensor:
        if K.shape[0] == bsz and im_ids is None:
            return K
        return K[torch.as_tensor(im_ids).to(K.device).long()]

    def _graph_flags(self) -> Tuple:
        """Model switches that change what a captured graph contains (part of the graph key)."""
        return (self.crop_tap_bits, str(self.compute_dtype), bool(self.use_direct_s2d), self.multiview_type,
                bool(self.views_inplane_rotations), bool(self.remove_TCO_rendering))

    def _direct_s2d_ok(self, images: torch.Tensor, return_debug_data: bool, cuda_timer: bool) -> bool:
        """The coarse / scoring forward may skip the float32 network input when nobody asks for it (debug data) and the
        network is the folded bf16 ResNet with the space-to-depth stem: RGB frame, rgb + normals renders, one view."""
        if return_debug_data or cuda_timer or self.debug or not images.is_cuda or not self.use_direct_s2d:
            return False
        if (self.input_depth or self.render_depth or not self.render_normals or self.n_rendered_views != 1
                or images.shape[1] != 3 or self.n_input_channels != 3 or self._n_single_render_channels != 6):
            return False
        h, w = self.render_size
        if h % 2 or w % 2:
            return False
        self._prepare_net(images)
        return self._folded is not None and self._folded.accepts_s2d and self._folded.s2d_channels >= 64

    def _s2d_buffer(self, bsz: int, device) -> torch.Tensor:
        """Persistent, zero-initialised stem-input buffer of the fused hand-off: only the rasteriser writes it (the first 48
        channels of every cell), so the zero padding channels never have to be rewritten.  One buffer per model, grown to
        the largest batch seen; smaller batches use its leading rows.  (Inside a captured CUDA graph the buffer was created
        by the eager warm-up, so the graph only holds its address.)"""
        h, w = self.render_size
        buf = getattr(self, "_s2d_buf", None)
        if buf is None or buf.shape[0] < bsz or buf.device != torch.device(device):
            buf = torch.empty((bsz, self._folded.s2d_channels, h // 2 + 3, w // 2 + 3), dtype=torch.bfloat16, device=device,
                              memory_format=torch.channels_last).zero_()
            self._s2d_buf = buf
        return buf[:bsz]

    def _alloc_input(self, bsz: int, device) -> torch.Tensor:
        C = self.n_input_channels + self._n_single_render_channels * self.n_rendered_views
        h, w = self.render_size
        return torch.empty((bsz, C, h, w), dtype=torch.float32, device=device)

    # ---- crop ----------------------------------------------------------------------------------
    def crop_inputs(self, images, K, TCO, tCR, labels, im_ids=None, out=None):
        """pose_rigid.py:199-277 -> (images_cropped, K_crop, boxes_rend, boxes_crop)."""
        bsz = TCO.shape[0]
        assert K.shape == (bsz, 3, 3)
        assert tCR.shape == (bsz, 3)
        assert TCO.shape == (bsz, 4, 4)
        assert len(labels) == bsz
        obj_ids, _, im_ids_t = self._ids(labels, im_ids, bsz, TCO.device)
        crops, K_crop, boxes_rend, boxes_crop = ops.crop(
            self._ctx(), images, im_ids_t, self.mesh_db.points_subset(2000), obj_ids, K, TCO, tCR,
            self.render_size, lamb=1.4, out=out, tap_bits=self.crop_tap_bits)
        return crops, K_crop, boxes_rend, boxes_crop

    def compute_crops_multiview(self, images, K, TCV_O, tCR, labels) -> torch.Tensor:
        """pose_rigid.py:279-337: K_crop of the extra views (200 points, no pixels resampled)."""
        bsz = len(labels)
        n_views = TCV_O.shape[1]
        assert tCR.shape == (bsz, n_views, 3)
        assert TCV_O.shape == (bsz, n_views, 4, 4)
        assert K.shape == (bsz, 3, 3)
        obj_ids = self.mesh_db.label_ids(labels, TCV_O.device).repeat_interleave(n_views)
        Kmv = K.unsqueeze(1).expand(bsz, n_views, 3, 3).reshape(-1, 3, 3)
        K_crop, _, _ = ops.crop_boxes(
            self._ctx(), images.shape[-2:], self.mesh_db.points_subset(200), obj_ids, Kmv, TCV_O.flatten(0, 1),
            tCR.flatten(0, 1), self.render_size, lamb=1.4)
        return K_crop.view(bsz, n_views, 3, 3)

    # ---- pose update / network -----------------------------------------------------------------
    def update_pose(self, TCO, K_crop, pose_outputs, tCR) -> torch.Tensor:
        assert pose_outputs.shape[-1] == 9
        return ops.pose_update(self._ctx(), TCO, K_crop, pose_outputs, tCR, _capi.POSE_MEGAPOSE)

    def _prepare_net(self, x: torch.Tensor) -> None:
        """Once, on first use: choose how the (unchanged) torch network is executed.  On CUDA with a reduced-precision
        compute dtype a torchvision-style ResNet is run through fast_resnet.FoldedResNet (batch-norm folded, fused
        cuDNN conv+bias+ReLU epilogues, channels_last); anything else runs the module as is (under autocast for a
        reduced-precision compute dtype -- the user's module is never cast in place).  The Linear heads always stay
        float32: the pose update multiplies depth by vz ~ 1.0, where bf16 spacing is 2^-8, and coarse logits near 13
        would tie at 0.0625 steps."""
        if self._net_ready:
            if self._folded is None or not self._folded.stale():
                return
            self._graphs.clear()  # graphs captured with the old weight snapshot
        self._folded = None
        if x.is_cuda and self.compute_dtype != torch.float32:
            self._folded = fast_resnet.try_fold(self.backbone, self.compute_dtype, self._ctx())
        self._net_ready = True

    def refold(self) -> None:
        """Re-snapshot the backbone (after load_state_dict / fine-tuning); also drops captured CUDA graphs."""
        self._net_ready = False
        self._folded = None
        self._graphs.clear()

    def _heads_forward(self, feat: torch.Tensor) -> Dict[str, torch.Tensor]:
        feat = feat.float()
        return {k: head(feat) for k, head in self.heads.items()}

    def net_forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """pose_rigid.py:352-374.  The torch backbone runs in bf16/channels_last; features and heads are float32."""
        self._prepare_net(x)
        if self._folded is not None:
            x = self._folded(x)  # packs the fp32 planar input to bf16 NHWC itself (space-to-depth for the 7x7 stem)
        elif x.is_cuda and self.compute_dtype != torch.float32:
            with torch.autocast("cuda", dtype=self.compute_dtype):
                x = self.backbone(x.contiguous(memory_format=torch.channels_last))
        else:
            x = self.backbone(x)
        if x.dim() == 4:
            x = x.float().flatten(2).mean(dim=-1)
        elif x.dim() != 2:
            raise ValueError
        return self._heads_forward(x)

    # ---- rendering -----------------------------------------------------------------------------
    def render_images_multiview(self, labels, TCV_O, KV, random_ambient_light: bool = False, out=None, mesh_ids=None,
                                mesh_ids_per_view: bool = False):
        """pose_rigid.py:376-453 -> renders [bsz, n_views*n_channels, H, W].
        With `out` (the network input) the renders are written into its channels after the crop.  `mesh_ids` are the
        renderer's ids per row, or (mesh_ids_per_view) already repeated per view."""
        bsz = TCV_O.shape[0]
        n_views = TCV_O.shape[1]
        assert isinstance(self.renderer, Panda3dBatchRenderer)
        device = TCV_O.device
        ambient = lights = None
        if mesh_ids is None:
            mesh_ids = self.renderer.mesh_ids(labels)
        if not mesh_ids_per_view:
            mesh_ids = mesh_ids.repeat_interleave(n_views) if n_views > 1 else mesh_ids
        if random_ambient_light:
            inten = np.random.uniform(0.7, 1.0, size=(bsz * n_views, 1)).astype(np.float32)
            ambient = torch.as_tensor(np.repeat(inten, 3, 1)).to(device)
        elif not self.render_normals:
            # make_scene_lights() (pose_rigid.py:421-422): ambient 0.1 + six point lights 0.4 at +-10 bounding radii
            lights = self.renderer.scene_light_rig(mesh_ids)
            ambient = torch.full((bsz * n_views, 3), 0.1, dtype=torch.float32, device=device)
        C_in = self.n_input_channels
        if out is None:
            out = torch.empty((bsz, C_in + n_views * self._n_single_render_channels) + tuple(self.render_size), dtype=torch.float32, device=device)
        self.renderer.render_into(
            mesh_ids, TCV_O.flatten(0, 1), KV.flatten(0, 1), self.render_size, out, C_in,
            render_normals=self.render_normals, render_depth=self.render_depth, views=n_views, ambient=ambient, lights=lights)
        return out[:, C_in:]

    # ---- depth normalisation -------------------------------------------------------------------
    def normalize_depth(self, depth: torch.Tensor, tCR: torch.Tensor) -> torch.Tensor:
        """pose_rigid.py:510-544 (functional form; the pipeline itself normalises in place on the network input)."""
        z = tCR[:, 2][(...,) + (None,) * (depth.ndim - 1)]
        kind = self.depth_normalization_type
        if kind == "tCR_scale":
            return depth / z
        if kind == "tCR_scale_clamp_center":
            return torch.clamp(depth / z, 0, 2) - 1
        if kind == "tCR_center_clamp":
            return torch.clamp(depth - z, -2, 2)
        if kind == "tCR_center_obj_diam":
            raise NotImplementedError("Not yet implemented")
        if kind == "none":
            return depth
        raise ValueError(f"Unknown depth_normalization_type = {kind}")

    def _depth_channels(self) -> List[int]:
        ch = list(self._input_depth_dims)
        if self.render_depth:
            ch += [self.n_input_channels + self._render_depth_dims[0] + self._n_single_render_channels * v for v in range(self.n_rendered_views)]
        return ch

    def normalize_images(self, images, renders, tCR, images_inplace=False, renders_inplace=False):
        """pose_rigid.py:455-508 (kept for API parity; returns new tensors)."""
        images = images if images_inplace else images.clone()
        renders = renders if renders_inplace else renders.clone()
        if self.input_depth:
            assert images.shape[1] == 4, "images must have C=4 channels if input_depth=True"
            images[:, self._input_depth_dims] = self.normalize_depth(images[:, self._input_depth_dims], tCR)
        if self.render_depth:
            dims = self._render_depth_dims[0] + self._n_single_render_channels * torch.arange(0, self.n_rendered_views)
            renders[:, dims] = self.normalize_depth(renders[:, dims], tCR)
        return images, renders

    def _normalize_input_(self, x: torch.Tensor, tCR: torch.Tensor) -> None:
        ch = self._depth_channels()
        if ch:
            kind = self.depth_normalization_type
            if kind == "tCR_center_obj_diam":
                raise NotImplementedError("Not yet implemented")
            ops.normalize_depth_(self._ctx(), x, ch, tCR, kind)

    # ---- refiner -------------------------------------------------------------------------------
    def forward(self, images, K, labels, TCO, n_iterations=1, random_ambient_light=False, im_ids=None) -> Dict[str, PosePredictorOutput]:
        bsz = TCO.shape[0]
        assert TCO.shape == (bsz, 4, 4)
        assert len(labels) == bsz
        if im_ids is None:
            assert images.shape[0] == bsz
        K = self._row_K(K, im_ids, bsz)
        assert K.shape == (bsz, 3, 3)
        obj_ids, mesh_ids, im_ids_t = self._ids(labels, im_ids, bsz, TCO.device)
        return self.forward_ids(images, K, obj_ids, mesh_ids, im_ids_t, TCO, n_iterations, random_ambient_light, labels=labels)

    def forward_ids(self, images, K, obj_ids, mesh_ids, im_ids, TCO, n_iterations=1, random_ambient_light=False, labels=None):
        """forward() on device-resident ids (no label lookups): obj_ids index mesh_db.points, mesh_ids the renderer's
        meshes, im_ids the frames of `images`; K is per row [bsz,3,3].

        Small batches (the refiner works on a handful of hypotheses, 5 dependent iterations each) are launch-bound:
        with use_cuda_graphs the whole n_iterations loop is captured once per batch shape and replayed."""
        bsz = TCO.shape[0]
        if (self.use_cuda_graphs and TCO.is_cuda and not self.debug and not random_ambient_light and 0 < bsz <= self.graph_max_batch
                and not torch.cuda.is_current_stream_capturing() and not ops.kernel_timer_active()):
            tensors = (images.contiguous(), K.contiguous().float(), obj_ids, mesh_ids, im_ids, TCO.contiguous().float())

            def fn(images_, K_, obj_ids_, mesh_ids_, im_ids_, TCO_):
                return self._forward_ids_impl(images_, K_, obj_ids_, mesh_ids_, im_ids_, TCO_, n_iterations, False, None)

            outputs, replayed = self._graphs.run(("refiner", n_iterations) + self._graph_flags(), fn, tensors)
            if replayed:  # static graph buffers: hand out copies of the small tensors, views of the big ones
                outputs = {k: self._detach_output(o, labels) for k, o in outputs.items()}
            return outputs
        return self._forward_ids_impl(images, K, obj_ids, mesh_ids, im_ids, TCO, n_iterations, random_ambient_light, labels)

    @staticmethod
    def _detach_output(o: "PosePredictorOutput", labels) -> "PosePredictorOutput":
        c = lambda t: t.clone()  # noqa: E731
        return PosePredictorOutput(
            renders=o.renders, images_crop=o.images_crop,  # views of graph memory, valid until the next replay
            TCO_input=c(o.TCO_input), TCO_output=c(o.TCO_output), TCV_O_input=c(o.TCV_O_input), tCR=c(o.tCR), labels=labels,
            K=c(o.K), K_crop=c(o.K_crop), KV_crop=c(o.KV_crop), network_outputs={k: c(v) for k, v in o.network_outputs.items()},
            boxes_rend=c(o.boxes_rend), boxes_crop=c(o.boxes_crop), renderings_logits=c(o.renderings_logits), timing_dict=o.timing_dict)

    def _forward_ids_impl(self, images, K, obj_ids, mesh_ids, im_ids, TCO, n_iterations, random_ambient_light, labels):
        timing_dict: Dict[str, float] = defaultdict(float)
        if not self.input_depth:
            images = images[:, :3]  # input_rgb_dims = [0, 1, 2]; a slice, not an index tensor (no copy, graph-capturable)
        bsz = TCO.shape[0]
        dtype, device = TCO.dtype, TCO.device
        ctx = self._ctx()
        n_views = self.n_rendered_views
        pts2000 = self.mesh_db.points_subset(2000)

        # loop invariant of the iterations: the renderer's mesh id of every (row, view)
        mesh_ids_v = mesh_ids.repeat_interleave(n_views) if (mesh_ids is not None and n_views > 1) else mesh_ids
        pts200 = self.mesh_db.points_subset(200)
        K = K.contiguous().float()

        outputs = {}
        TCO_input = TCO
        for n in range(n_iterations):
            # ONE launch: normalize_T, tCR (reference point = object origin, tOR = 0, pose_rigid.py:574-576), the extra views,
            # the row's crop geometry (2000 points) and the per-view K_crop (200 points), KV_crop[:, 0] = K_crop
            pro = ops.refiner_prologue(ctx, TCO_input, K, obj_ids, pts2000, pts200, images.shape[-2:], self.render_size,
                                       self.multiview_type, n_views, self.remove_TCO_rendering)
            TCO_input, tCR, TCV_O_input = pro["T_norm"], pro["tCR"], pro["TCV_O"]
            K_crop, boxes_rend, boxes_crop, KV_crop = pro["K_crop"], pro["boxes_rend"], pro["boxes_crop"], pro["KV_crop"]

            x = self._alloc_input(bsz, device)
            images_crop = ops.crop_pixels(ctx, images, im_ids, boxes_crop, self.render_size, out=x, tap_bits=self.crop_tap_bits)

            t = time.time()
            renders = self.render_images_multiview(labels, TCV_O_input, KV_crop, random_ambient_light, out=x, mesh_ids=mesh_ids_v,
                                                   mesh_ids_per_view=mesh_ids_v is not None)
            timing_dict["render"] = time.time() - t

            self._normalize_input_(x, tCR)
            network_outputs = self.net_forward(x)
            if self.predict_pose_update:
                TCO_output = self.update_pose(TCO_input, K_crop, network_outputs["pose"], tCR)
            else:
                TCO_output = TCO_input.detach().clone()
            if self.predict_rendered_views_logits:
                renderings_logits = network_outputs["renderings_logits"]
                assert not self.predict_pose_update
            else:
                renderings_logits = torch.empty(bsz, self.n_rendered_views, dtype=dtype, device=device)

            outputs[f"iteration={n+1}"] = PosePredictorOutput(
                renders=renders, images_crop=images_crop, TCO_input=TCO_input, TCO_output=TCO_output,
                TCV_O_input=TCV_O_input, tCR=tCR, labels=labels, K=K, K_crop=K_crop, KV_crop=KV_crop,
                network_outputs=network_outputs, boxes_rend=boxes_rend, boxes_crop=boxes_crop,
                renderings_logits=renderings_logits, timing_dict=timing_dict)
            if self.debug:
                self.debug_data.output = outputs[f"iteration={n+1}"]
            TCO_input = TCO_output
        return outputs

    # ---- coarse --------------------------------------------------------------------------------
    def forward_coarse_tensor(self, x: torch.Tensor, cuda_timer: bool = False) -> Dict[str, Union[torch.Tensor, float]]:
        """pose_rigid.py:676-706."""
        assert self.predict_rendered_views_logits, "Method only valid if coarse classification model"
        timer = CudaTimer(enabled=cuda_timer) if torch.cuda.is_available() else SimpleTimer()
        timer.start()
        logits = self.net_forward(x)["renderings_logits"]
        scores = torch.sigmoid(logits)
        timer.end()
        return {"logits": logits, "scores": scores, "time": timer.elapsed()}

    def forward_coarse(self, images, K, labels, TCO_input, cuda_timer=False, return_debug_data=False, im_ids=None) -> Dict[str, Any]:
        """pose_rigid.py:708-788: crop + render + score for a batch of (label, pose) hypotheses."""
        assert self.predict_rendered_views_logits, "Method only valid if coarse classification model"
        bsz = TCO_input.shape[0]
        assert TCO_input.shape == (bsz, 4, 4)
        assert len(labels) == bsz
        if im_ids is None:
            assert images.shape[0] == bsz
        K = self._row_K(K, im_ids, bsz)
        assert K.shape == (bsz, 3, 3)
        obj_ids, mesh_ids, im_ids_t = self._ids(labels, im_ids, bsz, TCO_input.device)
        return self.forward_coarse_ids(images, K, obj_ids, mesh_ids, im_ids_t, TCO_input, cuda_timer, return_debug_data)

    def forward_coarse_ids(self, images, K, obj_ids, mesh_ids, im_ids, TCO_input, cuda_timer=False, return_debug_data=False):
        """forward_coarse() on device-resident ids (see forward_ids)."""
        assert self.predict_rendered_views_logits, "Method only valid if coarse classification model"
        bsz = TCO_input.shape[0]
        if (self.use_cuda_graphs and TCO_input.is_cuda and not cuda_timer and not return_debug_data and not self.debug
                and 0 < bsz <= self.graph_max_batch and not torch.cuda.is_current_stream_capturing() and not ops.kernel_timer_active()):
            tensors = (images.contiguous(), K.contiguous().float(), obj_ids, mesh_ids, im_ids, TCO_input.contiguous().float())

            def fn(images_, K_, obj_ids_, mesh_ids_, im_ids_, TCO_):
                o = self._forward_coarse_ids_impl(images_, K_, obj_ids_, mesh_ids_, im_ids_, TCO_, False, False)
                return o["logits"], o["scores"]

            (logits, scores), replayed = self._graphs.run(("coarse",) + self._graph_flags(), fn, tensors)
            if replayed:
                logits, scores = logits.clone(), scores.clone()
            return {"logits": logits, "scores": scores, "time": 0.0, "render_time": 0.0, "model_time": 0.0}
        return self._forward_coarse_ids_impl(images, K, obj_ids, mesh_ids, im_ids, TCO_input, cuda_timer, return_debug_data)

    def _forward_coarse_ids_impl(self, images, K, obj_ids, mesh_ids, im_ids, TCO_input, cuda_timer, return_debug_data):
        if not self.input_depth:
            images = images[:, :3]  # input_rgb_dims = [0, 1, 2]; a slice, not an index tensor (no copy, graph-capturable)
        bsz = TCO_input.shape[0]
        ctx = self._ctx()
        device = TCO_input.device

        TCO_input = ops.normalize_T(ctx, TCO_input).detach()
        tCR = TCO_input[..., :3, 3].contiguous()
        if self._direct_s2d_ok(images, return_debug_data, cuda_timer):
            # fused hand-off: the rasteriser's resolve writes the stem's bf16 space-to-depth input itself (crop channels
            # read from the crop kernel's planes): no float32 [b,9,h,w] network input, no packing pass
            # ... and the crop travels as bf16 pixels (8 B instead of 12 B per pixel, one load in the resolve)
            crops, K_crop, _, _ = ops.crop_bf16x4(
                ctx, images, im_ids, self.mesh_db.points_subset(2000), obj_ids, K, TCO_input, tCR, self.render_size,
                tap_bits=self.crop_tap_bits)
            render_start = time.time()
            z = ops.render_s2d_bf16(ctx, mesh_ids, TCO_input, K_crop, crops, self._folded.s2d_channels,
                                    out=self._s2d_buffer(bsz, device), pad_prezeroed=True)
            render_time = time.time() - render_start
            start = time.time()
            feat = self._folded(z, packed_s2d=True)
            logits = self._heads_forward(feat)["renderings_logits"]
            out = {"logits": logits, "scores": torch.sigmoid(logits), "time": time.time() - start}
            out["render_time"] = render_time
            out["model_time"] = out["time"]
            return out
        x = self._alloc_input(bsz, device)
        images_crop, K_crop, boxes_rend, boxes_crop = ops.crop(
            ctx, images, im_ids, self.mesh_db.points_subset(2000), obj_ids, K, TCO_input, tCR, self.render_size, out=x,
            tap_bits=self.crop_tap_bits)

        render_timer = CudaTimer(enabled=cuda_timer) if torch.cuda.is_available() else SimpleTimer()
        render_start = time.time()
        render_timer.start()
        renders = self.render_images_multiview(None, TCO_input.unsqueeze(1), K_crop.unsqueeze(1), out=x, mesh_ids=mesh_ids)
        render_timer.end()
        render_time = render_timer.elapsed() if cuda_timer else time.time() - render_start

        self._normalize_input_(x, tCR)
        out = self.forward_coarse_tensor(x, cuda_timer=cuda_timer)
        out["render_time"] = render_time
        out["model_time"] = out["time"]
        if return_debug_data:
            out["images_crop"] = images_crop
            out["renders"] = renders
        return out
