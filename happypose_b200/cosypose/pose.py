"""CosyPose PosePredictor re-hosted on the B200 kernels (mirror of
happypose/pose_estimators/cosypose/cosypose/models/pose.py:33-199): single rendered view, RGB only (6-channel
network input), no normalize_T, reference point = object origin, 6-D or quaternion rotation output
(apply_imagespace_predictions, cosypose/lib3d/cosypose_ops.py:18-42).  Same crop / render kernels as MegaPose."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import nn

from .. import _capi, ops
from ..megapose.pose_rigid import PosePredictorOutputCosypose
from ..renderer.panda3d_batch_renderer import Panda3dBatchRenderer


class PosePredictor(nn.Module):
    def __init__(self, backbone, renderer, mesh_db, render_size=(240, 320), pose_dim=9, compute_dtype=torch.bfloat16):
        super().__init__()
        self.backbone = backbone
        self.renderer = renderer
        self.mesh_db = mesh_db
        self.render_size = render_size
        self.pose_dim = pose_dim
        self.compute_dtype = compute_dtype
        self.heads = {}
        self.pose_fc = nn.Linear(backbone.n_features, pose_dim, bias=True)
        self.heads["pose"] = self.pose_fc
        self.debug = False
        self.tmp_debug = {}
        self._net_ready = False

    def enable_debug(self):
        self.debug = True

    def disable_debug(self):
        self.debug = False

    def _ctx(self):
        return self.renderer._ctx

    def crop_inputs(self, images, K, TCO, labels, im_ids=None, out=None):
        """pose.py:58-93: crop centred on the projection of the object origin."""
        bsz = TCO.shape[0]
        assert K.shape == (bsz, 3, 3) and TCO.shape == (bsz, 4, 4) and len(labels) == bsz
        dev = TCO.device
        obj_ids = self.mesh_db.label_ids(labels, dev)
        if im_ids is None:
            im_ids = torch.arange(bsz, dtype=torch.int32, device=dev)
        crops, K_crop, boxes_rend, boxes_crop = ops.crop(
            self._ctx(), images, im_ids, self.mesh_db.points_subset(2000), obj_ids, K, TCO, TCO[:, :3, 3].contiguous(),
            self.render_size, lamb=1.4, out=out)
        return crops, K_crop.detach(), boxes_rend, boxes_crop

    def update_pose(self, TCO, K_crop, pose_outputs):
        if self.pose_dim == 9:
            variant = _capi.POSE_COSYPOSE_6D
        elif self.pose_dim == 7:
            variant = _capi.POSE_COSYPOSE_QUAT
        else:
            raise ValueError(f"pose_dim={self.pose_dim} not supported")
        return ops.pose_update(self._ctx(), TCO, K_crop, pose_outputs, None, variant)

    def net_forward(self, x):
        """pose.py:108-114.  Reduced-precision compute runs the unchanged module under autocast (never cast in place);
        the pooled features and the pose head stay float32."""
        if x.is_cuda and self.compute_dtype != torch.float32:
            with torch.autocast("cuda", dtype=self.compute_dtype):
                f = self.backbone(x.contiguous(memory_format=torch.channels_last))
        else:
            f = self.backbone(x)
        f = f.float().flatten(2).mean(dim=-1)
        return {k: head(f) for k, head in self.heads.items()}

    def forward(self, images, K, labels, TCO, n_iterations=1, im_ids=None) -> Dict[str, PosePredictorOutputCosypose]:
        bsz = TCO.shape[0]
        assert TCO.shape == (bsz, 4, 4) and len(labels) == bsz
        if im_ids is None:
            assert images.shape[0] == bsz
        else:
            K = K[torch.as_tensor(im_ids).to(K.device).long()] if K.shape[0] != bsz else K
        assert K.shape == (bsz, 3, 3)
        dev = TCO.device
        assert isinstance(self.renderer, Panda3dBatchRenderer), f"Renderer of type {type(self.renderer)} not supported"
        obj_ids = self.mesh_db.label_ids(labels, dev)
        mesh_ids = self.renderer.mesh_ids(labels)
        im_ids_t = torch.arange(bsz, dtype=torch.int32, device=dev) if im_ids is None else torch.as_tensor(im_ids).to(dev, torch.int32)
        images = images[:, :3]
        h, w = self.render_size
        pts = self.mesh_db.points_subset(2000)
        outputs = {}
        TCO_input = TCO
        for n in range(n_iterations):
            TCO_input = TCO_input.detach().float().contiguous()
            x = torch.empty((bsz, 6, h, w), dtype=torch.float32, device=dev)
            images_crop, K_crop, boxes_rend, boxes_crop = ops.crop(
                self._ctx(), images, im_ids_t, pts, obj_ids, K, TCO_input, TCO_input[:, :3, 3].contiguous(), self.render_size, out=x)
            self.renderer.render_into(mesh_ids, TCO_input, K_crop, self.render_size, x, 3, render_normals=False, render_depth=False)
            model_outputs = self.net_forward(x)
            TCO_output = self.update_pose(TCO_input, K_crop, model_outputs["pose"])
            outputs[f"iteration={n+1}"] = PosePredictorOutputCosypose(
                renders=x[:, 3:6], images_crop=images_crop, TCO_input=TCO_input, TCO_output=TCO_output, labels=labels,
                K=K, K_crop=K_crop, boxes_rend=boxes_rend, boxes_crop=boxes_crop, model_outputs=model_outputs)
            TCO_input = TCO_output
        return outputs
