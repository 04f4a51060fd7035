"""Model construction (happypose/pose_estimators/megapose/training/pose_models_cfg.py:88-142) and the two MegaPose
operating points (megapose/scripts/run_megapose_training.py:126-149: coarse = 1 rendered view + logits head,
refiner = 4 rendered views "TCO+front_3views" + pose head)."""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Optional, Tuple

import torch

from ..datasets.object_dataset import RigidObjectDataset
from ..lib3d.rigid_mesh_database import BatchedMeshes, MeshDataBase
from ..renderer.panda3d_batch_renderer import Panda3dBatchRenderer
from .backbones import make_backbone
from .pose_rigid import PosePredictor


@dataclass
class PoseModelConfig:
    """The model-shape knobs of TrainingConfig that change the hot path (training_config.py:43-145)."""

    backbone_str: str = "vanilla_resnet34"
    n_rendered_views: int = 1
    multiview_type: str = "TCO+front_3views"
    views_inplane_rotations: bool = False
    remove_TCO_rendering: bool = False
    predict_pose_update: bool = True
    predict_rendered_views_logits: bool = False
    render_normals: bool = True
    render_depth: bool = False
    input_depth: bool = False
    depth_normalization_type: str = "tCR_scale_clamp_center"


COARSE_RGB = PoseModelConfig(n_rendered_views=1, predict_pose_update=False, predict_rendered_views_logits=True)
REFINER_RGB = PoseModelConfig(n_rendered_views=4, multiview_type="TCO+front_3views")
COARSE_RGBD = replace(COARSE_RGB, render_depth=True, input_depth=True)
REFINER_RGBD = replace(REFINER_RGB, render_depth=True, input_depth=True)


def n_network_inputs(cfg: PoseModelConfig) -> int:
    """pose_models_cfg.py:93-103."""
    n_render = 3 + (3 if cfg.render_normals else 0) + (1 if cfg.render_depth else 0)
    return 3 + (1 if cfg.input_depth else 0) + n_render * cfg.n_rendered_views


def create_model_pose(cfg: PoseModelConfig, renderer: Panda3dBatchRenderer, mesh_db: BatchedMeshes) -> PosePredictor:
    backbone = make_backbone(cfg.backbone_str, n_network_inputs(cfg))
    model = PosePredictor(
        backbone=backbone,
        renderer=renderer,
        mesh_db=mesh_db,
        render_size=(240, 320),
        n_rendered_views=cfg.n_rendered_views,
        views_inplane_rotations=cfg.views_inplane_rotations,
        multiview_type=cfg.multiview_type,
        render_normals=cfg.render_normals,
        render_depth=cfg.render_depth,
        input_depth=cfg.input_depth,
        predict_rendered_views_logits=cfg.predict_rendered_views_logits,
        remove_TCO_rendering=cfg.remove_TCO_rendering,
        predict_pose_update=cfg.predict_pose_update,
        depth_normalization_type=cfg.depth_normalization_type,
    )
    model.cfg = cfg
    return model


def make_pose_models(
    object_dataset: RigidObjectDataset,
    coarse_cfg: PoseModelConfig = COARSE_RGB,
    refiner_cfg: PoseModelConfig = REFINER_RGB,
    device: Optional[torch.device] = None,
    seed: int = 0,
) -> Tuple[PosePredictor, PosePredictor, BatchedMeshes]:
    """load_pose_models (toolbox/inference/utils.py:84-161) without checkpoints: renderer + mesh database + the
    coarse and refiner PosePredictors with seeded random-init weights, in eval mode on `device`."""
    device = torch.device("cuda" if device is None else device)
    renderer = Panda3dBatchRenderer(object_dataset, n_workers=1, preload_cache=True, device=device)
    mesh_db = MeshDataBase.from_object_ds(object_dataset).batched().to(device)
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    coarse = create_model_pose(coarse_cfg, renderer, mesh_db).to(device).eval()
    refiner = create_model_pose(refiner_cfg, renderer, mesh_db).to(device).eval()
    torch.random.set_rng_state(gen_state)
    return coarse, refiner, mesh_db
