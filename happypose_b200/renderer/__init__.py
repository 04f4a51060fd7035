from .panda3d_batch_renderer import Panda3dBatchRenderer, make_scene_lights  # noqa: F401
from .types import BatchRenderOutput, Panda3dCameraData, Panda3dLightData, Resolution  # noqa: F401
