"""happypose_b200 -- B200-native (sm_100a) MegaPose / CosyPose render-and-compare hot path.

Hand-written CUDA kernels behind a C ABI (include/hpb200.h, happypose_b200/csrc/), called from a Python host layer
that mirrors happypose's API for this path (PoseEstimator, PosePredictor, Panda3dBatchRenderer, lib3d helpers).
There is no CPU fallback: the kernels run on the GPU or the call raises.
"""
__version__ = "0.1.0"
