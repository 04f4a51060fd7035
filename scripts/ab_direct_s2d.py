"""A/B in one process: coarse hand-off through the fused space-to-depth render (PosePredictor.use_direct_s2d) vs the
float32 network input + packing pass.  usage (GPU box): python scripts/ab_direct_s2d.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pandas as pd
import bench as B
from happypose_b200.datasets.object_dataset import RigidObject, RigidObjectDataset
from happypose_b200.inference.types import ObservationTensor
from happypose_b200.megapose.pose_estimator import PoseEstimator
from happypose_b200.megapose.pose_models_cfg import make_pose_models
from happypose_b200.utils.tensor_collection import PandasTensorCollection
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
ds = RigidObjectDataset([RigidObject(label=B.LABEL, mesh_path=B.MESH, mesh_units="mm")])
coarse, refiner, mesh_db = make_pose_models(ds, device=dev, seed=0)
with torch.no_grad():
    refiner.pose_fc.weight.mul_(1e-2); refiner.pose_fc.bias.copy_(torch.tensor([1.0, 0, 0, 0, 1, 0, 0, 0, 1]))
est = PoseEstimator(refiner_model=refiner, coarse_model=coarse, bsz_objects=16, bsz_images=576, SO3_grid_size=576)
est.use_cuda_graphs = True
boxes = torch.as_tensor(B.detections_arrays(1)).to(dev)
obs = ObservationTensor(torch.rand(1, 3, 480, 640, device=dev), torch.as_tensor(B.K_BBQ[None]).to(dev))
def step():
    det = PandasTensorCollection(infos=pd.DataFrame({"label": [B.LABEL], "batch_im_id": [0], "score": [1.0]}), bboxes=boxes)
    return est.run_inference_pipeline(obs, detections=det, n_refiner_iterations=5, n_pose_hypotheses=1)
def timed(n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for flag in (True, False):
    coarse.use_direct_s2d = flag
    coarse._graphs = type(coarse._graphs)()  # graphs were captured with the other hand-off
    for _ in range(4): step()
for rep in range(3):
    for flag in (True, False):
        coarse.use_direct_s2d = flag
        coarse._graphs = type(coarse._graphs)()
        for _ in range(2): step()
        print(f"rep {rep} direct_s2d={flag}: {timed():.3f} ms/step", flush=True)
