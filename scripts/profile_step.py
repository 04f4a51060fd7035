"""Where one bench step goes: torch.profiler kernel table + stage timings (run on the GPU box)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
import bench as B
from happypose_b200.datasets.object_dataset import RigidObject, RigidObjectDataset
from happypose_b200.inference.types import ObservationTensor
from happypose_b200.megapose.pose_estimator import PoseEstimator
from happypose_b200.megapose.pose_models_cfg import make_pose_models
from happypose_b200.utils.tensor_collection import PandasTensorCollection

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
ds = RigidObjectDataset([RigidObject(label=B.LABEL, mesh_path=B.MESH, mesh_units="mm")])
coarse, refiner, mesh_db = make_pose_models(ds, device=dev, seed=0)
with torch.no_grad():
    refiner.pose_fc.weight.mul_(1e-2); refiner.pose_fc.bias.copy_(torch.tensor([1.0, 0, 0, 0, 1, 0, 0, 0, 1]))
est = PoseEstimator(refiner_model=refiner, coarse_model=coarse, bsz_objects=16, bsz_images=576, SO3_grid_size=576)
est.use_cuda_graphs = os.environ.get('HPB_GRAPHS', '1') == '1'
n_det = int(sys.argv[1]) if len(sys.argv) > 1 else 1
boxes = torch.as_tensor(B.detections_arrays(n_det)).to(dev)
image = torch.rand(1, 3, 480, 640, device=dev)
obs = ObservationTensor(image, torch.as_tensor(B.K_BBQ[None]).to(dev))

def det():
    return PandasTensorCollection(infos=pd.DataFrame({"label": [B.LABEL] * n_det, "batch_im_id": [0] * n_det, "score": [1.0] * n_det}), bboxes=boxes)

def step():
    return est.run_inference_pipeline(obs, detections=det(), n_refiner_iterations=5, n_pose_hypotheses=1)

for _ in range(4):
    step()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    final, extra = step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t) / 5 * 1e3, extra["timing_str"])
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=70))
