"""CPU-enqueue vs GPU-execution timeline of ONE end-to-end bench step (run on the GPU box).
For every pipeline stage: when the host finished enqueueing it (perf_counter) and when the GPU finished it (CUDA event),
both relative to the start of the step.  Shows where the GPU idles waiting for the host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
import bench as B
from happypose_b200.datasets.object_dataset import RigidObject, RigidObjectDataset
from happypose_b200.inference.types import ObservationTensor
from happypose_b200.megapose import pose_estimator as PE
from happypose_b200.megapose.pose_models_cfg import make_pose_models
from happypose_b200.utils.tensor_collection import PandasTensorCollection

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
ds = RigidObjectDataset([RigidObject(label=B.LABEL, mesh_path=B.MESH, mesh_units="mm")])
coarse, refiner, mesh_db = make_pose_models(ds, device=dev, seed=0)
with torch.no_grad():
    refiner.pose_fc.weight.mul_(1e-2); refiner.pose_fc.bias.copy_(torch.tensor([1.0, 0, 0, 0, 1, 0, 0, 0, 1]))
est = PE.PoseEstimator(refiner_model=refiner, coarse_model=coarse, bsz_objects=16, bsz_images=576, SO3_grid_size=576)
est.use_cuda_graphs = True
n_det = int(sys.argv[1]) if len(sys.argv) > 1 else 1
boxes_host = torch.as_tensor(B.detections_arrays(n_det)).pin_memory()
image_host = torch.rand(1, 3, 480, 640).pin_memory()
K_host = torch.as_tensor(B.K_BBQ[None]).pin_memory()

marks = []
def mark(name):
    ev = torch.cuda.Event(enable_timing=True); ev.record()
    marks.append((name, time.perf_counter(), ev))

def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    def w(*a, **k):
        r = fn(*a, **k)
        mark(label or name)
        return r
    setattr(obj, name, w)

wrap(est, "forward_coarse_model"); wrap(est, "forward_refiner"); wrap(est, "forward_scoring_model")
_f = PE.filter_top_pose_estimates
def ftop(*a, **k):
    r = _f(*a, **k); mark("filter_top"); return r
PE.filter_top_pose_estimates = ftop
_a = PE.add_instance_id
def aid(*a, **k):
    r = _a(*a, **k); mark("add_instance_id"); return r
PE.add_instance_id = aid

def step():
    marks.clear()
    mark("start")
    o = ObservationTensor(image_host.to(dev, non_blocking=True), K_host.to(dev, non_blocking=True))
    det = PandasTensorCollection(infos=pd.DataFrame({"label": [B.LABEL] * n_det, "batch_im_id": [0] * n_det, "score": [1.0] * n_det}),
                                 bboxes=boxes_host.to(dev, non_blocking=True))
    mark("h2d+detections")
    final, _ = est.run_inference_pipeline(o, detections=det, n_refiner_iterations=5, n_pose_hypotheses=1)
    mark("pipeline_return")
    scores = final.infos["pose_score"].to_numpy()
    mark("scores_to_numpy")
    poses = final.poses.cpu()
    mark("poses_cpu")
    return poses, scores

for _ in range(5):
    step()
for rep in range(3):
    torch.cuda.synchronize()
    step()
    torch.cuda.synchronize()
    t0, e0 = marks[0][1], marks[0][2]
    print(f"--- step {rep}: stage, host-enqueue-done ms, gpu-done ms")
    for name, t, ev in marks:
        print(f"  {name:22s} host {1e3*(t-t0):7.2f}   gpu {e0.elapsed_time(ev):7.2f}")
