"""Host-side (Python) cost of one bench step: cProfile over a few run_inference_pipeline calls (run on the GPU box)."""
import cProfile, pstats, os, sys, io, time
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
est.use_cuda_graphs = True
n_det = int(sys.argv[1]) if len(sys.argv) > 1 else 1
boxes = torch.as_tensor(B.detections_arrays(n_det)).to(dev)
obs = ObservationTensor(torch.rand(1, 3, 480, 640, device=dev), torch.as_tensor(B.K_BBQ[None]).to(dev))

def det():
    return PandasTensorCollection(infos=pd.DataFrame({"label": [B.LABEL] * n_det, "batch_im_id": [0] * n_det, "score": [1.0] * n_det}), bboxes=boxes)

E2E = len(sys.argv) > 2 and sys.argv[2] == "e2e"
image_host = torch.rand(1, 3, 480, 640).pin_memory()
K_host = torch.as_tensor(B.K_BBQ[None]).pin_memory()


def step():
    if not E2E:
        return est.run_inference_pipeline(obs, detections=det(), n_refiner_iterations=5, n_pose_hypotheses=1)
    o = ObservationTensor(image_host.to(dev, non_blocking=True), K_host.to(dev, non_blocking=True))
    final, _ = est.run_inference_pipeline(o, detections=det(), n_refiner_iterations=5, n_pose_hypotheses=1)
    scores = final.infos["pose_score"].to_numpy()
    return final.poses.cpu(), scores

for _ in range(4):
    step()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10):
    step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t) / 10 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
torch.cuda.synchronize()
pr.disable()
for key in ("tottime", "cumtime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue()[:9000])
