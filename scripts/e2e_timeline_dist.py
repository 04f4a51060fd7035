"""Per-rank host/GPU timeline of ONE end-to-end bench step at N > 1 (run under torchrun on the GPU box):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/e2e_timeline_dist.py
For every pipeline stage and every collective: when the host finished enqueueing it and when the GPU finished it, relative
to the start of the step, on every rank (written to gpurun_out/timeline_rank<r>.txt).  Also times K back-to-back e2e steps
the way bench.py does (barrier + synchronize on both sides), with a few host-side knobs (env HPB_TL_VARIANT)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
import torch.distributed as dist
import bench as B
from happypose_b200 import distributed as hdist
from happypose_b200.datasets.object_dataset import RigidObject, RigidObjectDataset
from happypose_b200.inference.types import ObservationTensor
from happypose_b200.megapose import pose_estimator as PE
from happypose_b200.megapose.pose_models_cfg import make_pose_models
from happypose_b200.utils.tensor_collection import PandasTensorCollection

rank, local_rank, world = hdist.init_distributed_mode()
dev = torch.device("cuda", local_rank)
torch.backends.cudnn.benchmark = True
ds = RigidObjectDataset([RigidObject(label=B.LABEL, mesh_path=B.MESH, mesh_units="mm")])
coarse, refiner, mesh_db = make_pose_models(ds, device=dev, seed=0)
with torch.no_grad():
    refiner.pose_fc.weight.mul_(1e-2); refiner.pose_fc.bias.copy_(torch.tensor([1.0, 0, 0, 0, 1, 0, 0, 0, 1]))
est = PE.PoseEstimator(refiner_model=refiner, coarse_model=coarse, bsz_objects=16, bsz_images=576, SO3_grid_size=576, shard_across_ranks=True)
est.use_cuda_graphs = True
n_det = world
boxes_host = torch.as_tensor(B.detections_arrays(n_det)).pin_memory()
image_host = torch.as_tensor(np.random.RandomState(0).rand(1, 3, 480, 640).astype(np.float32)).pin_memory()
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
for nm in ("all_gather_rows", "all_gather_rows_packed"):
    wrap(hdist, nm)
_f = PE.filter_top_pose_estimates
def ftop(*a, **k):
    r = _f(*a, **k); mark("filter_top"); return r
PE.filter_top_pose_estimates = ftop

def step(record=True):
    if record:
        marks.clear(); mark("start")
    o = ObservationTensor(image_host.to(dev, non_blocking=True), K_host.to(dev, non_blocking=True))
    det = PandasTensorCollection(infos=pd.DataFrame({"label": [B.LABEL] * n_det, "batch_im_id": [0] * n_det, "score": [1.0] * n_det}),
                                 bboxes=boxes_host.to(dev, non_blocking=True))
    final, _ = est.run_inference_pipeline(o, detections=det, n_refiner_iterations=5, n_pose_hypotheses=1)
    if record: mark("pipeline_return")
    scores = final.infos["pose_score"].to_numpy()
    if record: mark("scores_to_numpy")
    poses = final.poses.cpu()
    if record: mark("poses_cpu")
    return poses, scores

def sync_all():
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier(); torch.cuda.synchronize(dev)

for _ in range(6):
    step()
out = open(os.path.join("gpurun_out", f"timeline_rank{rank}.txt"), "w")
print(f"rank {rank}/{world} pid {os.getpid()} affinity {sorted(os.sched_getaffinity(0))[:4]}..({len(os.sched_getaffinity(0))} cpus) "
      f"OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')} torch threads {torch.get_num_threads()}", file=out)
for rep in range(3):
    sync_all()
    step()
    torch.cuda.synchronize()
    t0, e0 = marks[0][1], marks[0][2]
    print(f"--- step {rep}: stage, host-enqueue-done ms, gpu-done ms", file=out)
    for name, t, ev in marks:
        print(f"  {name:24s} host {1e3*(t-t0):7.2f}   gpu {e0.elapsed_time(ev):7.2f}", file=out)
# bench-style loop
K = 20
sync_all()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(K):
    step(record=False)
e1.record(); sync_all()
print(f"bench-style loop: {e0.elapsed_time(e1)/K:.3f} ms/step (cuda events), {(time.perf_counter()-t0)*1e3/K:.3f} ms/step (wall)", file=out)
out.close()
if world > 1:
    dist.destroy_process_group()
