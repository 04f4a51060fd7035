"""CPU, world_size 2, gloo: the host-side sharding / gathering logic of the N>1 path (happypose_b200/distributed.py)."""
import os
import socket

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from happypose_b200 import distributed as hdist
    from happypose_b200.utils.tensor_collection import PandasTensorCollection

    r, _, w = hdist.init_distributed_mode(backend="gloo")
    assert (r, w) == (rank, world) and hdist.is_distributed()
    results = {}
    for n in (0, 1, 5, 576, 577, 1152):
        lo, hi = hdist.shard_range(n)
        full = torch.arange(n, dtype=torch.float32).reshape(n, 1).repeat(1, 3)
        got = hdist.all_gather_rows(full[lo:hi].contiguous(), n)
        assert got.shape == full.shape and torch.equal(got, full), n
        results[n] = (lo, hi)
    # the coarse-stage exchange: logits of this rank's slice -> full table -> the same top-1 on every rank
    n = 2 * 576
    torch.manual_seed(0)
    logits = torch.randn(n)
    lo, hi = hdist.shard_range(n)
    gathered = hdist.all_gather_rows(logits[lo:hi].reshape(-1, 1).contiguous(), n).reshape(-1)
    assert torch.equal(gathered, logits)
    best = gathered.reshape(2, 576).argmax(1)
    flag = torch.tensor([int(best[0]), int(best[1])])
    both = [torch.zeros_like(flag) for _ in range(world)]
    dist.all_gather(both, flag)
    assert torch.equal(both[0], both[1])
    # the refiner-stage exchange: a dict of row-aligned tensors through ONE collective
    nrows = 5
    lo, hi = hdist.shard_range(nrows)
    full = {("it", 1, "poses"): torch.arange(nrows * 16, dtype=torch.float32).reshape(nrows, 4, 4),
            ("it", 1, "K_crop"): torch.arange(nrows * 9, dtype=torch.float32).reshape(nrows, 3, 3) + 0.5,
            ("it", 2, "boxes"): -torch.arange(nrows * 4, dtype=torch.float32).reshape(nrows, 4)}
    got = hdist.all_gather_rows_packed({k: v[lo:hi].contiguous() for k, v in full.items()}, nrows)
    assert list(got.keys()) == list(full.keys())
    for k in full:
        assert got[k].shape == full[k].shape and torch.equal(got[k], full[k]), k
    # input agreement check used by PoseEstimator(shard_across_ranks=True)
    assert hdist.all_ranks_equal([3.0, 1.5, -2.0])
    assert not hdist.all_ranks_equal([3.0, float(rank)])
    # collections: file-free gather_distributed
    coll = PandasTensorCollection(pd.DataFrame({"rank": [rank] * (rank + 1)}), poses=torch.full((rank + 1, 4, 4), float(rank)))
    allc = coll.gather_distributed()
    assert len(allc) == 3 and list(allc.infos["rank"]) == [0, 1, 1] and allc.poses.shape == (3, 4, 4)
    hdist.barrier()
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array(sorted(results.items()), dtype=object), allow_pickle=True)
    dist.destroy_process_group()


def test_shard_bounds_cover_and_are_contiguous():
    from happypose_b200.distributed import shard_bounds

    for n in (0, 1, 7, 576, 138240):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(w))
            sizes = [b[i + 1] - b[i] for i in range(w)]
            assert max(sizes) == (n + w - 1) // w


def test_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok_0.npy") and os.path.exists(tmp_path / "ok_1.npy")
