"""torchrun worker of tests/test_gpu_bench_config.py::test_two_rank_nccl_pipeline_equals_single_rank: runs the sharded
pipeline on this rank's GPU and saves what it returned (every rank must return the full, identical result)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist

    from happypose_b200 import distributed as hdist
    from tests.test_gpu_bench_config import run_pipeline_for_nccl_test

    out_dir = sys.argv[1]
    rank, _, world = hdist.init_distributed_mode()
    assert world == 2 and hdist.is_distributed()
    out = run_pipeline_for_nccl_test()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **out)
    hdist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
