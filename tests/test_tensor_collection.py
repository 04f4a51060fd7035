"""Host logic of PandasTensorCollection (reference: happypose/toolbox/utils/tensor_collection.py:129-198) and of the
deferred column-dict frames the pipeline stages chain.  CPU only; results are compared with plain pandas."""
import numpy as np
import pandas as pd
import torch

from happypose_b200.utils import tensor_collection as tc
from happypose_b200.utils.tensor_collection import PandasTensorCollection


def _frame(n):
    return pd.DataFrame({"label": [f"obj_{i % 3:06d}" for i in range(n)], "batch_im_id": np.arange(n) // 4, "score": np.linspace(0, 1, n)})


def test_eager_collection_matches_reference_behaviour():
    df = _frame(8)
    c = PandasTensorCollection(infos=df, poses=torch.arange(8 * 16, dtype=torch.float32).reshape(8, 4, 4))
    assert len(c) == 8 and c.poses.shape == (8, 4, 4)
    sub = c[[1, 3, 5]]
    pd.testing.assert_frame_equal(sub.infos, df.iloc[[1, 3, 5]].reset_index(drop=True))
    assert torch.equal(sub.poses, c.poses[[1, 3, 5]])
    ids = torch.tensor([6, 0, 2])
    sub = c[ids]  # deferred frame
    assert len(sub) == 3
    pd.testing.assert_frame_equal(sub.infos, df.iloc[[6, 0, 2]].reset_index(drop=True))


def test_deferred_column_chain_equals_pandas():
    df = _frame(12)
    calls = []

    def thunk():
        calls.append(1)
        return tc.cols_assign(tc.cols_of(df), hypothesis_id=np.arange(12))

    c = PandasTensorCollection(thunk, n_rows=12, poses=torch.zeros(12, 4, 4))
    assert len(c) == 12 and not calls, "len() must not materialise a deferred frame"
    ids = torch.tensor([7, 2, 9])
    sub = c[ids]
    sub.map_cols(lambda cols: tc.cols_assign(cols, pose_logit=np.array([1.0, 2.0, 3.0], np.float32)))
    sub2 = sub[torch.tensor([2, 0])]
    assert not calls, "chaining must stay lazy"
    want = df.assign(hypothesis_id=np.arange(12)).iloc[[7, 2, 9]].reset_index(drop=True).assign(pose_logit=np.array([1.0, 2.0, 3.0], np.float32))
    pd.testing.assert_frame_equal(sub.infos, want)
    pd.testing.assert_frame_equal(sub2.infos, want.iloc[[2, 0]].reset_index(drop=True))
    pd.testing.assert_frame_equal(c.infos, df.assign(hypothesis_id=np.arange(12)))
    assert len(calls) == 1, "the base frame is built once and shared"


def test_map_infos_and_setter_still_work():
    df = _frame(5)
    c = PandasTensorCollection(lambda: df, n_rows=5, poses=torch.zeros(5, 4, 4))
    c.map_infos(lambda d: d.assign(x=np.arange(5)))
    pd.testing.assert_frame_equal(c.infos, df.assign(x=np.arange(5)))
    c.infos = df.iloc[:5].assign(y=1)
    assert "y" in c.infos and len(c) == 5
    c.map_cols(lambda cols: tc.cols_assign(cols, z=np.zeros(5)))
    assert list(c.infos.columns) == ["label", "batch_im_id", "score", "y", "z"]


def test_host_copy_passes_cpu_tensors_through():
    t = torch.arange(6).reshape(2, 3)
    assert np.array_equal(tc.HostCopy(t).numpy(), t.numpy())
