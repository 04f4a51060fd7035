"""TensorCollection / PandasTensorCollection with the reference's behaviour
(happypose/toolbox/utils/tensor_collection.py:28-230): a pandas DataFrame `infos` plus row-aligned tensors.

filter_top_pose_estimates keeps the reference's contract (rows in global descending score order, top_K per group)
but selects on the GPU with the segmented top-K kernel instead of pandas sort_values/groupby/head.
"""
from __future__ import annotations

from typing import List

import numpy as np
import pandas as pd
import torch


class TensorCollection:
    def __init__(self, **tensors):
        self.__dict__["_tensors"] = {}
        for name, t in tensors.items():
            self.register_tensor(name, t)

    # -- registry --------------------------------------------------------------------------------
    def register_tensor(self, name, tensor):
        self._tensors[name] = tensor

    def delete_tensor(self, name):
        del self._tensors[name]

    @property
    def tensors(self):
        return self._tensors

    @property
    def device(self):
        return next(iter(self._tensors.values())).device

    def __getattr__(self, name):
        tensors = self.__dict__.get("_tensors", {})
        if name in tensors:
            return tensors[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if "_tensors" not in self.__dict__:
            raise ValueError("Please call __init__")
        if name in self._tensors:
            self._tensors[name] = value
        else:
            self.__dict__[name] = value

    # -- container protocol ----------------------------------------------------------------------
    def __getitem__(self, ids):
        return TensorCollection(**{k: t[ids] for k, t in self._tensors.items()})

    def __repr__(self):
        body = "".join(f"    {k}: {t.shape} {t.dtype} {t.device},\n" for k, t in self._tensors.items())
        return f"{self.__class__.__name__}(\n{body})"

    def __getstate__(self):
        return {"tensors": self.tensors}

    def __setstate__(self, state):
        self.__init__(**state["tensors"])

    # -- conversions -----------------------------------------------------------------------------
    def to(self, torch_attr):
        for k, t in self._tensors.items():
            self._tensors[k] = t.to(torch_attr)
        return self

    def cuda(self):
        return self.to("cuda")

    def cpu(self):
        return self.to("cpu")

    def float(self):
        return self.to(torch.float)

    def double(self):
        return self.to(torch.double)

    def half(self):
        return self.to(torch.half)

    def clone(self):
        return TensorCollection(**{k: t.clone() for k, t in self._tensors.items()})


class PandasTensorCollection(TensorCollection):
    def __init__(self, infos: pd.DataFrame, **tensors):
        super().__init__(**tensors)
        self.infos = infos.reset_index(drop=True)
        self.meta = {}

    def merge_df(self, df, *args, **kwargs):
        infos = self.infos.merge(df, how="left", *args, **kwargs)
        assert len(infos) == len(self.infos)
        assert (infos.index == self.infos.index).all()
        return PandasTensorCollection(infos=infos, **self.tensors)

    def clone(self):
        return PandasTensorCollection(self.infos.copy(), **super().clone().tensors)

    def __repr__(self):
        body = "".join(f"    {k}: {t.shape} {t.dtype} {t.device},\n" for k, t in self._tensors.items())
        return f"{self.__class__.__name__}(\n{body}{'-' * 40}\n    infos:\n{self.infos!r}\n)"

    def __getitem__(self, ids):
        if isinstance(ids, torch.Tensor):
            rows = ids.detach().cpu().numpy()
        else:
            rows = ids
        infos = self.infos.iloc[rows].reset_index(drop=True)
        return PandasTensorCollection(infos, **super().__getitem__(ids).tensors)

    def __len__(self):
        return len(self.infos)

    def gather_distributed(self, tmp_dir=None):
        """Reference: pickle files in tmp_dir + barriers (tensor_collection.py:166-187).  Here: all_gather_object
        over the process group (NCCL/gloo), no files.  Every rank returns the concatenation."""
        from ..distributed import all_gather_collections

        return all_gather_collections(self)

    def __getstate__(self):
        state = super().__getstate__()
        state["infos"] = self.infos
        state["meta"] = self.meta
        return state

    def __setstate__(self, state):
        self.__init__(state["infos"], **state["tensors"])
        self.meta = state["meta"]


def concatenate(datas):
    """tensor_collection.py:28-42."""
    datas = [d for d in datas if len(d) > 0]
    if len(datas) == 0:
        return PandasTensorCollection(infos=pd.DataFrame())
    assert all(d.__class__ == datas[0].__class__ for d in datas)
    infos = pd.concat([d.infos for d in datas], axis=0, sort=False).reset_index(drop=True)
    tensors = {k: torch.cat([getattr(d, k) for d in datas], dim=0) for k in datas[0].tensors.keys()}
    return PandasTensorCollection(infos=infos, **tensors)


def group_ids_from_columns(df: pd.DataFrame, group_cols: List[str]) -> np.ndarray:
    """Dense int32 group id per row for the (batch_im_id, label, instance_id)-style grouping."""
    if len(df) == 0:
        return np.zeros((0,), np.int32)
    return df.groupby(group_cols, sort=False).ngroup().to_numpy().astype(np.int32)


def filter_top_pose_estimates(
    data_TCO: PandasTensorCollection,
    top_K: int,
    group_cols: List[str],
    filter_field: str,
    ascending: bool = False,
    scores_device: torch.Tensor = None,
) -> PandasTensorCollection:
    """tensor_collection.py:201-230: keep the top_K rows of every group, rows returned in global score order.

    The selection runs on the GPU (hpb_topk_segmented); `scores_device` lets the caller pass the logits that are
    already resident on the device instead of the DataFrame column.  Ties: lowest row index first.
    """
    from .. import ops
    from .._capi import Context

    df = data_TCO.infos
    if len(df) == 0:
        return data_TCO
    groups = group_ids_from_columns(df, group_cols)
    n_groups = int(groups.max()) + 1
    device = data_TCO.device if len(data_TCO.tensors) > 0 else torch.device("cuda")
    ctx = Context.get(device)
    if scores_device is None:
        scores_device = torch.as_tensor(df[filter_field].to_numpy(dtype=np.float32))
    scores_device = scores_device.reshape(-1).to(ctx.device, torch.float32)
    if ascending:
        scores_device = -scores_device
    keep = ops.topk_segmented(ctx, scores_device, torch.as_tensor(groups), n_groups, int(top_K))
    return data_TCO[keep.to(device)]
