"""TensorCollection / PandasTensorCollection with the reference's behaviour
(happypose/toolbox/utils/tensor_collection.py:28-230): a pandas DataFrame `infos` plus row-aligned tensors.

filter_top_pose_estimates keeps the reference's contract (rows in global descending score order, top_K per group)
but selects on the GPU with the segmented top-K kernel instead of pandas sort_values/groupby/head.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import pandas as pd
import torch


class HostCopy:
    """Device -> pinned-host copy enqueued NOW on the current stream, read later: `.numpy()` waits for this copy's
    event only, not for everything the caller has enqueued since.  Lets a stage's deferred DataFrame be built while the
    GPU is still busy with the next stages (CPU tensors pass straight through)."""

    def __init__(self, t: torch.Tensor):
        t = t.detach()
        if t.is_cuda:
            self.buf = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
            self.buf.copy_(t, non_blocking=True)
            self.event = torch.cuda.Event()
            self.event.record()
        else:
            self.buf, self.event = t, None

    def numpy(self) -> np.ndarray:
        if self.event is not None:
            self.event.synchronize()
            self.event = None
        return self.buf.numpy()


# Deferred frames travel between stages as plain {column: array} dicts ("cols"): a take / an added column is then a
# numpy operation, and ONE DataFrame is constructed when somebody looks at `.infos` (pandas spends ~0.4 ms per
# iloc / assign / constructor call on even a one-row frame, which used to be the tail of every pipeline call).
def cols_of(df: pd.DataFrame) -> dict:
    return {c: df[c]._values for c in df.columns}


def cols_take(cols: dict, ids: np.ndarray) -> dict:
    ids = np.asarray(ids)
    return {c: v.take(ids) for c, v in cols.items()}


def cols_assign(cols: dict, **new) -> dict:
    out = dict(cols)
    out.update(new)
    return out


def cols_to_frame(cols: dict, n_rows: int) -> pd.DataFrame:
    if not cols:
        return pd.DataFrame(index=pd.RangeIndex(n_rows))
    return pd.DataFrame(cols, copy=False)


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
    """`infos` DataFrame + row-aligned tensors (tensor_collection.py:129-198).

    Two additions that keep the device pipeline free of host synchronisation (the public behaviour is unchanged):
      * `infos` may be given as a zero-argument callable (+ `n_rows`): the DataFrame is then built on first access.  The
        stages of PoseEstimator.run_inference_pipeline chain such deferred frames, so all pandas work (and the
        device->host copies of scores / survivor indices it needs) happens after every kernel has been enqueued.
      * `row_tensors`: private row-aligned device tensors (mesh ids, frame ids, group ids) that follow the rows through
        __getitem__ without appearing in `.tensors`; they spare the stages the label -> id look-ups through pandas.
    """

    def __init__(self, infos, n_rows: Optional[int] = None, row_tensors: Optional[dict] = None, **tensors):
        super().__init__(**tensors)
        self.__dict__["_infos"] = None
        self.__dict__["_infos_thunk"] = None
        self.__dict__["_cols_cache"] = None
        self.__dict__["_n_rows"] = None
        self.__dict__["_row_tensors"] = dict(row_tensors) if row_tensors else {}
        if callable(infos) and not isinstance(infos, pd.DataFrame):
            assert n_rows is not None, "a deferred infos frame needs n_rows"
            self.__dict__["_infos_thunk"] = infos
            self.__dict__["_n_rows"] = int(n_rows)
        else:
            self.__dict__["_infos"] = infos.reset_index(drop=True)
        self.__dict__["meta"] = {}

    # -- deferred DataFrame -----------------------------------------------------------------------
    def _cols(self) -> dict:
        """Column dict of the frame (runs the deferred thunk once; thunks may return a DataFrame or a column dict)."""
        d = self.__dict__
        if d.get("_cols_cache") is None:
            if d["_infos"] is not None:
                d["_cols_cache"] = cols_of(d["_infos"])
            else:
                res = d["_infos_thunk"]()
                if isinstance(res, pd.DataFrame):
                    res = res.reset_index(drop=True)
                    assert len(res) == d["_n_rows"], (len(res), d["_n_rows"])
                    d["_infos"] = res
                    res = cols_of(res)
                d["_cols_cache"] = res
                d["_infos_thunk"] = None
        return d["_cols_cache"]

    @property
    def infos(self) -> pd.DataFrame:
        d = self.__dict__
        if d["_infos"] is None:
            cols = self._cols()
            if d["_infos"] is None:
                df = cols_to_frame(cols, d["_n_rows"])
                assert len(df) == d["_n_rows"], (len(df), d["_n_rows"])
                d["_infos"] = df
        return d["_infos"]

    @infos.setter
    def infos(self, df: pd.DataFrame) -> None:
        self.__dict__["_infos"] = df
        self.__dict__["_infos_thunk"] = None
        self.__dict__["_cols_cache"] = None
        self.__dict__["_n_rows"] = None

    def __setattr__(self, name, value):
        if name == "infos":
            PandasTensorCollection.infos.fset(self, value)
        else:
            super().__setattr__(name, value)

    def map_cols(self, fn) -> None:
        """In-place, deferred `columns = fn(columns)` on the column dict (the row count must not change)."""
        d = self.__dict__
        n = len(self)
        if d.get("_cols_cache") is not None:
            old = d["_cols_cache"]
            source = lambda: old  # noqa: E731
        elif d["_infos"] is not None:
            frame = d["_infos"]
            source = lambda: cols_of(frame.reset_index(drop=True))  # noqa: E731
        else:
            thunk = d["_infos_thunk"]

            def source():
                res = thunk()
                return cols_of(res.reset_index(drop=True)) if isinstance(res, pd.DataFrame) else res

        d["_infos"] = None
        d["_cols_cache"] = None
        d["_infos_thunk"] = lambda: fn(source())
        d["_n_rows"] = n

    def map_infos(self, fn) -> None:
        """In-place, deferred `self.infos = fn(self.infos)` (the row count must not change)."""
        n = len(self)
        self.map_cols(lambda cols: fn(cols_to_frame(cols, n)))

    @property
    def infos_ready(self) -> bool:
        return self.__dict__["_infos"] is not None or self.__dict__.get("_cols_cache") is not None

    @property
    def row_tensors(self) -> dict:
        return self.__dict__["_row_tensors"]

    def merge_df(self, df, *args, **kwargs):
        infos = self.infos.merge(df, how="left", *args, **kwargs)
        assert len(infos) == len(self.infos)
        assert (infos.index == self.infos.index).all()
        return PandasTensorCollection(infos=infos, **self.tensors)

    def clone(self):
        return PandasTensorCollection(self.infos.copy(), row_tensors=self.row_tensors, **super().clone().tensors)

    def to(self, torch_attr):
        super().to(torch_attr)
        if isinstance(torch_attr, (str, torch.device)):  # ids follow device moves, not dtype casts
            for k, t in self.row_tensors.items():
                self.row_tensors[k] = t.to(torch_attr)
        return self

    def __repr__(self):
        body = "".join(f"    {k}: {t.shape} {t.dtype} {t.device},\n" for k, t in self._tensors.items())
        return f"{self.__class__.__name__}(\n{body}{'-' * 40}\n    infos:\n{self.infos!r}\n)"

    def __getitem__(self, ids):
        tensors = super().__getitem__(ids).tensors
        if isinstance(ids, torch.Tensor):
            rt = {k: t[ids.to(t.device)] for k, t in self.row_tensors.items()}
            if ids.dtype != torch.bool:
                # deferred: the index tensor is read back only when somebody looks at the frame
                parent = self
                host_ids = HostCopy(ids)  # the copy is enqueued now, awaited when the frame is looked at
                return PandasTensorCollection(lambda: cols_take(parent._cols(), host_ids.numpy()), n_rows=int(ids.numel()),
                                              row_tensors=rt, **tensors)
            return PandasTensorCollection(self.infos.iloc[ids.detach().cpu().numpy()], row_tensors=rt, **tensors)
        rt = {}
        if self.row_tensors:
            idx = torch.as_tensor(np.asarray(ids))
            rt = {k: t[idx.to(t.device)] for k, t in self.row_tensors.items()}
        return PandasTensorCollection(self.infos.iloc[ids], row_tensors=rt, **tensors)

    def __len__(self):
        n = self.__dict__["_n_rows"]
        return n if n is not None else len(self.infos)

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
        self.__dict__["meta"] = state["meta"]


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
    """Dense int32 group id per row for the (batch_im_id, label, instance_id)-style grouping (first-appearance order,
    like groupby(sort=False).ngroup(), without building a GroupBy object: one factorize per column)."""
    if len(df) == 0:
        return np.zeros((0,), np.int32)
    key = np.zeros(len(df), np.int64)
    for c in group_cols:
        codes, uniques = pd.factorize(df[c].to_numpy(), use_na_sentinel=False)
        key = key * max(len(uniques), 1) + codes
    return pd.factorize(key)[0].astype(np.int32)


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

    When the collection carries device-resident group ids (row_tensors["group_ids"] + meta["n_groups"],
    meta["min_group_size"], set by PoseEstimator) and scores_device is given, nothing is read back to the host: the
    survivor count is n_groups * top_K (every group has at least top_K rows) and the result's `infos` is deferred.
    """
    from .. import ops
    from .._capi import Context

    if len(data_TCO) == 0:
        return data_TCO
    device = data_TCO.device if len(data_TCO.tensors) > 0 else torch.device("cuda")
    ctx = Context.get(device)
    groups_dev = data_TCO.row_tensors.get("group_ids") if group_cols == data_TCO.meta.get("group_cols") else None
    expected = None
    if groups_dev is not None and scores_device is not None:
        groups = groups_dev
        n_groups = int(data_TCO.meta["n_groups"])
        if int(data_TCO.meta.get("min_group_size", 0)) >= int(top_K):
            expected = n_groups * int(top_K)
    else:
        df = data_TCO.infos
        groups = torch.as_tensor(group_ids_from_columns(df, group_cols))
        n_groups = int(groups.max()) + 1
    if scores_device is None:
        scores_device = torch.as_tensor(data_TCO.infos[filter_field].to_numpy(dtype=np.float32))
    scores_device = scores_device.reshape(-1).to(ctx.device, torch.float32)
    if ascending:
        scores_device = -scores_device
    keep = ops.topk_segmented(ctx, scores_device, groups, n_groups, int(top_K), expected_count=expected)
    out = data_TCO[keep.to(device)]
    if groups_dev is not None:
        out.meta.update({"group_cols": group_cols, "n_groups": n_groups, "min_group_size": min(int(top_K), int(data_TCO.meta.get("min_group_size", 0)))})
    return out
