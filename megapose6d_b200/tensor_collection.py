"""DataFrame + same-length tensors, the container every stage of the pipeline exchanges.

Mirrors the interface of the reference's megapose.utils.tensor_collection
(src/megapose/utils/tensor_collection.py:27-197): `TensorCollection`, `PandasTensorCollection`
(`infos` DataFrame re-indexed 0..n-1, tensors reachable as attributes, integer / list / tensor
indexing selects rows of both) and `concatenate`.
"""
from __future__ import annotations

from typing import Dict, Iterable

import numpy as np
import pandas as pd
import torch


class TensorCollection:
    def __init__(self, **tensors: torch.Tensor):
        self.__dict__["_tensors"] = dict()
        for name, tensor in tensors.items():
            self.register_tensor(name, tensor)

    # -- registry
    def register_tensor(self, name: str, tensor: torch.Tensor) -> None:
        self._tensors[name] = tensor

    def delete_tensor(self, name: str) -> None:
        del self._tensors[name]

    @property
    def tensors(self) -> Dict[str, torch.Tensor]:
        return self._tensors

    @property
    def device(self) -> torch.device:
        return next(iter(self._tensors.values())).device

    # -- attribute plumbing: registered tensors behave like attributes
    def __getattr__(self, name: str):
        tensors = self.__dict__.get("_tensors", {})
        if name in tensors:
            return tensors[name]
        raise AttributeError(name)

    def __setattr__(self, name: str, value) -> None:
        if "_tensors" not in self.__dict__:
            raise ValueError("Please call __init__")
        if name in self._tensors:
            self._tensors[name] = value
        else:
            self.__dict__[name] = value

    def __getitem__(self, ids):
        return TensorCollection(**{k: v[ids] for k, v in self._tensors.items()})

    def __repr__(self) -> str:
        lines = [f"    {k}: {tuple(t.shape)} {t.dtype} {t.device}," for k, t in self._tensors.items()]
        return self.__class__.__name__ + "(\n" + "\n".join(lines) + "\n)"

    # -- pickling
    def __getstate__(self):
        return {"tensors": self.tensors}

    def __setstate__(self, state):
        self.__init__(**state["tensors"])

    # -- conversions (in place, return self, as the reference does)
    def to(self, torch_attr):
        for k, v in self._tensors.items():
            self._tensors[k] = v.to(torch_attr)
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
        return TensorCollection(**{k: v.clone() for k, v in self._tensors.items()})


class PandasTensorCollection(TensorCollection):
    def __init__(self, infos: pd.DataFrame, **tensors: torch.Tensor):
        super().__init__(**tensors)
        self.infos = infos.reset_index(drop=True)
        self.meta: dict = dict()

    @classmethod
    def _wrap(cls, infos: pd.DataFrame, tensors: Dict[str, torch.Tensor]) -> "PandasTensorCollection":
        """Constructor for a freshly made `infos` frame: re-labels its index in place instead of copying it."""
        out = cls.__new__(cls)
        TensorCollection.__init__(out, **tensors)
        infos.index = pd.RangeIndex(len(infos))
        out.infos = infos
        out.meta = dict()
        return out

    def __len__(self) -> int:
        return len(self.infos)

    def __getitem__(self, ids):
        if isinstance(ids, torch.Tensor):
            row_ids = ids.cpu().numpy()
        else:
            row_ids = ids
        infos = self.infos.iloc[row_ids]
        if infos is self.infos or not isinstance(infos, pd.DataFrame):
            infos = self.infos.iloc[list(np.atleast_1d(row_ids))]
        tensors = {k: v[ids] for k, v in self._tensors.items()}
        return PandasTensorCollection._wrap(infos, tensors)

    def merge_df(self, df: pd.DataFrame, *args, **kwargs) -> "PandasTensorCollection":
        infos = self.infos.merge(df, how="left", *args, **kwargs)
        assert len(infos) == len(self.infos)
        assert (infos.index == self.infos.index).all()
        return PandasTensorCollection(infos=infos, **self.tensors)

    def clone(self) -> "PandasTensorCollection":
        return PandasTensorCollection(self.infos.copy(), **{k: v.clone() for k, v in self._tensors.items()})

    def __repr__(self) -> str:
        head = super().__repr__()[:-2]
        return head + f"\n{'-' * 40}\n    infos:\n{self.infos!r}\n)"

    def __getstate__(self):
        state = super().__getstate__()
        state["infos"] = self.infos
        state["meta"] = self.meta
        return state

    def __setstate__(self, state):
        self.__init__(state["infos"], **state["tensors"])
        self.meta = state["meta"]


def concatenate(datas: Iterable[PandasTensorCollection]) -> PandasTensorCollection:
    datas = [d for d in datas if len(d) > 0]
    if len(datas) == 0:
        return PandasTensorCollection(infos=pd.DataFrame())
    assert all(d.__class__ == datas[0].__class__ for d in datas)
    infos = pd.concat([d.infos for d in datas], axis=0, sort=False).reset_index(drop=True)
    tensors = {k: torch.cat([getattr(d, k) for d in datas], dim=0) for k in datas[0].tensors.keys()}
    return PandasTensorCollection(infos=infos, **tensors)
