"""TEST INFRASTRUCTURE ONLY -- never imported by the product (happypose_b200/).

Import shim that loads the reference's *torch-only* leaf modules unchanged from
/root/reference so that golden vectors can be generated from the real reference
functions.  It only works in the build container (the GPU
box has no /root/reference); the only caller is tests/golden/generate_golden.py.

What it does:
  * patches importlib.metadata.metadata (happypose/__init__.py:8 asks for the
    installed distribution's metadata),
  * installs a sys.meta_path finder that stubs the third-party packages that are
    not installed here (panda3d, trimesh, roma, pinocchio, transforms3d, ...),
  * sets HAPPYPOSE_DATA_DIR / CUDA_VISIBLE_DEVICES the way megapose/config.py and
    megapose/__init__.py expect.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import importlib.metadata
import os
import sys
import tempfile
import types

REFERENCE_ROOT = "/root/reference"

_STUB_ROOTS = (
    "panda3d", "direct", "pybullet", "pybullet_data", "trimesh", "roma", "pinocchio",
    "transforms3d", "omegaconf", "plyfile", "bokeh", "webdataset", "bop_toolkit_lib",
    "pytest_order", "meshcat", "simplejson", "xarray", "imageio", "pypng", "png",
    "seaborn", "joblib_stub", "ipdb", "colorama", "httpx", "bs4", "open3d", "teaserpp_python",
    "pyarrow_stub", "dask", "distributed", "selenium", "geckodriver", "cosypose_cext",
)


class _Stub(types.ModuleType):
    """Module whose every attribute is another stub (callable, subscriptable)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        full = f"{self.__name__}.{name}"
        mod = sys.modules.get(full)
        if mod is None:
            mod = _StubAttr(full)
        return mod


class _StubAttr:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        return _StubAttr(self._name + "()")

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _StubAttr(f"{self._name}.{name}")

    def __getitem__(self, item):
        return _StubAttr(f"{self._name}[]")

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())

    def __mul__(self, other):
        return self

    __rmul__ = __mul__


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Make `import happypose...` resolve to /root/reference (leaf modules only)."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(
            "reference tree not present; golden vectors can only be regenerated in the build container"
        )
    os.environ.setdefault("HAPPYPOSE_DATA_DIR", tempfile.mkdtemp(prefix="hpdata_"))
    if not os.environ.get("CUDA_VISIBLE_DEVICES", "").strip():
        os.environ["CUDA_VISIBLE_DEVICES"] = "0"
    _orig_metadata = importlib.metadata.metadata

    def _metadata(name):
        if name == "happypose":
            return {"name": "happypose", "version": "0.0.0-ref", "license": "BSD-2-Clause", "author": "reference"}
        return _orig_metadata(name)

    importlib.metadata.metadata = _metadata
    sys.meta_path.append(_StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import numpy as np

    if not hasattr(np, "float_"):
        np.float_ = np.float64  # symmetries.py:36 uses the alias removed in NumPy 2
    _installed = True


def ref(module: str):
    """Import `happypose.<module>` from the reference tree."""
    install()
    return importlib.import_module("happypose." + module)
