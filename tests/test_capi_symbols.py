"""CPU: libhpb200.so builds for sm_100a, loads, and exports every entry point include/hpb200.h declares (no compute
calls -- there is no GPU here).  Also: the ctypes table binds exactly the declared set, and the product package never
imports the oracle."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hpb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hpb_[a-zA-Z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("hpb_create", "hpb_destroy", "hpb_mesh_upload", "hpb_render", "hpb_crop", "hpb_crop_boxes", "hpb_normalize_T",
                 "hpb_pose_update", "hpb_tco_init", "hpb_multiview", "hpb_topk_segmented", "hpb_last_error"):
        assert must in names


def test_library_builds_loads_and_exports_every_declared_symbol():
    from happypose_b200 import _build, _capi

    path = _build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/hpb200.h but not exported by libhpb200.so"
    assert sorted(_capi.SIGNATURES) == declared_symbols(), "ctypes table and header disagree"
    assert lib.hpb_version() == 100
    _capi.load_library(build_if_missing=False)


def test_library_contains_sm_100a_code():
    from happypose_b200 import _build

    path = _build.build()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_cpu_fallback_without_cuda():
    import torch
    from happypose_b200._capi import Context, HpbError

    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    with pytest.raises(HpbError):
        Context.get("cuda:0")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "happypose_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"
                assert "from .. import oracle" not in src
