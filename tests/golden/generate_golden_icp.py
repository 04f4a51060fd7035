"""Generates tests/golden/ref_icp.npz from the REAL reference functions of the ICP depth refiner's input stage
(happypose/pose_estimators/megapose/inference/icp_refiner.py: getXYZ :106-135, get_normal :33-103, the selection rules of
icp_refinement :138-188; refiner_utils.py: compute_masks).  icp_refiner.py itself cannot be imported here (it pulls the
Panda3D renderer in at module scope), so the two function definitions are taken out of the unmodified source file by name
(ast) and executed as they are; compute_masks is imported through oracle/ref_shim.py.  Build-container only."""
import ast
import os
import sys

import cv2
import numpy as np
from scipy import ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

SRC = os.path.join(ref_shim.REFERENCE_ROOT, "happypose", "pose_estimators", "megapose", "inference", "icp_refiner.py")


def reference_functions(names):
    tree = ast.parse(open(SRC).read())
    ns = {"np": np, "cv2": cv2, "ndimage": ndimage}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), SRC, "exec"), ns)
    return [ns[n] for n in names]


def synthetic_depths(rs, H, W):
    """A measured depth map (tilted plane + bumps, holes = 0, a few out-of-range values) and a rendered depth (an ellipse
    of an object surface near the measured one, 0 elsewhere)."""
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    dm = (0.6 + 0.0008 * xx + 0.0005 * yy + 0.02 * np.sin(xx / 7.0) * np.cos(yy / 5.0)).astype(np.float32)
    dm[rs.rand(H, W) < 0.05] = 0.0
    dm[5:9, 10:30] = 6.0      # beyond the 5 m validity limit
    dm[20:24, 40:60] = 0.15   # closer than 0.2 m
    inside = ((xx - W * 0.55) / (W * 0.3)) ** 2 + ((yy - H * 0.5) / (H * 0.35)) ** 2 < 1.0
    dr = np.where(inside, dm + 0.03 * np.cos(xx / 11.0) + rs.uniform(-0.02, 0.02, (H, W)), 0.0).astype(np.float32)
    dr[inside & (rs.rand(H, W) < 0.1)] += 0.2  # outliers beyond the 0.1 m mask threshold
    dr[inside & (dm == 0)] = 0.7
    return dm, dr.astype(np.float32)


def main():
    getXYZ, get_normal = reference_functions(["getXYZ", "get_normal"])
    compute_masks = ref_shim.ref("pose_estimators.megapose.inference.refiner_utils").compute_masks
    rs = np.random.RandomState(7)
    H, W = 120, 160
    out = {}
    for case, K in enumerate((np.array([[300.0, 0, 79.6], [0, 310.0, 58.3], [0, 0, 1]], np.float32),
                              np.array([[605.95, 0, 81.03], [0, 605.01, 60.68], [0, 0, 1]], np.float32))):
        dm, dr = synthetic_depths(rs, H, W)
        _, mask = compute_masks("threshold", depth_rendered=dr, depth_measured=dm, depth_delta_thresh=0.1)
        # icp_refinement :150-176 (the part in front of the ICP call)
        xyz_t = np.zeros((H, W, 3), np.float32)
        xyz_t[:] = getXYZ(dm, fx=K[0, 0], fy=K[1, 1], cx=K[0, 2], cy=K[1, 2])
        xyz_s = np.zeros((H, W, 3), np.float32)
        xyz_s[:] = getXYZ(dr, K[0, 0], K[1, 1], K[0, 2], K[1, 2])
        valid = np.logical_and(np.logical_and(dm > 0.2, dm < 5), mask)
        valid_src = np.logical_and(valid, dr > 0)
        n_t = np.zeros((H, W, 3), np.float32)
        n_t[:] = get_normal(dm, fx=K[0, 0], fy=K[1, 1], cx=K[0, 2], cy=K[1, 2], refine=True)
        out.update({f"K{case}": K, f"depth_measured{case}": dm, f"depth_rendered{case}": dr, f"mask{case}": mask,
                    f"points_tgt{case}": xyz_t[valid], f"points_src{case}": xyz_s[valid_src], f"normals_tgt{case}": n_t[valid]})
    np.savez_compressed(os.path.join(HERE, "ref_icp.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
