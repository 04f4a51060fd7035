"""Shared synthetic scenes for the tests (seeded; no reference files needed at run time)."""
import numpy as np


def quat_xyzw_to_mat(q):
    x, y, z, w = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )


def reference_test_scene():
    """Scene of the reference's tests/test_batch_renderer_panda3d.py:43-69:
    TWO = quat(0.5,0.5,-0.5,0.5), t=(0,0,0.3); camera at identity; K fx=fy=300, c=(320,240); 480x640."""
    T = np.eye(4)
    T[:3, :3] = quat_xyzw_to_mat((0.5, 0.5, -0.5, 0.5))
    T[:3, 3] = (0, 0, 0.3)
    K = np.array([[300.0, 0, 320], [0, 300, 240], [0, 0, 1]])
    return T, K, (480, 640)


def random_rotations(rs, n):
    q = rs.randn(n, 4)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.stack([quat_xyzw_to_mat(qi) for qi in q])


def random_crop_scene(rs, n, z_range=(0.3, 1.2), res=(240, 320)):
    """n random poses with a K that keeps the can roughly centred in a `res` image (like a K_crop)."""
    h, w = res
    T = np.tile(np.eye(4, dtype=np.float32), (n, 1, 1))
    T[:, :3, :3] = random_rotations(rs, n)
    z = rs.uniform(*z_range, n)
    T[:, 2, 3] = z
    T[:, 0, 3] = rs.uniform(-0.02, 0.02, n) * z
    T[:, 1, 3] = rs.uniform(-0.02, 0.02, n) * z
    K = np.tile(np.eye(3, dtype=np.float32), (n, 1, 1))
    f = (h / 0.2015 / 1.4) * z  # object diameter 0.2015 m fills ~1/1.4 of the height
    K[:, 0, 0] = f
    K[:, 1, 1] = f * rs.uniform(0.98, 1.02, n)
    K[:, 0, 2] = (w - 1) / 2 + rs.uniform(-3, 3, n)
    K[:, 1, 2] = (h - 1) / 2 + rs.uniform(-3, 3, n)
    return T.astype(np.float32), K.astype(np.float32)


def icosphere(subdiv=2, radius=0.05):
    """Small closed untextured mesh (metres) for edge-case tests."""
    t = (1 + 5**0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    v = np.asarray(v)
    return (v * radius).astype(np.float32), np.asarray(f, np.int32), v.astype(np.float32)
