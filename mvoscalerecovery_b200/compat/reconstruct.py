"""API surface of the reference's src/reconstruct.py that is computational: per-triangle plane model.  The drawing
and open3d export of the reference (reconstruct.py:91-198) are GUI code and not provided."""
import numpy as np


class Reconstruct:
    def __init__(self, cam=None):
        self.cam = cam

    def check_triangle(self, v, d):
        a = (v[0] - v[1]) * (d[0] - d[1]) > 0
        b = (v[0] - v[2]) * (d[0] - d[2]) > 0
        c = (v[1] - v[2]) * (d[1] - d[2]) > 0
        return [bool(a or b), bool(a or b or c), bool(c)]

    def triangle_model(self, feature3d, triangle_ids):
        """(T,4) rows [unit normal with n_y >= 0 flipped as reconstruct.py:83-85 does, height = 1/|n|] of n = P^-1 1."""
        f3, tri = np.asarray(feature3d, dtype=float), np.asarray(triangle_ids)
        p0, e1, e2 = f3[tri[:, 0]], f3[tri[:, 1]] - f3[tri[:, 0]], f3[tri[:, 2]] - f3[tri[:, 0]]
        c = np.cross(e1, e2)
        det = np.einsum("ij,ij->i", p0, c)
        n = c / det[:, None]
        ln = np.linalg.norm(n, axis=1)
        unit = n / ln[:, None]
        unit[unit[:, 1] < 0] *= -1
        return np.hstack([unit, (1.0 / ln)[:, None]])
