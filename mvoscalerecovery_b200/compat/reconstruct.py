"""API surface of the reference's src/reconstruct.py that is computational: the depth-order votes, the per-triangle plane
model and the dense depth map.  The drawing code and the open3d export (reconstruct.py:108-117,119-198) are GUI code and
are not provided."""
import numpy as np


class Reconstruct:
    def __init__(self, cam=None):
        self.threshold = 1
        self.cam = cam
        self.pixel = self.pixel_ori = None
        if cam is not None:
            # integer pixel grid (u, v) and its normalised image coordinates (reconstruct.py:23-36)
            v, u = np.mgrid[0:cam.height, 0:cam.width]
            self.pixel_ori = np.stack([u, v], -1).reshape(-1, 2)
            self.pixel = np.stack([(u - cam.cx) / cam.fx, (v - cam.cy) / cam.fy], -1).astype(float)

    def check_triangle(self, v, d):
        a = (v[0] - v[1]) * (d[0] - d[1]) > 0
        b = (v[0] - v[2]) * (d[0] - d[2]) > 0
        c = (v[1] - v[2]) * (d[1] - d[2]) > 0
        return [bool(a or b), bool(a or b or c), bool(c)]

    def find_outliers(self, feature3d, feature2d, triangle_ids):
        """1 - (number of triangles flagging the vertex) (reconstruct.py:58-69); votes on the GPU."""
        import _gpu
        flagged, _ = _gpu.triangle_votes(triangle_ids, np.asarray(feature2d)[:, 1], np.asarray(feature3d)[:, 2])
        return 1.0 - flagged

    def triangle_model(self, feature3d, triangle_ids):
        """(T,4) rows [unit normal of n = P^-1 1 flipped to n_y >= 0, height = +-1/|n| flipped with it] (reconstruct.py:70-90);
        planes on the GPU."""
        import _gpu
        normal, height, _ = _gpu.triangle_planes(triangle_ids, feature3d)
        unit = normal / np.sqrt(np.sum(normal * normal, 1))[:, None]
        flip = unit[:, 1] < 0
        unit[flip] *= -1
        return np.hstack([unit, np.where(flip, -height, height)[:, None]])

    def depth_generate(self, tri, datas, img=None):
        """Dense depth map: every pixel inside the triangulation gets the depth of its triangle's plane (reconstruct.py:91-107).
        ``tri``: a scipy.spatial.Delaunay-like object (``.simplices``, ``.points``) or a (simplices, points2d) pair; ``datas``:
        the rows of ``triangle_model``.  Returns the (H, W) depth map (0 outside the mesh) and stores ``self.pixel_tris``.
        The reference's point-cloud export (open3d) and imshow are GUI code and not provided."""
        import _gpu
        simplices, pts = (tri.simplices, tri.points) if hasattr(tri, "simplices") else tri
        depth, self.pixel_tris = _gpu.depth_from_mesh(self.cam, simplices, pts, datas)
        return depth
