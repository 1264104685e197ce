"""API surface of the reference's src/graph.py: the depth-order graph check.  On the per-frame path the check runs
inside the CUDA frame kernel (vote per star, gstar.cuh: consume_vote); this module gives the same results for
caller-supplied triangles as vectorised numpy, with the reference's names and signatures."""
import numpy as np


def triangle(edge_potential):
    """8x8 potential: rows = vertex states (v0,v1,v2), columns = edge observations (a:0-1, b:1-2, c:0-2) (graph.py:110-122)."""
    ep = np.asarray(edge_potential, dtype=float)
    bits = np.array([[(k >> 2) & 1, (k >> 1) & 1, k & 1] for k in range(8)])
    s, o = bits[:, None, :], bits[None, :, :]
    return ep[s[..., 0] * 2 + s[..., 1], o[..., 0]] * ep[s[..., 1] * 2 + s[..., 2], o[..., 1]] * ep[s[..., 0] * 2 + s[..., 2], o[..., 2]]


def check_triangle(v, d):
    """Index 4a+2b+c of the edge observations, each (v_i - v_j)(d_i - d_j) < 0 (graph.py:124-129)."""
    a = int((v[0] - v[1]) * (d[0] - d[1]) < 0)
    b = int((v[1] - v[2]) * (d[1] - d[2]) < 0)
    c = int((v[0] - v[2]) * (d[0] - d[2]) < 0)
    return a * 4 + b * 2 + c


def get_assemble_probability(probs):
    probs = np.asarray(probs)
    return np.sum(probs > 0.6) / len(probs)


def _vertex_table(tp):
    """p_k(idx) = sum of the potential over states with vertex k set / sum over all states (graph.py:134-145)."""
    sel = np.array([[(k >> 2) & 1, (k >> 1) & 1, k & 1] for k in range(8)], dtype=bool)
    z = tp.sum(0)
    return np.stack([tp[sel[:, k]].sum(0) / z for k in range(3)], 1)


def get_probability(v, d, tp):
    return list(_vertex_table(np.asarray(tp, dtype=float))[check_triangle(v, d)])


def bool2id(flag):
    return np.nonzero(np.asarray(flag))[0]


class GraphChecker:
    def __init__(self, edge_potential):
        self.triangle_potential = triangle(edge_potential)

    def find_inliers(self, feature3d, feature2d, triangle_ids):
        """keep[i] = (#incident triangles voting p > 0.6) / (#incident triangles) > 0.5; no triangle -> False (graph.py:18-36)."""
        f3, f2, tri = np.asarray(feature3d), np.asarray(feature2d), np.asarray(triangle_ids)
        n = f3.shape[0]
        if tri.size == 0:
            return np.zeros(n, dtype=bool)
        v, d = f2[tri, 1], f3[tri, 2]
        a = (v[:, 0] - v[:, 1]) * (d[:, 0] - d[:, 1]) < 0
        b = (v[:, 1] - v[:, 2]) * (d[:, 1] - d[:, 2]) < 0
        c = (v[:, 0] - v[:, 2]) * (d[:, 0] - d[:, 2]) < 0
        votes = _vertex_table(self.triangle_potential)[a * 4 + b * 2 + c] > 0.6
        total = np.bincount(tri.reshape(-1), minlength=n)
        passed = np.bincount(tri.reshape(-1), weights=votes.reshape(-1), minlength=n)
        return 2 * passed > total


class GraphGrow:
    """Region growing over the triangle mesh (graph.py:39-107): triangles are linked across shared edges when their pitch
    differs by less than ``threshold_angle`` degrees and their inverse height by less than 0.4 x the median inverse height;
    the largest linked region that contains a flat (< -85 deg), low triangle is returned.

    The reference grows from 100 seeds drawn with ``np.random.choice`` and keeps the largest region found.  The link test
    is symmetric, so the region of a seed is the connected component of the seed and does not depend on the traversal;
    here EVERY flat seed is tried, in ascending order (a superset of any random draw, and deterministic): the result is the
    largest component, the first one on ties, listed in depth-first order from its smallest seed.  Constructed by the live
    estimator (rescale.py:33); its call there is commented out (rescale.py:99)."""

    def __init__(self, threshold_angle=8):
        self.threshold_angle = threshold_angle
        self.threshold_height = 0.2
        self.graph = []
        self.proposal = []
        self.height_invs = []
        self.angles = []

    def graph_construction(self, triangle_ids):
        """Triangle adjacency across shared edges, each triangle linked to the EARLIER triangle on that edge (graph.py:47-71)."""
        tri = np.asarray(triangle_ids)
        graph = [[] for _ in range(tri.shape[0])]
        first = {}
        for i, (a, b, c) in enumerate(tri.tolist()):
            for e in ((a, b), (a, c), (b, c)):
                key = e if e[0] < e[1] else (e[1], e[0])
                j = first.get(key)
                if j is None:
                    first[key] = i
                else:
                    graph[i].append(j)
                    graph[j].append(i)
        self.graph = graph

    def check(self, i, j):
        return bool(np.abs(self.angles[i] - self.angles[j]) < self.threshold_angle
                    and np.abs(self.height_invs[i] - self.height_invs[j]) < self.threshold_height)

    def expend(self, i, proposal):
        """Depth-first growth from triangle i (graph.py:78-82), iterative (the reference recurses)."""
        stack = [(i, iter(self.graph[i]))]
        seen = set(proposal)
        while stack:
            node, it = stack[-1]
            for j in it:
                if j not in seen and self.check(node, j):
                    seen.add(j)
                    proposal.append(j)
                    stack.append((j, iter(self.graph[j])))
                    break
            else:
                stack.pop()

    def process(self, triangle_ids, heights, angles):
        self.graph_construction(triangle_ids)
        self.height_invs = 1 / np.asarray(heights, dtype=float)
        self.angles = np.asarray(angles, dtype=float)
        flat = self.angles < -85
        low = self.height_invs < np.median(self.height_invs[self.angles < -80])
        seeds = bool2id(flat & low)
        self.threshold_height = 0.4 * np.median(self.height_invs)
        best, done = [], set()
        for seed in seeds.tolist():
            if seed in done:
                continue
            proposal = [seed]
            self.expend(seed, proposal)
            done.update(proposal)
            if len(proposal) > len(best):
                best = proposal
        return best
