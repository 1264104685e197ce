"""``run_ransac`` with the signature, draw order and bookkeeping of the reference's src/thirdparty/Ransac/ransac.py:3-23
(Falcon Dai's generic RANSAC, MIT): hypotheses are evaluated one after the other, the first strictly larger inlier count is
kept, and the loop ends at the first count above ``goal_inliers`` (unless ``stop_at_goal`` is false).

Generic over the caller's ``estimate`` / ``is_inlier`` callbacks, hence host code.  The plane-fitting instance the estimator
uses runs on the GPU: per frame as stage 4 of the fused kernel, on explicit point lists as mvosr_ransac_planes
(estimate_road_norm.get_pitch_ransac); both draw from the Philox position stream, and this module's ``random`` global can be
replaced by a sampler with ``seed`` / ``sample`` methods to make the host loop follow the same stream."""
import random


def _count_inliers(model, items, is_inlier):
    return sum(bool(is_inlier(model, item)) for item in items)


def run_ransac(data, estimate, is_inlier, sample_size, goal_inliers, max_iterations, stop_at_goal=True, random_seed=None):
    random.seed(random_seed)                     # the reference re-seeds on every call (None: OS entropy)
    items = list(data)
    winner = (None, 0)                           # (model, inlier count)
    iteration = 0
    while iteration < max_iterations:
        iteration += 1
        candidate = estimate(random.sample(items, int(sample_size)))
        support = _count_inliers(candidate, items, is_inlier)
        if support <= winner[1]:
            continue
        winner = (candidate, support)
        if stop_at_goal and support > goal_inliers:
            break
    return winner
