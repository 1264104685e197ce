"""run_ransac with the reference's signature and bookkeeping (src/thirdparty/Ransac/ransac.py:3-23): sequential
hypotheses, keep the first strictly larger inlier count, stop at the first count above the goal.  Generic over the
caller's estimate / is_inlier callbacks, hence host code; the plane-fitting instance used per frame runs on the GPU
(stage 4 of the frame kernel) with the Philox position stream this module's ``random`` can be replaced by."""
import random


def run_ransac(data, estimate, is_inlier, sample_size, goal_inliers, max_iterations, stop_at_goal=True, random_seed=None):
    best_model, best_ic = None, 0
    random.seed(random_seed)
    items = list(data)
    for _ in range(max_iterations):
        model = estimate(random.sample(items, int(sample_size)))
        ic = sum(1 for x in items if is_inlier(model, x))
        if ic > best_ic:
            best_model, best_ic = model, ic
            if stop_at_goal and ic > goal_inliers:
                break
    return best_model, best_ic
