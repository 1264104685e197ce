"""Drop-in module set: put this directory FIRST on sys.path (PYTHONPATH) and the reference's
``src/main.py`` / ``src/main_offline.py`` import these ``rescale``, ``scale_calculator``, ``estimate_road_norm``,
``graph``, ``reconstruct``, ``param`` and ``thirdparty.Ransac.ransac`` modules instead of their own, unchanged.

``rescale.ScaleEstimator.scale_calculation`` -- the call both mains make once per frame -- runs the CUDA path
(libmvosr.so, one frame per launch); the batched entry points of ``mvoscalerecovery_b200.batch`` are the fast way in.
The remaining functions mirror the reference's API surface as small host-side numpy
helpers; none of them is on the per-frame path of either main."""
