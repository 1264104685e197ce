"""Module constants of the reference's src/param.py:30-39 (KITTI 00 calibration)."""
img_w = 1241.0
img_h = 376.0
img_fx = 718.856
img_fy = 718.856
img_cx = 607.1928
img_cy = 185.2157
camera_h = 1.75
minimum_feature_for_scale = 100
minimum_feature_for_tracking = 1500
fast_threshold = 25
