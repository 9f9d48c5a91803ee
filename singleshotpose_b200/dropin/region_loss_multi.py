"""Drop-in shim for multi_obj_pose_estimation/region_loss_multi.py."""
from singleshotpose_b200.region_loss_multi import *  # noqa: F401,F403
