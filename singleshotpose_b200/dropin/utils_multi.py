"""Drop-in shim for multi_obj_pose_estimation/utils_multi.py."""
from singleshotpose_b200.utils_multi import *  # noqa: F401,F403
