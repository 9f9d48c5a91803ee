"""Drop-in shim for multi_obj_pose_estimation/darknet_multi.py."""
from singleshotpose_b200.darknet_multi import *  # noqa: F401,F403
