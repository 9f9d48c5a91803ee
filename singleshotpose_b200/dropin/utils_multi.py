"""Drop-in shim for multi_obj_pose_estimation/utils_multi.py."""
from singleshotpose_b200.utils_multi import *  # noqa: F401,F403
from _reexports import *  # noqa: F401,F403,E402  (np, time, os, torch, Image, cv2 ...: the reference's star-import surface)
