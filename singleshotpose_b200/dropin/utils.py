"""Drop-in shim: lets the reference's unchanged scripts (`from utils import ...`) resolve to singleshotpose_b200.
Usage: PYTHONPATH=/path/to/repo/singleshotpose_b200/dropin:/path/to/repo python train.py ..."""
from singleshotpose_b200.utils import *  # noqa: F401,F403
from singleshotpose_b200.utils import (get_region_boxes, pnp, compute_projection, compute_transformation,  # noqa: F401,E402
                                        calcAngularDistance, get_3D_corners, get_camera_intrinsic, convert2cpu, convert2cpu_long)
from _reexports import *  # noqa: F401,F403,E402  (np, time, os, torch, Image, cv2 ...: the reference's star-import surface)
