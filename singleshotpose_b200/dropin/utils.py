"""Drop-in shim: lets the reference's unchanged scripts (`from utils import ...`) resolve to singleshotpose_b200.
Usage (from the reference checkout): PYTHONPATH=$REPO/singleshotpose_b200/dropin:$REPO:$PWD python -P train.py ...
(-P keeps the script's own directory, which holds the reference's modules of the same names, out of sys.path[0])"""
from singleshotpose_b200.utils import *  # noqa: F401,F403
from singleshotpose_b200.utils import (get_region_boxes, pnp, compute_projection, compute_transformation,  # noqa: F401,E402
                                        calcAngularDistance, get_3D_corners, get_camera_intrinsic, convert2cpu, convert2cpu_long)
from _reexports import *  # noqa: F401,F403,E402  (np, time, os, torch, Image, cv2 ...: the reference's star-import surface)
