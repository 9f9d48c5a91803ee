"""Drop-in shim: lets the reference's unchanged scripts (`from darknet import ...`) resolve to singleshotpose_b200.
Usage: PYTHONPATH=/path/to/repo/singleshotpose_b200/dropin:/path/to/repo python train.py ..."""
from singleshotpose_b200.darknet import *  # noqa: F401,F403
