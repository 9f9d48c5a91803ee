"""Drop-in shim: lets the reference's unchanged scripts (`from cfg import ...`) resolve to singleshotpose_b200.
Usage: PYTHONPATH=/path/to/repo/singleshotpose_b200/dropin:/path/to/repo python train.py ..."""
from singleshotpose_b200.cfg import *  # noqa: F401,F403
