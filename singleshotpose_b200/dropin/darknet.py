"""Drop-in shim: lets the reference's unchanged scripts (`from darknet import ...`) resolve to singleshotpose_b200.
Usage (from the reference checkout): PYTHONPATH=$REPO/singleshotpose_b200/dropin:$REPO:$PWD python -P train.py ...
(-P keeps the script's own directory, which holds the reference's modules of the same names, out of sys.path[0])"""
from singleshotpose_b200.darknet import *  # noqa: F401,F403
