"""Module-level names the reference's utils.py / utils_multi.py import at the top (utils.py:1-14) and thereby hand to every script
that does `from utils import *` -- valid_multi.py, for one, uses `np` and `time` without importing them.  Optional third-party
modules that are missing simply stay undefined, as they would make the reference's own `import utils` fail."""
import sys      # noqa: F401
import os       # noqa: F401
import time     # noqa: F401
import math     # noqa: F401
import struct   # noqa: F401

import numpy as np                          # noqa: F401
import torch                                # noqa: F401
import torch.nn.functional as F             # noqa: F401
from torch.autograd import Variable         # noqa: F401

try:
    from PIL import Image, ImageDraw, ImageFont   # noqa: F401
except ImportError:                               # pragma: no cover
    pass
try:
    import cv2                                    # noqa: F401
except ImportError:                               # pragma: no cover
    pass
try:
    from scipy import spatial                     # noqa: F401
except ImportError:                               # pragma: no cover
    pass
