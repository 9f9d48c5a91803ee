"""singleshotpose_b200 -- B200-native (sm_100a) implementation of the singleshotpose hot path:
Darknet-19/YOLO-pose conv stack forward/backward, RegionLoss head, decode and batched PnP, behind the
reference's Python surface (Darknet, RegionLoss, get_region_boxes, pnp), plus the training-image pipeline of image.py /
dataset.py on the GPU."""
from .darknet import Darknet          # noqa: F401
from .region_loss import RegionLoss   # noqa: F401
from .optim import FlatSGD            # noqa: F401
from .graph import GraphedTrainStep   # noqa: F401
from . import utils, utils_multi, cfg, cfgs, synth, darknet_multi, region_loss_multi  # noqa: F401
from . import image, dataset, checkpoint  # noqa: F401  (GPU image pipeline, its loader class, optimiser-state checkpoints)
