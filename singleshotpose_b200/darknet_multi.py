"""``Darknet`` for yolo-pose-multi.cfg -- drop-in for reference multi_obj_pose_estimation/darknet_multi.py (identical to
darknet.py except that the region block carries the multi-object RegionLoss and its anchors)."""
from .darknet import Darknet as _Darknet
from .region_loss_multi import RegionLoss


class Darknet(_Darknet):
    _region_loss_cls = RegionLoss
