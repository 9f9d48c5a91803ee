"""CPU oracle for the singleshotpose hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the timed
CPU baseline.  Nothing under ``singleshotpose_b200/`` imports it; the product path raises
if its CUDA library is missing.

Contents (each function cites the reference file:line it restates):
  darknet_ref.py      Darknet(cfg).forward in plain torch-CPU fp32   (darknet.py:59-249)
  region_loss_ref.py  build_targets + RegionLoss.forward             (region_loss.py:9-175)
  decode_ref.py       get_region_boxes                               (utils.py:216-296)
  pnp_ref.py          cv2.solvePnP(ITERATIVE)+Rodrigues restated in numpy fp64 (utils.py:86-100)
  pose_utils_ref.py   compute_projection / calcAngularDistance / get_3D_corners (utils.py:31-84)

Pinning: the reference has no tests or golden vectors.  The oracle
is pinned against outputs of the reference itself, run in the build container by
``tests/golden/make_golden.py`` (imports /root/reference + cv2) and committed under
``tests/golden/*.npz``; ``tests/test_oracle.py`` re-checks the oracle against them
everywhere (no /root/reference needed at test time).
"""
