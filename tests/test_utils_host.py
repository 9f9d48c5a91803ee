"""Host helpers of the drop-in `utils` module (singleshotpose_b200/utils_host.py) and the drop-in import surface.
tests/golden/make_golden.py check_host_helpers() asserts equality with the reference's utils.py functions while /root/reference is
available; here: known answers, the oracle's pinned corner-confidence restatement, and name resolution of everything the reference's
train.py / valid.py / dataset.py import."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import region_loss_ref as RL
from singleshotpose_b200 import utils as U

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_geometry_helpers_known_answers():
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, -3.0]])
    assert U.calc_pts_diameter(pts) == pytest.approx(np.sqrt(13.0))
    assert U.adi(pts, pts + 0.0) == 0.0 and U.adi(pts[:1], pts[1:2]) == pytest.approx(1.0)
    box = [0.5, 0.5, 0.2, 0.3, 0.8, 0.9]
    assert U.get_2d_bb(box, 10) == pytest.approx([5.0, 5.0, 6.0, 6.0])
    p2 = np.array([[100.0, 300.0, 200.0], [50.0, 250.0, 100.0]])
    assert U.compute_2d_bb(p2) == [200.0, 150.0, 200.0, 200.0]
    assert U.compute_2d_bb_from_orig_pix(p2, 13) == pytest.approx([200 / 640 * 13, 150 / 480 * 13, 200 / 640 * 13, 200 / 480 * 13])
    assert U.sigmoid(0.0) == 0.5 and torch.allclose(U.softmax(torch.tensor([1.0, 1.0])), torch.tensor([0.5, 0.5]))
    c = np.arange(18, dtype="float32").reshape(9, 2)
    assert U.fix_corner_order(c)[:, 0].tolist() == [0, 2, 6, 10, 14, 4, 8, 12, 16]
    assert U.scale_bboxes([[0.5, 0.5, 0.1, 0.2, 7]], 640, 480) == [[320.0, 240.0, 64.0, 96.0, 7]]


def test_corner_confidences_match_pinned_oracle():
    g = torch.Generator().manual_seed(0)
    gt, pr = torch.rand(18, 40, generator=g), torch.rand(18, 40, generator=g) * 0.2 + 0.4
    assert torch.equal(U.corner_confidences(gt.clone(), pr.clone()), RL.corner_confidences_ref(gt.clone(), pr.clone()))
    assert torch.equal(U.corner_confidence(gt[:, 3].clone(), pr[:, 3].clone()), RL.corner_confidence_ref(gt[:, 3].clone(), pr[:, 3].clone()))
    same = U.corner_confidence(gt[:, 0], gt[:, 0].clone())
    assert float(same) == pytest.approx(1.0, abs=1e-4)              # zero distance: (e^2-1)/(e^2-1+1e-5)


def test_file_helpers(tmp_path):
    d = tmp_path / "a" / "b"
    U.makedirs(str(d)); U.makedirs(str(d))
    (d / "x.txt").write_text("1 2 3\n4 5 6\n")
    (tmp_path / "a" / "y.names").write_text("ape \ncan\n")
    assert sorted(os.path.relpath(f, tmp_path) for f in U.get_all_files(str(tmp_path))) == [os.path.join("a", "b", "x.txt"), os.path.join("a", "y.names")]
    assert U.file_lines(str(d / "x.txt")) == 2 and U.load_class_names(str(tmp_path / "a" / "y.names")) == ["ape", "can"]
    cfg = tmp_path / "ape.data"
    cfg.write_text("train = LINEMOD/ape/train.txt\nvalid=LINEMOD/ape/test.txt\n\nmesh = LINEMOD/ape/ape.ply\ngpus = 0,1\n")
    assert U.read_data_cfg(str(cfg)) == {"gpus": "0,1", "num_workers": "10", "train": "LINEMOD/ape/train.txt", "valid": "LINEMOD/ape/test.txt",
                                         "mesh": "LINEMOD/ape/ape.ply"}
    lab = tmp_path / "l.txt"
    rows = np.arange(42, dtype=np.float64).reshape(2, 21) / 50
    np.savetxt(str(lab), rows)
    assert np.array_equal(U.read_truths(str(lab)), rows) and np.array_equal(U.read_pose(str(lab)), rows)
    assert np.array_equal(U.read_truths_args(str(lab)), rows[:, :19].reshape(-1))
    (tmp_path / "e.txt").write_text("")
    assert U.read_truths(str(tmp_path / "e.txt")).size == 0 and U.read_pose(str(tmp_path / "e.txt")).size == 0
    U.logging("hello")


def test_image_helpers(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    a = np.random.default_rng(0).integers(0, 256, (24, 40, 3), dtype=np.uint8)
    im = Image.fromarray(a)
    t = U.image2torch(im)
    assert t.shape == (1, 3, 24, 40) and torch.equal(t[0], torch.from_numpy(a).permute(2, 0, 1).float().div(255))
    for ext in ("png", "jpg", "gif"):
        f = str(tmp_path / ("x." + ext))
        im.save(f)
        assert tuple(U.get_image_size(f)) == (40, 24), ext
    (tmp_path / "junk.bin").write_bytes(b"0" * 64)
    (tmp_path / "short.png").write_bytes(b"\x89PNG")
    assert U.get_image_size(str(tmp_path / "junk.bin")) is None and U.get_image_size(str(tmp_path / "short.png")) is None


def test_dropin_modules_resolve_every_name_the_reference_scripts_use():
    """`from utils import *` / `from cfg import parse_cfg` / `from region_loss import RegionLoss` / `from darknet import Darknet` of
    train.py:18-23, valid.py:10-13 and dataset.py:12 through singleshotpose_b200/dropin (the reference's own dataset.py, image.py
    and MeshPly.py stay on the path behind it)."""
    code = (
        "from utils import *\n"
        "from utils import read_truths_args, read_truths, get_all_files\n"
        "from cfg import parse_cfg\n"
        "from region_loss import RegionLoss\n"
        "from darknet import Darknet\n"
        "names = ['makedirs', 'get_all_files', 'read_data_cfg', 'file_lines', 'logging', 'get_region_boxes', 'pnp', 'compute_projection',\n"
        "         'compute_transformation', 'calcAngularDistance', 'get_3D_corners', 'get_camera_intrinsic', 'convert2cpu', 'calc_pts_diameter',\n"
        "         'compute_2d_bb_from_orig_pix', 'fix_corner_order', 'corner_confidence', 'np', 'torch', 'time', 'os', 'math', 'Variable', 'F']\n"
        "missing = [n for n in names if n not in globals()]\n"
        "assert not missing, missing\n"
        "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(REPO, "singleshotpose_b200", "dropin"), REPO]))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=str(REPO))
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_multi_object_host_helpers(tmp_path):
    from singleshotpose_b200 import utils_multi as M
    assert M.bbox_iou([0.5, 0.5, 0.2, 0.2], [0.5, 0.5, 0.2, 0.2]) == pytest.approx(1.0)
    assert M.bbox_iou([0.2, 0.2, 0.2, 0.2], [0.8, 0.8, 0.2, 0.2]) == 0.0
    assert M.bbox_iou([0, 0, 2, 2], [1, 1, 3, 3], x1y1x2y2=True) == pytest.approx(1.0 / 7.0)
    boxes = [[0.5, 0.5, 0.2, 0.2, 0.9], [0.51, 0.5, 0.2, 0.2, 0.8], [0.1, 0.1, 0.1, 0.1, 0.7], [0.9, 0.9, 0.1, 0.1, 0.0]]
    kept = M.nms(boxes, 0.4)
    assert [b[4] for b in kept] == [0.9, 0.7] and boxes[1][4] == 0                 # the suppressed box is zeroed in place
    assert M.nms([], 0.4) == []
    f = tmp_path / "m.data"
    f.write_text("train = a\n")
    assert M.read_data_cfg(str(f))["gpus"] == "0,1,2,3"
    f.write_text("gpus = 5\ntrain = a\n")
    assert M.read_data_cfg(str(f))["gpus"] == "5"


def test_dropin_multi_modules_resolve():
    code = ("from darknet_multi import Darknet\nfrom utils_multi import *\nfrom cfg import parse_cfg\nfrom region_loss_multi import RegionLoss\n"
            "names = ['get_multi_region_boxes', 'fix_corner_order', 'read_data_cfg', 'logging', 'makedirs', 'get_all_files', 'file_lines', 'pnp',\n"
            "         'compute_projection', 'calcAngularDistance', 'get_3D_corners', 'get_camera_intrinsic', 'bbox_iou', 'nms', 'convert2cpu', 'np', 'time',\n"
            "         'os', 'sys', 'torch', 'calc_pts_diameter']\n"
            "missing = [n for n in names if n not in globals()]\nassert not missing, missing\nprint('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(REPO, "singleshotpose_b200", "dropin"), REPO]))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=str(REPO))
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_dropin_wins_over_the_script_directory(tmp_path):
    """the documented invocation: `python -P script.py` with PYTHONPATH = dropin : repo : checkout.  A fake checkout holds decoy
    utils.py / darknet.py (the names the drop-in replaces) and dataset.py (a name it does not): with -P the drop-in modules win and
    dataset.py still comes from the checkout; without -P the script directory shadows the drop-in."""
    ref = tmp_path / "checkout"
    ref.mkdir()
    for name in ("utils", "darknet", "dataset"):
        (ref / (name + ".py")).write_text("WHO = 'checkout'\n")
    (ref / "main.py").write_text("import utils, darknet, dataset\nprint(getattr(utils, 'WHO', 'dropin'), getattr(darknet, 'WHO', 'dropin'), dataset.WHO)\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(REPO, "singleshotpose_b200", "dropin"), REPO, str(ref)]))
    env.pop("PYTHONSAFEPATH", None)
    safe = subprocess.run([sys.executable, "-P", str(ref / "main.py")], env=env, capture_output=True, text=True)
    assert safe.returncode == 0 and safe.stdout.split() == ["dropin", "dropin", "checkout"], safe.stderr[-1500:]
    plain = subprocess.run([sys.executable, str(ref / "main.py")], env=env, capture_output=True, text=True)
    assert plain.returncode == 0 and plain.stdout.split() == ["checkout", "checkout", "checkout"]

