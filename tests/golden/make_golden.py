"""Generate the golden fixtures under tests/golden/ from the REFERENCE ITSELF.

Run in the build container only (needs /root/reference and cv2):
    python tests/golden/make_golden.py
It (1) checks that the oracle restatements agree with the reference's own functions on the
same seeded inputs, and (2) stores the reference outputs as small .npz files so that
tests/test_oracle.py can re-pin the oracle -- and tests/test_*_gpu.py the CUDA path --
on machines where /root/reference does not exist.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)

from oracle.darknet_ref import RefDarknet                      # noqa: E402
from oracle import region_loss_ref as RL                       # noqa: E402
from oracle.decode_ref import get_region_boxes_ref             # noqa: E402
from oracle.pnp_ref import pnp_ref                             # noqa: E402
from singleshotpose_b200 import synth                          # noqa: E402
from singleshotpose_b200.cfgs import write_cfg                 # noqa: E402


def ref_import():
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import darknet as ref_darknet
        import region_loss as ref_rl
        import utils as ref_utils
    os.chdir(cwd)
    return ref_darknet, ref_rl, ref_utils


def populate_and_eval(model, x):
    """Shared recipe (also used by the tests on the oracle model)."""
    bns = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for bn in bns:
        bn.reset_running_stats(); bn.momentum = None
    model.train()
    with torch.no_grad():
        for s in (0, 10, 11):
            model(synth.images(2, seed=s))
        model.eval()
        out = model(x).numpy().copy()
    for bn in bns:
        bn.momentum = 0.1
    model.train()
    return out


def main():
    torch.set_num_threads(os.cpu_count())
    ref_darknet, ref_rl, ref_utils = ref_import()

    # ---------------- network forward / backward ----------------
    torch.manual_seed(0)
    ref_model = ref_darknet.Darknet(os.path.join(REF, "cfg/yolo-pose.cfg"))
    torch.manual_seed(0)
    ora_model = RefDarknet(write_cfg())
    sd_ref, sd_ora = ref_model.state_dict(), ora_model.state_dict()
    assert list(sd_ref.keys()) == list(sd_ora.keys()), "state_dict key order differs"
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_ora[k]), k
    print("seeded init identical for %d tensors" % len(sd_ref))

    x = synth.images(2, seed=0)
    tgt = synth.targets(2, seed=1)
    ref_model.train(); ora_model.train()
    out_ref = ref_model(x)
    out_ora = ora_model(x)
    assert torch.equal(out_ref, out_ora), (out_ref - out_ora).abs().max()
    # loss + backward through the reference network with the restated loss head
    loss, info = RL.region_loss_ref(out_ref, tgt, 20, build_targets=lambda *a: ref_rl.build_targets(*a, 0))
    loss.backward()
    gnorm = np.array([p.grad.double().norm().item() for p in ref_model.parameters()])
    gsum = np.array([p.grad.double().sum().item() for p in ref_model.parameters()])
    names = [n for n, _ in ref_model.named_parameters()]
    first_w_grad = ref_model.models[0][0].weight.grad.numpy().copy()
    last_w_grad = ref_model.models[30][0].weight.grad.numpy().copy()
    train_logits = out_ref.detach().numpy().copy()
    # running-stat update rule (momentum 0.1, unbiased variance) after ONE train forward
    rm = ref_model.models[0][1].running_mean.numpy().copy()
    rv = ref_model.models[29][1].running_var.numpy().copy()
    # populate running stats with the cumulative average of 3 train forwards (momentum=None) so that the
    # eval-mode logits are O(1) (pristine (0,1) stats make eval output ~= the last bias)
    eval_logits = populate_and_eval(ref_model, x)
    assert torch.equal(torch.from_numpy(eval_logits), torch.from_numpy(populate_and_eval(ora_model, x)))
    np.savez_compressed(os.path.join(HERE, "net_b2.npz"), train_logits=train_logits, eval_logits=eval_logits,
                        loss=float(loss), grad_norms=gnorm, grad_sums=gsum, names=np.array(names),
                        first_w_grad=first_w_grad, last_w_grad=last_w_grad[:, :64],
                        running_mean0=rm, running_var29=rv)
    print("net golden: train absmax %.4f eval absmax %.4f loss %.6f" % (
        np.abs(train_logits).max(), np.abs(eval_logits).max(), float(loss)))

    # ---------------- region loss head ----------------
    g = torch.Generator().manual_seed(3)
    B = 6
    out = torch.randn(B, 20, 13, 13, generator=g) * 0.7
    tgt = synth.targets(B, seed=4)
    # make image 0's prediction good at the GT cell so tconf > 0, nCorrect > 0 and some conf_mask zeros exist
    t0 = tgt[0].numpy()
    gi, gj = int(t0[1] * 13), int(t0[2] * 13)
    for k in range(9):
        vx, vy = t0[1 + 2 * k] * 13 - gi, t0[2 + 2 * k] * 13 - gj
        if k == 0:
            vx, vy = np.log(vx / (1 - vx)), np.log(vy / (1 - vy))
        out[0, 2 * k, gj, gi] = float(vx); out[0, 2 * k + 1, gj, gi] = float(vy)
    res = {}
    for epoch in (0, 20):
        o = out.clone().requires_grad_(True)
        loss, info = RL.region_loss_ref(o, tgt, epoch, build_targets=lambda *a: ref_rl.build_targets(*a, 0))
        loss.backward()
        o2 = out.clone().requires_grad_(True)
        loss2, info2 = RL.region_loss_ref(o2, tgt, epoch)
        loss2.backward()
        assert torch.equal(loss, loss2) and torch.equal(o.grad, o2.grad), "restated build_targets differs"
        for k in ("tconf", "conf_mask", "coord_mask"):
            assert torch.equal(info[k], info2[k]), k
        assert (info["nGT"], info["nCorrect"]) == (info2["nGT"], info2["nCorrect"])
        res["loss_e%d" % epoch] = float(loss)
        res["grad_e%d" % epoch] = o.grad.numpy().copy()
        res["parts_e%d" % epoch] = np.array([float(info["loss_x"]), float(info["loss_y"]), float(info["loss_conf"])])
        res["counters_e%d" % epoch] = np.array([info["nGT"], info["nCorrect"], info["nProposals"]])
        if epoch == 20:
            res["tconf"] = info["tconf"].numpy().copy(); res["conf_mask_sqrt"] = info["conf_mask"].numpy().copy()
    # corner_confidence(s) direct pin
    pc = torch.rand(18, 169, generator=g); gt = torch.rand(18, 1, generator=g).repeat(1, 169)
    a = ref_utils.corner_confidences(pc.clone(), gt.clone()); b = RL.corner_confidences_ref(pc, gt)
    assert torch.equal(a, b)
    a1 = ref_utils.corner_confidence(list(gt[:, 0]), pc[:, 5].clone()); b1 = RL.corner_confidence_ref(list(gt[:, 0]), pc[:, 5])
    assert torch.equal(a1, b1)
    np.savez_compressed(os.path.join(HERE, "region_loss.npz"), output=out.numpy(), target=tgt.numpy(), **res)
    print("region golden:", {k: v for k, v in res.items() if k.startswith(("loss", "counters"))})

    # ---------------- decode (get_region_boxes): reference needs .cuda(); patch Tensor.cuda to identity ----
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            box_ref = ref_utils.get_region_boxes(out.clone(), 1, 9)
    finally:
        torch.Tensor.cuda = orig_cuda
    box_ora = get_region_boxes_ref(out.clone(), 1, 9)
    br = np.array([float(v) for v in box_ref]); bo = np.array([float(v) for v in box_ora])
    assert np.array_equal(br, bo), (br, bo)
    np.savez_compressed(os.path.join(HERE, "decode.npz"), output=out.numpy(), box=br)
    print("decode golden box conf %.4f" % br[18])

    # ---------------- PnP: reference utils.pnp (cv2) ----------------
    res = {}
    for sigma, tag in ((0.0, "s0"), (1.0, "s1")):
        pr = synth.pnp_problems(64, sigma=sigma, seed=7)
        Rs, ts = [], []
        worst = 0.0
        for i in range(64):
            R, t = ref_utils.pnp(pr["P3"], pr["uv"][i], pr["K"])
            Ro, to = pnp_ref(pr["P3"], pr["uv"][i], pr["K"])
            ang = np.degrees(np.arccos(np.clip((np.trace(R @ Ro.T) - 1) / 2, -1, 1)))
            worst = max(worst, ang, np.abs(t - to).max() * 1000)
            Rs.append(R); ts.append(t.reshape(3))
        assert worst < 1e-4, worst
        res["uv_" + tag] = pr["uv"]; res["R_" + tag] = np.array(Rs); res["t_" + tag] = np.array(ts)
        print("pnp golden sigma=%g: oracle vs cv2 worst (deg|mm) %.2e" % (sigma, worst))
    np.savez_compressed(os.path.join(HERE, "pnp.npz"), P3=pr["P3"], K=pr["K"], **res)


def main_multi():
    """multi-object head: reference build_targets / bbox_iou / get_multi_region_boxes vs the restatements."""
    from oracle import region_loss_multi_ref as RM
    from oracle.decode_multi_ref import get_multi_region_boxes_ref
    mdir = os.path.join(REF, "multi_obj_pose_estimation")
    sys.path.insert(0, mdir)
    cwd = os.getcwd(); os.chdir(mdir)
    with contextlib.redirect_stdout(io.StringIO()):
        import region_loss_multi as ref_rlm
        import utils_multi as ref_um
    os.chdir(cwd)
    A = synth.MULTI_ANCHORS
    g = torch.Generator().manual_seed(13)
    B = 5
    out = torch.randn(B, 160, 13, 13, generator=g) * 0.7
    tgt = synth.targets_multi(B, seed=14)
    tgt[1, 21:] = 0                                     # image 1: exactly one object
    tgt[3, 21 + 1:21 + 3] = tgt[3, 1:3]                 # image 3: two objects with the same centroid cell
    tgt[3, 21 + 19:21 + 21] = tgt[3, 19:21]             #          and the same extent => same anchor: later GT overwrites
    res = {}
    for epoch in (0, 20):
        o = out.clone().requires_grad_(True)
        loss, info = RM.region_loss_multi_ref(o, tgt, epoch, A, build_targets=lambda *a: ref_rlm.build_targets(*a, 0))
        loss.backward()
        o2 = out.clone().requires_grad_(True)
        loss2, info2 = RM.region_loss_multi_ref(o2, tgt, epoch, A)
        loss2.backward()
        assert torch.equal(loss, loss2) and torch.equal(o.grad, o2.grad), "restated multi build_targets differs"
        for k in ("tconf", "conf_mask", "coord_mask", "tcls"):
            assert torch.equal(info[k], info2[k]), k
        res["loss_e%d" % epoch] = float(loss); res["grad_e%d" % epoch] = o.grad.numpy().copy()
        res["parts_e%d" % epoch] = np.array([float(info[k]) for k in ("loss_x", "loss_y", "loss_conf", "loss_cls")])
        res["counters_e%d" % epoch] = np.array([info["nGT"], info["nCorrect"], info["nProposals"]])
    assert ref_um.bbox_iou([0, 0, 2.0, 3.0], [0, 0, 1.5, 4.0], x1y1x2y2=False) == RM.bbox_iou_ref([0, 0, 2.0, 3.0], [0, 0, 1.5, 4.0])
    np.savez_compressed(os.path.join(HERE, "region_loss_multi.npz"), output=out.numpy(), target=tgt.numpy(), anchors=np.array(A), **res)
    print("multi region golden:", {k: v for k, v in res.items() if k.startswith(("loss", "counters"))})
    # decode
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref_boxes = ref_um.get_multi_region_boxes(out.clone() * 3, 0.05, 13, 9, A, 5, 4, only_objectness=0)
    finally:
        torch.Tensor.cuda = orig_cuda
    ora_boxes = get_multi_region_boxes_ref(out.clone() * 3, 0.05, 13, 9, A, 5, 4, only_objectness=0)
    assert len(ref_boxes) == len(ora_boxes)
    flat, counts = [], []
    for rb, ob in zip(ref_boxes, ora_boxes):
        assert len(rb) == len(ob), (len(rb), len(ob))
        for r1, o1 in zip(rb, ob):
            a1 = np.array([float(v) for v in r1]); a2 = np.array([float(v) for v in o1])
            assert np.array_equal(a1, a2)
            flat.append(a1[:21])
        counts.append(len(rb))
    np.savez_compressed(os.path.join(HERE, "decode_multi.npz"), output=(out * 3).numpy(), boxes=np.array(flat), counts=np.array(counts),
                        anchors=np.array(A), conf_thresh=0.05, correspondingclass=4)
    print("multi decode golden: boxes per image", counts)


def main_multires():
    """the reference network at two of the multi-resolution training shapes (dataset.py:66-90) and the 672^2 test shape
    (yolo-pose.cfg:23-24): train-mode logits of a seeded random-init model, batch 1"""
    torch.set_num_threads(os.cpu_count())
    ref_darknet, _rl, _u = ref_import()
    torch.manual_seed(0)
    ref_model = ref_darknet.Darknet(os.path.join(REF, "cfg/yolo-pose.cfg"))
    torch.manual_seed(0)
    ora_model = RefDarknet(write_cfg())
    ref_model.train(); ora_model.train()
    out = {}
    for (h, w, seed) in ((352, 480, 5), (224, 224, 6), (672, 672, 7)):
        x = synth.images(1, h, w, seed=seed)
        with torch.no_grad():
            o = ref_model(x)
            assert torch.equal(o, ora_model(x))
        out["logits_%dx%d" % (h, w)] = o.numpy().copy()
        print("multires golden %dx%d ->" % (h, w), tuple(o.shape))
    np.savez_compressed(os.path.join(HERE, "net_multires.npz"), **out)


def main_dataset():
    """dataset.py: the reference's listDataset (train mode incl. the multi-resolution schedule, and test mode) run UNMODIFIED on a
    tiny synthetic dataset tree, seeded `random`; outputs = what the DataLoader's default collate would stack."""
    import random
    import tempfile
    from PIL import ImageMath
    if not hasattr(ImageMath, "eval"):
        ImageMath.eval = ImageMath.unsafe_eval
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import dataset as refdataset
    finally:
        os.chdir(cwd)
    to_tensor = lambda img: torch.from_numpy(np.asarray(img).copy()).permute(2, 0, 1).float().div(255)       # transforms.ToTensor
    out = {}
    with tempfile.TemporaryDirectory() as root:
        listfile, bgs = synth.write_linemod_like(root)
        for tag, seen in (("early", 0), ("late", 10 ** 6)):                 # "late": past 70 epochs -> randint(0,19)+7 cells
            random.seed(6)
            ds = refdataset.listDataset(listfile, shape=(96, 96), shuffle=True, transform=to_tensor, train=True, seen=seen, batch_size=2,
                                        num_workers=2, cell_size=8, bg_file_names=bgs)
            for i in range(4):
                img, label = ds[i]
                out["train_%s_img_%d" % (tag, i)] = (img * 255).round().to(torch.uint8).permute(1, 2, 0).numpy()
                assert torch.equal(torch.from_numpy(out["train_%s_img_%d" % (tag, i)]).permute(2, 0, 1).float().div(255), img)
                out["train_%s_label_%d" % (tag, i)] = label.numpy()
            out["train_%s_seen" % tag] = np.array(ds.seen)
        random.seed(9)
        ds = refdataset.listDataset(listfile, shape=(64, 48), shuffle=False, transform=to_tensor, train=False, num_workers=3)
        for i in range(4):
            img, label = ds[i]
            out["test_img_%d" % i] = (img * 255).round().to(torch.uint8).permute(1, 2, 0).numpy()
            out["test_label_%d" % i] = label.numpy()
    np.savez_compressed(os.path.join(HERE, "dataset.npz"), **out)
    print("dataset golden:", sorted(k for k in out if k.endswith("_0")), [out["train_late_img_%d" % i].shape for i in range(4)])


def check_host_helpers():
    """not a golden: asserts, while /root/reference is at hand, that the host helpers of singleshotpose_b200/utils_host.py return
    exactly what the reference's utils.py functions return on the same inputs (tests/test_utils_host.py re-checks known answers)."""
    import tempfile
    from PIL import Image
    from singleshotpose_b200 import utils_host as H
    _d, _rl, R = ref_import()
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(700, 3))
    assert H.calc_pts_diameter(pts) == R.calc_pts_diameter(pts)
    assert H.adi(pts[:300], pts[300:]) == R.adi(pts[:300], pts[300:])
    box = list(rng.random(18)); p2 = rng.random((2, 9)) * 400
    assert H.get_2d_bb(box, 13) == R.get_2d_bb(box, 13) and H.compute_2d_bb(p2) == R.compute_2d_bb(p2)
    assert H.compute_2d_bb_from_orig_pix(p2, 13) == R.compute_2d_bb_from_orig_pix(p2, 13)
    g, q = torch.rand(18, 50), torch.rand(18, 50) * 0.2 + 0.4
    assert torch.equal(H.corner_confidences(g.clone(), q.clone()), R.corner_confidences(g.clone(), q.clone()))
    assert torch.equal(H.corner_confidence(list(g[:, 0]), q[:, 0].clone()), R.corner_confidence(list(g[:, 0]), q[:, 0].clone()))
    assert H.sigmoid(0.3) == R.sigmoid(0.3) and torch.equal(H.softmax(g[:, 1]), R.softmax(g[:, 1]))
    c = rng.random((9, 2)).astype("float32")
    assert np.array_equal(H.fix_corner_order(c), R.fix_corner_order(c))
    with tempfile.TemporaryDirectory() as d:
        listfile, bgs = synth.write_linemod_like(d)
        lab = listfile.replace("train.txt", os.path.join("LINEMOD", "ape", "labels", "000001.txt"))
        assert np.array_equal(H.read_truths(lab), R.read_truths(lab)) and np.array_equal(H.read_truths_args(lab), R.read_truths_args(lab))
        assert np.array_equal(H.read_pose(lab), R.read_pose(lab))
        assert sorted(H.get_all_files(d)) == sorted(R.get_all_files(d)) and H.get_all_files(d) == R.get_all_files(d)
        assert H.file_lines(listfile) == R.file_lines(listfile) == 4 and H.load_class_names(listfile) == R.load_class_names(listfile)
        cfgf = os.path.join(d, "ape.data")
        with open(cfgf, "w") as f:
            f.write("train = a/b.txt\nvalid=c.txt\n\nmesh = m.ply\ngpus = 0,1\n")
        assert H.read_data_cfg(cfgf) == R.read_data_cfg(cfgf)
        im = Image.open(bgs[0]).convert("RGB")
        assert torch.equal(H.image2torch(im), R.image2torch(im))
        im.save(os.path.join(d, "x.jpg")); im.save(os.path.join(d, "x.gif"))
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for fn in (bgs[0], os.path.join(d, "x.jpg"), os.path.join(d, "x.gif"), cfgf):
                assert H.get_image_size(fn) == R.get_image_size(fn), fn
    bb = [[0.1, 0.2, 0.3, 0.4, 9], [0.5, 0.5, 0.1, 0.1, 7]]
    assert H.scale_bboxes(bb, 640, 480) == R.scale_bboxes(bb, 640, 480)
    # multi-object variants (multi_obj_pose_estimation/utils_multi.py): bbox_iou, nms, read_data_cfg with the 4-GPU default
    import copy
    sys.path.insert(0, os.path.join(REF, "multi_obj_pose_estimation"))
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "multi_obj_pose_estimation"))
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import utils_multi as RM
    finally:
        os.chdir(cwd)
    from singleshotpose_b200 import utils_multi as M
    for _ in range(2000):
        b1, b2 = list(rng.random(4)), list(rng.random(4))
        assert M.bbox_iou(b1, b2, False) == RM.bbox_iou(b1, b2, False)
        c1, c2 = [b1[0], b1[1], b1[0] + b1[2], b1[1] + b1[3]], [b2[0], b2[1], b2[0] + b2[2], b2[1] + b2[3]]
        assert M.bbox_iou(c1, c2, True) == RM.bbox_iou(c1, c2, True)
    for _ in range(20):
        boxes = [list(rng.random(4) * 0.5 + 0.25) + [float(rng.random())] for _ in range(12)]
        assert M.nms(copy.deepcopy(boxes), 0.3) == RM.nms(copy.deepcopy(boxes), 0.3)
    with tempfile.TemporaryDirectory() as d:
        for txt in ("train = a\nvalid = b\n", "gpus = 2\ntrain = a\n"):
            f = os.path.join(d, "x.data")
            with open(f, "w") as fh:
                fh.write(txt)
            assert M.read_data_cfg(f) == RM.read_data_cfg(f)
    print("host helpers identical to the reference's utils.py / utils_multi.py functions")


AUG_CASES = [  # seed, (ow, oh), (bw, bh), network shape
    (0, (160, 120), (100, 75), (96, 96)),
    (1, (160, 120), (211, 97), (128, 128)),
    (2, (320, 240), (250, 187), (224, 224)),
    (3, (96, 128), (64, 64), (160, 160)),
]


def main_augment():
    """image pipeline (image.py): the reference's change_background + data_augmentation + fill_truth_detection run UNMODIFIED on
    synthetic samples with a seeded `random`.  One shim: Pillow 12 renamed ImageMath.eval (image.py:124) to unsafe_eval."""
    import random
    import tempfile
    from PIL import Image, ImageMath
    from oracle import augment_ref as A
    if not hasattr(ImageMath, "eval"):
        ImageMath.eval = ImageMath.unsafe_eval
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        import image as refimage
    finally:
        os.chdir(cwd)
    out = {}
    for seed, (ow, oh), (bw, bh), shape in AUG_CASES:
        img, mask, bg = synth.photo_sample(seed, ow, oh, bw, bh)
        rows = synth.label_rows(seed, n=1 + seed % 2)
        random.seed(seed)
        comp = refimage.change_background(Image.fromarray(img), Image.fromarray(mask), Image.fromarray(bg))
        res, flip, dx, dy, sx, sy = refimage.data_augmentation(comp, shape, 0.2, 0.1, 1.5, 1.5)
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            np.savetxt(f, rows)
        label = refimage.fill_truth_detection(f.name, shape[0], shape[1], flip, dx, dy, 1. / sx, 1. / sy, 9, 50)
        os.unlink(f.name)
        res = np.asarray(res)
        # the oracle restatement, same seed
        o_img, o_flip, o_dx, o_dy, o_sx, o_sy = A.data_augmentation(A.change_background(img, mask, bg), shape, 0.2, 0.1, 1.5, 1.5,
                                                                    rng=random.Random(seed))
        assert np.array_equal(o_img, res) and (o_flip, o_dx, o_dy, o_sx, o_sy) == (flip, dx, dy, sx, sy), seed
        assert np.array_equal(A.fill_truth_detection(rows, flip, dx, dy, 1. / sx, 1. / sy, 9, 50), label)
        out["img_%d" % seed] = res
        out["comp_%d" % seed] = np.asarray(comp)
        out["xform_%d" % seed] = np.array([flip, dx, dy, sx, sy], np.float64)
        out["label_%d" % seed] = label
        print("augment golden seed %d: %s -> %s, oracle byte-identical" % (seed, (ow, oh), shape))
    import PIL
    out["pillow_version"] = np.array(PIL.__version__)
    np.savez_compressed(os.path.join(HERE, "augment.npz"), **out)


if __name__ == "__main__":
    if "--augment-only" in sys.argv:
        main_augment()
        sys.exit(0)
    if "--helpers-only" in sys.argv:
        check_host_helpers()
        sys.exit(0)
    if "--dataset-only" in sys.argv:
        main_dataset()
        sys.exit(0)
    if "--multires-only" in sys.argv:
        main_multires()
        sys.exit(0)
    if "--multi-only" not in sys.argv:
        main()
    main_multi()
    main_augment()
    main_multires()
    main_dataset()
    check_host_helpers()
