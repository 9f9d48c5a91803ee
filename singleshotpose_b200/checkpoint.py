"""Training checkpoint = Darknet `.weights` file (the reference's own format, darknet.py:350-394) + optimiser state.

The reference saves model weights only (train.py:409), so a resumed run restarts momentum from zero and the
learning-rate schedule from `model.seen` (train.py:345).  SURVEY 8f.4 asks for the optimiser state as well; it is kept
in a side file next to the `.weights` so that every reference tool still reads the weights unchanged.
"""
from __future__ import annotations

import os

import torch

_SUFFIX = ".optim.pt"


def save_checkpoint(model, optimizer, weightfile):
    """model.save_weights(weightfile) + `<weightfile>.optim.pt` (optimiser state_dict, model.seen, model.iter)."""
    model.save_weights(weightfile)
    sd = optimizer.state_dict()
    # torch.optim.SGD.state_dict() hands out the LIVE per-parameter dicts: build new ones instead of moving its buffers to the CPU
    state = {i: {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in ent.items()} for i, ent in sd.get("state", {}).items()}
    sd = {"state": state, "param_groups": sd["param_groups"]}
    tmp = weightfile + _SUFFIX + ".tmp"
    torch.save({"optimizer": sd, "seen": int(model.seen), "iter": int(getattr(model, "iter", 0))}, tmp)
    os.replace(tmp, weightfile + _SUFFIX)           # atomic: a crash mid-save never leaves a truncated state file


def load_checkpoint(model, optimizer, weightfile, strict=True):
    """model.load_weights(weightfile) and, if present, the optimiser state saved beside it.  Returns True when the optimiser
    state was restored.  strict: a missing state file is an error (otherwise momentum silently restarts from zero)."""
    model.load_weights(weightfile)
    side = weightfile + _SUFFIX
    if not os.path.exists(side):
        if strict:
            raise FileNotFoundError("no optimiser state %s beside the weights (pass strict=False to resume weights only)" % side)
        return False
    blob = torch.load(side, map_location="cpu", weights_only=True)
    optimizer.load_state_dict(blob["optimizer"])
    model.seen, model.iter = int(blob["seen"]), int(blob["iter"])
    return True
