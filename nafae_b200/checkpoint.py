"""Checkpoint files in the reference's format (model.py:1114-1126, net_utils.py:232-233):

    torch.save({'session', 'epoch', 'model': GroundModel.state_dict(), 'optimizer', 'pooling_mode'})

`GroundModel` (model.py:644-658) owns `fasterRCNN.*` (frozen detector), `vis_ebd.*`, `word_ebd.*`
and `DVSA.*`; the last group are the parameters of `slf_attn` / `position_enc` / `ffn`, which
`DVSA.forward` never uses (dead parameters, grad None).  The head of this package
(`bridge.VisEbd`, `bridge.WordEbd`, `grounding.DVSA`) keeps the same key names, so a reference
checkpoint loads into it and a checkpoint written here loads into the reference's `GroundModel`
(detector and dead `DVSA.*` tensors are carried through untouched when given).
"""
import collections

import torch

COMPONENTS = ("fasterRCNN", "vis_ebd", "word_ebd", "DVSA")
REQUIRED_KEYS = ("session", "epoch", "model")


def split_state_dict(state):
    """{'vis_ebd': {'fc1.weight': ...}, ...}: the model state dict grouped by GroundModel attribute."""
    out = collections.OrderedDict((c, collections.OrderedDict()) for c in COMPONENTS)
    for key, value in state.items():
        head, _, rest = key.partition(".")
        if head not in out or not rest:
            raise KeyError("unexpected key %r in a GroundModel state dict" % key)
        out[head][rest] = value
    return out


def merge_state_dict(parts):
    """Inverse of `split_state_dict` (component order as in GroundModel.__init__)."""
    out = collections.OrderedDict()
    for comp in COMPONENTS:
        for key, value in (parts.get(comp) or {}).items():
            out["%s.%s" % (comp, key)] = value
    return out


def load_checkpoint(path, map_location="cpu"):
    """Read a `vis_ground_{session}_{epoch}_{batch}.pth` file; returns the dict after checking the
    layout.  `pooling_mode`, when present, overrides cfg.POOLING_MODE (model.py:1045-1046)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    if not isinstance(ckpt, dict) or any(k not in ckpt for k in REQUIRED_KEYS):
        raise ValueError("not a NAFAE checkpoint: expected keys %s" % (REQUIRED_KEYS,))
    split_state_dict(ckpt["model"])  # validates the key prefixes
    return ckpt


def load_head(ckpt, vis_ebd, word_ebd, dvsa=None, resume=True):
    """Load `vis_ebd.*` / `word_ebd.*` (strictly) and, if given, whatever `DVSA.*` tensors `dvsa`
    actually has (this package's DVSA has none: the reference's are dead).  Returns (epoch, pooling
    mode or None): with `resume` (training, model.py:1042) the epoch to CONTINUE from, checkpoint
    epoch + 1; with `resume=False` (the val / test phases, model.py:1051) the checkpoint's own epoch."""
    parts = split_state_dict(ckpt["model"])
    vis_ebd.load_state_dict(parts["vis_ebd"], strict=True)
    word_ebd.load_state_dict(parts["word_ebd"], strict=True)
    if dvsa is not None:
        dvsa.load_state_dict(parts["DVSA"], strict=False)
    return int(ckpt["epoch"]) + (1 if resume else 0), ckpt.get("pooling_mode")


def save_checkpoint(path, session, epoch, vis_ebd, word_ebd, optimizer=None, pooling_mode="align",
                    detector_state=None, dvsa_state=None):
    """Write the reference's checkpoint layout.  `detector_state` / `dvsa_state`: the `fasterRCNN.*`
    and dead `DVSA.*` tensors (without prefix) to carry along, e.g. from `split_state_dict` of the
    checkpoint training started from -- required if the file is to load into the reference's
    GroundModel with strict key matching."""
    model = merge_state_dict({"fasterRCNN": detector_state or {}, "vis_ebd": vis_ebd.state_dict(),
                              "word_ebd": word_ebd.state_dict(), "DVSA": dvsa_state or {}})
    torch.save({"session": session, "epoch": epoch, "model": model,
                "optimizer": optimizer.state_dict() if optimizer is not None else None,
                "pooling_mode": pooling_mode}, path)
