"""Bridge layers and the checkpoint file format (host-side, no GPU): parameter names and arithmetic
against fixtures produced by the reference's own VisEbd / WordEbd classes
(tests/golden/make_bridge_golden.py), and the `vis_ground_*.pth` layout of model.py:1114-1126."""
import os
import types

import numpy as np
import pytest
import torch

from _golden import GOLDEN
from nafae_b200 import checkpoint
from nafae_b200.bridge import VisEbd, WordEbd

ARGS = types.SimpleNamespace(vis_fc_dim=64, glove_dim=20, word_ebd_dim=16, dropout_rate=0.0)


def _load(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


def test_bridge_layers_match_the_reference_classes():
    z = np.load(os.path.join(GOLDEN, "bridge_ebd.npz"))
    vis, word = VisEbd(ARGS), WordEbd(ARGS)
    # strict loading = identical parameter / buffer names (fc1.weight, bn.running_mean, ...)
    vis.load_state_dict(_load(z, "vis_ebd."), strict=True)
    word.load_state_dict(_load(z, "word_ebd."), strict=True)
    word.train()
    got = word(torch.from_numpy(z["x_word"])).detach().numpy()
    np.testing.assert_array_equal(got, z["word_train"])
    for k in ("running_mean", "running_var", "num_batches_tracked"):
        np.testing.assert_array_equal(word.state_dict()["bn." + k].numpy(), z["word_after." + k])
    vis.eval(), word.eval()
    np.testing.assert_array_equal(vis(torch.from_numpy(z["x_vis"])).detach().numpy(), z["vis_eval"])
    np.testing.assert_array_equal(word(torch.from_numpy(z["x_word"])).detach().numpy(), z["word_eval"])


def _reference_like_state(vis, word):
    """A GroundModel.state_dict() stand-in: detector, head and the dead DVSA.* tensors."""
    state = {"fasterRCNN.RCNN_base.0.weight": torch.randn(4, 3, 3, 3),
             "fasterRCNN.RCNN_top.0.bias": torch.randn(8)}
    state.update({"vis_ebd." + k: v.clone() for k, v in vis.state_dict().items()})
    state.update({"word_ebd." + k: v.clone() for k, v in word.state_dict().items()})
    state.update({"DVSA.slf_attn.w_qs.weight": torch.randn(4, 4), "DVSA.position_enc.weight": torch.randn(5, 4)})
    return state


def test_checkpoint_round_trip_in_the_reference_layout(tmp_path):
    torch.manual_seed(0)
    vis, word = VisEbd(ARGS), WordEbd(ARGS)
    state = _reference_like_state(vis, word)
    ref_path = str(tmp_path / "vis_ground_1_7_0.pth")
    opt = torch.optim.Adam(list(vis.parameters()) + list(word.parameters()), lr=1e-3, weight_decay=1e-5)
    torch.save({"session": 1, "epoch": 7, "model": state, "optimizer": opt.state_dict(),
                "pooling_mode": "align"}, ref_path)                      # what model.py:1118-1126 writes
    ckpt = checkpoint.load_checkpoint(ref_path)
    parts = checkpoint.split_state_dict(ckpt["model"])
    assert list(parts) == ["fasterRCNN", "vis_ebd", "word_ebd", "DVSA"]
    assert list(checkpoint.merge_state_dict(parts)) == list(state)     # lossless, same key order
    vis2, word2 = VisEbd(ARGS), WordEbd(ARGS)
    start_epoch, pooling = checkpoint.load_head(ckpt, vis2, word2)
    assert start_epoch == 8 and pooling == "align"                      # model.py:1042, 1045-1046
    for a, b in ((vis, vis2), (word, word2)):
        for k, v in a.state_dict().items():
            assert torch.equal(v, b.state_dict()[k])
    # written back, the file has the reference's layout and loads strictly by key
    out_path = str(tmp_path / "vis_ground_1_8_0.pth")
    checkpoint.save_checkpoint(out_path, 1, 8, vis2, word2, optimizer=opt, pooling_mode="align",
                               detector_state=parts["fasterRCNN"], dvsa_state=parts["DVSA"])
    back = torch.load(out_path, map_location="cpu", weights_only=False)
    assert sorted(back) == ["epoch", "model", "optimizer", "pooling_mode", "session"]
    assert list(back["model"]) == list(state)
    assert all(torch.equal(back["model"][k], state[k]) for k in state)


def test_checkpoint_rejects_foreign_files(tmp_path):
    p = str(tmp_path / "x.pth")
    torch.save({"weights": 1}, p)
    with pytest.raises(ValueError):
        checkpoint.load_checkpoint(p)
    torch.save({"session": 1, "epoch": 0, "model": {"backbone.w": torch.zeros(1)}}, p)
    with pytest.raises(KeyError):
        checkpoint.load_checkpoint(p)
