"""Generate tests/golden/dvsa_*.npz by executing the REFERENCE's own DVSA source.

Run in the build container only (needs /root/reference):

    python tests/golden/make_dvsa_golden.py

What it does: reads ``/root/reference/model.py`` lines 457-614 (postprocess, record_det, class
DVSA) -- nothing is copied into this repo -- and exec's them with the shims SURVEY.md section 8(c)
lists: ``torch.uint8 -> torch.bool`` (torch >= 1.2 rejects byte masks), stub globals ``cfg``,
``device``, ``EPS``, and dummy ``MultiHeadAttention`` / ``position_encoding_init`` (constructed by
``DVSA.__init__`` but never used by ``forward``).  It then runs forward + ``L1Loss(loss, 0)``
backward (model.py:768-772) on seeded inputs and stores inputs' *recipe* (seed, shapes) together
with the reference outputs.

Inputs are built from ``np.random.RandomState`` (bit-stable across numpy versions) using only
IEEE-exact operations (scale, clip, float32 rounding), so tests can rebuild them anywhere.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NAFAE_REFERENCE", "/root/reference")

CASES = [
    # name, Na, Ns, Nb, Ne, D, lens, Delta, vis_lam, phase, seed, dup_rows
    dict(name="small_train", Na=3, Ns=4, Nb=5, Ne=4, D=32, lens=[2, 0, 4], Delta=10.0,
         vis_lam=4.13, phase="train", seed=11, dup=0),
    dict(name="small_train_lowdelta", Na=3, Ns=4, Nb=5, Ne=4, D=32, lens=[1, 3, 2], Delta=0.05,
         vis_lam=1.0, phase="train", seed=12, dup=0),
    dict(name="ties_train", Na=2, Ns=3, Nb=6, Ne=3, D=16, lens=[3, 2], Delta=0.1,
         vis_lam=4.13, phase="train", seed=13, dup=3),
    dict(name="single_frame_train", Na=2, Ns=1, Nb=4, Ne=3, D=16, lens=[2, 1], Delta=10.0,
         vis_lam=4.13, phase="train", seed=14, dup=0),
    dict(name="eval_seg", Na=1, Ns=7, Nb=20, Ne=13, D=64, lens=[4], Delta=5.0,
         vis_lam=1.0, phase="eval", seed=15, dup=0),
    dict(name="cfg1_eval", Na=1, Ns=5, Nb=20, Ne=13, D=512, lens=[4], Delta=5.0,
         vis_lam=1.0, phase="eval", seed=16, dup=0),
    dict(name="cfg2_train", Na=8, Ns=5, Nb=20, Ne=13, D=512,
         lens=[2, 3, 0, 1, 5, 2, 13, 4], Delta=10.0, vis_lam=4.13, phase="train", seed=17,
         dup=0),
    dict(name="cfg4_eval", Na=1, Ns=32, Nb=100, Ne=13, D=512, lens=[6], Delta=5.0,
         vis_lam=1.0, phase="eval", seed=18, dup=0),
]


def make_inputs(case):
    """Seeded, platform-independent inputs (tanh-bounded like VisEbd/WordEbd outputs)."""
    rs = np.random.RandomState(case["seed"])
    R = case["Na"] * case["Ns"] * case["Nb"]
    vis = np.clip(rs.standard_normal((R, case["D"])) * 0.5, -1, 1).astype(np.float32)
    word = np.clip(rs.standard_normal((case["Na"] * case["Ne"], case["D"])) * 0.5, -1, 1)
    word = word.astype(np.float32)
    if case["dup"]:
        # zero-padded RoIs give identical feature rows: exact ties in the max over boxes
        Nb = case["Nb"]
        for f in range(case["Na"] * case["Ns"]):
            vis[f * Nb + Nb - case["dup"]: (f + 1) * Nb] = vis[f * Nb + Nb - case["dup"] - 1]
    return vis, word


def load_reference_dvsa(Nb):
    with open(os.path.join(REF, "model.py")) as fh:
        lines = fh.readlines()
    src = "".join(lines[456:614])  # model.py:457-614
    src = src.replace("torch.uint8", "torch.bool")

    class _Dummy(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    cfg = types.SimpleNamespace(TEST=types.SimpleNamespace(RPN_POST_NMS_TOP_N=Nb))
    ns = dict(torch=torch, nn=nn, np=np, F=F, cfg=cfg, device=torch.device("cpu"), EPS=1e-5,
              MultiHeadAttention=_Dummy,
              position_encoding_init=lambda n, d: torch.zeros(n, d))
    exec(compile(src, os.path.join(REF, "model.py"), "exec"), ns)
    return ns


def run_case(case):
    ns = load_reference_dvsa(case["Nb"])
    args = types.SimpleNamespace(batch_size=case["Na"], batch_size_val=case["Na"],
                                 max_ent_len=case["Ne"], Delta=case["Delta"],
                                 vis_lam=case["vis_lam"], n_head=1, word_ebd_dim=8, d_k=4, d_v=4,
                                 dropout_rate=0.1, n_position=4, sample_num=case["Ns"])
    torch.manual_seed(0)
    dvsa = ns["DVSA"](args, ns["cfg"])
    (dvsa.init_train if case["phase"] == "train" else dvsa.init_eval)()
    vis_np, word_np = make_inputs(case)
    vis = torch.from_numpy(vis_np.copy()).requires_grad_(True)
    word = torch.from_numpy(word_np.copy()).requires_grad_(True)
    D_ind, D_sim, loss = dvsa(vis, word, list(case["lens"]))
    nn.L1Loss()(loss, torch.zeros_like(loss)).backward()  # model.py:771-772
    D_pp, D_sim_pp = ns["postprocess"](D_ind.numpy(), D_sim.detach().numpy(), case["Na"],
                                       case["Ns"], case["Nb"], case["Ne"])
    out = dict(
        D_ind=D_ind.numpy().astype(np.int16),
        D_sim=D_sim.detach().numpy(),
        margin_loss=np.float32(loss.item()),
        post_D=D_pp.astype(np.int32),
        post_D_sim=D_sim_pp.astype(np.float32),
    )
    gv, gw = vis.grad.numpy(), word.grad.numpy()
    if gv.size + gw.size <= 20000:
        out["grad_vis"] = gv
        out["grad_word"] = gw
    else:
        # full-size cases: keep the fixture small -- row sums, abs sums and 4 fixed projections
        rs = np.random.RandomState(1000 + case["seed"])
        P = rs.standard_normal((case["D"], 4)).astype(np.float32)
        out["grad_proj"] = P
        out["grad_vis_proj"] = (gv.astype(np.float64) @ P).astype(np.float32)
        out["grad_word_proj"] = (gw.astype(np.float64) @ P).astype(np.float32)
        out["grad_vis_abs"] = np.abs(gv).sum(1).astype(np.float32)
        out["grad_word_abs"] = np.abs(gw).sum(1).astype(np.float32)
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree %s not found: fixtures can only be regenerated in the build "
                 "container" % REF)
    for case in CASES:
        out = run_case(case)
        meta = {k: np.asarray(v) for k, v in case.items() if k not in ("name", "phase")}
        meta["phase_train"] = np.asarray(case["phase"] == "train")
        path = os.path.join(HERE, "dvsa_%s.npz" % case["name"])
        np.savez_compressed(path, **meta, **out)
        print("%-24s loss=%.6f  %6.1f KB" % (case["name"], float(out["margin_loss"]),
                                             os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
