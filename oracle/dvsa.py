"""torch-fp32 CPU restatement of the reference scoring / loss head.  TEST INFRASTRUCTURE ONLY.

Follows reference ``model.py``:
  * ``DVSA.forward``   :517-614  (similarity + masks :523-551, clustering loss :553-577,
                                   frame weighting + margin ranking loss :579-606, picks :608-614)
  * ``postprocess``    :457-474
  * ``record_det``     :477-487

Pinned by ``tests/golden/dvsa_*.npz``: outputs (and autograd gradients) of the reference's own
``DVSA`` class source, exec'd in the build container by ``tests/golden/make_dvsa_golden.py``.
The op sequence below is kept the same as the reference wherever float rounding depends on it
(same ATen calls, same reduction axes), so the forward matches those fixtures bit-for-bit on the
same torch build.
"""
import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-5  # model.py:33


def dvsa_forward(vis_feats, word_feats, entities_length, Na, Nb, Ne, Delta, vis_lam, phase):
    """Returns (D_ind int64 (Na*Ns, Na*Ne), D_sim f32, margin_loss scalar, aux dict).

    vis_feats (Na*Ns*Nb, D), word_feats (Na*Ne, D); differentiable w.r.t. both.
    """
    assert phase in ("train", "eval")
    R, D = vis_feats.shape
    Ns = int(R / Na / Nb)
    dev = vis_feats.device
    lens = [int(x) for x in entities_length]
    div_vec = torch.tensor([1 if x == 0 else x for x in lens], dtype=torch.float, device=dev)

    # column (a', e) is masked for every row when e >= len[a']   (:534-538)
    col_mask = torch.zeros(Na, Ne, dtype=torch.bool, device=dev)
    for a, n in enumerate(lens):
        col_mask[a, n:] = True
    S_mask = col_mask.view(1, Na * Ne).expand(R, Na * Ne)

    S_ = vis_feats @ word_feats.permute(1, 0)  # :548
    S_ = S_.masked_fill(S_mask, 0)             # :551 (in place there; same values and grads)

    aux = {}
    vis_loss = None
    if phase == "train":
        # (:539-546) mask of the Ns x Ns Gram blocks: whole block for padded entities,
        # the diagonal for real ones
        G_mask = torch.zeros(Na, Ne, Ns, Ns, dtype=torch.bool, device=dev)
        eye = torch.eye(Ns, dtype=torch.bool, device=dev)
        for a, n in enumerate(lens):
            G_mask[a, n:] = True
            if n > 0:
                G_mask[a, :n] |= eye
        with torch.no_grad():  # :556-570
            S5 = S_.view(Na, Ns, Nb, Na, Ne)
            S_vis = torch.stack([S5[a, :, :, a, :] for a in range(Na)], 0)  # (Na,Ns,Nb,Ne)
            sim_scr, maxind = S_vis.max(2)
            # quirk: the box index is used WITHOUT its frame/segment row offset
            indarr = maxind.reshape(Na * Ns * Ne).to(torch.long)
            lo = sim_scr.min(1, True)[0]
            hi = sim_scr.max(1, True)[0]
            sim_scr = (sim_scr - lo) / (hi - lo + EPS)
            sim_scr = sim_scr.view(Na, Ns, Ne, -1)
        V = torch.index_select(vis_feats, 0, indarr).view(Na, Ns, Ne, -1)  # :571
        V = V / (torch.norm(V, 2, 3, True) + EPS)
        V = V * sim_scr
        V1 = V.permute(0, 2, 1, 3).contiguous().view(Na * Ne, Ns, -1)
        V2 = V.permute(0, 2, 3, 1).contiguous().view(Na * Ne, -1, Ns)
        G = 1 - torch.bmm(V1, V2).view(Na, Ne, Ns, Ns)
        G = G.masked_fill(G_mask, 0)
        dem = G.nonzero().shape[0]  # data-dependent python int (:576)
        vis_loss = G.sum() / dem
        aux.update(dem=dem, vis_loss=vis_loss.detach().clone(), maxind=maxind, sim_scr=sim_scr)

    # frame weighting + ranking loss (:579-606)
    S = S_.view(Na * Ns, Nb, Na * Ne).max(1)[0].view(Na, Ns, Na * Ne)
    lo = S.min(1, True)[0]
    hi = S.max(1, True)[0]
    S_att = (S - lo) / (hi - lo + EPS)  # carries gradient (the no_grad at :586 is commented out)
    S = S * S_att
    Sf = S.view(Na, Ns, Na, Ne).sum(-1) / div_vec
    Sf_diag = torch.stack([Sf[a, :, a] for a in range(Na)], 0).unsqueeze(2)  # (Na,Ns,1)
    frame_score = (F.relu(Sf - Sf_diag.permute(2, 1, 0) + Delta).mean(0).permute(1, 0)
                   + F.relu(Sf - Sf_diag + Delta).mean(2))
    if phase == "train":
        margin_loss = (frame_score.mean() + vis_lam * vis_loss) * 10
    else:
        margin_loss = frame_score.mean() * 10
    aux.update(frame_score=frame_score.detach().clone(), Sf=Sf.detach().clone())

    D_sim, D_ind = S_.view(Na * Ns, -1, Na * Ne).max(1)  # :608-612
    return D_ind, D_sim, margin_loss, aux


def dvsa_forward_backward(vis_feats, word_feats, entities_length, Na, Nb, Ne, Delta, vis_lam,
                          phase="train"):
    """One reference training step's loss part: L1Loss(margin_loss, 0).backward()
    (model.py:768-772).  Returns numpy dict with picks, loss and dL/dvis, dL/dword."""
    v = torch.as_tensor(vis_feats, dtype=torch.float32).clone().requires_grad_(True)
    w = torch.as_tensor(word_feats, dtype=torch.float32).clone().requires_grad_(True)
    D_ind, D_sim, loss, aux = dvsa_forward(v, w, entities_length, Na, Nb, Ne, Delta, vis_lam,
                                           phase)
    F.l1_loss(loss, torch.zeros_like(loss)).backward()
    return {
        "D_ind": D_ind.numpy().copy(),
        "D_sim": D_sim.detach().numpy().copy(),
        "margin_loss": np.float32(loss.item()),
        "grad_vis": v.grad.numpy().copy(),
        "grad_word": w.grad.numpy().copy(),
        "aux": aux,
    }


def postprocess(D, D_sim, Na, Ns, Nb, Ne):
    """model.py:457-474: keep the a'==a blocks, turn box index into a global row."""
    D4 = np.asarray(D).reshape(Na, Ns, Na, Ne)
    S4 = np.asarray(D_sim).reshape(Na, Ns, Na, Ne)
    out = np.zeros((Na, Ns, Ne), dtype=int)
    out_sim = np.zeros((Na, Ns, Ne))
    for a in range(Na):
        for s in range(Ns):
            out[a, s] = D4[a, s, a] + a * Ns * Nb + s * Nb
            out_sim[a, s] = S4[a, s, a]
    return out, out_sim


def record_det(Nb, vid_entities, D, D_sim, img_ids, infer_boxes):
    """model.py:477-487 -> (img_inds, obj_labels, obj_bboxes, obj_confs) lists."""
    img_inds, labels, boxes, confs = [], [], [], []
    Na, Ns, Ne = D.shape
    for a, ents in enumerate(vid_entities):
        for s in range(Ns):
            for e, ent in enumerate(ents):
                row = D[a][s][e]
                img_inds.append(img_ids[row // Nb])
                labels.append(ent)
                boxes.append(infer_boxes[row])
                confs.append(D_sim[a][s][e])
    return img_inds, labels, boxes, confs
