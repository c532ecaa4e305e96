"""Step-wrapper parity (reference model.py:762-774): the fused clip + Adam
kernels against torch's own clip_grad_norm_ + torch.optim.Adam, and three HeadTrainer steps against
the same loop written with the oracle's DVSA and stock PyTorch."""
import copy
import types

import numpy as np
import pytest
import torch

gpu = pytest.mark.gpu


def _torch_clip_adam(p, g, steps, lr, wd, clip):
    """Reference: parameters split into uneven tensors like a real model, stock torch ops."""
    cuts = [0, 1000, 1000 + 4096 * 3, p.numel()]
    params = [torch.nn.Parameter(p[a:b].clone()) for a, b in zip(cuts[:-1], cuts[1:])]
    opt = torch.optim.Adam(params, lr=lr, weight_decay=wd)
    norms = []
    for k in range(steps):
        for q, (a, b) in zip(params, zip(cuts[:-1], cuts[1:])):
            q.grad = g[k][a:b].clone()
        norms.append(float(torch.nn.utils.clip_grad_norm_(params, clip)))
        opt.step()
    return torch.cat([q.detach() for q in params]), norms


@gpu
@pytest.mark.parametrize("clip", [100.0, 0.05])
def test_clip_adam_kernels_match_torch(clip):
    from nafae_b200 import _C
    dev = torch.device("cuda:0")
    n = 1000 + 4096 * 3 + 777  # not a multiple of 4: exercises the scalar tail
    gen = torch.Generator(device="cpu").manual_seed(5)
    p0 = torch.randn(n, generator=gen).to(dev)
    grads = [(torch.randn(n, generator=gen) * (0.01 * (k + 1))).to(dev) for k in range(4)]
    want, norms = _torch_clip_adam(p0, grads, 4, 1e-3, 1e-5, clip)
    P = _C.ptr
    p, m, v = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    ws = torch.zeros(int(_C.lib.nafae_clip_adam_workspace_bytes()) // 8, dtype=torch.int64, device=dev)
    for k in range(4):
        g = grads[k].clone()
        st = _C.lib.nafae_clip_adam_step(P(p), P(g), P(m), P(v), n, 1e-3, 0.9, 0.999, 1e-8, 1e-5, clip,
                                         P(ws), ws.numel() * 8, _C.stream())
        assert st == 1, _C.last_error()
        torch.cuda.synchronize()
        got_norm = float(ws.view(torch.float32)[2])
        assert abs(got_norm - norms[k]) <= 1e-5 * norms[k]
        coef = min(1.0, clip / (norms[k] + 1e-6))
        assert torch.allclose(g, grads[k] * coef, rtol=1e-5, atol=0)  # .grad scaled in place
    assert int(ws.view(torch.int32)[1]) == 4  # step count
    assert torch.allclose(p, want, rtol=2e-5, atol=2e-7)
    assert (p - p0).abs().max() > 1e-3


@gpu
def test_head_trainer_three_steps_match_a_stock_pytorch_loop():
    from nafae_b200.bridge import VisEbd, WordEbd
    from nafae_b200.train_step import HeadTrainer
    from oracle import dvsa as odvsa
    dev = torch.device("cuda:0")
    args = types.SimpleNamespace(vis_fc_dim=256, glove_dim=64, word_ebd_dim=128, dropout_rate=0.0)
    Na, Ns, Nb, Ne = 4, 3, 6, 5
    torch.manual_seed(3)
    vis_ebd, word_ebd = VisEbd(args), WordEbd(args)
    ref_vis, ref_word = copy.deepcopy(vis_ebd), copy.deepcopy(word_ebd)
    vis_ebd, word_ebd = vis_ebd.to(dev), word_ebd.to(dev)
    tr = HeadTrainer(vis_ebd, word_ebd, Na, Nb, Ne, Delta=10.0, vis_lam=4.13, lr=1e-3, weight_decay=1e-5,
                     clip=100.0)
    ref_params = list(ref_word.parameters()) + list(ref_vis.parameters())
    opt = torch.optim.Adam([{"params": ref_word.parameters()}, {"params": ref_vis.parameters()}],
                           lr=1e-3, weight_decay=1e-5)
    gen = torch.Generator(device="cpu").manual_seed(11)
    for it in range(3):
        fc = torch.randn(Na * Ns * Nb, 256, generator=gen) * 30
        gl = torch.randn(Na * Ne, 64, generator=gen) * 0.4
        lens = [int(x) for x in torch.randint(1, Ne + 1, (Na,), generator=gen)]
        D_ind, D_sim, loss = tr.step(fc.to(dev), gl.to(dev), lens)
        opt.zero_grad()
        o_ind, o_sim, o_loss, _ = odvsa.dvsa_forward(ref_vis(fc), ref_word(gl), lens, Na, Nb, Ne, 10.0,
                                                     4.13, "train")
        l1 = torch.nn.functional.l1_loss(o_loss, torch.zeros_like(o_loss))
        l1.backward()
        torch.nn.utils.clip_grad_norm_(ref_params, 100.0)
        opt.step()
        assert abs(float(loss) - float(l1)) <= 1e-4 * abs(float(l1))
        live = np.zeros((Na, Ne), bool)
        for a, k in enumerate(lens):
            live[a, :k] = True
        # the embeddings come from cuBLAS here and from the CPU GEMM there (~1e-6 apart): a pick may flip only
        # where the two best boxes tie to that precision, and then the similarities still agree
        lv = live.reshape(-1)
        ours, theirs = D_ind.cpu().numpy()[:, lv], o_ind.numpy()[:, lv]
        flip = ours != theirs
        assert flip.mean() <= 0.02, flip.mean()
        np.testing.assert_allclose(D_sim.detach().cpu().numpy()[:, lv], o_sim.detach().numpy()[:, lv], rtol=1e-4, atol=2e-5)
    got = tr.flat_param.cpu()
    want = torch.cat([q.detach().reshape(-1) for q in ref_params])
    # Adam's first steps move every weight by ~lr * sign-like(m / sqrt(v)) whatever the gradient's
    # size, so a weight whose gradient is rounding noise may legitimately differ by ~lr; everything
    # else must agree to a few 1e-5 of an update that is ~3e-3 after three steps
    diff = (got - want).abs()
    assert float(diff.median()) < 2e-6, float(diff.median())
    assert float((diff > 5e-5).float().mean()) < 2e-3, float((diff > 5e-5).float().mean())
    assert float(diff.max()) < 7e-3
    # the modules' parameters are views of the flat buffer the kernels update
    assert word_ebd.fc1.weight.data_ptr() == tr.flat_param.data_ptr()
