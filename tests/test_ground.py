"""Similarity / loss head parity: fused CUDA kernels vs the reference-generated fixtures and the
torch CPU restatement (oracle/dvsa.py)."""
import types

import numpy as np
import pytest
import torch

from _golden import check_grads_against_fixture, dvsa_case_names, dvsa_inputs, load_dvsa_case
from nafae_b200 import synth
from oracle import dvsa as odvsa

gpu = pytest.mark.gpu
# fp32 FMA path; north_star: <= 1e-4 relative for losses / similarities / gradients
RTOL = 1e-4


def _run(case, vis, word, lens=None):
    from nafae_b200.grounding import ground
    dev = torch.device("cuda:0")
    v = torch.from_numpy(vis).to(dev).requires_grad_(True)
    w = torch.from_numpy(word).to(dev).requires_grad_(True)
    D_ind, D_sim, loss = ground(v, w, case["lens"] if lens is None else lens, case["Na"],
                                case["Nb"], case["Ne"], case["Delta"], case["vis_lam"],
                                case["phase"] == "train")
    torch.nn.functional.l1_loss(loss, torch.zeros_like(loss)).backward()  # model.py:771-772
    return (D_ind.cpu().numpy(), D_sim.cpu().numpy(), float(loss), v.grad.cpu().numpy(),
            w.grad.cpu().numpy())


def _live_mask(case):
    m = np.zeros((case["Na"], case["Ne"]), bool)
    for a, n in enumerate(case["lens"]):
        m[a, :n] = True
    return m.reshape(-1)


@gpu
@pytest.mark.parametrize("name", dvsa_case_names())
def test_matches_reference_fixture(name):
    case, z = load_dvsa_case(name)
    vis, word = dvsa_inputs(case)
    D_ind, D_sim, loss, gv, gw = _run(case, vis, word)
    live = _live_mask(case)
    # picks: bit-exact on every column (masked columns are all-zero ties -> index 0 in both)
    np.testing.assert_array_equal(D_ind, z["D_ind"].astype(np.int64))
    np.testing.assert_allclose(D_sim, z["D_sim"], rtol=RTOL, atol=1e-5)
    assert not D_sim[:, ~live].any()
    if np.isnan(z["margin_loss"]):
        assert np.isnan(loss)
        return
    np.testing.assert_allclose(loss, z["margin_loss"], rtol=RTOL)
    check_grads_against_fixture(z, gv, gw, rtol=RTOL, atol_scale=1e-5)


@gpu
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_batches_match_oracle(seed):
    """cfg2 shapes with histogram-sampled query counts (incl. empty segments)."""
    c = synth.CONFIGS["cfg2"]
    rs = np.random.RandomState(100 + seed)
    lens = synth.entity_lengths(rs, c["Na"], c["Ne"])
    vis = synth.embeddings(rs, c["Na"] * c["Ns"] * c["Nb"], c["D"])
    word = synth.embeddings(rs, c["Na"] * c["Ne"], c["D"])
    case = dict(Na=c["Na"], Nb=c["Nb"], Ne=c["Ne"], Delta=c["Delta"], vis_lam=c["vis_lam"],
                phase="train", lens=lens)
    D_ind, D_sim, loss, gv, gw = _run(case, vis, word)
    ref = odvsa.dvsa_forward_backward(vis, word, lens, c["Na"], c["Nb"], c["Ne"], c["Delta"],
                                      c["vis_lam"], "train")
    live = _live_mask(case)
    np.testing.assert_array_equal(D_ind[:, live], ref["D_ind"][:, live])
    np.testing.assert_allclose(D_sim, ref["D_sim"], rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(loss, ref["margin_loss"], rtol=RTOL)
    for got, want in ((gv, ref["grad_vis"]), (gw, ref["grad_word"])):
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-5 * np.abs(want).max())


@gpu
@pytest.mark.parametrize("Na,Ns,lens", [(1, 800, [9]), (2, 64, [13, 2])])
def test_eval_shapes_the_reference_runs(Na, Ns, lens):
    """Evaluation runs whole videos: up to 800 frames of one segment (model.py:851), detector features in
    64-frame chunks (model.py:436).  Eval phase: picks bit-exact, similarities within 1e-4."""
    from nafae_b200.grounding import ground
    Nb, Ne, D = 20, 13, 512
    rs = np.random.RandomState(Ns)
    vis = synth.embeddings(rs, Na * Ns * Nb, D)
    word = synth.embeddings(rs, Na * Ne, D)
    dev = torch.device("cuda:0")
    with torch.no_grad():
        D_ind, D_sim, loss = ground(torch.from_numpy(vis).to(dev), torch.from_numpy(word).to(dev), lens, Na, Nb, Ne,
                                    10.0, 4.13, False)
    ref = odvsa.dvsa_forward_backward(vis, word, lens, Na, Nb, Ne, 10.0, 4.13, "eval")
    live = _live_mask(dict(Na=Na, Ne=Ne, lens=lens))
    np.testing.assert_array_equal(D_ind.cpu().numpy()[:, live], ref["D_ind"][:, live])
    np.testing.assert_allclose(D_sim.cpu().numpy(), ref["D_sim"], rtol=RTOL, atol=1e-5)


@gpu
def test_train_phase_limits_are_clear_errors():
    """Shapes outside the kernels' limits fail with a message, never with wrong results."""
    from nafae_b200 import _C
    from nafae_b200.grounding import ground
    dev = torch.device("cuda:0")

    def call(Na, Ns, Nb, Ne, D, train):
        v = torch.zeros((Na * Ns * Nb, D), device=dev)
        w = torch.zeros((Na * Ne, D), device=dev)
        return ground(v, w, [1] * Na, Na, Nb, Ne, 10.0, 4.13, train)
    with pytest.raises(_C.NafaeError, match="at most 64 frames"):
        call(1, 65, 4, 3, 64, True)
    call(1, 65, 4, 3, 64, False)  # the eval phase has no such limit
    with pytest.raises(_C.NafaeError, match="max_ent_len"):
        call(1, 2, 4, 17, 64, True)
    with pytest.raises(_C.NafaeError, match="exceeds"):
        call(200, 1, 2, 16, 64, True)  # Na * Ne = 3200 query slots
    with pytest.raises(_C.NafaeError, match="multiple of 4"):
        call(1, 2, 4, 3, 66, True)


@gpu
def test_small_delta_activates_and_deactivates_hinges():
    rs = np.random.RandomState(7)
    Na, Ns, Nb, Ne, D = 4, 3, 6, 5, 64
    lens = [2, 5, 1, 3]
    vis = synth.embeddings(rs, Na * Ns * Nb, D)
    word = synth.embeddings(rs, Na * Ne, D)
    for Delta in (0.0, 0.3, 2.0):
        case = dict(Na=Na, Nb=Nb, Ne=Ne, Delta=Delta, vis_lam=0.7, phase="train", lens=lens)
        D_ind, D_sim, loss, gv, gw = _run(case, vis, word)
        ref = odvsa.dvsa_forward_backward(vis, word, lens, Na, Nb, Ne, Delta, 0.7, "train")
        np.testing.assert_allclose(loss, ref["margin_loss"], rtol=RTOL)
        for got, want in ((gv, ref["grad_vis"]), (gw, ref["grad_word"])):
            np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-5 * np.abs(want).max())


@gpu
def test_module_surface_and_postprocess():
    from nafae_b200.grounding import DVSA, postprocess, record_det
    case, z = load_dvsa_case("small_train")
    vis, word = dvsa_inputs(case)
    args = types.SimpleNamespace(batch_size=case["Na"], batch_size_val=1, max_ent_len=case["Ne"],
                                 Delta=case["Delta"], vis_lam=case["vis_lam"])
    cfg = types.SimpleNamespace(TEST=types.SimpleNamespace(RPN_POST_NMS_TOP_N=case["Nb"]))
    dvsa = DVSA(args, cfg).cuda()
    with pytest.raises(RuntimeError):
        dvsa(torch.from_numpy(vis).cuda(), torch.from_numpy(word).cuda(), case["lens"])
    dvsa.init_train()
    D_ind, D_sim, loss = dvsa(torch.from_numpy(vis).cuda(), torch.from_numpy(word).cuda(),
                              case["lens"])
    assert D_ind.dtype == torch.int64 and D_ind.shape == (case["Na"] * case["Ns"],
                                                          case["Na"] * case["Ne"])
    assert loss.dim() == 0
    np.testing.assert_allclose(float(loss), z["margin_loss"], rtol=RTOL)
    # device postprocess and the numpy calling convention of the reference (model.py:457-474)
    Dp, Sp = postprocess(D_ind, D_sim, case["Na"], case["Ns"], case["Nb"], case["Ne"])
    np.testing.assert_array_equal(Dp.cpu().numpy(), z["post_D"])
    Dn, Sn = postprocess(D_ind.cpu().numpy(), D_sim.cpu().numpy(), case["Na"], case["Ns"],
                         case["Nb"], case["Ne"])
    np.testing.assert_array_equal(Dn, z["post_D"])
    np.testing.assert_allclose(Sn, z["post_D_sim"], rtol=RTOL, atol=1e-5)
    # record_det bookkeeping equals the oracle's
    ents = [["a", "b"], [], ["c", "d", "e", "f"]]
    boxes = np.arange(case["Na"] * case["Ns"] * case["Nb"] * 4).reshape(-1, 4)
    ids = ["img%d" % i for i in range(case["Na"] * case["Ns"])]
    got = ([], [], [], [])
    record_det(got[0], got[1], got[2], got[3], case["Nb"], ents, Dn, Sn, ids, boxes)
    want = odvsa.record_det(case["Nb"], ents, z["post_D"], z["post_D_sim"], ids, boxes)
    assert got[0] == want[0] and got[1] == want[1]
    np.testing.assert_array_equal(np.asarray(got[2]), np.asarray(want[2]))


@gpu
def test_repeated_steps_reuse_workspace_cleanly():
    """The kernels must leave counters / accumulators zeroed: step N+1 equals step 1."""
    case, z = load_dvsa_case("cfg2_train")
    vis, word = dvsa_inputs(case)
    first = _run(case, vis, word)
    for _ in range(3):
        again = _run(case, vis, word)
        np.testing.assert_array_equal(again[0], first[0])
        assert again[2] == first[2]
        np.testing.assert_allclose(again[3], first[3], rtol=1e-5, atol=1e-7)


@gpu
def test_no_grad_forward_returns_the_workspace_and_second_backward_is_a_clear_error():
    """under torch.no_grad() with requires-grad inputs (validation over live modules) the
    forward must not keep the workspace for a backward that never comes; a second backward through
    one forward (retain_graph) raises instead of passing a NULL workspace."""
    import torch
    from nafae_b200.grounding import _WorkspacePool, ground
    dev = torch.device("cuda:0")
    pool = _WorkspacePool()
    vis = torch.randn(2 * 3 * 4, 64, device=dev, requires_grad=True)
    word = torch.randn(2 * 5, 64, device=dev, requires_grad=True)
    with torch.no_grad():
        for _ in range(3):
            ground(vis, word, [2, 3], 2, 4, 5, 10.0, 1.0, False, pool)
    assert sum(len(v) for v in pool._free.values()) == 1  # one workspace, recycled every time
    D_ind, D_sim, loss = ground(vis, word, [2, 3], 2, 4, 5, 10.0, 1.0, True, pool)
    loss.backward(retain_graph=True)
    assert vis.grad is not None and sum(len(v) for v in pool._free.values()) == 1
    with pytest.raises(RuntimeError, match="second backward"):
        loss.backward()


@gpu
@pytest.mark.parametrize("Na,Ns,Nb,Ne,D,train", [
    (8, 5, 20, 13, 512, True),      # cfg2
    (1, 32, 100, 13, 512, False),   # cfg4: one frame per 128-row tile
    (3, 4, 7, 5, 64, True),         # small, ragged tiles (18 frames of 7 rows per tile)
    (24, 2, 20, 13, 512, False),    # 312 query columns: three column tiles
])
def test_tensor_core_contraction_matches_the_fp32_path(Na, Ns, Nb, Ne, D, train):
    """nafae_ground_forward_tc (tcgen05, tf32 x 3) vs nafae_ground_forward (fp32 FMA): identical picks
    on every live column, D_sim / margin_loss within the looser tensor-core bound (2e-4 relative,
    1e-5 absolute), and gradients from the shared backward agree to the same bound."""
    from nafae_b200.grounding import ground
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(Na * 31 + Nb)
    vis = torch.from_numpy(synth.embeddings(rs, Na * Ns * Nb, D)).to(dev)
    word = torch.from_numpy(synth.embeddings(rs, Na * Ne, D)).to(dev)
    lens = [int(x) for x in rs.randint(0, Ne + 1, Na)]
    lens[0] = max(lens[0], 1)
    # exact ties (a duplicated box row: first index wins in both paths) and a copy that differs by about
    # one ulp per element: an fp32 near-tie whose winner depends on the summation order
    vis[1] = vis[0]
    vis[2] = vis[0] * (1 + 2 ** -22)
    outs = []
    for tc in (False, True):
        v = vis.clone().requires_grad_(train)
        w = word.clone().requires_grad_(train)
        D_ind, D_sim, loss = ground(v, w, lens, Na, Nb, Ne, 10.0, 4.13, train, tensor_cores=tc)
        if train:
            loss.backward()
        torch.cuda.synchronize()
        outs.append((D_ind, D_sim, loss.detach(), v.grad, w.grad))
    live = np.zeros((Na, Ne), bool)
    for a, n in enumerate(lens):
        live[a, :n] = True
    live = torch.from_numpy(live.reshape(-1)).to(dev)
    # identical picks -- except where two boxes are closer than fp32 summation noise (the planted
    # near-tie in frame 0): there either order of summation is "the fp32 pick", and D_sim must agree
    diff = (outs[0][0] != outs[1][0]) & live.unsqueeze(0)
    assert int(diff[1:].sum()) == 0 and int(diff[0].sum()) <= int(live.sum())
    if int(diff.sum()):
        a, b = outs[0][1][diff], outs[1][1][diff]
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-6), (a, b)
        rows_fma, rows_tc = outs[0][0][diff], outs[1][0][diff]
        assert set(rows_fma.tolist()) | set(rows_tc.tolist()) <= {0, 1, 2}  # only the planted rows
    assert torch.equal(outs[1][0][:, ~live], torch.zeros_like(outs[1][0][:, ~live]))
    assert torch.allclose(outs[0][1], outs[1][1], rtol=2e-4, atol=1e-5)
    assert torch.allclose(outs[0][2], outs[1][2], rtol=2e-4, atol=1e-5)
    if train and int(diff.sum()) == 0:  # (a flipped near-tie routes the same gradient to the twin row)
        for g0, g1 in ((outs[0][3], outs[1][3]), (outs[0][4], outs[1][4])):
            assert torch.allclose(g0, g1, rtol=2e-4, atol=2e-4 * float(g0.abs().max()))


@gpu
@pytest.mark.parametrize("name", ["cfg2_train", "ties_train", "cfg4_eval"])
def test_tensor_core_path_matches_the_reference_fixture(name):
    """The tensor-core forward against the reference's own DVSA outputs (tests/golden/dvsa_*.npz),
    including the fixture whose frames hold duplicated box rows (exact ties -> first index)."""
    case, z = load_dvsa_case(name)
    vis, word = dvsa_inputs(case)
    from nafae_b200.grounding import ground
    dev = torch.device("cuda:0")
    D_ind, D_sim, loss = ground(torch.from_numpy(vis).to(dev), torch.from_numpy(word).to(dev), case["lens"],
                                case["Na"], case["Nb"], case["Ne"], case["Delta"], case["vis_lam"],
                                case["phase"] == "train", tensor_cores=True)
    np.testing.assert_array_equal(D_ind.cpu().numpy(), z["D_ind"])  # all columns, masked ones included
    np.testing.assert_allclose(D_sim.cpu().numpy(), z["D_sim"], rtol=2e-4, atol=1e-5)
    np.testing.assert_allclose(float(loss), float(z["margin_loss"]), rtol=2e-4)
