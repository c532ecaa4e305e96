"""Dev / measurement tool: the scoring kernel's contraction on the FMA pipe vs on the tensor cores
(nafae_ground_forward vs nafae_ground_forward_tc), forward alone, CUDA events, graph replays.
    python tools/time_head.py            # cfg2, cfg4, and a 520-live-column stress shape
Used for the A/B table in profiles/RESULTS.md; run under ncu for the tensor-pipe utilisation."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nafae_b200 import synth, _C  # noqa: E402

dev = torch.device("cuda:0")
SHAPES = [("cfg2 (R=800, 104 slots, ~17 live)", 8, 5, 20, 13, 512, None, True),
          ("cfg4 (R=3200, Nb=100, 13 slots, 4 live)", 1, 32, 100, 13, 512, [4], False),
          ("stress (R=4000, 520 slots, all live)", 40, 5, 20, 13, 512, [13] * 40, False)]
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, Na, Ns, Nb, Ne, D, lens, train in SHAPES:
    if only and only not in name:
        continue
    rs = np.random.RandomState(1)
    vis = torch.from_numpy(synth.embeddings(rs, Na * Ns * Nb, D)).to(dev)
    word = torch.from_numpy(synth.embeddings(rs, Na * Ne, D)).to(dev)
    if lens is None:
        lens = synth.entity_lengths(rs, Na, Ne)
    lt = torch.tensor(lens, dtype=torch.int32, device=dev)
    F, NQ = Na * Ns, Na * Ne
    D_ind = torch.empty((F, NQ), dtype=torch.int64, device=dev)
    D_sim = torch.empty((F, NQ), device=dev)
    loss = torch.zeros((), device=dev)
    ws = torch.zeros(int(_C.lib.nafae_ground_workspace_bytes(Na, Ns, Nb, Ne, D)) // 4, dtype=torch.int32, device=dev)
    res = {}
    for label, fn in (("fma", _C.lib.nafae_ground_forward), ("tcgen05 tf32x3", _C.lib.nafae_ground_forward_tc)):
        def run():
            st = fn(_C.ptr(vis), _C.ptr(word), _C.ptr(lt), Na, Ns, Nb, Ne, D, 10.0, 4.13, int(train), _C.ptr(D_ind),
                    _C.ptr(D_sim), _C.ptr(loss), _C.ptr(ws), ws.numel() * 4, _C.stream())
            assert st == 1, _C.last_error()
        run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        for _ in range(10):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res[label] = (e0.elapsed_time(e1) / 200 * 1e3, D_ind.clone(), D_sim.clone(), float(loss))
    live = torch.tensor([(c % Ne) < lens[c // Ne] for c in range(NQ)], device=dev)
    same = bool(torch.equal(res["fma"][1][:, live], res["tcgen05 tf32x3"][1][:, live]))
    err = float((res["fma"][2] - res["tcgen05 tf32x3"][2]).abs().max())
    print("%-42s fma %7.1f us | tcgen05 tf32x3 %7.1f us | picks identical %s, max |dD_sim| %.2e, live columns %d" % (
        name, res["fma"][0], res["tcgen05 tf32x3"][0], same, err, int(live.sum())), flush=True)
