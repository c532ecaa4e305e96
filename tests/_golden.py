"""Helpers shared by the parity tests: load committed fixtures and rebuild their inputs."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dvsa_case_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "dvsa_*.npz")))


def load_dvsa_case(name):
    z = np.load(os.path.join(GOLDEN, "dvsa_%s.npz" % name))
    case = dict(Na=int(z["Na"]), Ns=int(z["Ns"]), Nb=int(z["Nb"]), Ne=int(z["Ne"]), D=int(z["D"]),
                lens=[int(x) for x in z["lens"]], Delta=float(z["Delta"]),
                vis_lam=float(z["vis_lam"]), phase="train" if bool(z["phase_train"]) else "eval",
                seed=int(z["seed"]), dup=int(z["dup"]), name=name)
    return case, z


def dvsa_inputs(case):
    """Same recipe as tests/golden/make_dvsa_golden.py::make_inputs (IEEE-exact ops only)."""
    rs = np.random.RandomState(case["seed"])
    R = case["Na"] * case["Ns"] * case["Nb"]
    vis = np.clip(rs.standard_normal((R, case["D"])) * 0.5, -1, 1).astype(np.float32)
    word = np.clip(rs.standard_normal((case["Na"] * case["Ne"], case["D"])) * 0.5, -1, 1)
    word = word.astype(np.float32)
    if case["dup"]:
        Nb = case["Nb"]
        for f in range(case["Na"] * case["Ns"]):
            vis[f * Nb + Nb - case["dup"]: (f + 1) * Nb] = vis[f * Nb + Nb - case["dup"] - 1]
    return vis, word


def check_grads_against_fixture(z, gv, gw, rtol, atol_scale=1e-6):
    """Full grads for small cases, projections / abs-sums for the full-size ones."""
    if "grad_vis" in z:
        for got, ref in ((gv, z["grad_vis"]), (gw, z["grad_word"])):
            np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol_scale * np.abs(ref).max())
    else:
        P = z["grad_proj"].astype(np.float64)
        for got, proj, ab in ((gv, z["grad_vis_proj"], z["grad_vis_abs"]),
                              (gw, z["grad_word_proj"], z["grad_word_abs"])):
            np.testing.assert_allclose(got.astype(np.float64) @ P, proj, rtol=rtol,
                                       atol=10 * atol_scale * np.abs(proj).max())
            np.testing.assert_allclose(np.abs(got).sum(1), ab, rtol=rtol,
                                       atol=atol_scale * np.abs(ab).max())
