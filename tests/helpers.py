"""Shared helpers for the parity tests (numpy side)."""
import os

import numpy as np

from oracle import siren_oracle as so

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name, tag="f64"):
    return dict(np.load(os.path.join(GOLDEN, "%s_%s.npz" % (name, tag)), allow_pickle=False))


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.sqrt((b * b).sum())
    return float(np.sqrt(((a - b) ** 2).sum()) / max(den, 1e-300))


def sample(a, stride=97):
    return np.ascontiguousarray(np.asarray(a).reshape(-1)[::stride])


def check_grads(prefix, g, dWs, dbs, tol):
    """Compare full oracle/CUDA gradients with the sampled golden record."""
    worst = 0.0
    for l in range(len(dWs)):
        key = "%s_dW%d" % (prefix, l)
        if key in g:
            e = rel_l2(dWs[l], g[key])
        else:
            e = rel_l2(sample(dWs[l]), g[key + "_sample"])
            l2 = float(np.sqrt((np.asarray(dWs[l], np.float64) ** 2).sum()))
            e = max(e, abs(l2 - float(g[key + "_l2"])) / max(float(g[key + "_l2"]), 1e-300))
        worst = max(worst, e)
        assert e <= tol, "%s: rel err %.3e > %.1e" % (key, e, tol)
        ref_b = g["%s_db%d" % (prefix, l)]
        if np.abs(ref_b).max() == 0:
            assert np.abs(dbs[l]).max() <= 1e-12 + tol
        else:
            e = rel_l2(dbs[l], ref_b)
            worst = max(worst, e)
            assert e <= tol, "%s_db%d: rel err %.3e > %.1e" % (prefix, l, e, tol)
    return worst


def case_inputs(g, tasks=0):
    d, o, n, seed = int(g["d"]), int(g["o"]), int(g["n"]), int(g["seed"])
    Ws, bs = so.make_params(d, 256, 3, o, seed=seed, tasks=tasks)
    x = so.make_coords(max(tasks, 1), n, d, seed=seed + 100)
    return d, o, n, Ws, bs, x


def loss_adjoints_gradmse(J, gt_grad):
    """dL/dJ for loss_functions.gradients_mse (loss_functions.py:330-335)."""
    g = so.gradient(J)
    T, N = g.shape[:2]
    gg = 2.0 * (g - gt_grad) / (T * N)
    return np.broadcast_to(gg[..., None, :], J.shape).copy()


def loss_adjoints_lapmse(D, gt_lap):
    """dL/dD for loss_functions.laplace_mse (loss_functions.py:350-355)."""
    lap = so.laplace(D)
    gl = 2.0 * (lap - gt_lap) / lap.size
    return np.broadcast_to(gl[..., None], D.shape).copy() * (np.arange(D.shape[-2]) == 0)[:, None]
