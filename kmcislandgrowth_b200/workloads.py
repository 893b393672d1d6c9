"""Synthetic inputs for the configurations BASELINE.json names, in the reference's
terms: a W x Hs hexagonal substrate (2DGrowth.py:18-37) and a film of adatoms placed as perturbed
hexagonal monolayers with a partially filled top (islands / surface atoms).  Seeds are fixed.
"""
import types

import numpy as np

from . import Constants


def constants_for(W, Hs=10, Ha=10, E_a=2.083, E_s=2.083, r_a=2.7, r_s=2.7, temperature=300, deposition_rate=1.0,
                  aa_mode=0, hop_index_mode=0, deposit_mode=0):
    """A Constants-like namespace (same attribute names as the reference's `gv`)."""
    ns = types.SimpleNamespace()
    Constants.configure(E_a, E_s, r_a, r_s, W, "_Images/", Hs, Ha, temperature, deposition_rate, target=ns)
    ns.aa_mode, ns.hop_index_mode, ns.deposit_mode = aa_mode, hop_index_mode, deposit_mode
    return ns


def wrap_x(x, L):
    """Periodic.py:8-20 vectorised (numpy's % is Python's float modulo)."""
    x = np.mod(x, L)
    return np.where(np.abs(x) > L / 2, x - L, x)


def substrate(gv):
    """2DGrowth.py:18-37 vectorised."""
    i = np.arange(gv.Hs)
    k = np.arange(gv.W)
    X = k[None, :] * gv.r_s + np.where(i[:, None] % 2 == 1, gv.r_s / 2, 0.0)
    Y = np.broadcast_to((-i * gv.r_s * gv.sqrt3 / 2)[:, None], X.shape)
    R = np.empty(2 * gv.W * gv.Hs)
    R[0::2] = wrap_x(X.reshape(-1), gv.L)
    R[1::2] = Y.reshape(-1)
    return R


def film(gv, n_adatoms, seed, sigma=0.02, shuffle=True):
    """n_adatoms adatoms: full monolayers continuing the substrate's stacking, the last two layers
    randomly filled; Gaussian jitter sigma (Angstrom); x wrapped; order shuffled (the reference's
    adatom order is deposition order, not spatial)."""
    rng = np.random.default_rng(seed)
    W = gv.W
    full = max(n_adatoms // W - 1, 0)
    rest = n_adatoms - full * W
    pts = []
    for l in range(1, full + 1):
        x = np.arange(W) * gv.r_a + (gv.r_a / 2 if l % 2 == 1 else 0.0)
        pts.append(np.column_stack([x, np.full(W, l * gv.r_a * gv.sqrt3 / 2)]))
    if rest > 0:
        nl = max(2, int(np.ceil(rest / (0.6 * W))))
        sites = []
        for l in range(full + 1, full + 1 + nl):
            x = np.arange(W) * gv.r_a + (gv.r_a / 2 if l % 2 == 1 else 0.0)
            sites.append(np.column_stack([x, np.full(W, l * gv.r_a * gv.sqrt3 / 2)]))
        sites = np.concatenate(sites)
        # lower layers fill first on average
        w = np.exp(-(sites[:, 1] - sites[:, 1].min()) / (2 * gv.r_a))
        pick = rng.choice(len(sites), size=rest, replace=False, p=w / w.sum())
        pts.append(sites[pick])
    P = np.concatenate(pts) if pts else np.zeros((0, 2))
    P = P + rng.normal(0.0, sigma, size=P.shape)
    if shuffle:
        P = P[rng.permutation(len(P))]
    R = np.empty(2 * len(P))
    R[0::2] = wrap_x(P[:, 0], gv.L)
    R[1::2] = P[:, 1]
    return R


# the configurations of BASELINE.json
CONFIGS = {
    # name: (W, Hs, Ha, n_adatoms, seed, aa_mode)
    "default": (18, 10, 10, 100, 1, 0),
    "cfg2_256x256_10k": (256, 256, 64, 10000, 2, 1),
    "cfg3_1024x1024_500k": (1024, 1024, 512, 500000, 3, 1),
    "cfg5_4096x4096": (4096, 4096, 10, 8192, 5, 1),
}


def make(name, **over):
    W, Hs, Ha, n, seed, aa = CONFIGS[name]
    gv = constants_for(W, Hs, Ha, aa_mode=aa, hop_index_mode=1 if aa else 0, deposit_mode=1 if aa else 0, **over)
    return gv, substrate(gv), film(gv, n, seed)


# ---- the bench protocol of configs[1] (bench.py, tests/test_gpu_parity.py::test_headline_*) ---------------------------
BENCH_SEED = 1234               # uniform stream of the bench trajectory
BENCH_PRERELAX_CAP = 200000     # line-search bound of the untimed set-up relaxation (it converges long before)


def bench_uniforms(n_events):
    """the explicit uniform stream of the bench trajectory: u[s] = (selection, placement) of event s"""
    return np.random.default_rng(BENCH_SEED).uniform(0, 1, (n_events, 2))


def bench_prerelax(ctx, sbins, cap=BENCH_PRERELAX_CAP):
    """untimed set-up: relax the synthetic film with the reference's own Relaxation (2DGrowth.py:97-166) to ITS convergence
    criterion, so that the timed events start from a relaxed state; -> stats of the last relaxation, total line searches"""
    total = 0
    st = None
    while total < cap:
        st = ctx.relax(sbins, 16, 1e-4, min(20000, cap - total))
        total += int(st.line_searches)
        if int(st.status) == 0:
            break
    return st, total
