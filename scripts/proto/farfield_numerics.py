"""Numerical prototype (CPU, numpy FP64/FP32) of a far-field scheme for K2: lines whose centre lies far from a warp's
span are evaluated at m Chebyshev nodes of the span and interpolated, instead of at every point.  Prints the error of the
interpolated far-field sum relative to the exact total k at the span's points, for cfg2-like and atmosphere-like cells,
on a handful of RANDOM spans (the worst case over all spans is several times larger: oracle/farfield_model.py and
tests/test_farfield_model.py pin it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import physics as ph
from pyrad_b200 import workloads

def cheb_nodes(m, a, b):
    k = np.arange(m)
    return 0.5 * (a + b) + 0.5 * (b - a) * np.cos((2 * k + 1) * np.pi / (2 * m))

def lagrange_matrix(nodes, x):
    L = np.ones((len(x), len(nodes)))
    for k in range(len(nodes)):
        for j in range(len(nodes)):
            if j != k:
                L[:, k] *= (x - nodes[j]) / (nodes[k] - nodes[j])
    return L

def run(P, T, span, m, R, n_lines=60000, rmax=360.0, seed_spans=8, f32=False):
    w = workloads.cfg2(n_lines, rmax)
    w["P"], w["T"] = P, T
    cutoff = P / 1013.25 * 5
    res = w["res"]
    n = ph.grid_len(w["range_min"], w["range_max"], res)
    W = ph.window_len(cutoff, res)
    wm = max(W - 2, 0)
    A, B, G, C, F = [], [], [], [], []
    for g, sp in enumerate(w["species"]):
        ln = w["per_group_lines"][g]
        lp = ph.LineParams(ln, T, P, w["conc"][g], sp.molmass, sp.q(T), sp.q296)
        wt = float(ph.abs_coef(1.0, w["conc"][g], P, T))
        idx = ph.line_index(ln["nu"], w["range_min"], res)
        f, eta = ph.voigt_f_eta(lp.gD, lp.gL)
        hL = np.where(lp.regime == ph.VOIGT, f / 2, lp.gL)
        hG = np.where(lp.regime == ph.VOIGT, f / 2, lp.gD)
        cL = np.where(lp.regime == ph.VOIGT, eta, np.where(lp.regime == ph.LORENTZ, 1.0, 0.0)) * lp.S * hL / np.pi * wt
        cG = np.where(lp.regime == ph.VOIGT, 1 - eta, np.where(lp.regime == ph.GAUSS, 1.0, 0.0)) * lp.S / (hG * np.sqrt(np.pi)) * wt
        A.append(cL / res ** 2); B.append((hL / res) ** 2); G.append(cG); C.append(-(res / hG) ** 2); F.append(idx)
    A, B, G, C, F = (np.concatenate(v) for v in (A, B, G, C, F))
    rng = np.random.default_rng(1)
    worst = 0.0
    fracs = []
    for s0 in rng.integers(0, n // span, seed_spans) * span:
        pts = np.arange(s0, s0 + span)
        centre = s0 + (span - 1) / 2
        full = (F >= s0 + span - 1 - wm) & (F <= s0 + wm)          # window covers the whole span
        part = ~full & (F + wm >= s0) & (F - wm <= s0 + span - 1)
        far = full & (np.abs(F - centre) > R)
        near = (full & ~far) | part
        d = pts[:, None] - F[None, near]
        exact_near = np.where(np.abs(d) <= wm, A[near] / (d * d + B[near]) + G[near] * np.exp(C[near] * d * d), 0).sum(axis=1)
        df = pts[:, None] - F[None, far]
        exact_far = (A[far] / (df * df + B[far]) + G[far] * np.exp(C[far] * df * df)).sum(axis=1)
        nodes = cheb_nodes(m, s0 - 0.5, s0 + span - 0.5)
        dn = nodes[:, None] - F[None, far]
        if f32:
            dn32 = dn.astype(np.float32); 
            vals = (A[far].astype(np.float32) / (dn32 * dn32 + B[far].astype(np.float32))).astype(np.float64).sum(axis=1)
        else:
            vals = (A[far] / (dn * dn + B[far])).sum(axis=1)
        approx_far = lagrange_matrix(nodes, pts.astype(np.float64)) @ vals
        tot = exact_near + exact_far
        err = np.abs(approx_far - exact_far) / tot
        worst = max(worst, err.max())
        fracs.append(far.sum() / max(full.sum() + part.sum(), 1))
    print("P=%7.2f W-2=%5d span=%d m=%d R=%d: far fraction %.2f  max rel err of total %.2e" % (P, wm, span, m, R, np.mean(fracs), worst))

if __name__ == "__main__":
    # shipped geometry (radius = 2 spans from the centre, 8 nodes) on the window classes it runs on, then alternatives
    print("# shipped: span 256 (k2_line_sum<8>), span 128 (k2_line_sum<4>), radius 2 spans, 8 nodes")
    for (P, T) in ((1013.25, 296), (353.4, 250), (250.0, 230)):
        run(P, T, 256, 8, 512)
    for (P, T) in ((150.0, 225), (100.0, 215), (60.0, 215)):
        run(P, T, 128, 8, 256)
    print("# alternatives for the next round: nodes x radius, and 64-point spans for the narrow wide-kernel class")
    for (P, T) in ((1013.25, 296), (250.0, 230)):
        for m, R in ((8, 384), (6, 512), (10, 384), (12, 320), (8, 768)):
            run(P, T, 256, m, R)
    for (P, T) in ((150.0, 225), (60.0, 215)):
        for m, R in ((10, 192), (12, 160), (6, 256)):
            run(P, T, 128, m, R)
    for (P, T) in ((40.0, 215), (25.0, 220)):
        for m, R in ((8, 128), (10, 96)):
            run(P, T, 64, m, R)
