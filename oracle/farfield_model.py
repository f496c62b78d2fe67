"""FP64 numpy model of the far-field variant of K2 (PRB_K2_FARFIELD; pyrad_b200/csrc/k2_line_sum.cuh).

TEST INFRASTRUCTURE, like the rest of oracle/: it is the checker for the variant's ALGORITHM -- the class thresholds, the
Chebyshev nodes (rounded to FP32 as the kernel holds them), the Lagrange table -- against the exact oracle
(physics.cross_section, which follows pyradClasses.py:361-400), so that the approximation's error bound is pinned on CPU,
independently of the CUDA kernel's FP32 arithmetic.  Nothing in the product imports it.
"""
import numpy as np

from . import physics as ph

NODES = 16
RADIUS_SPANS = 1
LEVEL2_SPANS = 8          # K2_FAR2_SPANS: spans per level-2 domain (the whole tile)
LEVEL2_MIN_DOMAINS = 4    # K2_FAR2_MIN_DOMAINS: level 2 runs for windows of at least this many domain lengths


def node_offsets(span, nodes=NODES):
    """Chebyshev nodes of [-0.5, span - 0.5] as offsets from the span's first point, rounded to FP32 (api.cu build_far_table)."""
    k = np.arange(nodes)
    return (0.5 * (span - 1) + 0.5 * span * np.cos((2 * k + 1) * np.pi / (2 * nodes))).astype(np.float32).astype(np.float64)


def lagrange_table(span, nodes=NODES):
    """w[i, k]: weight of node k at point i of the span (built from the rounded offsets, as the kernel's table is)."""
    x = node_offsets(span, nodes)
    i = np.arange(span, dtype=np.float64)
    w = np.ones((span, nodes))
    for k in range(nodes):
        for j in range(nodes):
            if j != k:
                w[:, k] *= (i - x[j]) / (x[k] - x[j])
    return w


def line_records(lines, T, P, conc, molmass, qT, q296, range_min, res, weight=1.0):
    """Per-line (idx, A, B, G, C) of the unified form  A/(d^2 + B) + G exp(C d^2),  d = i - idx  in grid units
    (derived form of pyradLineshape.py:32-76 with the regime select pyradClasses.py:378-387)."""
    lp = ph.LineParams(lines, T, P, conc, molmass, qT, q296)
    f, eta = ph.voigt_f_eta(lp.gD, lp.gL)
    voigt, lor, gau = lp.regime == ph.VOIGT, lp.regime == ph.LORENTZ, lp.regime == ph.GAUSS
    hL = np.where(voigt, f / 2, lp.gL)
    hG = np.where(voigt, f / 2, lp.gD)
    cL = np.where(voigt, eta, np.where(lor, 1.0, 0.0)) * lp.S * hL / np.pi * weight
    cG = np.where(voigt, 1 - eta, np.where(gau, 1.0, 0.0)) * lp.S / (hG * np.sqrt(np.pi)) * weight
    with np.errstate(divide="ignore", invalid="ignore"):
        C = np.where(cG != 0, -(res / hG) ** 2, 0.0)
    return ph.line_index(lines["nu"], range_min, res), cL / res ** 2, (hL / res) ** 2, cG, C


def _far_mask(idx, first, length, wm, radius):
    """The kernel's integer far test for the block [first, first + length): window covers the whole block and the
    centre lies more than `radius` points from the block's centre."""
    last = first + length - 1
    full = (idx >= last - wm) & (idx <= first + wm)
    return full & ((idx < first + (length - 1) // 2 - radius) | (idx > first + length // 2 + radius))


def line_sum(idx, A, B, G, C, n, window, span, farfield=True, nodes=NODES, radius_spans=RADIUS_SPANS, radius_points=None,
             level2_spans=LEVEL2_SPANS, level2_min_domains=LEVEL2_MIN_DOMAINS):
    """k[0..n) with the kernel's class logic per span of `span` points: lines whose window |d| <= W-2 covers the whole span
    and whose index lies beyond the integer far thresholds are summed at the nodes (Lorentz term only; their Gaussian
    cores still point by point) and interpolated; every other (line, point) pair exactly.  With level2_spans > 0 the same
    test is first applied to domains of that many spans: a line that is far from a whole domain is summed at the DOMAIN's
    nodes (and skipped by the domain's spans).  Returns (k, far pair fraction)."""
    wm = max(int(window) - 2, 0)
    radius = int(radius_spans * span) if radius_points is None else int(radius_points)   # design studies: any radius
    out = np.zeros(n)
    lag = lagrange_table(span, nodes)
    xn = node_offsets(span, nodes)
    far_pairs = all_pairs = 0
    dom = span * level2_spans if (farfield and level2_spans and wm >= level2_min_domains * level2_spans * span) else 0
    if dom:
        lag2 = lagrange_table(dom, nodes)
        xn2 = node_offsets(dom, nodes)
    for first in range(0, n, span):
        last = first + span - 1
        pts = np.arange(first, min(last, n - 1) + 1)
        reach = (idx + wm >= first) & (idx - wm <= last)
        far = np.zeros(len(idx), dtype=bool)
        far2 = np.zeros(len(idx), dtype=bool)
        if farfield:
            far = _far_mask(idx, first, span, wm, radius)
            if dom:
                dfirst = (first // dom) * dom
                far2 = _far_mask(idx, dfirst, dom, wm, int(radius_spans * dom) if radius_points is None else radius * level2_spans)
                assert not (far2 & ~far).any()                             # level 2 is a subset of level 1
                far = far & ~far2
        near = reach & ~far & ~far2
        d = pts[:, None] - idx[None, near]
        inside = np.abs(d) <= wm
        val = A[near] / (d * d + B[near]) + G[near] * np.exp(C[near] * d * d)
        k = np.where(inside, val, 0.0).sum(axis=1)
        all_pairs += int(inside.sum())
        if far.any():
            dn = (first - idx[far])[None, :] + xn[:, None]                 # (wb - idx) + node offset, as the kernel forms it
            node_sum = (A[far] / (dn * dn + B[far])).sum(axis=1)
            k += lag[: len(pts)] @ node_sum
        if far2.any():
            dn = (dfirst - idx[far2])[None, :] + xn2[:, None]
            node_sum = (A[far2] / (dn * dn + B[far2])).sum(axis=1)
            k += lag2[first - dfirst: first - dfirst + len(pts)] @ node_sum
        anyfar = far | far2
        if anyfar.any():
            dg = pts[:, None] - idx[None, anyfar]
            k += (G[anyfar] * np.exp(C[anyfar] * dg * dg)).sum(axis=1)     # Gaussian cores are never far-fielded
            far_pairs += int(anyfar.sum()) * len(pts)
            all_pairs += int(anyfar.sum()) * len(pts)
        out[first:first + len(pts)] = k
    return out, far_pairs / max(all_pairs, 1)
