"""Shared helpers for the parity tests: oracle evaluation of a workload and the error metrics."""
import numpy as np

from oracle import physics as ph
from pyrad_b200 import engine as eng

#: north_star tolerances
K_REL_TOL = 1e-5        # relative error on k(nu) / sigma(nu)
T_ABS_TOL = 1e-6        # absolute error on transmittance
#: the scaled-FP32 evaluation has a floor ~2^-148 of the strongest line peak (DESIGN.md, K2 numerics)
K_FLOOR_REL = 1e-40


def k_rel_err(out, ref):
    ref = np.asarray(ref)
    floor = K_FLOOR_REL * np.max(np.abs(ref)) if ref.size else 0.0
    den = np.maximum(np.abs(ref), floor)
    den = np.where(den == 0, 1.0, den)
    return np.abs(np.asarray(out) - ref) / den


def group_lines(w, g):
    return w["per_group_lines"][g]


def kept(lines, rmin, rmax, cutoff):
    """The reference keeps lines with effMin < nu < effMax strictly (pyradUtilities.py:437-438)."""
    lo, hi = ph.effective_range(rmin, rmax, cutoff)
    m = (lines["nu"] > lo) & (lines["nu"] < hi)
    return {k: np.asarray(v)[m] for k, v in lines.items()}


def oracle_sigma_groups(w, T=None, P=None, conc=None, cutoff=None, points=None):
    """Per-group cross sections from the oracle (scatter form, or gather form at `points`)."""
    T = w["T"] if T is None else T
    P = w["P"] if P is None else P
    conc = w["conc"] if conc is None else conc
    cutoff = w["cutoff"] if cutoff is None else cutoff
    out = []
    for g, sp in enumerate(w["species"]):
        ln = group_lines(w, g)
        args = (ln, T, P, conc[g], sp.molmass, sp.q(T), sp.q296, w["range_min"], w["range_max"], w["res"], cutoff)
        if points is None:
            out.append(ph.cross_section(*args))
        else:
            out.append(ph.cross_section_at(points, *args))
    return np.array(out)


def engine_setup(e, w, i_begin=0, i_end=None):
    n = eng.grid_len(w["range_min"], w["range_max"], w["res"])
    e.upload_lines(w["lines"], n_groups=len(w["species"]))
    e.set_grid(w["range_min"], w["res"], n, i_begin, i_end)
    return n


def engine_prepass(e, w, weights=None, T=None, P=None, conc=None, cutoff=None):
    T = w["T"] if T is None else T
    P = w["P"] if P is None else P
    conc = w["conc"] if conc is None else conc
    cutoff = w["cutoff"] if cutoff is None else cutoff
    sp = w["species"]
    e.layer_prepass(T, P, conc, [s.molmass for s in sp], [s.q(T) for s in sp], [s.q296 for s in sp],
                    eng.window_len(cutoff, w["res"]), weights)
