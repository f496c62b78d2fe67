"""GPU parity at BASELINE.json's full sizes: the oracle cannot scatter 5e9+ pairs on a CPU in seconds, so
full-size runs are checked (a) against the oracle's gather form at sampled grid points, and (b) through
size-independent properties: linearity in S, additivity over line subsets, shard invariance."""
import numpy as np
import pytest

from oracle import physics as ph
from pyrad_b200 import engine as eng
from pyrad_b200 import workloads
from tests import helpers as H

pytestmark = pytest.mark.gpu


def sample_points(n, k, seed, tile=2048):
    """>= k sampled grid indices: random ones plus both sides of tile and warp-span edges (where the kernels'
    masked / unmasked / far classes flip) and the ends of the grid."""
    return H.boundary_points(n, k, seed, tile=tile)


def check_sampled(engine, w, weights_mode, P=None, T=None, n_pts=1000, seed=0, variant=eng.K2_CLASSED):
    P = w["P"] if P is None else P
    T = w["T"] if T is None else T
    cutoff = P / 1013.25 * 5 if P != w["P"] else w["cutoff"]
    n = H.engine_setup(engine, w)
    wts = [eng.number_density_weight(c, P, T) for c in w["conc"]] if weights_mode else None
    engine.set_k2_variant(variant, 0)
    try:
        H.engine_prepass(engine, w, weights=wts, T=T, P=P, cutoff=cutoff)
        out = engine.line_sum()
    finally:
        engine.set_k2_variant(eng.K2_CLASSED, 0)
    pts = sample_points(n, n_pts, seed)
    assert len(pts) >= n_pts
    if weights_mode:
        ref = H.oracle_layer_k_at(w, pts, T, P, w["conc"], cutoff)
    else:
        ref = H.oracle_sigma_groups(w, T=T, P=P, cutoff=cutoff, points=pts).sum(axis=0)
    floor = H.K_FLOOR_REL * np.abs(out).max()
    err = np.abs(out[pts] - ref) / np.maximum(np.abs(ref), floor)
    assert err.max() <= H.K_REL_TOL, (err.max(), pts[err.argmax()])
    return out


def test_cfg2_full_size_sampled_against_oracle(engine):
    """cfg2: 3.0M points, 500k lines, W = 5000 (the bench workload)."""
    w = workloads.cfg2()
    out = check_sampled(engine, w, weights_mode=True)
    assert engine.pair_count() > 4.9e9
    assert np.all(np.isfinite(out)) and np.all(out >= 0)


@pytest.mark.parametrize("P,T", [(353.4, 250), (44.28, 230), (5.315, 260)])
def test_atmosphere_layer_shapes_full_size_sampled(engine, P, T):
    """cfg4-sized line list (5M lines, 5M points) at three layer pressures: wide, mid and narrow windows."""
    w = workloads.cfg5(cutoff=5.0)
    check_sampled(engine, w, weights_mode=True, P=P, T=T, seed=int(P))


def test_cfg5_stress_sweep_full_size_sampled(engine):
    """cfg5: 5M lines, 5M points, 25 cm-1 cutoff (W = 25 000, ~2.5e11 accumulations) against the oracle's gather form
    at sampled points, plus the exact accumulate count."""
    w = workloads.cfg5()
    out = check_sampled(engine, w, weights_mode=True, seed=5)
    n = eng.grid_len(w["range_min"], w["range_max"], w["res"])
    idx = ph.line_index(w["lines"]["nu"], w["range_min"], w["res"])
    assert engine.pair_count() == ph.pair_count(idx, n, eng.window_len(w["cutoff"], w["res"])) > 2.4e11
    assert np.all(np.isfinite(out))       # (not >= 0: next to 0 cm-1 the reference's negative Doppler widths show up)


def test_cfg1_size_properties(engine):
    """cfg1 (30 000 points, 50k lines): linearity in S, additivity over a line split, T=296 identity."""
    w = workloads.cfg1()
    n = H.engine_setup(engine, w)
    H.engine_prepass(engine, w)
    base = engine.line_sum()
    ref_pts = sample_points(n, 1000, 3)
    sig = H.oracle_sigma_groups(w, points=ref_pts)[0]
    assert H.k_rel_err(base[ref_pts], sig).max() <= H.K_REL_TOL
    # linearity: S -> 4 S (a power of two: bitwise equal thanks to the power-of-two scaling)
    L = dict(w["lines"]); L["sw"] = w["lines"]["sw"] * 4.0
    w4 = dict(w); w4["lines"] = L
    H.engine_setup(engine, w4)
    H.engine_prepass(engine, w4)
    assert np.array_equal(engine.line_sum(), 4.0 * base)
    # additivity over odd/even line subsets
    parts = []
    for sel in (slice(0, None, 2), slice(1, None, 2)):
        ws = dict(w); ws["lines"] = {k: np.ascontiguousarray(v[sel]) for k, v in w["lines"].items()}
        H.engine_setup(engine, ws)
        H.engine_prepass(engine, ws)
        parts.append(engine.line_sum())
    assert H.k_rel_err(parts[0] + parts[1], base).max() <= 2e-6


def test_cfg2_full_size_farfield_variant_sampled_against_oracle(engine):
    """cfg2 with the opt-in far-field variant of K2: same tolerance against the oracle as the exact kernel."""
    w = workloads.cfg2()
    out = check_sampled(engine, w, weights_mode=True, variant=eng.K2_FARFIELD)
    assert np.all(np.isfinite(out)) and np.all(out >= 0)


@pytest.mark.parametrize("P,T", [(353.4, 250), (150.0, 225)])
def test_atmosphere_layer_shapes_full_size_farfield_variant_sampled(engine, P, T):
    """cfg4-sized line list under the far-field variant: a 256-point-span layer (P = 8) and a 128-point-span layer (P = 4)."""
    w = workloads.cfg5(cutoff=5.0)
    check_sampled(engine, w, weights_mode=True, P=P, T=T, seed=int(P), variant=eng.K2_FARFIELD)


def test_cfg5_full_size_farfield_variant_sampled_against_oracle(engine):
    """cfg5 under the far-field variant, where both of its levels carry most of the lines (W = 25 000)."""
    w = workloads.cfg5()
    out = check_sampled(engine, w, weights_mode=True, seed=5, variant=eng.K2_FARFIELD)
    assert np.all(np.isfinite(out))


def test_cfg3_full_size_line_by_line_plus_xsc_tables(engine, tmp_path):
    """cfg3 (30 000 points, 50k CO2 + H2O lines, CFC-11 and HCFC-22 xsc tables on the same grid) through the host mirror on
    an on-disk data tree, every spectrum against the oracle on the same inputs."""
    from oracle import ref_harness as rh          # only its data-tree writers (test infrastructure)
    from pyrad_b200 import classes as C
    w = workloads.cfg3()
    root = str(tmp_path)
    C.set_engine(engine)
    C.DATA_ROOT = root
    C.Layer.hasAtmosphere = False
    try:
        for g, sp in enumerate(w["species"]):
            ln = dict(w["per_group_lines"][g])
            rh.write_params(root, sp.global_iso, sp.name, sp.mol_id, 1, 0.99, sp.q296, 1, sp.molmass)
            rh.write_q_table(root, sp.global_iso, range(100, 501), [sp.q(t) for t in range(100, 501)])
            rh.write_line_segments(root, sp.global_iso, sp.mol_id, 1, ln, int(ln["nu"].min() / 100) * 100, ln["nu"].max() + 101)
        layer = C.Layer(w["depth_cm"], w["T"], w["P"], w["range_min"], w["range_max"])
        xms = []
        for x in w["xsc"]:
            f = rh.write_xsc_file(root, x["name"], x["T"], x["torr"], x["range_min"], x["range_max"], x["res"],
                                  x["wavenumber"], x["intensity"])
            xms.append(layer.addMolecule({x["name"]: f}, concentration=x["conc"]))
        mols = [layer.addMolecule(sp.name, concentration=c) for sp, c in zip(w["species"], w["conc"])]
        assert layer.T == w["T"] and layer.P == w["P"]          # the tables' T and P are the cell's: nothing was overwritten
        xa = ph.x_axis(w["range_min"], w["range_max"], w["res"])
        np.testing.assert_array_equal(layer.xAxis, xa)
        k_ref = np.zeros(len(xa))
        for x, xm in zip(w["xsc"], xms):
            sig = ph.xsc_cross_section(xa, x["wavenumber"], x["intensity"], x["range_min"], x["range_max"], x["res"])
            np.testing.assert_allclose(C.getCrossSection(xm), sig, rtol=1e-14, atol=0)
            assert np.count_nonzero(sig) > 5000
            k_ref += ph.abs_coef(sig, x["conc"], w["P"], w["T"])
        for g, (sp, m) in enumerate(zip(w["species"], mols)):
            ln = H.kept(w["per_group_lines"][g], w["range_min"], w["range_max"], w["cutoff"])
            assert len(m[0]) == len(ln["nu"]) > 20_000
            sig = ph.cross_section(ln, w["T"], w["P"], w["conc"][g], sp.molmass, sp.q(w["T"]), sp.q296,
                                   w["range_min"], w["range_max"], w["res"], w["cutoff"])
            assert H.k_rel_err(C.getCrossSection(m[0]), sig).max() <= H.K_REL_TOL
            k_ref += ph.abs_coef(sig, w["conc"][g], w["P"], w["T"])
        assert H.k_rel_err(C.getAbsCoef(layer), k_ref).max() <= H.K_REL_TOL
        t_ref = ph.transmittance(k_ref, w["depth_cm"])
        assert np.abs(C.getTransmittance(layer) - t_ref).max() <= H.T_ABS_TOL
        surf = ph.planck_wavenumber(xa, 288)
        rad_ref = ph.transmission(t_ref, surf, ph.planck_wavenumber(xa, w["T"]))
        np.testing.assert_allclose(layer.transmission(surf), rad_ref, rtol=2e-5)
    finally:
        C.DATA_ROOT = None


@pytest.mark.parametrize("xsc_scale", [1.0, 1e6])
def test_cfg3_full_size_through_one_engine_call(engine, xsc_scale):
    """cfg3 at full size as ONE prb_atmosphere call through the C ABI: CO2 + H2O line lists (two groups) and the CFC-11 /
    HCFC-22 xsc tables resident on the device (prb_xsc_resident: the native-resolution table placed, the 0.05 cm-1 table
    re-gridded with np.interp's arithmetic), added to the line sum in K2's epilogue.  k, transmittance and radiance at
    EVERY grid point against the oracle (pyradClasses.py:466-505, 581-587, 707-716, 784-787).  At cfg3's own mole fractions
    (250 / 230 ppt) the tables' optical depth is 5e-7 -- below the transmittance tolerance --, so the cell is also run with
    them a million times more abundant (optical depth 0.5), where a missing or misplaced table fails every assertion."""
    import torch
    from pyrad_b200 import classes as C
    from pyrad_b200 import distributed as pd
    w = workloads.cfg3()
    for x in w["xsc"]:
        x["conc"] = x["conc"] * xsc_scale
    n = H.engine_setup(engine, w)
    xa = ph.x_axis(w["range_min"], w["range_max"], w["res"])
    k_ref = np.zeros(n)
    engine.xsc_clear()
    try:
        for slot, x in enumerate(w["xsc"]):
            grid01 = np.arange(x["range_min"], x["range_max"], .01)
            dst0, src0, count, out_len = C._merge_plan(xa, grid01)
            assert out_len == n
            coarse = x["res"] > .01
            engine.xsc_resident(slot, n, dst0, src0, count, x["wavenumber"] if coarse else None, x["intensity"], coarse,
                                ax0=float(grid01[0]), adelta=float(grid01[1] - grid01[0]))
            sig = ph.xsc_cross_section(xa, x["wavenumber"], x["intensity"], x["range_min"], x["range_max"], x["res"])
            assert np.count_nonzero(sig) > 5000
            k_ref += ph.abs_coef(sig, x["conc"], w["P"], w["T"])
        engine.set_xsc_conc([[x["conc"] for x in w["xsc"]]])
        rad, tr = _column(engine, w)
        assert engine.atmosphere_launches() == 4                  # two table resamplings + K1 + K2 (fused): one engine call
        kp, ld = engine.atmosphere_kmatrix_dev()
        k = pd.device_tensor(kp, ld)[:n].cpu().numpy().astype(np.float64)
        # a second call reuses the resampled tables
        rad2, tr2 = _column(engine, w)
        assert engine.atmosphere_launches() == 2 and np.array_equal(tr, tr2) and np.array_equal(rad, rad2, equal_nan=True)
    finally:
        engine.xsc_clear()
    for g, sp in enumerate(w["species"]):
        ln = H.kept(w["per_group_lines"][g], w["range_min"], w["range_max"], w["cutoff"])
        sig = ph.cross_section(ln, w["T"], w["P"], w["conc"][g], sp.molmass, sp.q(w["T"]), sp.q296, w["range_min"],
                               w["range_max"], w["res"], w["cutoff"])
        k_ref += ph.abs_coef(sig, w["conc"][g], w["P"], w["T"])
    assert H.k_rel_err(k, k_ref).max() <= H.K_REL_TOL             # (FP32 row of the k matrix: 6e-8 of rounding on top)
    t_ref = ph.transmittance(k_ref, w["depth_cm"])
    assert np.abs(tr - t_ref).max() <= H.T_ABS_TOL
    rad_ref = ph.transmission(t_ref, ph.planck_wavenumber(xa, 288.0), ph.planck_wavenumber(xa, w["T"]))
    np.testing.assert_allclose(rad, rad_ref, rtol=2e-5)
    # without the tables the call is a different spectrum
    _, tr0 = _column(engine, w)
    assert (np.abs(tr0 - t_ref).max() > 1e-3) == (xsc_scale > 1)


def _column(engine, w):
    sp = w["species"]
    win = [eng.window_len(c, w["res"]) for c in np.atleast_1d(w["cutoff"])]
    T, P = np.atleast_1d(w["T"]), np.atleast_1d(w["P"])
    qt = np.array([[s.q(t) for s in sp] for t in T])
    conc = np.atleast_2d(w["conc"])
    depth = np.broadcast_to(np.asarray(w["depth_cm"], dtype=np.float64), (len(T),))
    engine.atmosphere(depth, T, P, conc, [s.molmass for s in sp], qt, [s.q296 for s in sp], win,
                      w.get("t_surface", 288.0), w["range_max"])
    return engine.atmosphere_read()


def _as_column(w):
    """A single-layer gas-cell workload in the column layout of H.oracle_column_at."""
    c = dict(w)
    c["T"], c["P"] = [w["T"]], [w["P"]]
    c["conc"], c["cutoff"], c["depth_cm"] = [w["conc"]], [w["cutoff"]], [w["depth_cm"]]
    return c


@pytest.mark.parametrize("depth_scale", [1.0, 2e-5])
def test_cfg4_full_size_column_sampled(engine, depth_scale):
    """The north-star configuration itself: the 100-layer, 5 M-line, 5 M-point atmosphere through prb_atmosphere
    (K1 for the whole column, the batched line sums of every kernel class, the K3 fold), radiance and TOTAL
    transmittance against the oracle's layer-by-layer fold at > 256 grid points: both sides of tile and warp-span
    edges, the first points next to 0 cm-1, the last points, random points.  Reference: pyradClasses.py:707-716, 784-787.
    The bench column (depth_scale 1) is opaque everywhere (median optical depth 2.4e4: its total transmittance is 0 and its
    radiance is the emission of the upper layers), so the same column -- same lines, same k matrix -- is also run with
    1.4 cm layers (median optical depth ~0.5), where the 100-term optical-depth sum decides the total transmittance."""
    w = workloads.atmosphere()
    w["depth_cm"] = w["depth_cm"] * depth_scale
    n = H.engine_setup(engine, w)
    assert n == 5_000_000 and len(w["T"]) == 100 and len(w["lines"]["nu"]) >= 4_900_000
    rad, tr = _column(engine, w)
    pts = H.boundary_points(n, 200, 44, n_tiles=16)
    assert len(pts) >= 256
    rad_ref, tr_ref = H.oracle_column_at(w, pts, w["t_surface"])
    phys = np.isfinite(tr_ref) & (tr_ref <= 1.0)                  # next to 0 cm-1 the reference's negative Doppler widths
    assert (~phys).sum() <= 2                                     # give k < 0 and "transmittances" above 1 (or inf)
    err_t = np.abs(tr[pts][phys] - tr_ref[phys])
    assert err_t.max() <= H.T_ABS_TOL, (err_t.max(), pts[phys][err_t.argmax()])
    ok = np.isfinite(rad_ref) & phys                              # nu = 0: NaN in the reference (0/0)
    assert np.isnan(rad[0]) and (~np.isfinite(rad)).sum() <= 16   # (FP32 overflow where k < 0 next to 0 cm-1)
    np.testing.assert_allclose(rad[pts][ok], rad_ref[ok], rtol=2e-5)
    if depth_scale < 1:                                           # the thin column has structure at these points
        assert (tr_ref[phys] < 0.45).sum() > 20 and (tr_ref[phys] > 0.6).sum() > 20


@pytest.mark.parametrize("cfg", ["cfg2", "cfg5"])
def test_fused_epilogue_full_size_transmittance_and_radiance(engine, cfg):
    """cfg2 / cfg5 as ONE engine call (prb_atmosphere with one layer: K2's fused epilogue turns the finished k tile into
    transmittance and radiance): |dT| <= 1e-6 and radiance rtol 2e-5 against the oracle at >= 1000 boundary-inclusive
    points -- k within 1e-5 does not by itself bound T."""
    w = workloads.cfg2() if cfg == "cfg2" else workloads.cfg5()
    n = H.engine_setup(engine, w)
    rad, tr = _column(engine, w)
    assert engine.atmosphere_launches() == 2                       # K1 + K2 (fused), no separate K3
    pts = sample_points(n, 1000, 21)
    rad_ref, tr_ref = H.oracle_column_at(_as_column(w), pts, 288.0)
    phys = np.isfinite(tr_ref) & (tr_ref <= 1.0)                  # k < 0 next to 0 cm-1 (negative Doppler widths)
    assert (~phys).sum() <= 4
    err_t = np.abs(tr[pts][phys] - tr_ref[phys])
    assert err_t.max() <= H.T_ABS_TOL, (err_t.max(), pts[phys][err_t.argmax()])
    ok = np.isfinite(rad_ref) & phys
    np.testing.assert_allclose(rad[pts][ok], rad_ref[ok], rtol=2e-5)
    assert 0.05 < np.median(tr_ref[phys]) < 0.95
