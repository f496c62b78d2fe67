"""Seeded random sweep of gas cells through the C ABI against the oracle: pressure (window length and line-shape regime),
temperature, grid resolution, cutoff, line density, species mix, the wavenumber chunk a rank would own, the K2 variant
(exact / far-field) and line-range parts all drawn together, so that the class boundaries of the kernels (span sizes,
window classes, far-field thresholds, narrow-window kernels, chunk edges) are met in combinations the hand-picked
parity cases do not list.  The oracle is evaluated in its gather form at boundary-inclusive sample points, so windows
of tens of thousands of points stay cheap on the CPU.  Tolerance: the north_star's rel <= 1e-5 on k(nu)."""
import os

import numpy as np
import pytest

from oracle import physics as ph
from pyrad_b200 import engine as eng
from pyrad_b200 import workloads
from tests import helpers as H

pytestmark = pytest.mark.gpu

#: cases per sweep in the suite; PRB_FUZZ_SCALE=25 runs the long sweep (1500 cells, 500 columns: a few minutes of oracle time)
SCALE = int(os.environ.get("PRB_FUZZ_SCALE", "1"))
TILE = 2048
SPECIES = ["h2o", "co2", "ch4", "o3"]
CONC = {"h2o": 0.01, "co2": 400e-6, "ch4": 1.8e-6, "o3": 5e-8}


def cluster_lines(w, rng):
    """Replace the jittered lattice of line positions by an uneven one: a few dense clusters (band heads: many lines on
    the same few grid points), repeated wavenumbers, and empty stretches in between."""
    from pyrad_b200 import synth
    lo = max(w["range_min"] - w["cutoff"], 0.0) + 1e-6
    hi = w["range_max"] + w["cutoff"] - 1e-6
    for ln in w["per_group_lines"]:
        m = len(ln["nu"])
        if m < 4:
            continue
        k = int(rng.integers(1, 6))
        centre = rng.uniform(lo, hi, k)
        width = w["res"] * np.exp(rng.uniform(np.log(0.5), np.log(3000.0), k))
        which = rng.integers(0, k, m)
        nu = centre[which] + width[which] * rng.standard_normal(m)
        stray = rng.random(m) < 0.2
        nu[stray] = rng.uniform(lo, hi, int(stray.sum()))
        nu[rng.random(m) < 0.05] = centre[0]                          # a pile of lines on one wavenumber
        ln["nu"] = np.sort(np.clip(np.round(nu, 6), lo, hi))
    w["lines"] = synth.merge_species_lines(w["per_group_lines"])
    w.pop("_idx_cache", None)


def draw_case(seed):
    rng = np.random.default_rng(1000 + seed)
    res = float(rng.choice([0.1, 0.01, 0.005, 0.002, 0.001]))
    P = float(np.exp(rng.uniform(np.log(0.05), np.log(3000.0))))
    T = int(rng.integers(180, 330))
    # the layer's own cutoff (5 cm-1 * P / p0) most of the time, else a cutoff drawn for a window of 1 .. 40 000 points
    cutoff = None if rng.random() < 0.6 else float(res * np.exp(rng.uniform(0.0, np.log(40000.0))))
    n_grid = int(np.exp(rng.uniform(np.log(300.0), np.log(400000.0))))
    rmin = float(rng.choice([0.0, 2.5, 600.0, 1000.0, 2349.0]))
    rmax = rmin + n_grid * res
    names = list(rng.choice(SPECIES, size=int(rng.integers(1, 5)), replace=False))
    # mean line spacing of 0.2 .. 50 grid points, bounded so that the whole case stays small
    n_lines = int(min(max(n_grid / np.exp(rng.uniform(np.log(0.2), np.log(50.0))), 8), 120000))
    w = workloads.gas_cell(names, n_lines, rmin, rmax, res, T, P, [CONC[s] for s in names], 10.0, 500 + seed, cutoff=cutoff)
    if rng.random() < 0.4:
        cluster_lines(w, rng)
    n = eng.grid_len(rmin, rmax, res)
    # chunk: the whole grid, or a tile-aligned slice as a rank of a sharded run owns it
    lo, hi = 0, n
    if n > 3 * TILE and rng.random() < 0.5:
        t = np.sort(rng.choice(np.arange(0, n // TILE + 1), size=2, replace=False))
        lo, hi = int(t[0]) * TILE, min(int(t[1]) * TILE, n)
    variant = eng.K2_FARFIELD if rng.random() < 0.5 else eng.K2_CLASSED
    split = bool(rng.random() < 0.3)
    return w, n, lo, hi, variant, split


@pytest.mark.parametrize("seed", range(60 * SCALE))
def test_random_cell_matches_oracle(engine, seed):
    w, n, lo, hi, variant, split = draw_case(seed)
    H.engine_setup(engine, w, lo, hi)
    wts = [eng.number_density_weight(c, w["P"], w["T"]) for c in w["conc"]]
    engine.set_k2_variant(variant, 0)
    engine.set_option(eng.OPT_SPLIT_TILES, 1 if split else 0)
    try:
        H.engine_prepass(engine, w, weights=wts)
        out = engine.line_sum()
    finally:
        engine.set_option(eng.OPT_SPLIT_TILES, 0)
        engine.set_k2_variant(eng.K2_CLASSED, 0)
    assert out.shape == (hi - lo,)
    pts = H.boundary_points(n, 150, seed, n_tiles=8)
    pts = np.unique(np.concatenate([pts, [lo, min(lo + 1, hi - 1), hi - 2 if hi - 2 >= lo else lo, hi - 1]]))
    pts = pts[(pts >= lo) & (pts < hi)]
    ref = H.oracle_layer_k_at(w, pts, w["T"], w["P"], w["conc"], w["cutoff"])
    # (the sample of a chunk may hold no line centre: the FP32 floor is taken from the line list's tallest core)
    err = H.k_rel_err(out[pts - lo], ref, peak=max(H.gaussian_peak(w, wts), float(np.abs(ref).max())))
    info = dict(seed=seed, P=w["P"], T=w["T"], res=w["res"], cutoff=w["cutoff"], n=n, chunk=(lo, hi),
                lines=len(w["lines"]["nu"]), variant=variant, split=split, window=eng.window_len(w["cutoff"], w["res"]))
    assert err.max() <= H.K_REL_TOL, (info, float(err.max()), int(pts[err.argmax()]))
    # the metric's numerator on the owned chunk (integer work: exact)
    if lo == 0 and hi == n:
        idx = ph.line_index(w["lines"]["nu"], w["range_min"], w["res"])
        assert engine.pair_count() == ph.pair_count(idx, n, eng.window_len(w["cutoff"], w["res"])), info


def draw_column(seed):
    rng = np.random.default_rng(7000 + seed)
    res = float(rng.choice([0.01, 0.005, 0.002, 0.001]))
    n_layers = int(rng.integers(1, 41))
    top_km = float(rng.uniform(15.0, 90.0))
    n_grid = int(np.exp(rng.uniform(np.log(2000.0), np.log(200000.0))))
    rmin = float(rng.choice([2.5, 600.0, 1000.0, 2349.0, 4300.0]))
    rmax = rmin + n_grid * res
    n_lines = int(min(max(n_grid / np.exp(rng.uniform(np.log(0.5), np.log(40.0))), 40), 100000))
    fixed = None if rng.random() < 0.7 else float(res * np.exp(rng.uniform(np.log(3.0), np.log(20000.0))))
    w = workloads.atmosphere(n_layers=n_layers, n_lines=n_lines, rmin=rmin, rmax=rmax, res=res, top_km=top_km,
                             fixed_cutoff=fixed, seed=900 + seed)
    # layer depth scaled so that columns from transparent to opaque are met (the fold's thin and thick forms)
    w["depth_cm"] = w["depth_cm"] * float(np.exp(rng.uniform(np.log(1e-7), np.log(1e-2))))
    w["t_surface"] = float(rng.uniform(220.0, 320.0))
    variant = eng.K2_FARFIELD if rng.random() < 0.5 else eng.K2_CLASSED
    split = bool(rng.random() < 0.3)
    return w, variant, split


@pytest.mark.parametrize("seed", range(20 * SCALE))
def test_random_column_matches_oracle(engine, seed):
    """The column path (one batched K1, K2 per layer class, K3 fold or the fused epilogue) on random columns: total
    transmittance to 1e-6 absolute, radiance to 2e-5 relative, at boundary-inclusive sample points."""
    w, variant, split = draw_column(seed)
    n = H.engine_setup(engine, w)
    sp = w["species"]
    win = [eng.window_len(c, w["res"]) for c in w["cutoff"]]
    qt = np.array([[s.q(T) for s in sp] for T in w["T"]])
    engine.set_k2_variant(variant, 0)
    engine.set_option(eng.OPT_SPLIT_TILES, 1 if split else 0)
    try:
        engine.atmosphere(w["depth_cm"], w["T"], w["P"], w["conc"], [s.molmass for s in sp], qt, [s.q296 for s in sp],
                          win, w["t_surface"], w["range_max"])
        rad, tr = engine.atmosphere_read()
    finally:
        engine.set_option(eng.OPT_SPLIT_TILES, 0)
        engine.set_k2_variant(eng.K2_CLASSED, 0)
    pts = H.boundary_points(n, 120, seed, n_tiles=6)
    rad_ref, tr_ref = H.oracle_column_at(w, pts, w["t_surface"])
    info = dict(seed=seed, layers=len(w["T"]), res=w["res"], n=n, lines=len(w["lines"]["nu"]), variant=variant, split=split,
                windows=(min(win), max(win)), rmin=w["range_min"], depth=float(w["depth_cm"][0]))
    dt = np.abs(tr[pts] - tr_ref)
    assert dt.max() <= H.T_ABS_TOL, (info, float(dt.max()), int(pts[dt.argmax()]))
    np.testing.assert_allclose(rad[pts], rad_ref, rtol=2e-5, atol=0, err_msg=str(info))


@pytest.mark.parametrize("seed", range(24 * SCALE))
def test_random_line_groups_rows_match_oracle(engine, seed):
    """prb_upload_line_groups / prb_line_sum_groups on random groupings: every species' list cut into a random number of
    isotopologue lists (some of them empty or a single line), per-group cross-section rows against the oracle."""
    rng = np.random.default_rng(3000 + seed)
    w, n, lo, hi, variant, _ = draw_case(5000 + seed)
    groups, meta = [], []
    for g, sp in enumerate(w["species"]):
        ln = w["per_group_lines"][g]
        m = len(ln["nu"])
        parts = int(rng.integers(1, 4))
        tag = rng.integers(0, parts, m) if rng.random() < 0.8 else np.zeros(m, dtype=np.int64)   # (then the others are empty)
        for q in range(parts):
            sel = tag == q
            groups.append({k: np.ascontiguousarray(np.asarray(v)[sel]) for k, v in ln.items()})
            meta.append(g)
    sp = [w["species"][g] for g in meta]
    conc = [w["conc"][g] for g in meta]
    win = eng.window_len(w["cutoff"], w["res"])
    engine.upload_line_groups(groups)
    engine.set_grid(w["range_min"], w["res"], n, lo, hi)
    engine.set_k2_variant(variant, 0)
    try:
        engine.layer_prepass(w["T"], w["P"], conc, [s.molmass for s in sp], [s.q(w["T"]) for s in sp], [s.q296 for s in sp], win)
        rows = engine.line_sum_groups()
    finally:
        engine.set_k2_variant(eng.K2_CLASSED, 0)
    assert rows.shape == (len(groups), hi - lo)
    pts = H.boundary_points(n, 100, seed, n_tiles=6)
    pts = np.unique(np.concatenate([pts, [lo, hi - 1]]))
    pts = pts[(pts >= lo) & (pts < hi)]
    wm = max(win - 2, 0)
    for j, (ln, s) in enumerate(zip(groups, sp)):
        if len(ln["nu"]) == 0:
            assert not rows[j].any(), j
            continue
        idx = ph.line_index(ln["nu"], w["range_min"], w["res"])
        sub = H._lines_near(idx, ln, pts, wm)
        ref = np.zeros(len(pts))
        if len(sub["nu"]):
            ref = ph.cross_section_at(pts, sub, w["T"], w["P"], conc[j], s.molmass, s.q(w["T"]), s.q296, w["range_min"],
                                      w["range_max"], w["res"], w["cutoff"])
        p = ph.LineParams(ln, w["T"], w["P"], conc[j], s.molmass, s.q(w["T"]), s.q296)
        ok = p.gD > 0
        peak = max(float(np.max(np.abs(p.S[ok]) / p.gD[ok])) / np.sqrt(np.pi) if ok.any() else 0.0, float(np.abs(ref).max()))
        err = H.k_rel_err(rows[j][pts - lo], ref, peak=peak)
        assert err.max() <= H.K_REL_TOL, (seed, j, len(ln["nu"]), float(err.max()), int(pts[err.argmax()]), win, variant)


@pytest.mark.parametrize("seed", range(12 * SCALE))
def test_random_cell_from_host_buffers_matches_oracle(engine, seed):
    """prb_gas_cell_host (pipelined upload, fused epilogue, results stored into host buffers) on random cells and chunks:
    transmittance to 1e-6 absolute against the oracle, and bitwise what upload + set_grid + atmosphere return."""
    import torch
    rng = np.random.default_rng(4000 + seed)
    w, n, lo, hi, variant, _ = draw_case(9000 + seed)
    sp = w["species"]
    win = eng.window_len(w["cutoff"], w["res"])
    depth = float(np.exp(rng.uniform(np.log(1e-2), np.log(1e5))))
    mol, q296, qt = [s.molmass for s in sp], [s.q296 for s in sp], [s.q(w["T"]) for s in sp]
    h_rad = torch.zeros(hi - lo, dtype=torch.float32).pin_memory()
    h_tr = torch.zeros(hi - lo, dtype=torch.float32).pin_memory()
    engine.set_k2_variant(variant, 0)
    engine.set_result_host(h_rad.numpy(), h_tr.numpy())
    try:
        engine.gas_cell_host(w["lines"], len(sp), w["range_min"], w["res"], n, lo, hi, depth, w["T"], w["P"], w["conc"],
                             mol, qt, q296, win, 288.0, w["range_max"])
        engine.synchronize()
        tr = h_tr.numpy().copy(); rad = h_rad.numpy().copy()
        engine.set_result_host()
        engine.upload_lines(w["lines"], n_groups=len(sp))
        engine.set_grid(w["range_min"], w["res"], n, lo, hi)
        engine.atmosphere([depth], [w["T"]], [w["P"]], [w["conc"]], mol, [qt], q296, [win], 288.0, w["range_max"])
        rad2 = np.empty(hi - lo, dtype=np.float32); tr2 = np.empty(hi - lo, dtype=np.float32)
        engine.atmosphere_read_f32(rad2, tr2)
    finally:
        engine.set_result_host()
        engine.set_k2_variant(eng.K2_CLASSED, 0)
    assert np.array_equal(tr, tr2) and np.array_equal(rad, rad2, equal_nan=True)
    pts = H.boundary_points(n, 100, seed, n_tiles=6)
    pts = pts[(pts >= lo) & (pts < hi)]
    k = H.oracle_layer_k_at(w, pts, w["T"], w["P"], w["conc"], w["cutoff"])
    t_ref = ph.transmittance(k, depth)
    got = tr[pts - lo]
    # (lines piled next to 0 cm-1 with a negative pressure shift get a negative Doppler width in the reference and with it
    # a negative k: exp(-k u) overflows there, in the reference's FP64 and in the engine's FP32 result alike)
    huge = ~np.isfinite(t_ref) | (t_ref > 1e30)
    assert np.all(~np.isfinite(got[huge]) | (got[huge] > 1e30))
    # (T > 1 where k < 0: an FP32 result, compared to FP32 resolution of exp(|k u|))
    dt = np.abs(got[~huge] - t_ref[~huge]) / np.maximum(1.0, 10.0 * np.abs(t_ref[~huge]))
    assert dt.size == 0 or dt.max() <= H.T_ABS_TOL, (seed, float(dt.max()), int(pts[~huge][dt.argmax()]), win, variant, depth)


@pytest.mark.parametrize("seed", range(24 * SCALE))
def test_random_xsc_tables_through_the_mirror_match_oracle(engine, tmp_path, seed):
    """xsc molecules through the host mirror (file on disk -> device parse -> np.interp onto the 0.01 grid when coarser ->
    aligned placement, pyradClasses.py:466-505) for random table ranges and resolutions relative to the layer's range:
    inside it, around it, on and off the 0.01 lattice, disjoint.  Where the reference's mergeArray is undefined (partial
    overlap: wrong-length result; off-lattice edge: ValueError) the mirror must raise, never return a guess."""
    from oracle import ref_harness as rh
    from pyrad_b200 import classes as C
    from pyrad_b200 import synth
    rng = np.random.default_rng(6000 + seed)
    C.set_engine(engine)
    old = (C.DATA_ROOT, C.BASE_RESOLUTION, C.Layer.hasAtmosphere)
    C.DATA_ROOT, C.BASE_RESOLUTION, C.Layer.hasAtmosphere = str(tmp_path), 0.01, False
    try:
        l_min = float(rng.choice([600.0, 800.0, 1000.5, 1234.25]))
        l_max = l_min + float(rng.choice([20.0, 57.5, 100.0, 300.0]))
        kind = rng.integers(0, 4)
        if kind == 0:      # inside the layer range
            f_min = l_min + round(float(rng.uniform(1.0, 5.0)), int(rng.integers(0, 3)))
            f_max = l_max - round(float(rng.uniform(1.0, 5.0)), int(rng.integers(0, 3)))
        elif kind == 1:    # around it
            f_min = l_min - round(float(rng.uniform(0.0, 40.0)), int(rng.integers(0, 3)))
            f_max = l_max + round(float(rng.uniform(0.0, 40.0)), int(rng.integers(0, 3)))
        elif kind == 2:    # overlapping one end
            f_min = l_min - round(float(rng.uniform(1.0, 30.0)), 1)
            f_max = l_min + round(float(rng.uniform(1.0, 15.0)), 1)
        else:              # disjoint
            f_min = l_max + round(float(rng.uniform(1.0, 30.0)), 1)
            f_max = f_min + round(float(rng.uniform(5.0, 40.0)), 1)
        f_min = max(f_min, 0.5)
        f_res = float(rng.choice([0.01, 0.01, 0.02, 0.05, 0.1]))
        fx, fy = synth.make_xsc_table(f_min, f_max, f_res, 70 + seed)
        fname = rh.write_xsc_file(str(tmp_path), "CFC11", 296.0, 760.0, f_min, f_max, f_res, fx, fy)
        layer = C.Layer(10.0, 296, 1013.25, l_min, l_max)
        try:
            want = ph.xsc_cross_section(layer.xAxis, fx, fy, f_min, f_max, f_res)
        except (ValueError, IndexError):
            want = None
        info = (seed, int(kind), l_min, l_max, f_min, f_max, f_res)
        if want is None or len(want) != len(layer.xAxis):
            from pyrad_b200 import _lib
            with pytest.raises((ValueError, IndexError, _lib.EngineError)):
                layer.addMolecule({"CFC11": fname}, concentration=1e-9)
            return
        xm = layer.addMolecule({"CFC11": fname}, concentration=1e-9)
        got = C.getCrossSection(xm)
        assert got.shape == want.shape, info
        np.testing.assert_allclose(got, want, rtol=1e-14, atol=0, err_msg=str(info))
    finally:
        C.DATA_ROOT, C.BASE_RESOLUTION, C.Layer.hasAtmosphere = old


@pytest.mark.parametrize("seed", range(16 * SCALE))
def test_random_mirror_layers_follow_the_reference_through_mutations(engine, tmp_path, seed):
    """The host mirror on an on-disk data tree (device CSV ingestion, kept-line filter, one-pass per-isotopologue rows,
    resident-state keys) through random changeTemperature / changePressure / changeDepth / setPPM / changeRange sequences:
    k to 1e-5 relative and T to 1e-6 after every step against tests.helpers.LayerModel, which
    tests/test_oracle_vs_reference.py holds the REAL reference to on the same cases."""
    from oracle import ref_harness as rh
    from pyrad_b200 import classes as C
    case = H.random_layer_case(seed if seed < 8 else 100 + seed, max_points=40000 if seed >= 8 else 2500,
                               max_lines=2000 if seed >= 8 else 60)
    C.set_engine(engine)
    old = (C.DATA_ROOT, C.BASE_RESOLUTION, C.Layer.hasAtmosphere)
    C.DATA_ROOT, C.BASE_RESOLUTION, C.Layer.hasAtmosphere = str(tmp_path), 0.01, False
    try:
        H.seed_layer_case(str(tmp_path), case, rh)
        model = H.LayerModel(case["species"], case["lines"], case["conc"], case["depth"], case["T"], case["P"], case["rmin"],
                             case["rmax"])

        def check(layer, model, tag):
            k, t = C.getAbsCoef(layer), C.getTransmittance(layer)
            assert layer.resolution == model.res and layer.distanceFromCenter == model.cutoff, tag
            k_ref = model.abs_coef()
            assert k.shape == k_ref.shape, tag
            assert H.k_rel_err(k, k_ref).max() <= H.K_REL_TOL, (seed, tag)
            tm = model.transmittance()
            assert np.abs(t - tm).max() <= H.T_ABS_TOL, (seed, tag)
            # the line survey made at load time (bit exact: S296 summed per bin in file order) and the derived spectra
            assert np.array_equal(layer.lineSurvey, model.line_survey()), (seed, tag)
            with np.errstate(all="ignore"):
                tau_ref, ab_ref = ph.optical_depth(tm), ph.absorbance(tm)
            fin = np.isfinite(tau_ref)
            np.testing.assert_allclose(C.getOpticalDepth(layer)[fin], tau_ref[fin], rtol=2e-5, atol=1e-6, err_msg=str((seed, tag)))
            np.testing.assert_allclose(C.getAbsorbance(layer)[fin], ab_ref[fin], rtol=2e-5, atol=1e-6, err_msg=str((seed, tag)))
            em = C.getEmissivity(layer)
            assert np.abs(em - ph.emissivity(tm)).max() <= H.T_ABS_TOL, (seed, tag)
            assert C.integrateSpectrum(em, res=.01) == pytest.approx(ph.integrate_spectrum(em, res=.01), rel=1e-12)

        H.drive_layer_case(C, case, model, check)
    finally:
        C.DATA_ROOT, C.BASE_RESOLUTION, C.Layer.hasAtmosphere = old


@pytest.mark.parametrize("seed", range(10 * SCALE))
def test_random_mirror_columns_match_the_layer_by_layer_chain(engine, tmp_path, seed):
    """Atmosphere.columnSpectrum (ONE prb_atmosphere call) on random stacks of mirror layers against the reference's own
    way of doing a column -- layer.transmission(spectrum) chained bottom to top (pyradClasses.py:784-787) -- evaluated
    with the oracle state model, and against the mirror's own chained calls."""
    from oracle import ref_harness as rh
    from pyrad_b200 import classes as C
    rng = np.random.default_rng(8000 + seed)
    case = H.random_layer_case(300 + seed, max_points=20000, max_lines=800)
    C.set_engine(engine)
    old = (C.DATA_ROOT, C.BASE_RESOLUTION, C.Layer.hasAtmosphere)
    C.DATA_ROOT, C.BASE_RESOLUTION, C.Layer.hasAtmosphere = str(tmp_path), 0.01, False
    try:
        H.seed_layer_case(str(tmp_path), case, rh)
        atm = C.Atmosphere("random column")
        models = []
        for _ in range(int(rng.integers(2, 9))):
            T = int(rng.integers(200, 310))
            P = float(np.exp(rng.uniform(np.log(0.2), np.log(1100.0))))
            depth = float(np.exp(rng.uniform(np.log(1.0), np.log(3e4))))
            conc = [c * float(rng.uniform(0.2, 2.0)) for c in case["conc"]]
            layer = atm.addLayer(depth, T, P, case["rmin"], case["rmax"], dynamicResolution=False)
            for n, c in zip(case["names"], conc):
                layer.addMolecule(n, concentration=c)
            models.append(H.LayerModel(case["species"], case["lines"], conc, depth, T, P, case["rmin"], case["rmax"],
                                       dynamic=False))
        t_surf = float(rng.uniform(230.0, 320.0))
        rad, tot = atm.columnSpectrum(t_surf)
        xa = ph.x_axis(case["rmin"], case["rmax"], .01)
        rad_ref, tot_ref = ph.planck_wavenumber(xa, t_surf), np.ones(len(xa))
        chain = atm[0].planck(t_surf)
        for layer, m in zip(atm, models):
            t = m.transmittance()
            rad_ref = ph.transmission(t, rad_ref, ph.planck_wavenumber(xa, m.T))
            tot_ref = tot_ref * t
            chain = layer.transmission(chain)
        info = (seed, len(models), [round(m.P, 2) for m in models])
        assert np.abs(tot - tot_ref).max() <= H.T_ABS_TOL, info
        np.testing.assert_allclose(rad, rad_ref, rtol=2e-5, atol=0, err_msg=str(info))
        np.testing.assert_allclose(chain, rad_ref, rtol=2e-5, atol=0, err_msg=str(info))
    finally:
        C.DATA_ROOT, C.BASE_RESOLUTION, C.Layer.hasAtmosphere = old


def make_csv_case(seed):
    """Random HITRAN-online CSV text: (text, data rows as the reference's reader sees them, lo, hi)."""
    from pyrad_b200 import synth
    rng = np.random.default_rng(12000 + seed)
    n = int(rng.choice([0, 1, 2, 7, 100, 1000, 4000]))
    ln = synth.make_lines(max(n, 1), 480.0, 830.0, 7 * seed + 1)
    n = min(n, len(ln["nu"]))
    fmts = [repr, lambda v: "%.6f" % v, lambda v: "%.3E" % v, lambda v: "%.10e" % v, lambda v: "%+.8E" % v,
            lambda v: " %r " % v, lambda v: "%.17g" % v]
    rows = []
    for j in range(n):
        cells = ["2", "1"] + [fmts[int(rng.integers(0, len(fmts)))](float(ln[k][j])) for k in
                              ("nu", "sw", "a", "elower", "gamma_air", "gamma_self", "delta_air", "n_air")]
        rows.append(",".join(cells))
    for _ in range(int(rng.integers(0, 4)) if n else 0):               # repeated wavenumbers with another last cell
        j = int(rng.integers(0, n))
        rows.append(rows[j].rsplit(",", 1)[0] + ",0.%03d" % int(rng.integers(0, 999)))
    order = rng.integers(0, 3)
    if order == 1:
        rows = [rows[i] for i in rng.permutation(len(rows))]
    elif order == 2 and len(rows) > 4:                                  # two ascending runs
        cut = int(rng.integers(1, len(rows) - 1))
        rows = rows[cut:] + rows[:cut]
    with_comments = list(rows)
    for _ in range(int(rng.integers(0, 3))):
        with_comments.insert(int(rng.integers(0, len(with_comments) + 1)), "# a comment row")
    eol = "\r\n" if rng.random() < 0.3 else "\n"
    tail = eol if rng.random() < 0.5 else ""
    text = "# header" + eol + (eol.join(with_comments) + tail if with_comments else "")     # (a blank row crashes the reference)
    lo = float(rng.choice([0.0, 500.0, float(ln["nu"][0]), 600.123456]))
    hi = float(rng.choice([800.0, 5000.0, float(ln["nu"][n - 1]) if n else 700.0, 650.0]))
    return text, [r + ("\r" if eol == "\r\n" else "") for r in rows], lo, hi


@pytest.mark.parametrize("seed", range(20 * SCALE))
def test_random_hitran_csv_text_is_ingested_like_the_reference_reader(engine, seed):
    """HITRAN-online CSV text with everything the reference's reader tolerates (pyradUtilities.py:421-448) thrown together
    at random -- cell formats (repr, fixed, %E with few or many digits, a leading '+', padding blanks), rows in file order,
    shuffled or partly shuffled, repeated wavenumbers (the last row wins wherever it stands), comment rows, LF or CRLF,
    with and without a final newline, range bounds that fall on rows -- parsed on the device: the same lines as the
    oracle's dict-keyed reader, ascending, every column bit for bit."""
    from pyrad_b200 import hitran_io
    text, data_rows, lo, hi = make_csv_case(seed)
    ref = ph.read_hitran_online_rows(data_rows, lo, hi) if data_rows else {k: np.zeros(0) for k in hitran_io.LINE_COLUMNS}
    asc = np.argsort(ref["nu"], kind="stable")
    got_n = engine.ingest_csv(text.encode(), lo, hi)
    assert got_n == len(ref["nu"]), (seed, got_n, len(ref["nu"]))
    if got_n:
        got = engine.download_lines()
        for k in hitran_io.LINE_COLUMNS:
            np.testing.assert_array_equal(got[k], ref[k][asc], err_msg="%s seed %d" % (k, seed))


def make_xsc_text(seed):
    """Random two-column xsc table text: good rows in several number formats and spacings mixed with everything
    returnXscFileContents skips (pyradUtilities.py:680-696)."""
    rng = np.random.default_rng(15000 + seed)
    n = int(rng.choice([0, 1, 3, 50, 2000]))
    fmts = [repr, lambda v: "%.6f" % v, lambda v: "%.4E" % v, lambda v: "%.12e" % v, lambda v: "%+.5E" % v]
    rows = []
    for j in range(n):
        a, b = 700.0 + 0.02 * j, float(10.0 ** rng.uniform(-24, -17))
        kind = rng.random()
        fa, fb = fmts[int(rng.integers(0, len(fmts)))](a), fmts[int(rng.integers(0, len(fmts)))](b)
        sep = " " * int(rng.integers(1, 8))
        if kind < 0.75:
            rows.append(" " * int(rng.integers(0, 3)) + fa + sep + fb + " " * int(rng.integers(0, 3)))
        elif kind < 0.80:
            rows.append(fa + "\t" + fb)                 # a tab is not a separator: one token, skipped
        elif kind < 0.85:
            rows.append(fa + sep + fb + sep + "7")      # three tokens: skipped
        elif kind < 0.90:
            rows.append("abc" + sep + fb)               # not a number: skipped
        elif kind < 0.94:
            rows.append("")                             # blank: skipped
        elif kind < 0.97:
            rows.append("# a comment" if j else fa + sep + fb)    # (a LEADING '#' row is dropped before the parse)
        else:
            rows.append(fa + sep + fb + "\t")           # trailing tab: stripped
    eol = "\r\n" if rng.random() < 0.3 else "\n"
    return "# header line" + eol + eol.join(rows) + (eol if rows and rng.random() < 0.5 else ""), eol


@pytest.mark.parametrize("seed", range(12 * SCALE))
def test_random_xsc_table_text_is_parsed_like_the_reference_reader(engine, seed):
    text, eol = make_xsc_text(seed)
    ref_rows = text.split("\n")[1:]                      # (the reader sees rows split at LF; a CR stays and is stripped)
    wn_ref, xs_ref = ph.read_xsc_rows(ref_rows)
    wn, xs = engine.parse_xsc_text(text.encode())
    np.testing.assert_array_equal(wn, wn_ref, err_msg=str(seed))
    np.testing.assert_array_equal(xs, xs_ref, err_msg=str(seed))
