"""GPU parity tests proper: the CUDA path through the C ABI versus the CPU oracle on the same
seeded inputs.  Tolerances are the north_star's: rel <= 1e-5 on k(nu), abs <= 1e-6 on transmittance."""
import numpy as np
import pytest

from oracle import physics as ph
from pyrad_b200 import _lib
from pyrad_b200 import engine as eng
from pyrad_b200 import workloads
from tests import helpers as H

pytestmark = pytest.mark.gpu


def small_cell(P=1013.0, T=296, res=0.01, rmin=600.0, rmax=700.0, n_lines=3000, names=("co2",), conc=(400e-6,),
               seed=7, cutoff=None):
    return workloads.gas_cell(list(names), n_lines, rmin, rmax, res, T, P, list(conc), 10.0, seed, cutoff=cutoff)


def test_k1_line_params_match_oracle(engine):
    w = small_cell(P=500.0, T=250, names=("co2", "h2o"), conc=(400e-6, 0.01))
    H.engine_setup(engine, w)
    H.engine_prepass(engine, w)
    d = engine.debug_line_params()
    L = w["lines"]
    grp = L["group"]
    for g, sp in enumerate(w["species"]):
        m = grp == g
        sub = {k: v[m] for k, v in L.items()}
        p = ph.LineParams(sub, w["T"], w["P"], w["conc"][g], sp.molmass, sp.q(w["T"]), sp.q296)
        np.testing.assert_allclose(d["nu_shift"][m], p.nu_shift, rtol=1e-14)
        np.testing.assert_allclose(d["gamma_l"][m], p.gL, rtol=1e-13)
        np.testing.assert_allclose(d["gamma_d"][m], p.gD, rtol=1e-13)
        np.testing.assert_allclose(d["s_t"][m], p.S, rtol=1e-12)
        assert np.array_equal(d["regime"][m], p.regime)
        assert np.array_equal(d["index"][m], ph.line_index(sub["nu"], w["range_min"], w["res"]))   # bit exact


@pytest.mark.parametrize("variant", [eng.K2_GENERAL, eng.K2_CLASSED])
@pytest.mark.parametrize("ppt", [2, 4, 8, 16])
def test_k2_cross_section_small(engine, variant, ppt):
    w = small_cell()
    H.engine_setup(engine, w)
    engine.set_k2_variant(variant, ppt)
    try:
        H.engine_prepass(engine, w)
        out = engine.line_sum()
    finally:
        engine.set_k2_variant(eng.K2_CLASSED, 0)
    ref = H.oracle_sigma_groups(w).sum(axis=0)
    err = H.k_rel_err(out, ref)
    assert err.max() <= H.K_REL_TOL, (err.max(), int(err.argmax()))


@pytest.mark.parametrize("P,T", [(1013.25, 296), (300.0, 230), (50.0, 220), (5.0, 250), (0.5, 270), (0.05, 200),
                                 (3000.0, 320)])
def test_k2_pressure_regimes(engine, P, T):
    """Lorentz-dominated -> Voigt -> Gaussian-dominated; windows from thousands of points down to W = 1."""
    w = small_cell(P=P, T=T, n_lines=2000, names=("co2", "h2o"), conc=(400e-6, 0.01), seed=11)
    H.engine_setup(engine, w)
    H.engine_prepass(engine, w)
    out = engine.line_sum()
    ref = H.oracle_sigma_groups(w).sum(axis=0)
    err = H.k_rel_err(out, ref)
    assert err.max() <= H.K_REL_TOL, (P, T, err.max(), int(err.argmax()))
    n = eng.grid_len(w["range_min"], w["range_max"], w["res"])
    idx = ph.line_index(w["lines"]["nu"], w["range_min"], w["res"])
    assert engine.pair_count() == ph.pair_count(idx, n, eng.window_len(w["cutoff"], w["res"]))


def test_k2_fine_grid_and_weights(engine):
    """0.001 cm-1 grid, 4 species, absorption-coefficient weights folded into K1."""
    w = workloads.gas_cell(["h2o", "co2", "ch4", "o3"], 4000, 1000.0, 1030.0, 0.001, 280, 800.0,
                           [0.01, 400e-6, 1.8e-6, 5e-8], 10.0, 23)
    H.engine_setup(engine, w)
    wts = [eng.number_density_weight(c, w["P"], w["T"]) for c in w["conc"]]
    H.engine_prepass(engine, w, weights=wts)
    out = engine.line_sum()
    sig = H.oracle_sigma_groups(w)
    ref = sum(ph.abs_coef(sig[g], w["conc"][g], w["P"], w["T"]) for g in range(4))
    err = H.k_rel_err(out, ref)
    assert err.max() <= H.K_REL_TOL, err.max()


def test_k2_chunk_partition_is_bitwise_invariant(engine):
    """Multi-GPU sharding property on one GPU: tile-aligned chunks reproduce the unchunked result bit for bit."""
    w = small_cell(n_lines=4000, rmin=600.0, rmax=760.0)
    n = H.engine_setup(engine, w)
    H.engine_prepass(engine, w)
    full = engine.line_sum()
    cuts = [0, 4096, 12288, n]
    parts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        H.engine_setup(engine, w, a, b)
        H.engine_prepass(engine, w)
        parts.append(engine.line_sum())
    assert np.array_equal(np.concatenate(parts), full)


def test_k2_edge_cases(engine):
    # no lines at all
    w = small_cell(n_lines=2000)
    empty = {k: v[:0] for k, v in w["lines"].items()}
    engine.upload_lines(empty, 1)
    engine.set_grid(600.0, 0.01, 10000)
    engine.layer_prepass(296, 1013.0, [4e-4], [43.98983], [286.09], [286.09], 500)
    assert np.array_equal(engine.line_sum(), np.zeros(10000))
    # a single line just below rangeMin lands on index 0 (truncation toward zero, pyradClasses.py:390)
    one = {k: np.array([v]) for k, v in dict(nu=599.995, sw=1e-20, gamma_air=.07, gamma_self=.09, elower=100.,
                                             n_air=.7, delta_air=-.002).items()}
    engine.upload_lines(one, 1)
    n_r = eng.grid_len(600.0, 607.77, 0.01)     # ragged: not a multiple of any tile
    engine.set_grid(600.0, 0.01, n_r)
    engine.layer_prepass(296, 1013.0, [4e-4], [43.98983], [286.09], [286.09], eng.window_len(ph.layer_cutoff(1013.0), .01))
    out = engine.line_sum()
    ref = ph.cross_section(one, 296, 1013.0, 4e-4, 43.98983, 286.09, 286.09, 600.0, 607.77, 0.01, ph.layer_cutoff(1013.0))
    assert len(ref) == n_r
    assert H.k_rel_err(out, ref).max() <= H.K_REL_TOL
    assert engine.debug_line_params()["index"][0] == 0


def test_k3_layer_stream_and_planck(engine):
    rng = np.random.default_rng(5)
    n = 30000
    sigma = 10.0 ** rng.uniform(-26, -19, (3, n))
    conc = [400e-6, 0.01, 1.8e-6]
    P, T, depth = 900.0, 275, 1000.0
    wts = [eng.number_density_weight(c, P, T) for c in conc]
    axis = eng.linspace_axis(500.0, 800.0, n)
    xa = ph.x_axis(500.0, 800.0, 0.01)
    surf = ph.planck_wavenumber(xa, 288)
    k, t, r = engine.layer_stream(sigma, wts, depth, T, axis, surf)
    k_ref = sum(ph.abs_coef(sigma[m], conc[m], P, T) for m in range(3))
    t_ref = ph.transmittance(k_ref, depth)
    r_ref = ph.transmission(t_ref, surf, ph.planck_wavenumber(xa, T))
    np.testing.assert_allclose(k, k_ref, rtol=1e-13)
    assert np.abs(t - t_ref).max() <= 1e-12
    np.testing.assert_allclose(r, r_ref, rtol=1e-11)
    np.testing.assert_allclose(engine.planck(axis, 288), surf, rtol=1e-12)
    # nu = 0 -> NaN like the reference (0/0 with errors silenced, pyradPlanck.py:2)
    p0 = engine.planck(eng.linspace_axis(0.0, 10.0, 1000), 250)
    assert np.isnan(p0[0]) and np.all(np.isfinite(p0[1:]))


def test_atmosphere_small(engine):
    """8-layer column: device K1+K2 per layer + FP32 fold versus the oracle's layer-by-layer transmission."""
    w = workloads.atmosphere(n_layers=8, n_lines=6000, rmin=600.0, rmax=640.0, res=0.001, top_km=40.0)
    n = H.engine_setup(engine, w)
    sp = w["species"]
    L = len(w["T"])
    win = [eng.window_len(c, w["res"]) for c in w["cutoff"]]
    qt = np.array([[s.q(T) for s in sp] for T in w["T"]])
    engine.atmosphere(w["depth_cm"], w["T"], w["P"], w["conc"], [s.molmass for s in sp], qt, [s.q296 for s in sp],
                      win, w["t_surface"], w["range_max"])
    rad, tr = engine.atmosphere_read()
    xa = ph.x_axis(w["range_min"], w["range_max"], w["res"])
    I = ph.planck_wavenumber(xa, w["t_surface"])
    ttot = np.ones(n)
    for l in range(L):
        sig = H.oracle_sigma_groups(w, T=w["T"][l], P=w["P"][l], conc=w["conc"][l], cutoff=w["cutoff"][l])
        k = sum(ph.abs_coef(sig[g], w["conc"][l][g], w["P"][l], w["T"][l]) for g in range(len(sp)))
        t = ph.transmittance(k, w["depth_cm"][l])
        I = ph.transmission(t, I, ph.planck_wavenumber(xa, w["T"][l]))
        ttot = ttot * t
    assert np.abs(tr - ttot).max() <= H.T_ABS_TOL, np.abs(tr - ttot).max()
    np.testing.assert_allclose(rad, I, rtol=2e-5)


@pytest.mark.parametrize("P", [300.0, 150.0, 30.0, 2.0])
def test_k2_narrow_kernel_matches_wide_kernel_and_oracle(engine, P):
    """Upper-atmosphere windows take the thread-per-point kernels (k2_point with table-driven line ranges up to
    W-2 = 511, here P <= 150; k2_narrow with binary searches above that, here P = 300 with the threshold forced);
    forcing the wide kernel on the same inputs must agree, and both must match the oracle.  Shard invariance holds
    for them as well."""
    w = workloads.gas_cell(["co2", "h2o"], 6000, 600.0, 660.0, 0.0025, 240, P, [400e-6, 0.01], 10.0, 31)
    n = H.engine_setup(engine, w)
    ref = H.oracle_sigma_groups(w).sum(axis=0)
    try:
        engine.set_narrow_threshold(0)                       # wide kernel only
        H.engine_prepass(engine, w)
        wide = engine.line_sum()
        engine.set_narrow_threshold(1 << 20)                 # narrow kernel only
        H.engine_prepass(engine, w)
        narrow = engine.line_sum()
        parts = []
        for a, b in [(0, 8192), (8192, n)]:
            H.engine_setup(engine, w, a, b)
            H.engine_prepass(engine, w)
            parts.append(engine.line_sum())
    finally:
        engine.set_narrow_threshold(-1)
    assert H.k_rel_err(wide, ref).max() <= H.K_REL_TOL
    assert H.k_rel_err(narrow, ref).max() <= H.K_REL_TOL
    assert np.array_equal(np.concatenate(parts), narrow)


def test_k2_narrow_kernel_many_chunks_per_tile(engine):
    """Dense band: ~10 lines per grid point, so a narrow-kernel tile stages many 2048-record chunks."""
    w = workloads.gas_cell(["co2"], 60000, 600.0, 615.0, 0.0025, 240, 150.0, [400e-6], 10.0, 37)
    H.engine_setup(engine, w)
    ref = H.oracle_sigma_groups(w).sum(axis=0)
    try:
        engine.set_narrow_threshold(1 << 20)
        H.engine_prepass(engine, w)
        narrow = engine.line_sum()
        engine.set_narrow_threshold(0)
        H.engine_prepass(engine, w)
        wide = engine.line_sum()
    finally:
        engine.set_narrow_threshold(-1)
    assert H.k_rel_err(wide, ref).max() <= H.K_REL_TOL
    assert H.k_rel_err(narrow, ref).max() <= H.K_REL_TOL


@pytest.mark.parametrize("narrow", [False, True])
def test_lines_next_to_zero_wavenumber_with_negative_shift(engine, narrow):
    """Reference quirk: nu* = nu0 + delta*P/p0 can go negative next to 0 cm-1, which makes the Doppler width --
    and the Gaussian-regime contribution -- negative (pyradClasses.py:252-263, 378-381).  Both kernels follow."""
    w = workloads.gas_cell(["co2", "h2o"], 3000, 0.0, 4.0, 0.001, 250, 353.4, [400e-6, 0.01], 10.0, 41)
    H.engine_setup(engine, w)
    ref = H.oracle_sigma_groups(w).sum(axis=0)
    assert ref.min() < 0                                   # the quirk is actually exercised
    try:
        engine.set_narrow_threshold((1 << 20) if narrow else 0)
        H.engine_prepass(engine, w)
        out = engine.line_sum()
    finally:
        engine.set_narrow_threshold(-1)
    floor = H.K_FLOOR_REL * np.abs(ref).max()
    err = np.abs(out - ref) / np.maximum(np.abs(ref), floor)
    assert err.max() <= H.K_REL_TOL, (err.max(), int(err.argmax()))


def _run_atm(engine, w, win=None):
    sp = w["species"]
    win = [eng.window_len(c, w["res"]) for c in w["cutoff"]] if win is None else win
    qt = np.array([[s.q(T) for s in sp] for T in w["T"]])
    engine.atmosphere(w["depth_cm"], w["T"], w["P"], w["conc"], [s.molmass for s in sp], qt, [s.q296 for s in sp],
                      win, w["t_surface"], w["range_max"])
    rad = np.empty(engine.n_chunk, dtype=np.float32)
    tr = np.empty(engine.n_chunk, dtype=np.float32)
    engine.atmosphere_read_f32(rad, tr)
    return rad, tr


def test_atmosphere_batched_launches_are_bitwise_equal_to_per_layer_launches(engine):
    """The multi-layer launches (one K1 + one K2 per kernel class, layers sorted by window) only change WHEN a
    (layer, tile) item runs, never its arithmetic: spectra must be identical to the layer-at-a-time schedule,
    also when the record budget forces several batches."""
    w = workloads.atmosphere(n_layers=24, n_lines=8000, rmin=600.0, rmax=660.0, res=0.001, top_km=60.0)
    w["lines"] = {k: v[:-1] for k, v in w["lines"].items()}
    H.engine_setup(engine, w)
    assert engine.n_lines % 4 != 0                                  # per-layer record slices must stay TMA aligned
    try:
        engine.set_option(eng.OPT_BATCH_LAYERS, 0)
        rad0, tr0 = _run_atm(engine, w)
        n0 = engine.atmosphere_launches()
        engine.set_option(eng.OPT_BATCH_LAYERS, 1)
        rad1, tr1 = _run_atm(engine, w)
        n1 = engine.atmosphere_launches()
        engine.set_option(eng.OPT_RECORD_BUDGET_MB, 2)          # 8000 lines x 36 B -> 7 layers per batch
        rad2, tr2 = _run_atm(engine, w)
        n2 = engine.atmosphere_launches()
    finally:
        engine.set_option(eng.OPT_BATCH_LAYERS, 1)
        engine.set_option(eng.OPT_RECORD_BUDGET_MB, 0)
    assert np.array_equal(rad0, rad1) and np.array_equal(tr0, tr1)
    assert np.array_equal(rad0, rad2) and np.array_equal(tr0, tr2)
    assert n0 >= 2 * 24 + 1 and n1 <= 1 + 5 + 2 + 1 and n1 < n2 < n0   # K1 + (ppt 8, 4, 2, point, narrow) + 2 tile-bound passes + K3


def test_single_layer_fused_epilogue_equals_k3_fold(engine):
    """Gas cell (one layer): K2's fused epilogue (k -> T, Planck, transmission) is the same arithmetic as the
    separate K3 fold -- bitwise -- and one launch fewer."""
    w = workloads.atmosphere(n_layers=1, n_lines=5000, rmin=600.0, rmax=650.0, res=0.001, top_km=2.0)
    H.engine_setup(engine, w)
    try:
        engine.set_option(eng.OPT_FUSE_SINGLE_LAYER, 0)
        rad0, tr0 = _run_atm(engine, w)
        n0 = engine.atmosphere_launches()
        engine.set_option(eng.OPT_FUSE_SINGLE_LAYER, 1)
        rad1, tr1 = _run_atm(engine, w)
        n1 = engine.atmosphere_launches()
    finally:
        engine.set_option(eng.OPT_FUSE_SINGLE_LAYER, 1)
    assert (n0, n1) == (3, 2)
    assert np.array_equal(rad0, rad1) and np.array_equal(tr0, tr1)
    assert 0 <= tr1.min() < tr1.max() <= 1 and np.isfinite(rad1[1:]).all()


def test_peer_gather_single_rank_roundtrip(engine):
    """world = 1 exercises the whole peer path on one GPU: IPC export, slot addressing, double buffering by
    epoch, the signal/wait kernel -- the gathered row must equal the rank's own spectra, step after step."""
    w = workloads.atmosphere(n_layers=3, n_lines=4000, rmin=600.0, rmax=630.0, res=0.001, top_km=30.0)
    H.engine_setup(engine, w)
    rad_ref, tr_ref = _run_atm(engine, w)
    import ctypes as C
    try:
        h = engine.peer_alloc(0, 1, engine.n_chunk)
        assert len(h) == eng.PEER_HANDLE_BYTES
        engine.peer_connect([h])
        for step in range(3):
            rad, tr = _run_atm(engine, w)
            assert np.array_equal(rad, rad_ref) and np.array_equal(tr, tr_ref)
            rp, tp, ld = engine.peer_gathered_dev()
            assert ld >= engine.n_chunk and rp and tp
            own_r, own_t = engine.atmosphere_result_dev()
            assert (own_r, own_t) == (rp, tp)                       # rank 0's slot is row 0 of the gather buffer
    finally:
        engine.peer_disconnect()
    rad, tr = _run_atm(engine, w)                                   # back to the local result arrays
    assert np.array_equal(rad, rad_ref)


def test_atmosphere_integrals_on_device(engine):
    """integrateSpectrum of the device-resident radiance (K4 reduction) versus numpy on the read-back arrays."""
    w = workloads.atmosphere(n_layers=4, n_lines=4000, rmin=600.0, rmax=640.0, res=0.001, top_km=20.0)
    H.engine_setup(engine, w)
    rad, tr = _run_atm(engine, w)
    integ, tsum = engine.atmosphere_integrate(np.pi, w["res"])
    assert integ == pytest.approx(float(np.sum(np.nan_to_num(rad.astype(np.float64))) * np.pi * w["res"]), rel=1e-12)
    assert tsum == pytest.approx(float(np.sum(tr.astype(np.float64))), rel=1e-12)


@pytest.mark.parametrize("n_layers,top", [(1, 2.0), (5, 40.0)])
def test_zero_copy_host_results_equal_read_back(engine, n_layers, top):
    """prb_set_result_host: the kernels store the finished spectra into pinned host memory themselves (K2's fused
    epilogue for one layer, K3 otherwise); the bytes must equal an explicit read-back."""
    import torch
    w = workloads.atmosphere(n_layers=n_layers, n_lines=6000, rmin=600.0, rmax=650.0, res=0.001, top_km=top)
    H.engine_setup(engine, w)
    rad_ref, tr_ref = _run_atm(engine, w)
    h_rad = torch.zeros(engine.n_chunk, dtype=torch.float32).pin_memory()
    h_tr = torch.zeros(engine.n_chunk, dtype=torch.float32).pin_memory()
    engine.set_result_host(h_rad.numpy(), h_tr.numpy())
    try:
        rad, tr = _run_atm(engine, w)
        assert np.array_equal(h_rad.numpy(), rad_ref) and np.array_equal(h_tr.numpy(), tr_ref)
        assert np.array_equal(rad, rad_ref) and np.array_equal(tr, tr_ref)      # the device copy is still there
        with pytest.raises(Exception):
            engine.set_result_host(np.zeros(engine.n_chunk, dtype=np.float32), np.zeros(engine.n_chunk, dtype=np.float32))
    finally:
        engine.set_result_host()


def test_pipelined_gas_cell_equals_separate_calls(engine):
    """prb_gas_cell_host: line columns uploaded in wavenumber pieces on a copy stream, one wave of K2 tiles per piece,
    results stored into pinned host buffers by K2 itself -- bitwise equal to upload_lines + set_grid + atmosphere."""
    import torch
    w = workloads.gas_cell(["h2o", "co2", "ch4", "o3"], 60000, 0.0, 1300.0, 0.001, 296, 1013.25,
                           [0.01, 400e-6, 1.8e-6, 5e-8], 10.0, 77)
    sp = w["species"]
    n = eng.grid_len(w["range_min"], w["range_max"], w["res"])
    assert n > 2 * 296 * 2048                                       # more than two waves of tiles: really pipelined
    win = eng.window_len(w["cutoff"], w["res"])
    mol, q296, qt = [s.molmass for s in sp], [s.q296 for s in sp], [s.q(w["T"]) for s in sp]
    engine.upload_lines(w["lines"], n_groups=len(sp))
    engine.set_grid(w["range_min"], w["res"], n)
    engine.atmosphere([w["depth_cm"]], [w["T"]], [w["P"]], [w["conc"]], mol, [qt], q296, [win], 288.0, w["range_max"])
    rad_ref = np.empty(n, dtype=np.float32); tr_ref = np.empty(n, dtype=np.float32)
    engine.atmosphere_read_f32(rad_ref, tr_ref)
    pairs_ref = None
    h_rad = torch.zeros(n, dtype=torch.float32).pin_memory()
    h_tr = torch.zeros(n, dtype=torch.float32).pin_memory()
    engine.set_result_host(h_rad.numpy(), h_tr.numpy())
    try:
        for a, b in ((0, n), (4096 * 20, n - 4096 * 30)):           # the whole grid, then an interior chunk
            h_rad.zero_(); h_tr.zero_()
            engine.gas_cell_host(w["lines"], len(sp), w["range_min"], w["res"], n, a, b, w["depth_cm"], w["T"], w["P"],
                                 w["conc"], mol, qt, q296, win, 288.0, w["range_max"])
            assert np.array_equal(h_rad.numpy()[: b - a], rad_ref[a:b], equal_nan=True) and np.array_equal(h_tr.numpy()[: b - a], tr_ref[a:b])
            rad = np.empty(b - a, dtype=np.float32); tr = np.empty(b - a, dtype=np.float32)
            engine.atmosphere_read_f32(rad, tr)
            assert np.array_equal(rad, rad_ref[a:b], equal_nan=True) and np.array_equal(tr, tr_ref[a:b])
        # the engine is left as after the separate calls: the prepass / pair count work on the resident list
        engine.layer_prepass(w["T"], w["P"], w["conc"], mol, qt, q296, win)
        assert engine.pair_count() > 0
        bad = dict(w["lines"]); bad["nu"] = bad["nu"][::-1].copy()
        with pytest.raises(Exception):
            engine.gas_cell_host(bad, len(sp), w["range_min"], w["res"], n, 0, n, w["depth_cm"], w["T"], w["P"],
                                 w["conc"], mol, qt, q296, win, 288.0, w["range_max"])
    finally:
        engine.set_result_host()


@pytest.mark.parametrize("n_layers,rmax", [(24, 660.0), (9, 607.777), (8, 601.003)])
def test_tma_staged_fold_equals_register_fold(engine, n_layers, rmax):
    """k3_fold_tma (k matrix through a shared-memory ring fed by TMA bulk copies, Planck term interpolated across each
    thread's four points) against k3_fold_f32 (register-held loads, per-point Planck): the transmittance is the same
    arithmetic (bitwise), the radiance agrees to FP32 rounding -- full strips, ragged last strips, ragged layer groups."""
    w = workloads.atmosphere(n_layers=n_layers, n_lines=6000, rmin=600.0, rmax=rmax, res=0.001, top_km=50.0)
    H.engine_setup(engine, w)
    try:
        engine.set_option(eng.OPT_FOLD_TMA, 0)
        rad0, tr0 = _run_atm(engine, w)
        engine.set_option(eng.OPT_FOLD_TMA, 1)
        rad1, tr1 = _run_atm(engine, w)
    finally:
        engine.set_option(eng.OPT_FOLD_TMA, 1)
    assert np.array_equal(tr0, tr1)
    np.testing.assert_allclose(rad1, rad0, rtol=2e-6, atol=0)
    assert np.isfinite(rad1).all() and 0 <= tr1.min() <= tr1.max() <= 1


def test_error_behaviour_is_loud_and_specific(engine):
    """Nothing is silently clamped or degraded: call-order violations, bad arguments, grids too large for exact FP32
    offsets and non-finite line data each come back as their own error code with a message."""
    from pyrad_b200 import _lib
    w = small_cell(n_lines=500)
    L = w["lines"]

    def code_of(fn):
        with pytest.raises(_lib.EngineError) as ei:
            fn()
        assert str(ei.value)
        return ei.value.code

    engine.upload_lines(L, 1)
    assert code_of(lambda: engine.layer_prepass(296, 1013.0, [4e-4], [44.0], [286.0], [286.0], 500)) == -3   # grid not set
    engine.set_grid(600.0, 0.01, 10000)
    assert code_of(engine.line_sum) == -3                                                  # no prepass yet
    assert code_of(lambda: engine.set_grid(600.0, -0.01, 10000)) == -2
    assert code_of(lambda: engine.set_grid(600.0, 0.01, 10000, 5000, 4000)) == -2
    assert code_of(lambda: engine.layer_prepass(296, 1013.0, [4e-4], [44.0], [286.0], [286.0], 0)) == -2   # window < 1 sample
    assert code_of(lambda: engine.layer_prepass(-5.0, 1013.0, [4e-4], [44.0], [286.0], [286.0], 500)) == -2
    unsorted = {k: v[::-1].copy() for k, v in L.items()}
    assert code_of(lambda: engine.upload_lines(unsorted, 1)) == -2
    grp = dict(L); grp["group"] = np.full(len(L["nu"]), 3, dtype=np.int32)
    assert code_of(lambda: engine.upload_lines(grp, 2)) == -2                              # group id out of range
    # a chunk of 2^24 points cannot keep integer offsets exact in FP32: the caller is told to shard
    engine.upload_lines(L, 1)
    engine.set_grid(600.0, 0.00001, 1 << 24)
    assert code_of(lambda: engine.layer_prepass(296, 1013.0, [4e-4], [44.0], [286.0], [286.0], 500)) == -4
    # non-finite line data surfaces when the sum is read, not as NaNs in the spectrum
    bad = {k: v.copy() for k, v in L.items()}
    bad["gamma_air"][7] = np.nan
    engine.upload_lines(bad, 1)
    engine.set_grid(600.0, 0.01, 10000)
    engine.layer_prepass(296, 1013.0, [4e-4], [44.0], [286.0], [286.0], 500)
    assert code_of(engine.line_sum) == -4
    # and the engine is still usable afterwards
    engine.upload_lines(L, 1)
    engine.set_grid(600.0, 0.01, 10000)
    engine.layer_prepass(296, 1013.0, [4e-4], [44.0], [286.0], [286.0], 500)
    assert np.isfinite(engine.line_sum()).all()


# ---------------------------------------------------------------- line-range parts (PRB_OPT_SPLIT_TILES)
@pytest.mark.parametrize("variant", [eng.K2_CLASSED, eng.K2_FARFIELD])
def test_split_tiles_regroups_the_same_sum(engine, variant):
    """Short launches with PRB_OPT_SPLIT_TILES: every tile's line range is summed in parts by different CTAs and the FP64
    partials are added in part order by whichever CTA finishes last.  Same sum, regrouped: equal to the unsplit result to
    FP64 rounding (far below the FP32 evaluation noise), identical from run to run, and through the fused epilogue."""
    w = workloads.gas_cell(["h2o", "co2", "ch4", "o3"], 40000, 1000.0, 1040.0, 0.001, 296, 1013.25,
                           [0.01, 400e-6, 1.8e-6, 5e-8], 10.0, 47)
    n = H.engine_setup(engine, w)
    wts = [eng.number_density_weight(c, w["P"], w["T"]) for c in w["conc"]]
    sp = w["species"]
    win = eng.window_len(w["cutoff"], w["res"])
    col = ([w["depth_cm"]], [w["T"]], [w["P"]], [w["conc"]], [s.molmass for s in sp], [[s.q(w["T"]) for s in sp]],
           [s.q296 for s in sp], [win], 288.0, w["range_max"])
    engine.set_k2_variant(variant, 0)
    try:
        H.engine_prepass(engine, w, weights=wts)
        plain = engine.line_sum()
        engine.atmosphere(*col)
        rad0, tr0 = engine.atmosphere_read()
        engine.set_option(eng.OPT_SPLIT_TILES, 1)
        H.engine_prepass(engine, w, weights=wts)
        split = engine.line_sum()
        again = engine.line_sum()
        engine.atmosphere(*col)
        rad1, tr1 = engine.atmosphere_read()
    finally:
        engine.set_option(eng.OPT_SPLIT_TILES, 0)
        engine.set_k2_variant(eng.K2_CLASSED, 0)
    assert np.array_equal(split, again)                              # deterministic: fixed part order
    assert not np.array_equal(split, plain)                          # (the tiles really were split: 20 tiles on 296 CTA slots)
    assert H.k_rel_err(split, plain).max() <= 1e-13
    assert np.abs(tr1 - tr0).max() <= 2e-7
    np.testing.assert_allclose(rad1[1:], rad0[1:], rtol=2e-6, atol=0)
    pts = H.boundary_points(n, 200, 3, n_tiles=6)
    ref = H.oracle_layer_k_at(w, pts, w["T"], w["P"], w["conc"], w["cutoff"])
    assert H.k_rel_err(split[pts], ref).max() <= H.K_REL_TOL


# ---------------------------------------------------------------- per-group rows in one pass
@pytest.mark.parametrize("P,T", [(1013.25, 296), (60.0, 230), (3.0, 250)])
def test_line_sum_groups_equals_one_run_per_group(engine, P, T):
    """prb_upload_line_groups + ONE prepass + ONE line-sum launch give every isotopologue's cross-section row
    (pyradClasses.py:498-503, 566-576: the per-isotope / per-molecule spectra the menus plot); each row is bitwise what a
    separate upload + prepass + line sum of that group alone gives -- in the wide, table-driven and binary-search kernel
    classes -- and the device-resident k / T / transmission built from the rows match the FP64 host-buffer path."""
    w = workloads.gas_cell(["h2o", "co2", "ch4", "o3"], 6000, 1000.0, 1030.0, 0.001, T, P, [0.01, 400e-6, 1.8e-6, 5e-8], 10.0, 41)
    sp = w["species"]
    n = eng.grid_len(w["range_min"], w["range_max"], w["res"])
    win = eng.window_len(w["cutoff"], w["res"])
    engine.upload_line_groups(w["per_group_lines"])
    engine.set_grid(w["range_min"], w["res"], n)
    engine.layer_prepass(T, P, w["conc"], [s.molmass for s in sp], [s.q(T) for s in sp], [s.q296 for s in sp], win)
    pairs = engine.pair_count()
    rows = engine.line_sum_groups()
    assert rows.shape == (4, n)
    with pytest.raises(_lib.EngineError):
        engine.line_sum()                                             # the summed row needs the merged list
    wts = [eng.number_density_weight(c, P, T) for c in w["conc"]]
    axis = eng.linspace_axis(w["range_min"], w["range_max"], n)
    surf = ph.planck_wavenumber(ph.x_axis(w["range_min"], w["range_max"], w["res"]), 288)
    k, t, r = engine.layer_spectra_resident(wts, w["depth_cm"], T, w["range_max"], radiance_in=surf)
    k2, t2, r2 = engine.layer_stream(rows, wts, w["depth_cm"], T, axis, surf)
    np.testing.assert_allclose(k, k2, rtol=1e-15, atol=0)
    np.testing.assert_allclose(t, t2, rtol=1e-14, atol=0)
    np.testing.assert_allclose(r[1:], r2[1:], rtol=1e-14, atol=0)
    total = 0
    for g, s in enumerate(sp):
        engine.upload_lines(w["per_group_lines"][g], 1)
        engine.set_grid(w["range_min"], w["res"], n)
        engine.layer_prepass(T, P, [w["conc"][g]], [s.molmass], [s.q(T)], [s.q296], win)
        total += engine.pair_count()
        assert np.array_equal(engine.line_sum(), rows[g]), g
    assert total == pairs
    sig = H.oracle_sigma_groups(w)
    for g in range(4):
        assert H.k_rel_err(rows[g], sig[g]).max() <= H.K_REL_TOL


# ---------------------------------------------------------------- opt-in far-field variant (PRB_K2_FARFIELD)
def _farfield_cell(P, T):
    return workloads.gas_cell(["h2o", "co2", "ch4", "o3"], 6000, 1000.0, 1040.0, 0.001, T, P,
                              [0.01, 400e-6, 1.8e-6, 5e-8], 10.0, 29)


@pytest.mark.parametrize("P,T", [(1013.25, 296), (353.4, 250), (220.0, 225), (150.0, 220), (100.0, 215)])
def test_k2_farfield_variant_matches_exact_paths_and_oracle(engine, P, T):
    """Lorentz wings of far lines summed at 16 Chebyshev nodes per 128/256-point span and interpolated: within 1e-6 of the
    exact per-point kernel (two FP32 evaluations of the same sum; the interpolation itself is bounded at 2e-8 by the FP64
    model of the algorithm, tests/test_farfield_model.py) and within the north_star tolerance of the oracle."""
    w = _farfield_cell(P, T)
    H.engine_setup(engine, w)
    wts = [eng.number_density_weight(c, P, T) for c in w["conc"]]
    H.engine_prepass(engine, w, weights=wts)
    exact = engine.line_sum()
    engine.set_k2_variant(eng.K2_FARFIELD, 0)
    try:
        H.engine_prepass(engine, w, weights=wts)
        far = engine.line_sum()
        again = engine.line_sum()
    finally:
        engine.set_k2_variant(eng.K2_CLASSED, 0)
    assert np.array_equal(far, again)                               # deterministic
    d = H.k_rel_err(far, exact)
    assert 0 < d.max() <= 1e-6, d.max()                             # two FP32 evaluations of the same spectrum
    sig = H.oracle_sigma_groups(w)
    ref = sum(ph.abs_coef(sig[g], w["conc"][g], P, T) for g in range(4))
    err = H.k_rel_err(far, ref)
    assert err.max() <= H.K_REL_TOL, err.max()


def test_k2_farfield_level2_on_a_stress_sweep_window(engine):
    """cfg5's 25 cm-1 cutoff (W = 25 000 >= four tile lengths) switches level 2 on: lines far from a whole 2048-point tile
    are summed once per tile by the eight warps together.  Same bounds as level 1, deterministic, and tile-aligned shards
    reproduce the unsharded result bit for bit."""
    w = workloads.gas_cell(["h2o", "co2", "ch4", "o3"], 30000, 1000.0, 1020.0, 0.001, 296, 1013.25,
                           [0.01, 400e-6, 1.8e-6, 5e-8], 10.0, 33, cutoff=25.0)
    n = H.engine_setup(engine, w)
    wts = [eng.number_density_weight(c, w["P"], w["T"]) for c in w["conc"]]
    H.engine_prepass(engine, w, weights=wts)
    exact = engine.line_sum()
    engine.set_k2_variant(eng.K2_FARFIELD, 0)
    try:
        H.engine_prepass(engine, w, weights=wts)
        far = engine.line_sum()
        assert np.array_equal(far, engine.line_sum())
        parts = []
        for a, b in ((0, 4096), (4096, 14336), (14336, n)):
            H.engine_setup(engine, w, a, b)
            H.engine_prepass(engine, w, weights=wts)
            parts.append(engine.line_sum())
        assert np.array_equal(np.concatenate(parts), far)
    finally:
        engine.set_k2_variant(eng.K2_CLASSED, 0)
    d = H.k_rel_err(far, exact)
    assert 0 < d.max() <= 1e-6, d.max()
    pts = H.boundary_points(n, 300, 7, n_tiles=6)
    ref = H.oracle_layer_k_at(w, pts, w["T"], w["P"], w["conc"], w["cutoff"])
    assert H.k_rel_err(far[pts], ref).max() <= H.K_REL_TOL


def test_k2_farfield_variant_is_shard_invariant_and_feeds_the_fused_paths(engine):
    """Node positions are tile-relative, so tile-aligned shards reproduce the unsharded far-field result bit for bit;
    the single-layer fused epilogue (atmosphere call) and the FP32 k row see the same sums."""
    w = _farfield_cell(1013.25, 296)
    sp = w["species"]
    n = H.engine_setup(engine, w)
    win = eng.window_len(w["cutoff"], w["res"])
    mol, q296, qt = [s.molmass for s in sp], [s.q296 for s in sp], [s.q(w["T"]) for s in sp]
    args = ([w["depth_cm"]], [w["T"]], [w["P"]], [w["conc"]], mol, [qt], q296, [win], 288.0, w["range_max"])
    engine.atmosphere(*args)
    _, tr_exact = engine.atmosphere_read()
    engine.set_k2_variant(eng.K2_FARFIELD, 0)
    try:
        H.engine_prepass(engine, w)
        full = engine.line_sum()
        parts = []
        for a, b in ((0, 8192), (8192, 28672), (28672, n)):
            H.engine_setup(engine, w, a, b)
            H.engine_prepass(engine, w)
            parts.append(engine.line_sum())
        assert np.array_equal(np.concatenate(parts), full)
        H.engine_setup(engine, w)
        engine.atmosphere(*args)
        _, tr_far = engine.atmosphere_read()
    finally:
        engine.set_k2_variant(eng.K2_CLASSED, 0)
    assert np.abs(tr_far - tr_exact).max() <= 1e-6


def test_measured_peaks_are_plausible_roofline_denominators(engine):
    """prb_measure_peaks (the denominators of bench.py's rooflines): the packed-FFMA2 and scalar-FFMA streams both reach the
    FP32 pipe's rate (148 SMs x 128 lanes x clock: 98 % measured; the bound leaves room for a box that runs below its maximum
    clock), the float4 copy a few TB/s."""
    info = engine.device_info()
    p = engine.measure_peaks()
    nominal = info["sm_count"] * 128 * info["sm_clock_khz"] * 1e3
    assert 0.5 * nominal <= p["ffma2_lane_fma_per_s"] <= 1.05 * nominal, (p, nominal)
    assert 0.5 * nominal <= p["ffma_lane_fma_per_s"] <= 1.05 * nominal, (p, nominal)
    assert abs(p["ffma2_lane_fma_per_s"] / p["ffma_lane_fma_per_s"] - 1) < 0.1       # two forms of the same pipe
    assert 2e12 <= p["copy_bytes_per_s"] <= 9e12, p


def test_result_pool_hands_out_and_takes_back_page_locked_arrays(engine):
    """engine.ResultPool (the mirror's result arrays): a dropped array's block is reused by the next request of that size,
    arrays alive at the same time never share memory, views keep a block leased, small results stay ordinary arrays."""
    import gc
    pool = eng.ResultPool(max_free=2)
    a = pool.array(1 << 18)                                         # 2 MB of doubles
    assert a.shape == (1 << 18,) and a.dtype == np.float64 and not a.flags.owndata and a.flags.writeable
    addr = a.ctypes.data
    a[:] = 1.0
    view = a[10:20]
    del a
    gc.collect()
    b = pool.array(1 << 18)
    assert b.ctypes.data != addr and view[0] == 1.0                 # the view still leases the first block
    del view
    gc.collect()
    c = pool.array(1 << 18)
    assert c.ctypes.data == addr                                    # handed back, reused
    m = pool.array((4, 1 << 16))
    assert m.shape == (4, 1 << 16) and m.ctypes.data not in (b.ctypes.data, c.ctypes.data)
    assert pool.array(10).flags.owndata                             # small: plain numpy
    del b, c, m
    gc.collect()
    assert sum(len(v) for v in pool.free.values()) == 2             # max_free bounds what is kept
    # the engine fills pooled arrays like any other host buffer
    w = small_cell()
    H.engine_setup(engine, w)
    H.engine_prepass(engine, w)
    ref = engine.line_sum()
    engine.result_pool = pool
    try:
        axis = eng.linspace_axis(w["range_min"], w["range_max"], len(ref))
        k, t, _ = engine.layer_stream(ref[None, :], [1.0], 10.0, 296, axis)
    finally:
        engine.result_pool = None
    assert np.isfinite(k).all() and np.isfinite(t).all()


def test_layer_line_ranges_keep_a_column_on_the_references_per_layer_line_sets(engine):
    """ONE line list for a column whose layers have different cutoffs (prb_set_layer_line_range): the reference loads
    each layer's lines from that layer's own effective range (pyradUtilities.py:437-438), and a layer whose cutoff is
    shorter than a grid step (W = 1: centre samples only) must not pick up the line that sits within one grid step below
    rangeMin -- int() truncation puts it on index 0 (pyradClasses.py:390) -- although a wider layer needs it."""
    sp = workloads.synth.species("co2")
    rmin, rmax, res = 600.0, 610.0, 0.01
    n = eng.grid_len(rmin, rmax, res)
    ln = workloads.synth.make_lines(40, 598.0, 612.0, 5)
    ln["nu"] = np.sort(np.concatenate([ln["nu"][:-1], [rmin - 0.004]]))         # 0.4 grid steps below rangeMin
    P = [800.0, 0.5]                                                            # cutoffs 3.95 and 0.0025 cm-1: W = 395 and 1
    T = [280, 230]
    win = [eng.window_len(ph.layer_cutoff(p), res) for p in P]
    assert win[1] == 1
    conc = [[400e-6], [400e-6]]
    depth = [1e4, 1e6]
    engine.upload_lines(ln, 1)
    engine.set_grid(rmin, res, n)
    args = (depth, T, P, conc, [sp.molmass], [[sp.q(t)] for t in T], [sp.q296], win, 288.0, rmax)
    engine.atmosphere(*args)
    _, tr_all = engine.atmosphere_read()
    lo = [max(rmin - ph.layer_cutoff(p), 0) for p in P]
    hi = [rmax + ph.layer_cutoff(p) for p in P]
    engine.set_layer_line_range(lo, hi)
    try:
        engine.atmosphere(*args)
        _, tr = engine.atmosphere_read()
        with pytest.raises(_lib.EngineError):
            engine.atmosphere(depth[:1], T[:1], P[:1], conc[:1], [sp.molmass], [[sp.q(T[0])]], [sp.q296], win[:1], 288.0, rmax)
    finally:
        engine.set_layer_line_range()
    want = np.ones(n)
    for l in range(2):
        m = (ln["nu"] > lo[l]) & (ln["nu"] < hi[l])
        sub = {k: v[m] for k, v in ln.items()}
        sig = ph.cross_section(sub, T[l], P[l], conc[l][0], sp.molmass, sp.q(T[l]), sp.q296, rmin, rmax, res, ph.layer_cutoff(P[l]))
        want = want * ph.transmittance(ph.abs_coef(sig, conc[l][0], P[l], T[l]), depth[l])
    assert np.abs(tr - want).max() <= H.T_ABS_TOL
    assert np.abs(tr_all[0] - want[0]) > 1e-4 and np.abs(tr_all[1:] - want[1:]).max() <= H.T_ABS_TOL   # (the line matters)
    engine.atmosphere(*args)                                                    # cleared: every line takes part again
    assert np.array_equal(engine.atmosphere_read()[1], tr_all)
