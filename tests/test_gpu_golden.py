"""GPU tests: the drop-in host mirror (pyrad_b200.classes, same API as the reference's pyradClasses) driven
through the C ABI on the same on-disk data tree the reference read, compared with the golden vectors the
REAL reference produced (tests/golden/make_golden.py).  Reads like the reference's own usage (main.py:35-46)."""
import numpy as np
import pytest

from oracle import physics as ph
from oracle import ref_harness as rh          # only its data-tree writers (test infrastructure)
from pyrad_b200 import classes as C
from pyrad_b200 import synth
from tests import golden_util as G
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture()
def data_root(tmp_path, engine):
    C.set_engine(engine)
    C.DATA_ROOT = str(tmp_path)
    C.Layer.hasAtmosphere = False
    old = C.BASE_RESOLUTION
    yield str(tmp_path)
    C.BASE_RESOLUTION = old
    C.DATA_ROOT = None


def seed(root, g, group, name):
    sp = synth.species(name)
    ln = G.lines_of(g, group)
    rh.write_params(root, sp.global_iso, sp.name, sp.mol_id, 1, 0.99, sp.q296, 1, sp.molmass)
    rh.write_q_table(root, sp.global_iso, range(100, 501), [sp.q(t) for t in range(100, 501)])
    rh.write_line_segments(root, sp.global_iso, sp.mol_id, 1, ln, int(ln["nu"].min() / 100) * 100, ln["nu"].max() + 101)


@pytest.mark.parametrize("name", G.CELL_CASES)
def test_mirror_matches_reference_gas_cell(name, data_root):
    g = G.load(name)
    species = [str(s) for s in g["species"]]
    for i, s in enumerate(species):
        seed(data_root, g, i, s)
    C.BASE_RESOLUTION = float(g["base"])
    T = int(g["T"])
    layer = C.Layer(float(g["depth"]), T, float(g["P"]), float(g["range_min"]), float(g["range_max"]),
                    dynamicResolution=bool(g["dynamic"]))
    mols = [layer.addMolecule(s, concentration=float(c)) for s, c in zip(species, g["conc"])]
    assert layer.resolution == float(g["res"]) and layer.distanceFromCenter == float(g["cutoff"])
    np.testing.assert_array_equal(layer.xAxis, g["xaxis"])
    for i, m in enumerate(mols):
        np.testing.assert_array_equal([l.wavenumber for l in m[0]], g["kept_nu_%d" % i])
        sig = C.getCrossSection(m[0])
        assert H.k_rel_err(sig, g["sigma_%d" % i]).max() <= H.K_REL_TOL
        assert H.k_rel_err(C.getAbsCoef(m), g["abscoef_%d" % i]).max() <= H.K_REL_TOL
    assert H.k_rel_err(C.getAbsCoef(layer), g["layer_abscoef"]).max() <= H.K_REL_TOL
    assert np.abs(C.getTransmittance(layer) - g["layer_transmittance"]).max() <= H.T_ABS_TOL
    np.testing.assert_allclose(layer.planck(T), g["layer_planck"], rtol=1e-12)
    out = layer.transmission(g["surface"])
    np.testing.assert_allclose(out, g["layer_transmission"], rtol=2e-5)


@pytest.mark.parametrize("name", G.CELL_CASES)
def test_mirror_survey_derived_and_integral_match_reference(name, data_root, engine):
    """SURVEY section 8(f) rows through the mirror (device kernels k4_*): the line survey is bit-exact; derived
    spectra follow the transmittance's own tolerance; the integral is a pairwise device reduction."""
    g = G.load(name)
    species = [str(s) for s in g["species"]]
    for i, s in enumerate(species):
        seed(data_root, g, i, s)
    C.BASE_RESOLUTION = float(g["base"])
    layer = C.Layer(float(g["depth"]), int(g["T"]), float(g["P"]), float(g["range_min"]), float(g["range_max"]),
                    dynamicResolution=bool(g["dynamic"]))
    mols = [layer.addMolecule(s, concentration=float(c)) for s, c in zip(species, g["conc"])]
    for i, m in enumerate(mols):
        np.testing.assert_array_equal(m[0].createLineSurvey(), g["survey_%d" % i])
    np.testing.assert_array_equal(layer.lineSurvey, g["layer_survey"])
    with np.errstate(all="ignore"):
        tau, tau_ref = C.getOpticalDepth(layer), g["layer_optical_depth"]
        ab, ab_ref = C.getAbsorbance(layer), g["layer_absorbance"]
    fin = np.isfinite(tau_ref) & (g["layer_transmittance"] > 1e-300)
    np.testing.assert_allclose(tau[fin], tau_ref[fin], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(ab[fin], ab_ref[fin], rtol=2e-5, atol=1e-6)
    assert np.abs(C.getEmissivity(layer) - g["layer_emissivity"]).max() <= H.T_ABS_TOL
    # the kernels themselves on the reference's own arrays: numpy-level agreement
    d = engine.derived_spectra(g["layer_transmittance"])
    np.testing.assert_array_equal(d["emissivity"], g["layer_emissivity"])
    np.testing.assert_allclose(d["optical_depth"][fin], tau_ref[fin], rtol=1e-14, atol=1e-300)
    np.testing.assert_allclose(d["absorbance"][fin], ab_ref[fin], rtol=1e-14, atol=1e-300)
    for key, ref in (("layer_transmission", "integrated_transmission"), ("surface", "integrated_surface")):
        v = C.integrateSpectrum(g[key], res=float(g["base"]))
        assert v == pytest.approx(float(g[ref]), rel=1e-13)
    assert C.integrateSpectrum(np.array([1.0, np.nan, 2.0, 1e300]), res=1.0) == pytest.approx((3.0 + 1e300) * np.pi, rel=1e-15)


@pytest.mark.parametrize("name", G.XSC_CASES)
def test_mirror_matches_reference_xsc(name, data_root):
    g = G.load(name)
    seed(data_root, g, 0, "co2")
    fname = rh.write_xsc_file(data_root, "CFC11", 296.0, 760.0, float(g["file_rmin"]), float(g["file_rmax"]),
                              float(g["file_res"]), g["file_x"], g["file_y"])
    layer = C.Layer(100.0, 250, 500.0, float(g["layer_rmin"]), float(g["layer_rmax"]))
    xm = layer.addMolecule({"CFC11": fname}, concentration=250e-12)
    assert layer.T == int(g["T_after"]) and layer.P == float(g["P_after"])     # forced to the file's T, P
    m = layer.addMolecule("co2", concentration=400e-6)
    np.testing.assert_allclose(C.getCrossSection(xm), g["xsc_sigma"], rtol=1e-14, atol=0)
    assert H.k_rel_err(C.getCrossSection(m[0]), g["co2_sigma"]).max() <= H.K_REL_TOL
    assert H.k_rel_err(C.getAbsCoef(layer), g["layer_abscoef"]).max() <= H.K_REL_TOL
    assert np.abs(C.getTransmittance(layer) - g["layer_transmittance"]).max() <= H.T_ABS_TOL


def test_mirror_matches_reference_cfg3_miniature(data_root):
    """cfg3 in miniature: CO2 + H2O line by line next to two xsc tables (one native 0.01 cm-1, one 0.05 cm-1 re-gridded
    on the device), every spectrum against what the real reference produced on the same files."""
    g = G.load("cfg3_mini")
    names = [str(s) for s in g["species"]]
    for i, s in enumerate(names):
        seed(data_root, g, i, s)
    fnames = [rh.write_xsc_file(data_root, str(g["xsc_names"][i]), 296.0, 760.0, float(g["xsc_rmin"][i]),
                                float(g["xsc_rmax"][i]), float(g["xsc_res"][i]), g["xsc_x_%d" % i], g["xsc_y_%d" % i])
              for i in range(2)]
    layer = C.Layer(float(g["depth"]), 280, 900.0, float(g["range_min"]), float(g["range_max"]))
    xms = [layer.addMolecule({str(g["xsc_names"][i]): fnames[i]}, concentration=float(g["xsc_conc"][i])) for i in range(2)]
    mols = [layer.addMolecule(s, concentration=float(c)) for s, c in zip(names, g["conc"])]
    assert layer.T == int(g["T_after"]) and layer.P == float(g["P_after"])
    np.testing.assert_array_equal(layer.xAxis, g["xaxis"])
    for i, xm in enumerate(xms):
        np.testing.assert_allclose(C.getCrossSection(xm), g["xsc_sigma_%d" % i], rtol=1e-14, atol=0)
    for i, m in enumerate(mols):
        assert H.k_rel_err(C.getCrossSection(m[0]), g["sigma_%d" % i]).max() <= H.K_REL_TOL
    assert H.k_rel_err(C.getAbsCoef(layer), g["layer_abscoef"]).max() <= H.K_REL_TOL
    assert np.abs(C.getTransmittance(layer) - g["layer_transmittance"]).max() <= H.T_ABS_TOL
    surf = layer.planck(int(g["surface_T"]))
    np.testing.assert_allclose(layer.transmission(surf), g["layer_transmission"], rtol=2e-5)


def test_cfg3_miniature_through_one_engine_call(data_root):
    """The same cell as ONE engine call (Atmosphere.columnSpectrum -> prb_atmosphere): both xsc tables resident on the
    device (prb_xsc_resident), resampled onto the grid there, and added to the line sum inside K2's epilogue -- against
    the real reference's transmittance and Layer.transmission on the same files (pyradClasses.py:466-505, 707-716, 784-787).
    (At the golden's ppt mole fractions the tables are a 1e-7 effect; the column test below and the full-size cfg3 test
    run them abundant enough to fail on a missing table.)"""
    g = G.load("cfg3_mini")
    names = [str(s) for s in g["species"]]
    for i, s in enumerate(names):
        seed(data_root, g, i, s)
    fnames = [rh.write_xsc_file(data_root, str(g["xsc_names"][i]), 296.0, 760.0, float(g["xsc_rmin"][i]),
                                float(g["xsc_rmax"][i]), float(g["xsc_res"][i]), g["xsc_x_%d" % i], g["xsc_y_%d" % i])
              for i in range(2)]
    atm = C.Atmosphere("cfg3")
    layer = atm.addLayer(float(g["depth"]), 280, 900.0, float(g["range_min"]), float(g["range_max"]), dynamicResolution=False)
    for i in range(2):
        layer.addMolecule({str(g["xsc_names"][i]): fnames[i]}, concentration=float(g["xsc_conc"][i]))
    for s, c in zip(names, g["conc"]):
        layer.addMolecule(s, concentration=float(c))
    assert layer.T == int(g["T_after"]) and layer.P == float(g["P_after"])
    e = C.engine()
    rad, trans = atm.columnSpectrum(int(g["surface_T"]))
    assert e.atmosphere_launches() == 4                            # two table resamplings, K1, K2 (fused epilogue): no K3
    assert np.abs(trans - g["layer_transmittance"]).max() <= H.T_ABS_TOL
    np.testing.assert_allclose(rad, g["layer_transmission"], rtol=2e-5)


def test_mirror_error_behaviour(data_root):
    g = G.load("cell_co2_1atm")
    seed(data_root, g, 0, "co2")
    layer = C.Layer(10.0, 296.5, 1013.0, 600.0, 700.0)            # non-integer T: Q lookup fails like the reference
    m = layer.addMolecule("co2", ppm=400)
    with pytest.raises(KeyError):
        C.getCrossSection(m[0])
    layer.changeTemperature(296)
    assert H.k_rel_err(C.getCrossSection(m[0]), g["sigma_0"]).max() <= H.K_REL_TOL
    layer.changePressure(500.0)                                    # invalidates and recomputes (lazy cache)
    assert not m[0].progressCrossSection


def test_mirror_column_spectrum_equals_layer_by_layer_transmission(data_root):
    """Atmosphere.columnSpectrum (one prb_atmosphere call: batched K1/K2 + FP32 fold on the device) against the fold of
    Layer.transmission through the host-buffer FP64 path, on the same data tree."""
    g = G.load("cell_fine_grid")
    species = [str(s) for s in g["species"]]
    for i, s in enumerate(species):
        seed(data_root, g, i, s)
    C.BASE_RESOLUTION = float(g["base"])
    atm = C.Atmosphere("column")
    rmin, rmax = float(g["range_min"]), float(g["range_max"])
    for depth, T, P in ((2e4, 280, 800.0), (3e4, 250, 300.0), (5e4, 220, 40.0), (8e4, 230, 3.0)):
        layer = atm.addLayer(depth, T, P, rmin, rmax, dynamicResolution=False)
        for s, c in zip(species, g["conc"]):
            layer.addMolecule(s, concentration=float(c))
    rad, trans = atm.columnSpectrum(288)
    surface = atm[0].planck(288)
    ref = atm.transmission(surface)
    t_ref = np.ones_like(ref)
    for layer in atm:
        t_ref = t_ref * C.getTransmittance(layer)
    assert np.abs(trans - t_ref).max() <= H.T_ABS_TOL
    np.testing.assert_allclose(rad, ref, rtol=3e-5)


def test_mirror_column_spectrum_carries_xsc_molecules(data_root):
    """A column whose layers carry xsc molecules -- one table in every layer at different mole fractions, a second one
    in a single layer -- through ONE engine call, against the layer-by-layer fold of Layer.transmission (FP64 host path,
    which the goldens pin against the real reference)."""
    g = G.load("cfg3_mini")
    names = [str(s) for s in g["species"]]
    for i, s in enumerate(names):
        seed(data_root, g, i, s)
    fnames = [rh.write_xsc_file(data_root, str(g["xsc_names"][i]), 296.0, 760.0, float(g["xsc_rmin"][i]),
                                float(g["xsc_rmax"][i]), float(g["xsc_res"][i]), g["xsc_x_%d" % i], g["xsc_y_%d" % i])
              for i in range(2)]
    atm = C.Atmosphere("column with xsc")
    rmin, rmax = float(g["range_min"]), float(g["range_max"])
    for k, depth in enumerate((30.0, 50.0, 80.0)):
        layer = atm.addLayer(depth, 280, 900.0, rmin, rmax, dynamicResolution=False)
        layer.addMolecule({str(g["xsc_names"][0]): fnames[0]}, concentration=float(g["xsc_conc"][0]) * (k + 1) * 2e5)
        if k == 1:
            layer.addMolecule({str(g["xsc_names"][1]): fnames[1]}, concentration=float(g["xsc_conc"][1]) * 1e6)
        for s, c in zip(names, g["conc"]):
            layer.addMolecule(s, concentration=float(c))
    rad, trans = atm.columnSpectrum(288)
    ref = atm.transmission(atm[0].planck(288))
    t_ref = np.ones_like(ref)
    for layer in atm:
        t_ref = t_ref * C.getTransmittance(layer)
    assert np.abs(trans - t_ref).max() <= H.T_ABS_TOL
    np.testing.assert_allclose(rad, ref, rtol=3e-5)
    # and the engine is clean afterwards: the same column without its xsc molecules differs
    for layer in atm:
        layer[:] = [m for m in layer if not m.exotic]
    _, trans2 = atm.columnSpectrum(288)
    assert np.abs(trans2 - t_ref).max() > 1e-4


# ---------------------------------------------------------------- xsc file utilities
def _seed_folder(root, folder, files):
    import os
    d = os.path.join(root, "data", "xsc", folder)
    os.makedirs(d)
    for name, blob in files.items():
        with open(os.path.join(d, name), "wb") as f:
            f.write(blob)
    return d


def _folder(d):
    import os
    return {n: open(os.path.join(d, n), "rb").read() for n in sorted(os.listdir(d))}


def test_change_res_xsc_file_is_byte_identical_to_the_reference(data_root):
    from pyrad_b200 import xsc_files as xf
    g = G.load("xsc_files")
    d = _seed_folder(data_root, "CFC11", G.files_of(g, "res_in"))
    written = xf.changeResFolder("CFC11")
    want = G.files_of(g, "res_out")
    got = _folder(d)
    assert sorted(written) == sorted(want) == sorted(got)
    for name in want:
        assert got[name] == want[name], name
    # the rewritten table is a usable xsc molecule of the mirror (same path as test_mirror_matches_reference_xsc)
    layer = C.Layer(100.0, 250, 500.0, 800.0, 900.0)
    xm = layer.addMolecule({"CFC11": "CFC11_296.0K-760.0Torr_830.0-860.0_0.01_air_00_00.txt"}, concentration=250e-12)
    assert layer.T == 296 and np.count_nonzero(C.getCrossSection(xm)) > 2900


def test_merge_xsc_is_byte_identical_to_the_reference(data_root):
    from pyrad_b200 import xsc_files as xf
    g = G.load("xsc_files")
    d = _seed_folder(data_root, "HCFC22", G.files_of(g, "merge_in"))
    written = xf.mergeXsc("HCFC22")
    want = G.files_of(g, "merge_out")
    got = _folder(d)
    assert sorted(written) == sorted(want) == sorted(got)
    for name in want:
        assert got[name] == want[name], name


def test_mirror_resident_rows_follow_the_layer_state(data_root):
    """The engine holds ONE layer's per-isotopologue rows at a time; the mirror re-uploads whenever another layer, another
    line list, another mole fraction or another T / P is asked for -- and only then."""
    g = G.load("cell_fine_grid")
    species = [str(s) for s in g["species"]]
    for i, s in enumerate(species):
        seed(data_root, g, i, s)
    C.BASE_RESOLUTION = float(g["base"])
    rmin, rmax = float(g["range_min"]), float(g["range_max"])
    a = C.Layer(float(g["depth"]), int(g["T"]), float(g["P"]), rmin, rmax, dynamicResolution=False)
    b = C.Layer(3.0 * float(g["depth"]), 250, 300.0, rmin, rmax, dynamicResolution=False)
    for layer in (a, b):
        for s, c in zip(species, g["conc"]):
            layer.addMolecule(s, concentration=float(c))
    ta = C.getTransmittance(a)
    assert np.abs(ta - g["layer_transmittance"]).max() <= H.T_ABS_TOL
    key_a = C._RESIDENT_KEY
    tb = C.getTransmittance(b)
    assert C._RESIDENT_KEY != key_a and not np.array_equal(ta, tb)
    assert np.array_equal(C.getTransmittance(a), ta)                      # a again: uploaded again, same spectrum
    assert np.array_equal(a.transmittance, ta) and C._RESIDENT_KEY == key_a  # ... and now straight from the resident rows
    # per-isotopologue rows come from the same launch
    for i, m in enumerate(a):
        assert H.k_rel_err(C.getCrossSection(m[0]), g["sigma_%d" % i]).max() <= H.K_REL_TOL
    # a new mole fraction invalidates; the result is the one of a fresh layer built with it
    a[0].setPPM(a[0].concentration * 2e6)
    t2 = C.getTransmittance(a)
    fresh = C.Layer(float(g["depth"]), int(g["T"]), float(g["P"]), rmin, rmax, dynamicResolution=False)
    for k, (s, c) in enumerate(zip(species, g["conc"])):
        fresh.addMolecule(s, concentration=float(c) * (2 if k == 0 else 1))
    assert np.array_equal(C.getTransmittance(fresh), t2) and not np.array_equal(t2, ta)
    # a new line list of the same length invalidates too (tokens, not ids or lengths, identify a list)
    iso = a[1][0]
    cols = {k: v.copy() for k, v in iso._cols.items()}
    cols["sw"] = cols["sw"] * 3.0
    iso.setLines(cols)
    assert not np.array_equal(C.getTransmittance(a), t2)
