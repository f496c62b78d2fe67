"""CPU tests pinning the oracle (oracle/physics.py): known answers and golden vectors produced by the REAL
reference (tests/golden/make_golden.py), plus the oracle's internal consistency."""
import numpy as np
import pytest

from oracle import physics as ph
from tests import golden_util as G


def test_known_answers_from_survey():
    """Literal values probed from the real modules during the survey."""
    assert ph.c2 == 1.4387773538277202
    assert ph.boltzmann_factors(100, 250) == pytest.approx(0.914445405949774, rel=1e-15)
    assert ph.stimulated_emissions(667.661, 250) == pytest.approx(1.0182252250288204, rel=1e-15)
    assert ph.intensity_factor(3.0e-19, 667.661, 250, 100.0, 220.0, 286.09) == pytest.approx(3.6324771066759466e-19, rel=1e-15)
    m = ph.mol_mass_kg(43.98983)
    assert m == pytest.approx(7.304683009675408e-26, rel=1e-15)
    g = ph.gaussian_hw(667.661, 250, m)
    assert g == pytest.approx(0.0006846382933486595, rel=1e-15)
    l = ph.lorentz_hw(.07, .09, 500., 250, 4e-4, .7)
    assert l == pytest.approx(0.038881879896491875, rel=1e-15)
    x = np.array([0, .01, .1, 1])
    np.testing.assert_allclose(ph.lorentz_shape(l, x),
                               [8.186586837652113, 7.67867122343346, 1.0751130263113933, 0.01235780422881055], rtol=1e-14)
    np.testing.assert_allclose(ph.pseudo_voigt_shape(g, l, x),
                               [8.184656492951845, 7.677377672916272, 1.0751163912564965, 0.012359532679778377], rtol=1e-14)
    assert ph.gaussian_shape(g, 0) == pytest.approx(824.0695693316071, rel=1e-14)
    l2 = ph.lorentz_hw(.07, .09, 10., 250, 4e-4, .7)
    assert l2 == pytest.approx(0.0007776375979298374, rel=1e-15)
    np.testing.assert_allclose(ph.pseudo_voigt_shape(g, l2, np.array([0, .001, .002, .01])),
                               [325.6183264783162, 178.3370178221835, 58.18159783494326, 2.688302266244672], rtol=1e-14)
    np.testing.assert_allclose(ph.planck_wavenumber(np.array([1., 667., 2000.]), 288),
                               [2.3781599011426146e-06, 0.13090536054881546, 0.004362946597573447], rtol=1e-14)
    assert int((599.995 - 600) / .01) == 0 and ph.line_index([599.995], 600, .01)[0] == 0
    assert int((599.985 - 600) / .01) == -1 and ph.line_index([599.985], 600, .01)[0] == -1
    assert ph.window_len(5 * 1013 / 1013.25, .01) == 500
    assert ph.window_len(5 * 1013 / 1013.25, .001) == 4999
    assert ph.window_len(25, .001) == 25000


def test_kat_file_matches_oracle():
    k = G.load("kat")
    assert ph.c2 == float(k["c2"])
    assert ph.boltzmann_factors(100, 250) == pytest.approx(float(k["boltz"]), rel=1e-15)
    assert ph.stimulated_emissions(667.661, 250) == pytest.approx(float(k["stim"]), rel=1e-15)
    g, l, l2 = float(k["gHW"]), float(k["lHW"]), float(k["lHW2"])
    np.testing.assert_allclose(ph.lorentz_shape(l, k["x"]), k["lorentz"], rtol=1e-15)
    np.testing.assert_allclose(ph.pseudo_voigt_shape(g, l, k["x"]), k["voigt"], rtol=1e-15)
    np.testing.assert_allclose(ph.gaussian_shape(g, k["x2"]), k["gauss"], rtol=1e-15)
    np.testing.assert_allclose(ph.pseudo_voigt_shape(g, l2, k["x2"]), k["voigt2"], rtol=1e-15)
    np.testing.assert_allclose(ph.planck_wavenumber(np.array([1., 667., 2000.]), 288), k["planck"], rtol=1e-15)


@pytest.mark.parametrize("name", G.CELL_CASES)
def test_oracle_reproduces_reference_gas_cell(name):
    g = G.load(name)
    T, P = float(g["T"]), float(g["P"])
    rmin, rmax, base = float(g["range_min"]), float(g["range_max"]), float(g["base"])
    res = ph.layer_resolution(P, base, bool(g["dynamic"]))
    cutoff = ph.layer_cutoff(P)
    assert res == float(g["res"]) and cutoff == float(g["cutoff"])
    np.testing.assert_array_equal(ph.x_axis(rmin, rmax, base), g["xaxis"])
    k_layer = np.zeros(ph.grid_len(rmin, rmax, base))
    for i in range(len(g["conc"])):
        ln = G.lines_of(g, i, cutoff, rmin, rmax)
        np.testing.assert_array_equal(ln["nu"], g["kept_nu_%d" % i])         # same strict range filter
        sig = ph.cross_section(ln, T, P, float(g["conc"][i]), float(g["molmass"][i]), float(g["qT"][i]),
                               float(g["q296"][i]), rmin, rmax, res, cutoff)
        if res != base:
            sig = ph.interp_to_base(sig, rmin, rmax, res, base)
        np.testing.assert_allclose(sig, g["sigma_%d" % i], rtol=1e-12, atol=0)
        k = ph.abs_coef(sig, float(g["conc"][i]), P, T)
        np.testing.assert_allclose(k, g["abscoef_%d" % i], rtol=1e-12, atol=0)
        k_layer += k
    np.testing.assert_allclose(k_layer, g["layer_abscoef"], rtol=1e-12)
    t = ph.transmittance(k_layer, float(g["depth"]))
    assert np.abs(t - g["layer_transmittance"]).max() <= 1e-13
    np.testing.assert_allclose(ph.planck_wavenumber(g["xaxis"], int(g["surface_T"])), g["surface"], rtol=1e-14)
    np.testing.assert_allclose(ph.planck_wavenumber(g["xaxis"], T), g["layer_planck"], rtol=1e-14)
    np.testing.assert_allclose(ph.transmission(t, g["surface"], g["layer_planck"]), g["layer_transmission"], rtol=1e-12)


@pytest.mark.parametrize("name", G.CELL_CASES)
def test_oracle_reproduces_reference_survey_and_derived_spectra(name):
    """line survey, optical depth / absorbance / emissivity, integrateSpectrum."""
    g = G.load(name)
    rmin, rmax, base, res = float(g["range_min"]), float(g["range_max"]), float(g["base"]), float(g["res"])
    cutoff = ph.layer_cutoff(float(g["P"]))
    n_out = ph.grid_len(rmin, rmax, base)
    total = np.zeros(n_out)
    for i in range(len(g["conc"])):
        ln = G.lines_of(g, i, cutoff, rmin, rmax)
        sv = ph.line_survey(ln["nu"], ln["sw"], rmin, res, n_out)
        np.testing.assert_array_equal(sv, g["survey_%d" % i])                 # same adds in the same order
        total += sv
    np.testing.assert_array_equal(total, g["layer_survey"])
    t = g["layer_transmittance"]
    np.testing.assert_array_equal(ph.optical_depth(t), g["layer_optical_depth"])
    np.testing.assert_array_equal(ph.absorbance(t), g["layer_absorbance"])
    np.testing.assert_array_equal(ph.emissivity(t), g["layer_emissivity"])
    assert ph.integrate_spectrum(g["layer_transmission"], res=base) == float(g["integrated_transmission"])
    assert ph.integrate_spectrum(g["surface"], res=base) == float(g["integrated_surface"])


@pytest.mark.parametrize("name", G.XSC_CASES)
def test_oracle_reproduces_reference_xsc(name):
    g = G.load(name)
    sig = ph.xsc_cross_section(g["xaxis"], g["file_x"], g["file_y"], float(g["file_rmin"]), float(g["file_rmax"]),
                               float(g["file_res"]))
    np.testing.assert_array_equal(sig, g["xsc_sigma"])
    # the xsc branch forces the layer to the file's T and P (pyradClasses.py:488-491)
    T, P = float(g["T_after"]), float(g["P_after"])
    cutoff = ph.layer_cutoff(P)
    # quirk kept by the reference: changePressure (pyradClasses.py:745-752) updates the cutoff but NOT
    # effectiveRangeMin/Max, so the kept lines still follow the layer's ORIGINAL pressure (500 hPa here)
    ln = G.lines_of(g, 0, ph.layer_cutoff(500.0), float(g["layer_rmin"]), float(g["layer_rmax"]))
    co2 = ph.cross_section(ln, T, P, float(g["conc_co2"]), float(g["molmass"]), float(g["qT"]), float(g["q296"]),
                           float(g["layer_rmin"]), float(g["layer_rmax"]), float(g["res"]), cutoff)
    np.testing.assert_allclose(co2, g["co2_sigma"], rtol=1e-12)
    k = ph.abs_coef(sig, float(g["conc_xsc"]), P, T) + ph.abs_coef(co2, float(g["conc_co2"]), P, T)
    np.testing.assert_allclose(k, g["layer_abscoef"], rtol=1e-12)
    assert np.abs(ph.transmittance(k, float(g["depth"])) - g["layer_transmittance"]).max() <= 1e-13


def test_oracle_reproduces_reference_cfg3_miniature():
    """cfg3 in miniature (two line-by-line molecules + two xsc tables on one grid), outputs of the real reference."""
    g = G.load("cfg3_mini")
    rmin, rmax = float(g["range_min"]), float(g["range_max"])
    T, P = float(g["T_after"]), float(g["P_after"])
    assert (T, P) == (296, 760.0 / 0.75006)
    k = np.zeros(len(g["xaxis"]))
    for i in range(2):
        sig = ph.xsc_cross_section(g["xaxis"], g["xsc_x_%d" % i], g["xsc_y_%d" % i], float(g["xsc_rmin"][i]),
                                   float(g["xsc_rmax"][i]), float(g["xsc_res"][i]))
        np.testing.assert_array_equal(sig, g["xsc_sigma_%d" % i])
        k = k + ph.abs_coef(sig, float(g["xsc_conc"][i]), P, T)
    for i in range(2):
        # kept lines follow the layer's ORIGINAL pressure (900 hPa), see test_oracle_reproduces_reference_xsc
        ln = G.lines_of(g, i, ph.layer_cutoff(900.0), rmin, rmax)
        sig = ph.cross_section(ln, T, P, float(g["conc"][i]), float(g["molmass"][i]), float(g["qT"][i]),
                               float(g["q296"][i]), rmin, rmax, float(g["res"]), ph.layer_cutoff(P))
        np.testing.assert_allclose(sig, g["sigma_%d" % i], rtol=1e-12)
        k = k + ph.abs_coef(sig, float(g["conc"][i]), P, T)
    np.testing.assert_allclose(k, g["layer_abscoef"], rtol=1e-12)
    t = ph.transmittance(k, float(g["depth"]))
    assert np.abs(t - g["layer_transmittance"]).max() <= 1e-13
    rad = ph.transmission(t, ph.planck_wavenumber(g["xaxis"], 288), ph.planck_wavenumber(g["xaxis"], T))
    np.testing.assert_allclose(rad, g["layer_transmission"], rtol=1e-12)


def test_scatter_scalar_and_gather_forms_agree():
    from pyrad_b200 import synth
    ln = synth.make_lines(150, 598.0, 612.0, 3)
    args = (ln, 250, 300.0, 4e-4, 43.98983, 250.0, 286.09, 600.0, 610.0, 0.01, ph.layer_cutoff(300.0))
    a = ph.cross_section(*args)
    b = ph.cross_section_scalar(*args)
    np.testing.assert_array_equal(a, b)                       # slice-add == literal loop, bit for bit
    pts = np.array([0, 1, 17, 500, 998, 999])
    c = ph.cross_section_at(pts, *args)
    np.testing.assert_allclose(c, a[pts], rtol=1e-13)
    # properties: linearity in S, additivity over line subsets
    ln2 = dict(ln); ln2["sw"] = ln["sw"] * 3.0
    np.testing.assert_allclose(ph.cross_section(ln2, *args[1:]), 3.0 * a, rtol=1e-14)
    h = {k: v[:70] for k, v in ln.items()}; t = {k: v[70:] for k, v in ln.items()}
    np.testing.assert_allclose(ph.cross_section(h, *args[1:]) + ph.cross_section(t, *args[1:]), a, rtol=1e-13)
    # T = 296 with Q(T) = Q(296) leaves S unchanged
    p = ph.LineParams(ln, 296, 300.0, 4e-4, 43.98983, 286.09, 286.09)
    np.testing.assert_allclose(p.S, ln["sw"], rtol=1e-15)


# ---------------------------------------------------------------- xsc file utilities
def test_oracle_reproduces_reference_change_res_xsc_file():
    g = G.load("xsc_files")
    src, want = G.files_of(g, "res_in"), G.files_of(g, "res_out")
    got = dict(ph.change_res_xsc_file(name, blob) for name, blob in src.items())
    assert sorted(got) == sorted(want)
    for name in want:
        assert got[name] == want[name], name                     # byte for byte


def test_oracle_reproduces_reference_merge_xsc():
    g = G.load("xsc_files")
    src, want = G.files_of(g, "merge_in"), G.files_of(g, "merge_out")
    got = ph.merge_xsc(src)
    assert sorted(got) == sorted(want)
    for name in want:
        assert got[name] == want[name], name
    # a group that mixes resolutions is refused (pyradUtilities.py:583-586)
    mixed = dict(src)
    k = sorted(mixed)[0]
    mixed[k.replace("_0.01_", "_0.05_")] = mixed.pop(k)
    assert ph.merge_xsc(mixed) is False


def test_xsc_file_name_fields():
    p = ph.parse_xsc_file_name("HCFC22_296.0K-760.0Torr_800.0-840.0_0.01_N2_07_19.txt")
    assert p == {"RANGE": "800.0-840.0", "MOLECULE_SHORT_NAME": "HCFC22", "TEMP": "296.0", "PRESSURE": "760.0",
                 "RES": "0.01", "ID": "07-19", "BROADENER": "N2"}
