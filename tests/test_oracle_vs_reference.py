"""CPU test, build container only: the oracle against the REAL reference executed live through
oracle/ref_harness.py.  Skipped where /root/reference is absent (the GPU box); the committed golden vectors
(tests/test_oracle.py) carry the same pin there."""
import numpy as np
import pytest

from oracle import physics as ph
from oracle import ref_harness as rh
from pyrad_b200 import synth

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference sources not mounted")


def test_real_reference_loop_equals_oracle(tmp_path):
    wd = str(tmp_path)
    sp = synth.species("co2")
    lines = synth.make_lines(250, 595.2, 704.8, 1)
    rh.seed_workdir(wd)
    rh.write_params(wd, sp.global_iso, "co2", sp.mol_id, 626, 0.98, sp.q296, 1, sp.molmass)
    rh.write_q_table(wd, sp.global_iso, range(100, 400), [sp.q(t) for t in range(100, 400)])
    rh.write_line_segments(wd, sp.global_iso, sp.mol_id, 1, lines, 500, 800)
    ref = rh.load_reference(wd)
    C = ref.classes
    with rh.quiet():
        layer = C.Layer(10, 250, 500.0, 600, 700)
        mol = layer.addMolecule("co2", ppm=400)
        sig = C.getCrossSection(mol[0])
        tr = C.getTransmittance(layer)
        surf = ref.planck.planckWavenumber(layer.xAxis, 288)
        rad = layer.transmission(surf)
    lo, hi = ph.effective_range(600, 700, ph.layer_cutoff(500.0))
    m = (lines["nu"] > lo) & (lines["nu"] < hi)
    sub = {k: v[m] for k, v in lines.items()}
    assert np.array_equal([l.wavenumber for l in mol[0]], sub["nu"])
    args = (sub, 250, 500.0, 400e-6, sp.molmass, sp.q(250), sp.q296, 600, 700, 0.01, ph.layer_cutoff(500.0))
    o = ph.cross_section(*args)
    np.testing.assert_allclose(o, sig, rtol=1e-13, atol=0)
    pts = np.array([0, 1, 5, 100, 5000, 9998, 9999])
    np.testing.assert_allclose(ph.cross_section_at(pts, *args), sig[pts], rtol=1e-13)
    t = ph.transmittance(ph.abs_coef(o, 400e-6, 500.0, 250), 10)
    np.testing.assert_allclose(t, tr, rtol=1e-13)
    xa = ph.x_axis(600, 700, .01)
    np.testing.assert_array_equal(xa, layer.xAxis)
    np.testing.assert_allclose(ph.transmission(t, ph.planck_wavenumber(xa, 288), ph.planck_wavenumber(xa, 250)), rad, rtol=1e-13)
    # accumulation count of the reference loop == the metric's pair definition
    n = ph.grid_len(600, 700, .01)
    W = ph.window_len(ph.layer_cutoff(500.0), .01)
    idx = ph.line_index(sub["nu"], 600, .01)
    brute = sum(sum(1 for d in range(-(W - 2), W - 1) if 0 <= i + d <= n - 1) for i in idx)
    assert ph.pair_count(idx, n, W) == brute
