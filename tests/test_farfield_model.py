"""CPU pin of the far-field variant's ALGORITHM (oracle/farfield_model.py: the kernel's class thresholds, FP32-rounded
Chebyshev nodes and Lagrange table in FP64 numpy) against the exact oracle: the approximation error the design states
(16 nodes at a radius of one span -- below 2e-8 of k in the worst case, a valley of the spectrum
dominated by one strong line just beyond the far threshold, 1e-11 typically) holds on the window classes the variant runs on.  The CUDA kernel itself is compared with
the exact kernel and the oracle in tests/test_gpu_parity.py / test_gpu_fullsize.py."""
import numpy as np
import pytest

from oracle import farfield_model as fm
from oracle import physics as ph
from pyrad_b200 import partition as pt
from pyrad_b200 import workloads


def _cell(P, T, n_lines=2500, rmin=1000.0, rmax=1012.0):
    return workloads.gas_cell(["h2o", "co2", "ch4", "o3"], n_lines, rmin, rmax, 0.001, T, P,
                              [0.01, 400e-6, 1.8e-6, 5e-8], 10.0, 31)


def _records(w):
    parts = []
    for g, sp in enumerate(w["species"]):
        wt = float(ph.abs_coef(1.0, w["conc"][g], w["P"], w["T"]))
        parts.append(fm.line_records(w["per_group_lines"][g], w["T"], w["P"], w["conc"][g], sp.molmass, sp.q(w["T"]), sp.q296,
                                     w["range_min"], w["res"], wt))
    idx, A, B, G, C = (np.concatenate([p[i] for p in parts]) for i in range(5))
    order = np.argsort(idx, kind="stable")
    return idx[order], A[order], B[order], G[order], C[order]


def test_unified_form_reproduces_the_oracle_exactly():
    """farfield=False: the model is the oracle's sum in gather form (same shapes, same window)."""
    w = _cell(353.4, 250, n_lines=600, rmax=1003.0)
    n = ph.grid_len(w["range_min"], w["range_max"], w["res"])
    win = ph.window_len(w["cutoff"], w["res"])
    k, frac = fm.line_sum(*_records(w), n, win, 256, farfield=False)
    assert frac == 0
    ref = np.zeros(n)
    for g, sp in enumerate(w["species"]):
        sig = ph.cross_section(w["per_group_lines"][g], w["T"], w["P"], w["conc"][g], sp.molmass, sp.q(w["T"]), sp.q296,
                               w["range_min"], w["range_max"], w["res"], w["cutoff"])
        ref += ph.abs_coef(sig, w["conc"][g], w["P"], w["T"])
    np.testing.assert_allclose(k, ref, rtol=1e-11, atol=1e-300)


@pytest.mark.parametrize("P,T,span", [(1013.25, 296, 256), (353.4, 250, 256), (250.0, 230, 256), (150.0, 225, 128),
                                      (100.0, 215, 128), (60.0, 215, 128)])
def test_farfield_interpolation_error_is_below_2e8_of_k(P, T, span):
    w = _cell(P, T)
    n = ph.grid_len(w["range_min"], w["range_max"], w["res"])
    win = ph.window_len(w["cutoff"], w["res"])
    rec = _records(w)
    exact, _ = fm.line_sum(*rec, n, win, span, farfield=False)
    far, frac = fm.line_sum(*rec, n, win, span, farfield=True)
    err = np.abs(far - exact) / np.maximum(np.abs(exact), 1e-40 * np.abs(exact).max())
    assert err.max() <= 2e-8, (P, span, err.max())          # 16 nodes, far = beyond one span length (measured <= 9.2e-9)
    assert np.median(err) <= 1e-10, (P, span, np.median(err))
    assert frac > 0.3, frac                                  # the far class is not empty on these windows
    # the host-side accounting (partition.farfield_work) counts the same far pairs as the model
    ex_pairs, node_evals = pt.farfield_work(rec[0], 0, n, win, span)
    idx = rec[0]
    total = ph.pair_count(idx, n, win)
    assert ex_pairs == round(total * (1 - frac)) or abs(ex_pairs - total * (1 - frac)) <= 1
    assert node_evals % fm.NODES == 0 and node_evals > 0


def test_farfield_level2_interpolation_error_on_a_stress_sweep_window():
    """Level 2 (lines far from a whole 2048-point tile summed at the tile's 16 nodes) switches on for windows of at least
    four tile lengths: cfg5's 25 cm-1 cutoff (W = 25 000) on a 6 cm-1 cell.  Same error bound as level 1; most of the far
    lines are level-2 lines, so the node evaluations drop well below level 1 alone."""
    w = workloads.gas_cell(["h2o", "co2", "ch4", "o3"], 6000, 1000.0, 1006.0, 0.001, 296, 1013.25,
                           [0.01, 400e-6, 1.8e-6, 5e-8], 10.0, 32, cutoff=25.0)
    n = ph.grid_len(w["range_min"], w["range_max"], w["res"])
    win = ph.window_len(w["cutoff"], w["res"])
    assert win - 2 >= fm.LEVEL2_MIN_DOMAINS * fm.LEVEL2_SPANS * 256
    rec = _records(w)
    exact, _ = fm.line_sum(*rec, n, win, 256, farfield=False)
    far, frac = fm.line_sum(*rec, n, win, 256, farfield=True)
    err = np.abs(far - exact) / np.maximum(np.abs(exact), 1e-40 * np.abs(exact).max())
    assert err.max() <= 2e-8, err.max()
    assert frac > 0.9, frac
    _, evals2 = pt.farfield_work(rec[0], 0, n, win, 256)
    _, evals1 = pt.farfield_work(rec[0], 0, n, win, 256, level2_spans=0)
    assert 0 < evals2 < 0.4 * evals1, (evals2, evals1)


def test_lagrange_table_is_a_partition_of_unity_and_exact_on_polynomials():
    for span in (128, 256, 2048):
        w = fm.lagrange_table(span)
        np.testing.assert_allclose(w.sum(axis=1), 1.0, rtol=0, atol=1e-12)
        x = fm.node_offsets(span)
        i = np.arange(span, dtype=np.float64)
        for deg in range(fm.NODES):
            np.testing.assert_allclose(w @ x ** deg, i ** deg, rtol=0, atol=1e-10 * float(span) ** deg)
