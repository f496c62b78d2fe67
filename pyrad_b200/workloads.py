"""The BASELINE.json configurations as seeded synthetic workloads.

A workload is a plain dict: merged SoA `lines` (with int32 `group`), per-group species info, the
grid (range_min, range_max, res) and the layer(s).  No physics here -- only inputs.
"""
import numpy as np

from . import synth

P0 = 1013.25


def _layer_cutoff(P):
    return P / 1013.25 * 5            # Layer.__init__ pyradClasses.py:655


def gas_cell(names, n_lines_total, rmin, rmax, res, T, P, conc, depth_cm, seed, margin=None, cutoff=None):
    """Single layer, one isotopologue (group) per named molecule."""
    cutoff = _layer_cutoff(P) if cutoff is None else cutoff
    margin = cutoff if margin is None else margin
    lo, hi = max(rmin - margin, 0.0), rmax + margin
    per = []
    sp = [synth.species(n) for n in names]
    for g, s in enumerate(sp):
        per.append(synth.make_lines(n_lines_total // len(names), lo, hi, seed + 101 * g))
    lines = synth.merge_species_lines(per)
    return {
        "lines": lines, "per_group_lines": per, "species": sp,
        "range_min": rmin, "range_max": rmax, "res": res,
        "T": T, "P": P, "conc": list(conc), "depth_cm": depth_cm, "cutoff": cutoff,
    }


def cfg1(n_lines=50_000):
    """single CO2 gas cell 10 cm, 296 K, 1013 hPa, 400 ppm, 500-800 cm-1 at 0.01 cm-1."""
    return gas_cell(["co2"], n_lines, 500.0, 800.0, 0.01, 296, 1013.0, [400e-6], 10.0, synth.SEED0 + 1)


def cfg2(n_lines=500_000, rmax=3000.0):
    """mixed H2O+CO2+CH4+O3 gas cell 0-3000 cm-1 at 0.001 cm-1 (~3M points, ~500k lines)."""
    return gas_cell(["h2o", "co2", "ch4", "o3"], n_lines, 0.0, rmax, 0.001, 296, 1013.25,
                    [0.01, 400e-6, 1.8e-6, 5e-8], 10.0, synth.SEED0 + 2)


def cfg3(n_lines=50_000):
    """gas cell combining line-by-line CO2 / H2O with xsc cross-section tables CFC-11 (native 0.01 cm-1 spacing) and
    HCFC-22 (0.05 cm-1, re-gridded by np.interp) on the same 500-800 cm-1 grid at 0.01 cm-1.  Both tables are
    296 K / 760 Torr files, and an xsc molecule forces the layer to its file's T and P (pyradClasses.py:488-491), so
    the cell sits at 760 Torr / 0.75006."""
    P = 760.0 / 0.75006
    w = gas_cell(["co2", "h2o"], n_lines, 500.0, 800.0, 0.01, 296, P, [400e-6, 0.01], 10.0, synth.SEED0 + 3)
    w["xsc"] = []
    for name, res, a, b, conc, k in (("CFC11", 0.01, 560.0, 620.0, 250e-12, 31), ("HCFC22", 0.05, 700.0, 780.0, 230e-12, 32)):
        x, y = synth.make_xsc_table(a, b, res, synth.SEED0 + k)
        w["xsc"].append({"name": name, "res": res, "range_min": a, "range_max": b, "conc": conc, "T": 296.0,
                         "torr": 760.0, "wavenumber": x, "intensity": y})
    return w


def cfg5(n_lines=5_000_000, rmax=5000.0, res=0.001, cutoff=25.0):
    """stress sweep: 5M synthetic lines, 0-5000 cm-1 at 0.001 cm-1, 25 cm-1 cutoff."""
    return gas_cell(["h2o", "co2", "ch4", "o3"], n_lines, 0.0, rmax, res, 296, 1013.25,
                    [0.01, 400e-6, 1.8e-6, 5e-8], 10.0, synth.SEED0 + 5, cutoff=cutoff)


def atmosphere(n_layers=100, n_lines=5_000_000, rmin=0.0, rmax=5000.0, res=0.001, top_km=70.0,
               fixed_cutoff=None, seed=synth.SEED0 + 4):
    """cfg4: L-layer US-standard-atmosphere column, per-layer T/p, reference cutoff 5*P/p0."""
    depth_cm, T, P = synth.atmosphere_profile(n_layers, top_km)
    names = ["h2o", "co2", "ch4", "o3"]
    sp = [synth.species(n) for n in names]
    cut_max = fixed_cutoff if fixed_cutoff is not None else _layer_cutoff(P.max())
    lo, hi = max(rmin - cut_max, 0.0), rmax + cut_max
    per = [synth.make_lines(n_lines // len(names), lo, hi, seed + 101 * g) for g in range(len(names))]
    lines = synth.merge_species_lines(per)
    # simple composition profile: H2O falls off with height, the others well mixed
    z = (np.arange(n_layers) + 0.5) * top_km / n_layers
    conc = np.stack([np.maximum(0.01 * np.exp(-z / 2.0), 4e-6), np.full(n_layers, 400e-6),
                     np.full(n_layers, 1.8e-6), 5e-8 + 8e-6 * np.exp(-((z - 25.0) / 8.0) ** 2)], axis=1)
    cutoff = np.full(n_layers, fixed_cutoff) if fixed_cutoff is not None else _layer_cutoff(P)
    return {
        "lines": lines, "per_group_lines": per, "species": sp,
        "range_min": rmin, "range_max": rmax, "res": res,
        "depth_cm": np.full(n_layers, depth_cm), "T": T, "P": P, "conc": conc, "cutoff": cutoff,
        "t_surface": 288.0,
    }


def cfg2_shard(rank, world, n_lines_per_chunk=500_000, chunk_width=3000.0, res=0.001):
    """Weak-scaling form of cfg2: the spectrum is `world` consecutive 3000 cm-1 chunks of the cfg2 cell
    (same line density, T, P and mixing ratios), rank r owns chunk r.  Returns the lines rank r needs
    (its chunk plus the cutoff margin on both sides) and the global grid description."""
    names = ["h2o", "co2", "ch4", "o3"]
    sp = [synth.species(n) for n in names]
    P, T = 1013.25, 296
    cutoff = _layer_cutoff(P)
    lo, hi = chunk_width * rank, chunk_width * (rank + 1)
    per = []
    for g in range(len(names)):
        parts = []
        for c in (rank - 1, rank, rank + 1):
            if c < 0:
                continue
            d = synth.make_lines(n_lines_per_chunk // len(names), chunk_width * c, chunk_width * (c + 1),
                                 synth.SEED0 + 2 + 1009 * c + 101 * g)
            m = (d["nu"] > max(lo - cutoff, 0.0)) & (d["nu"] < hi + cutoff)
            parts.append({k: v[m] for k, v in d.items()})
        per.append({k: np.concatenate([p[k] for p in parts]) for k in parts[0]})
    lines = synth.merge_species_lines(per)
    n_chunk = int(round(chunk_width / res))
    return {
        "lines": lines, "per_group_lines": per, "species": sp,
        "range_min": 0.0, "range_max": chunk_width * world, "res": res,
        "n_total": n_chunk * world, "i_begin": n_chunk * rank, "i_end": n_chunk * (rank + 1),
        "T": T, "P": P, "conc": [0.01, 400e-6, 1.8e-6, 5e-8], "depth_cm": 10.0, "cutoff": cutoff,
        "t_surface": 288.0,
    }
