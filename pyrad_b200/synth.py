"""Seeded synthetic HITRAN-format line lists, Q tables and xsc tables.

There is no network, so every workload in BASELINE.json is built from these
generators.  The column set is the one the reference
parses from HITRAN-online CSV (reference pyradUtilities.py:421-448): nu, sw,
a, elower, gamma_air, gamma_self, delta_air, n_air.
"""
from dataclasses import dataclass

import numpy as np

SEED0 = 20261018

#: name -> (HITRAN molecule id, global id of first isotopologue, molar mass g/mol, Q(296 K))
SPECIES = {
    "h2o": (1, 1, 18.010565, 174.58),
    "co2": (2, 7, 43.98983, 286.09),
    "o3": (3, 16, 47.984745, 3483.7),
    "ch4": (6, 32, 16.0313, 590.48),
}


@dataclass
class SpeciesInfo:
    name: str
    mol_id: int
    global_iso: int
    molmass: float
    q296: float

    def q(self, T):
        """Synthetic partition sum, tabulated at integer T like the reference's q<iso>.txt
        (the reference looks Q up by integer T with no interpolation, pyradClasses.py:389)."""
        return self.q296 * (float(int(T)) / 296.0) ** 1.5


def species(name):
    m, g, mass, q = SPECIES[name]
    return SpeciesInfo(name, m, g, mass, q)


def make_lines(n, nu_lo, nu_hi, seed):
    """n lines on a jittered lattice strictly inside (nu_lo, nu_hi), ascending, unique at 1e-6."""
    rng = np.random.default_rng(seed)
    cell = (nu_hi - nu_lo) / n
    nu = nu_lo + (np.arange(n) + rng.uniform(0.05, 0.95, n)) * cell
    nu = np.round(nu, 6)
    nu = np.clip(nu, nu_lo + 1e-6, nu_hi - 1e-6)
    nu = np.unique(nu)
    n = nu.size
    return {
        "nu": nu,
        "sw": 10.0 ** rng.uniform(-30.0, -19.0, n),
        "a": np.ones(n),
        "elower": rng.uniform(0.0, 5000.0, n),
        "gamma_air": rng.uniform(0.03, 0.12, n),
        "gamma_self": rng.uniform(0.05, 0.15, n),
        "delta_air": rng.uniform(-0.01, 0.005, n),
        "n_air": rng.uniform(0.4, 0.9, n),
    }


def merge_species_lines(per_species):
    """per_species: list of line dicts, one per group (isotopologue).  Returns SoA arrays sorted by
    nu ascending (stable) plus the int32 group id of every line."""
    keys = ("nu", "sw", "elower", "gamma_air", "gamma_self", "delta_air", "n_air")
    cat = {k: np.concatenate([np.asarray(d[k], dtype=np.float64) for d in per_species]) for k in keys}
    group = np.concatenate([np.full(len(d["nu"]), g, dtype=np.int32) for g, d in enumerate(per_species)])
    order = np.argsort(cat["nu"], kind="stable")
    out = {k: np.ascontiguousarray(v[order]) for k, v in cat.items()}
    out["group"] = np.ascontiguousarray(group[order])
    return out


def make_xsc_table(rmin, rmax, spacing, seed, peak=5e-18):
    """Synthetic cross-section table: a few Gaussian bumps on [rmin, rmax) at `spacing`."""
    rng = np.random.default_rng(seed)
    n = int(round((rmax - rmin) / spacing))
    x = rmin + np.arange(n) * spacing
    y = np.zeros(n)
    for _ in range(4):
        c = rng.uniform(rmin, rmax)
        w = rng.uniform(0.02, 0.15) * (rmax - rmin)
        y += peak * rng.uniform(0.2, 1.0) * np.exp(-((x - c) / w) ** 2)
    return x, y


# US Standard Atmosphere 1976, geopotential-height layers up to 71 km.
_USSA_HB = np.array([0.0, 11.0, 20.0, 32.0, 47.0, 51.0, 71.0])          # km
_USSA_LB = np.array([-6.5, 0.0, 1.0, 2.8, 0.0, -2.8, -2.0])             # K/km
_USSA_TB = np.array([288.15, 216.65, 216.65, 228.65, 270.65, 270.65, 214.65])
_USSA_PB = np.array([1013.25, 226.3206, 54.74889, 8.680187, 1.109063, 0.6693887, 0.03956420])  # hPa
_G0_M_R = 9.80665 * 0.0289644 / 8.3144598 * 1000.0                      # K/km


def us_standard_atmosphere(z_km):
    """T [K], P [hPa] at altitude z (treated as geopotential height), 0..71 km."""
    z = float(z_km)
    b = int(np.searchsorted(_USSA_HB, z, side="right") - 1)
    b = min(max(b, 0), len(_USSA_HB) - 1)
    dz = z - _USSA_HB[b]
    Tb, Lb, Pb = _USSA_TB[b], _USSA_LB[b], _USSA_PB[b]
    if Lb == 0.0:
        return Tb, Pb * np.exp(-_G0_M_R * dz / Tb)
    T = Tb + Lb * dz
    return T, Pb * (Tb / T) ** (_G0_M_R / Lb)


def atmosphere_profile(n_layers=100, top_km=70.0):
    """Mid-layer T (rounded to integer K, because the reference's Q lookup needs integer T)
    and P for n_layers equal slabs from the surface to top_km.  Returns depth_cm, T[], P[]."""
    dz = top_km / n_layers
    T = np.empty(n_layers)
    P = np.empty(n_layers)
    for l in range(n_layers):
        t, p = us_standard_atmosphere((l + 0.5) * dz)
        T[l] = float(int(round(t)))
        P[l] = p
    return dz * 1e5, T, P
