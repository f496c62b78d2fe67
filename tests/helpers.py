"""Shared helpers for the parity tests: oracle evaluation of a workload and the error metrics."""
import numpy as np

from oracle import physics as ph
from pyrad_b200 import engine as eng

#: north_star tolerances
K_REL_TOL = 1e-5        # relative error on k(nu) / sigma(nu)
T_ABS_TOL = 1e-6        # absolute error on transmittance
#: K2 evaluates a Gaussian core as G * ex2(C d^2) in FP32 with flush-to-zero: a tail below 2^-126 (1.2e-38) of ITS OWN
#: line's peak becomes exactly 0, so a value of the spectrum can miss up to 1.2e-38 of the strongest peak in absolute
#: terms.  The 1e-5 relative tolerance is therefore measured against max(|ref|, 2e-33 max|ref|) (1.2e-38 / 1e-5); only
#: far tails of pure-Gaussian lines on coarse grids get that low (found by tests/test_gpu_fuzz.py)
K_FLOOR_REL = 2e-33


def k_rel_err(out, ref, peak=None):
    """|out - ref| / max(|ref|, floor).  The floor is K_FLOOR_REL of the strongest peak: max|ref| when `ref` is a whole
    spectrum, or `peak` (see gaussian_peak) when `ref` is a sample that may not hold any line centre."""
    ref = np.asarray(ref)
    top = (np.max(np.abs(ref)) if ref.size else 0.0) if peak is None else peak
    den = np.maximum(np.abs(ref), K_FLOOR_REL * top)
    den = np.where(den == 0, 1.0, den)
    return np.abs(np.asarray(out) - ref) / den


def gaussian_peak(w, weights):
    """Upper bound of the tallest Gaussian core of a gas-cell workload in the units of k: max over lines of
    S w / (gD sqrt(pi)) (a pseudo-Voigt core is (1 - eta) / (h sqrt(pi)) with h >= ~gD)."""
    top = 0.0
    for g, sp in enumerate(w["species"]):
        ln = w["per_group_lines"][g]
        if len(ln["nu"]) == 0:
            continue
        p = ph.LineParams(ln, w["T"], w["P"], w["conc"][g], sp.molmass, sp.q(w["T"]), sp.q296)
        ok = p.gD > 0
        if ok.any():
            top = max(top, float(np.max(np.abs(p.S[ok]) / p.gD[ok])) * weights[g] / np.sqrt(np.pi))
    return top


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


def boundary_points(n, k_random, seed, tile=2048, span=256, n_tiles=24):
    """Deterministic sample of grid indices that stresses the kernels' class boundaries: the first and last points,
    both sides of `n_tiles` tile edges (k*tile-1, k*tile) and warp-span edges (k*span-1, k*span, k*span+1) spread over
    the grid, plus `k_random` seeded random points."""
    rng = np.random.default_rng(seed)
    edge = [0, 1, 2, 3, n // 2, n - 3, n - 2, n - 1]
    tiles = np.unique(rng.integers(1, max(n // tile, 2), n_tiles))
    for t in tiles:
        edge += [t * tile - 1, t * tile]
        s = t * tile + span * int(rng.integers(1, tile // span))
        edge += [s - 1, s, s + 1]
    pts = np.concatenate([np.array(edge, dtype=np.int64), rng.integers(0, n, k_random)])
    return np.unique(pts[(pts >= 0) & (pts < n)])


def _lines_near(idx, lines, points, wm):
    """The sub-list of `lines` (ascending) whose index lies within wm of at least one of `points`."""
    lo = np.searchsorted(idx, points - wm, side="left")
    hi = np.searchsorted(idx, points + wm, side="right")
    mark = np.zeros(len(idx) + 1, dtype=np.int64)
    np.add.at(mark, lo, 1)
    np.add.at(mark, hi, -1)
    sel = np.nonzero(np.cumsum(mark[:-1]) > 0)[0]
    return {k: np.asarray(v)[sel] for k, v in lines.items()}


def oracle_layer_k_at(w, points, T, P, conc, cutoff):
    """k(nu) of one layer at the grid indices `points` (oracle gather form, only the lines that can reach them)."""
    points = np.asarray(points, dtype=np.int64)
    wm = max(ph.window_len(cutoff, w["res"]) - 2, 0)
    k = np.zeros(len(points))
    for g, sp in enumerate(w["species"]):
        ln = w["per_group_lines"][g]
        idx = w.setdefault("_idx_cache", {}).get(g)
        if idx is None:
            idx = w["_idx_cache"][g] = ph.line_index(ln["nu"], w["range_min"], w["res"])
        sub = _lines_near(idx, ln, points, wm)
        if len(sub["nu"]) == 0:
            continue
        sig = ph.cross_section_at(points, sub, T, P, conc[g], sp.molmass, sp.q(T), sp.q296, w["range_min"],
                                  w["range_max"], w["res"], cutoff)
        k += ph.abs_coef(sig, conc[g], P, T)
    return k


def oracle_column_at(w, points, t_surface):
    """Radiance and total transmittance of a column workload at the grid indices `points`: per layer the oracle's
    k -> exp(-k u) -> T I + (1 - T) B fold (pyradClasses.py:707-716, 784-787), bottom to top from B(nu, t_surface)."""
    points = np.asarray(points, dtype=np.int64)
    xa = ph.x_axis(w["range_min"], w["range_max"], w["res"])[points]
    rad = ph.planck_wavenumber(xa, t_surface)
    total = np.ones(len(points))
    layers = np.atleast_1d(w["T"])
    for l in range(len(layers)):
        T, P = w["T"][l], w["P"][l]
        k = oracle_layer_k_at(w, points, T, P, w["conc"][l], w["cutoff"][l])
        t = ph.transmittance(k, w["depth_cm"][l])
        rad = ph.transmission(t, rad, ph.planck_wavenumber(xa, T))
        total = total * t
    return rad, total


class LayerModel:
    """What a reference Layer holds and computes, restated with the oracle (pyradClasses.py:640-780): the state the
    change* methods move, the line subset getData keeps (effective range as it stood when the data was last loaded --
    changePressure does not update it, changeRange does), and k / T of the layer on the base grid."""

    def __init__(self, species, lines, conc, depth, T, P, rmin, rmax, base=.01, dynamic=True):
        self.species, self.lines, self.conc = species, lines, list(conc)
        self.depth, self.T, self.P, self.rmin, self.rmax, self.base, self.dynamic = depth, T, P, rmin, rmax, base, dynamic
        self.cutoff = ph.layer_cutoff(P)
        self.eff = ph.effective_range(rmin, rmax, self.cutoff)
        self.res = ph.layer_resolution(P, base, dynamic)

    def change_temperature(self, T):
        self.T = T

    def change_pressure(self, P):
        self.P = P
        self.cutoff = ph.layer_cutoff(P)                 # effectiveRange* stay as they were (:746-753)
        self.res = ph.layer_resolution(P, self.base, self.dynamic)

    def change_range(self, rmin, rmax):
        self.rmin, self.rmax = rmin, rmax
        self.eff = ph.effective_range(rmin, rmax, self.cutoff)

    def change_depth(self, depth):
        self.depth = depth

    def kept(self, g):
        ln = self.lines[g]
        m = (ln["nu"] > self.eff[0]) & (ln["nu"] < self.eff[1])
        return {k: np.asarray(v)[m] for k, v in ln.items()}

    def sigma(self, g):
        sp = self.species[g]
        o = ph.cross_section(self.kept(g), self.T, self.P, self.conc[g], sp.molmass, sp.q(self.T), sp.q296, self.rmin,
                             self.rmax, self.res, self.cutoff)
        return o if self.res == self.base else ph.interp_to_base(o, self.rmin, self.rmax, self.res, self.base)

    def abs_coef(self):
        return sum(ph.abs_coef(self.sigma(g), self.conc[g], self.P, self.T) for g in range(len(self.species)))

    def line_survey(self):
        """Layer.lineSurvey (pyradClasses.py:409-428, 589-594, 691-696): made when the data is loaded, from the kept lines."""
        n = ph.grid_len(self.rmin, self.rmax, self.base)
        out = np.zeros(n)
        for g in range(len(self.species)):
            k = self.kept(g)
            out += ph.line_survey(k["nu"], k["sw"], self.rmin, self.res, n)
        return out

    def transmittance(self):
        return ph.transmittance(self.abs_coef(), self.depth)


def random_layer_case(seed, max_points=2500, max_lines=60):
    """A small random layer and a sequence of mutations for the mirror / reference state tests."""
    from pyrad_b200 import synth
    rng = np.random.default_rng(1234 + seed)
    names = [str(s) for s in rng.choice(["co2", "h2o", "ch4", "o3"], size=int(rng.integers(1, 4)), replace=False)]
    conc = [float(c) for c in rng.choice([400e-6, 0.01, 1.8e-6, 5e-6], size=len(names))]
    P = float(np.exp(rng.uniform(np.log(0.5), np.log(1500.0))))
    T = int(rng.integers(200, 320))
    rmin = float(rng.choice([600.0, 1000.0, 2349.0]))
    rmax = rmin + float(rng.integers(4, max_points // 100))
    species = [synth.species(n) for n in names]
    # (lines over everything a later changeRange / changePressure can reach)
    lines = [synth.make_lines(int(rng.integers(10, max_lines)), max(rmin - 30.0, 0.0), rmin + max_points // 100 + 30.0, 11 * seed + g)
             for g in range(len(names))]
    steps = []
    for _ in range(int(rng.integers(2, 5))):
        kind = str(rng.choice(["T", "P", "depth", "ppm", "range"]))
        if kind == "T":
            steps.append((kind, int(rng.integers(200, 320))))
        elif kind == "P":
            steps.append((kind, float(np.exp(rng.uniform(np.log(0.5), np.log(1500.0))))))
        elif kind == "depth":
            steps.append((kind, float(np.exp(rng.uniform(np.log(1.0), np.log(1e4))))))
        elif kind == "ppm":
            steps.append((kind, (int(rng.integers(0, len(names))), float(np.round(np.exp(rng.uniform(np.log(1.0), np.log(2e4))), 3)))))
        else:
            a = rmin + float(rng.integers(-3, 4))
            steps.append((kind, (a, a + float(rng.integers(3, max_points // 100)))))
    return dict(names=names, conc=conc, P=P, T=T, rmin=rmin, rmax=rmax, depth=float(rng.choice([10.0, 100.0, 1000.0])),
                species=species, lines=lines, steps=steps)


def seed_layer_case(root, case, rh):
    """Write the case's species files (params, Q table, 100 cm-1 line segments) into a data tree."""
    for sp, ln in zip(case["species"], case["lines"]):
        rh.write_params(root, sp.global_iso, sp.name, sp.mol_id, 1, 0.99, sp.q296, 1, sp.molmass)
        rh.write_q_table(root, sp.global_iso, range(100, 501), [sp.q(t) for t in range(100, 501)])
        rh.write_line_segments(root, sp.global_iso, sp.mol_id, 1, ln, int(max(ln["nu"].min() - 100, 0) / 100) * 100,
                               ln["nu"].max() + 201)


def drive_layer_case(C, case, model, check, dynamic=True):
    """Build the layer through an object model `C` (the real reference's pyradClasses or the mirror), apply the case's
    mutations to it and to the LayerModel in step, and call check(layer, model, tag) after the build and every step."""
    layer = C.Layer(case["depth"], case["T"], case["P"], case["rmin"], case["rmax"], dynamicResolution=dynamic)
    mols = [layer.addMolecule(n, concentration=c) for n, c in zip(case["names"], case["conc"])]
    check(layer, model, "built")
    for kind, v in case["steps"]:
        if kind == "T":
            layer.changeTemperature(v); model.change_temperature(v)
        elif kind == "P":
            layer.changePressure(v); model.change_pressure(v)
        elif kind == "depth":
            layer.changeDepth(v); model.change_depth(v)
        elif kind == "ppm":
            mols[v[0]].setPPM(v[1]); model.conc[v[0]] = v[1] * 10 ** -6
        else:
            layer.changeRange(*v); model.change_range(*v)
        check(layer, model, (kind, v))
    return layer
