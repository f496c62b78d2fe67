"""Wavenumber-chunk sharding of the grid across ranks.

Every grid point is independent given the line list, so the path shards with NO exchange step:
each rank owns a contiguous chunk of grid points, reads the lines whose cutoff window reaches the
chunk (read-only duplication at chunk edges) and the only collective is one all-gather of the
finished spectra.  Chunks are
  * aligned to ALIGN grid points (the largest K2 tile), which makes the sharded result bitwise
    identical to the unsharded one (per-point summation order is a function of the tile only), and
  * balanced by (line, grid point) pair count -- not by width -- because real line density is very
    non-uniform in wavenumber.
Pure numpy; no device code here.
"""
import numpy as np

ALIGN = 4096          # 8 consumer warps x 32 lanes x 16 points: the largest K2 tile


def line_index(nu0, range_min, res):
    """arrayIndex of pyradClasses.py:390 (FP64 divide, truncation toward zero) -- host copy used only to
    plan the sharding; the engine recomputes it on the device."""
    return np.trunc((np.asarray(nu0, dtype=np.float64) - range_min) / res).astype(np.int64)


def block_pair_cost(idx, n_total, windows, block=ALIGN):
    """cost[b] = number of (line, point) accumulations falling into grid block b, summed over the layers'
    windows (W = len(arange(0, cutoff, res)); |d| <= max(W-2, 0))."""
    idx = np.asarray(idx, dtype=np.int64)
    windows = [int(w) for w in np.atleast_1d(windows)]
    nb = (n_total + block - 1) // block
    cost = np.zeros(nb, dtype=np.float64)
    if idx.size == 0 or n_total == 0:
        return cost
    wmax = max(max(w - 2, 0) for w in windows)
    off = wmax + 2
    size = n_total + 2 * off
    hist = np.bincount(np.clip(idx + off, 0, size - 1), minlength=size)
    # lines clipped onto the borders lie outside every window of in-range points only if they are
    # further than wmax away; drop them instead of piling them on the border cells
    far = (idx + off < 0) | (idx + off > size - 1)
    if far.any():
        hist = np.bincount((idx + off)[~far], minlength=size)
    C = np.cumsum(hist)                       # C[j] = #lines with idx+off <= j
    D = np.concatenate([[0], np.cumsum(C)])   # D[j+1] = sum_{t<=j} C[t]
    a = np.arange(nb, dtype=np.int64) * block            # block start (grid index)
    b = np.minimum(a + block, n_total)                   # block end (exclusive)
    for w in set(windows):
        mult = windows.count(w)
        wm = max(w - 2, 0)
        # sum_{i=a}^{b-1} [C(i+wm) - C(i-wm-1)]   (arguments shifted by off)
        hi = D[b + wm + off] - D[a + wm + off]
        lo = D[b - wm - 1 + off] - D[a - wm - 1 + off]
        cost += mult * (hi - lo)
    return cost


#: Measured B200 cost model of K2 per kernel class (profiles/r01_k2_experiments.txt): seconds per (line, point)
#: accumulation and per grid point of a layer, by window W-2.  Only the RATIOS matter for balancing.
def _class_cost(wm):
    if wm >= 1024:
        return 1 / 4.9e12, 0.0            # k2_line_sum<8>
    if wm >= 256:
        return 1 / 3.1e12, 0.0            # k2_line_sum<4>
    if wm >= 100:
        return 1 / 1.7e12, 0.0            # k2_line_sum<2>
    if wm >= 16:
        return 1 / 1.5e12, 4e-12          # k2_point
    return 1 / 1.5e12, 1.5e-11            # k2_narrow: a few lines per point, per-point overhead dominates


#: far-field variant (PRB_K2_FARFIELD): a node evaluation costs about two exact pair evaluations (measured,
#: profiles/r02_k2_farfield.txt), 16 nodes per span (level 1) or per 2048-point tile (level 2, windows >= 8192)
FAR_NODES, FAR_NODE_COST, FAR_TILE, FAR_L2_MIN_WM = 16, 2.0, 2048, 8192


def _farfield_pair_equivalents(idx, n_total, w, block):
    """What the far-field kernel evaluates for one layer of window w, per grid block, in exact-pair equivalents: the lines
    within 1.5 spans of a point and the partially covering ones (about one more span at either window edge) point by
    point, every other line of the window at the span's nodes -- or at the tile's nodes beyond 1.5 tiles (level 2)."""
    wm = max(w - 2, 0)
    span = 256 if wm >= 1024 else 128
    near_w = min(w, int(2.5 * span) + 2)
    near = block_pair_cost(idx, n_total, [near_w], block)
    all_pairs = block_pair_cost(idx, n_total, [w], block)
    if wm >= FAR_L2_MIN_WM:
        mid_w = min(w, int(1.5 * FAR_TILE) + FAR_TILE // 2 + 2)
        mid = block_pair_cost(idx, n_total, [mid_w], block)
        far = (mid - near) * FAR_NODES / span + (all_pairs - mid) * FAR_NODES / FAR_TILE
    else:
        far = (all_pairs - near) * FAR_NODES / span
    return near + FAR_NODE_COST * np.maximum(far, 0.0)


def block_time_cost(idx, n_total, windows, block=ALIGN, farfield=False):
    """Estimated K2 seconds per grid block over all layers: the pair counts of block_pair_cost weighted by the
    measured cost of the kernel class each window runs on, plus the per-point overhead of the thread-per-point
    kernels.  Balancing on this instead of the raw pair count evens out ranks whose chunks differ in width (edge
    chunks see one-sided windows, so equal pair counts give them more points -- and more narrow-layer work).
    farfield=True: the wide classes (W-2 >= 256) run the far-field variant, whose far pairs are nearly free."""
    windows = [int(w) for w in np.atleast_1d(windows)]
    nb = (n_total + block - 1) // block
    a = np.arange(nb, dtype=np.int64) * block
    pts = (np.minimum(a + block, n_total) - a).astype(np.float64)
    cost = np.zeros(nb, dtype=np.float64)
    by_class = {}
    for w in windows:
        by_class.setdefault(_class_cost(max(w - 2, 0)), []).append(w)
    for (c_pair, c_point), ws in by_class.items():
        if farfield and max(ws[0] - 2, 0) >= 256:
            for w in set(ws):
                cost += ws.count(w) * c_pair * _farfield_pair_equivalents(idx, n_total, w, block)
        else:
            cost += c_pair * block_pair_cost(idx, n_total, ws, block) + c_point * len(ws) * pts
    return cost


def balanced_chunks(cost, n_total, nranks, block=ALIGN):
    """Contiguous, block-aligned chunks [(i_begin, i_end)] with near-equal cost.  A rank may get an
    empty chunk when there are fewer blocks than ranks."""
    nb = len(cost)
    total = float(np.sum(cost))
    if total <= 0:
        cum = np.arange(1, nb + 1, dtype=np.float64)
        total = float(nb)
    else:
        cum = np.cumsum(cost)
    cuts = [0]
    for r in range(1, nranks):
        target = total * r / nranks
        k = int(np.searchsorted(cum, target, side="left")) + 1   # blocks [0, k) reach the target
        # pick the nearer of k-1 / k
        if k - 1 >= 1 and abs(cum[k - 2] - target) <= abs(cum[min(k - 1, nb - 1)] - target):
            k -= 1
        cuts.append(min(max(k, cuts[-1]), nb))
    cuts.append(nb)
    return [(min(c0 * block, n_total), min(c1 * block, n_total)) for c0, c1 in zip(cuts[:-1], cuts[1:])]


def rebalanced_chunks(cost, chunks, measured, n_total, block=ALIGN):
    """One feedback step: `measured[r]` is what rank r actually took on `chunks[r]`.  The model's cost of every block
    is rescaled by its rank's measured/predicted ratio (the model cannot know, e.g., that high-wavenumber lines carry
    wider Doppler cores and so more Gaussian work per pair) and the cuts are recomputed on the corrected cost."""
    cost = np.asarray(cost, dtype=np.float64)
    corrected = cost.copy()
    for (a, b), t in zip(chunks, measured):
        b0, b1 = a // block, (b + block - 1) // block
        pred = cost[b0:b1].sum()
        if pred > 0 and t > 0:
            corrected[b0:b1] *= t / pred
    return balanced_chunks(corrected, n_total, len(chunks), block), corrected


def equal_chunks(n_total, nranks, block=ALIGN):
    nb = (n_total + block - 1) // block
    return balanced_chunks(np.ones(nb), n_total, nranks, block)


def lines_for_chunk(nu0, range_min, res, i_begin, i_end, wmax):
    """Slice [l0, l1) of the (sorted) line list whose windows can reach [i_begin, i_end) with |d| <= wmax."""
    idx = line_index(nu0, range_min, res)
    l0 = int(np.searchsorted(idx, i_begin - wmax, side="left")) & ~3      # four-line alignment: see ShardPlan
    l1 = int(np.searchsorted(idx, i_end - 1 + wmax, side="right"))
    return l0, max(l1, l0)


def assemble(gathered, chunks):
    """gathered: (nranks, max_chunk) padded rows from the all-gather -> the full spectrum."""
    parts = [np.asarray(gathered[r])[: b - a] for r, (a, b) in enumerate(chunks)]
    return np.concatenate(parts) if parts else np.zeros(0)


def farfield_work(idx, i_begin, i_end, window, span, radius_spans=1, nodes=16, level2_spans=8, level2_min_domains=4):
    """What the far-field variant of K2 (PRB_K2_FARFIELD) actually evaluates on the chunk [i_begin, i_end)
    for one layer of window W: (pairs evaluated point by point, node evaluations).  A warp span starts at
    i_begin + k * span; a line is summed at the span's `nodes` Chebyshev nodes instead of at its points when its window
    covers the whole span (idx in [last - wm, first + wm]) and its centre lies beyond the far thresholds
    idx < first + (span-1)//2 - radius_spans*span  or  idx > first + span//2 + radius_spans*span  -- the kernel's own
    integer tests (k2_line_sum.cuh).  Level 2 applies the same test to domains of `level2_spans` spans: a line far from
    a whole domain costs `nodes` evaluations per DOMAIN and none in the domain's spans; the kernel switches level 2 on for
    windows of at least `level2_min_domains` domain lengths.  Host-side accounting only
    (roofline of the far-field kernel, cost models)."""
    idx = np.sort(np.asarray(idx, dtype=np.int64))
    wm = max(int(window) - 2, 0)
    n_chunk = int(i_end) - int(i_begin)
    if n_chunk <= 0 or idx.size == 0:
        return 0, 0
    count = lambda lo, hi: np.maximum(np.searchsorted(idx, hi, side="left") - np.searchsorted(idx, lo, side="left"), 0)

    def far_counts(length):
        """per block of `length` points from i_begin: (far lines, real points of the block)"""
        first = int(i_begin) + np.arange((n_chunk + length - 1) // length, dtype=np.int64) * length
        last = first + length - 1
        pts = np.minimum(last, int(i_end) - 1) - first + 1           # real points of the (possibly ragged) last block
        full_lo, full_hi = last - wm, first + wm + 1                 # full cover: full_lo <= idx < full_hi
        tfl = first + (length - 1) // 2 - radius_spans * length      # far left:  idx < tfl
        tfr = first + length // 2 + radius_spans * length            # far right: idx > tfr
        far = count(full_lo, np.minimum(tfl, full_hi)) + count(np.maximum(tfr + 1, full_lo), full_hi)
        return np.where(full_lo >= full_hi, 0, far), pts

    # every accumulation of the reference on these points: lines with |i - idx| <= wm
    i = np.arange(int(i_begin), int(i_end), dtype=np.int64)
    pairs_all = int((np.searchsorted(idx, i + wm, side="right") - np.searchsorted(idx, i - wm, side="left")).sum())
    far1, pts1 = far_counts(span)
    far_pairs = int((far1 * pts1).sum())                             # level-2 lines are level-1 far lines of every span
    node_evals = int(far1.sum()) * nodes
    if level2_spans and wm >= level2_min_domains * level2_spans * span:
        far2, _ = far_counts(span * level2_spans)
        # a level-2 line is skipped by each of its domain's spans (level2_spans of them, fewer in a ragged last domain)
        spans_in_dom = np.minimum(level2_spans, (n_chunk + span - 1) // span - np.arange(len(far2)) * level2_spans)
        node_evals += int(far2.sum()) * nodes - int((far2 * spans_in_dom).sum()) * nodes
    return pairs_all - far_pairs, node_evals
