"""Dose comparison metrics shared by the parity tests and bench.py (test infrastructure)."""
import numpy as np


def r80_mm(idd, dz=1.0):
    """Distal 80 % range [mm from the entry face].  idd is indexed like the reference output:
    k = nz-1 is the entry slab, depth of slab k centre = (nz-1-k+0.5)*dz."""
    d = np.asarray(idd, dtype=np.float64)[::-1]   # now index = depth bin
    depth = (np.arange(d.size) + 0.5) * dz
    ipk = int(d.argmax())
    lvl = 0.8 * d[ipk]
    for i in range(ipk, d.size - 1):
        if d[i] >= lvl > d[i + 1]:
            t = (d[i] - lvl) / (d[i] - d[i + 1])
            return depth[i] + t * dz
    return float("nan")


def gamma_1d(ref, ev, spacing_mm, dd=0.01, dta_mm=1.0, cut=0.10, upsample=10, window_mm=3.0):
    """Global 1-D gamma index of `ev` against `ref` (same grid).  Returns (pass_rate, gamma array on
    the evaluated points, mask)."""
    ref = np.asarray(ref, dtype=np.float64)
    ev = np.asarray(ev, dtype=np.float64)
    n = ref.size
    x = np.arange(n) * spacing_mm
    xf = np.arange((n - 1) * upsample + 1) * (spacing_mm / upsample)
    evf = np.interp(xf, x, ev)
    dmax = ref.max()
    mask = ref > cut * dmax
    g = np.full(n, np.nan)
    w = int(round(window_mm / (spacing_mm / upsample)))
    for i in np.nonzero(mask)[0]:
        c = i * upsample
        lo, hi = max(0, c - w), min(xf.size, c + w + 1)
        dist2 = ((xf[lo:hi] - x[i]) / dta_mm) ** 2
        dose2 = ((evf[lo:hi] - ref[i]) / (dd * dmax)) ** 2
        g[i] = np.sqrt((dist2 + dose2).min())
    return float((g[mask] <= 1.0).mean()), g, mask


def gamma_2d(ref, ev, spacing_mm, dd=0.01, dta_mm=1.0, cut=0.10, window_mm=2.0, upsample=4):
    """Global 2-D gamma (arrays [n0][n1], spacing per axis).  Brute force over an upsampled window."""
    from scipy.ndimage import zoom
    ref = np.asarray(ref, dtype=np.float64)
    ev = np.asarray(ev, dtype=np.float64)
    s0, s1 = spacing_mm
    dmax = ref.max()
    mask = ref > cut * dmax
    evf = zoom(ev, upsample, order=1, grid_mode=False, mode="nearest")
    # zoom with grid_mode False maps index i -> i*(N*up-1)/(N-1); use explicit interpolation grid instead
    n0, n1 = ref.shape
    f0 = np.linspace(0, n0 - 1, (n0 - 1) * upsample + 1)
    f1 = np.linspace(0, n1 - 1, (n1 - 1) * upsample + 1)
    from scipy.interpolate import RegularGridInterpolator
    itp = RegularGridInterpolator((np.arange(n0), np.arange(n1)), ev)
    F0, F1 = np.meshgrid(f0, f1, indexing="ij")
    evf = itp(np.stack([F0.ravel(), F1.ravel()], axis=-1)).reshape(F0.shape)
    w0 = int(np.ceil(window_mm / s0 * upsample))
    w1 = int(np.ceil(window_mm / s1 * upsample))
    best = np.full(ref.shape, np.inf)
    I0, I1 = np.nonzero(mask)
    c0, c1 = I0 * upsample, I1 * upsample
    for a in range(-w0, w0 + 1):
        p0 = np.clip(c0 + a, 0, evf.shape[0] - 1)
        d0 = ((p0 - c0) * s0 / upsample / dta_mm) ** 2
        for b in range(-w1, w1 + 1):
            p1 = np.clip(c1 + b, 0, evf.shape[1] - 1)
            d1 = ((p1 - c1) * s1 / upsample / dta_mm) ** 2
            g2 = d0 + d1 + ((evf[p0, p1] - ref[I0, I1]) / (dd * dmax)) ** 2
            best[I0, I1] = np.minimum(best[I0, I1], g2)
    g = np.sqrt(best)
    return float((g[mask] <= 1.0).mean()), g, mask


def gamma_3d(ref, ev, spacing_mm, dd=0.01, dta_mm=1.0, cut=0.10, step_mm=0.25, ref_max=None):
    """Global 3-D gamma index of `ev` against `ref` (arrays [n0][n1][n2] on the same grid, spacing per axis) in the
    voxels where ref > cut * max(ref): for every such reference voxel the minimum over the evaluated distribution,
    linearly interpolated on a lattice of `step_mm` inside the distance-to-agreement sphere, of
    sqrt((dose difference / (dd * max))^2 + (distance / dta)^2).  Voxels that agree at zero distance are settled
    first; only the others are searched.  Returns (pass_rate, gamma array (nan outside the mask), mask)."""
    from scipy.ndimage import map_coordinates
    ref = np.asarray(ref, dtype=np.float64)
    ev = np.asarray(ev, dtype=np.float64)
    dmax = float(ref.max()) if ref_max is None else float(ref_max)
    mask = ref > cut * dmax
    tol = dd * dmax
    g2 = np.full(ref.shape, np.nan)
    g2[mask] = ((ev[mask] - ref[mask]) / tol) ** 2
    pend = np.argwhere(mask & ~(g2 <= 1.0))   # nan-safe
    if pend.size:
        offs = []
        r = [int(np.floor(dta_mm / step_mm))] * 3
        for a in range(-r[0], r[0] + 1):
            for b in range(-r[1], r[1] + 1):
                for c in range(-r[2], r[2] + 1):
                    d2 = (a * a + b * b + c * c) * step_mm ** 2
                    if 0 < d2 < dta_mm ** 2:
                        offs.append((d2 / dta_mm ** 2, a * step_mm / spacing_mm[0], b * step_mm / spacing_mm[1], c * step_mm / spacing_mm[2]))
        offs.sort()
        best = g2[tuple(pend.T)]
        rv = ref[tuple(pend.T)]
        alive = np.arange(len(pend))
        for d2, a, b, c in offs:
            if alive.size == 0:
                break
            p = pend[alive].T.astype(np.float64)
            val = map_coordinates(ev, [p[0] + a, p[1] + b, p[2] + c], order=1, mode="nearest")
            cand = d2 + ((val - rv[alive]) / tol) ** 2
            best[alive] = np.minimum(best[alive], cand)
            alive = alive[best[alive] > 1.0]
        g2[tuple(pend.T)] = best
    g = np.sqrt(g2)
    return float((g[mask] <= 1.0).mean()), g, mask


def fraction_within_sigma(a, a_se, b, b_se, nsig=2.0, cut=0.10):
    """Fraction of voxels (above cut * max) whose difference is within nsig combined standard errors."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    s = np.sqrt(np.asarray(a_se, dtype=np.float64) ** 2 + np.asarray(b_se, dtype=np.float64) ** 2)
    mask = (a > cut * a.max()) & (s > 0)
    z = np.abs(a - b)[mask] / s[mask]
    return float((z <= nsig).mean()), z


def reduce_dose(d, rebin):
    """Same reductions as oracle/ref_run.py: depth dose, projections, lateral rebin."""
    nz, ny, nx = d.shape
    return {
        "idd": d.sum(axis=(1, 2)),
        "xz": d.sum(axis=1),
        "yz": d.sum(axis=2),
        "reb": d.reshape(nz, ny // rebin, rebin, nx // rebin, rebin).sum(axis=(2, 4)),
        "total": d.sum(),
    }
