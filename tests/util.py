import numpy as np

# north_star tolerance: per-bin model values and chi-square within 1e-10 relative in fp64.
# xi changes sign across the grid, so "relative" is taken against max(|xi_i|, 1e-3 * max|xi|)
# ("1e-10 relative per bin needs a floor").
RTOL = 1e-10


def rel_err_bins(got, ref):
    got, ref = np.asarray(got), np.asarray(ref)
    scale = np.maximum(np.abs(ref), 1e-3 * np.max(np.abs(ref), axis=0, keepdims=True))
    return float(np.max(np.abs(got - ref) / scale))


def rel_err(got, ref):
    got, ref = np.asarray(got), np.asarray(ref)
    return float(np.max(np.abs(got - ref) / np.abs(ref)))
