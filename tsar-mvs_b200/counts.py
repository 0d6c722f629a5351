"""Closed-form work counts of the per-view depthmap path, in pure Python so that callers that
must not load the CUDA library (the reference arm of bench.py) can use them.  Mirrors tsar_eval_count (context.cu)."""
import numpy as np


def refinement_rounds(max_disparity):
    """R of planeRefinement_cu: for (deltaZ = max_disparity / 2; deltaZ >= 0.01; deltaZ /= 10) in float (gipuma.cu:644)."""
    r, dz = 0, np.float32(max_disparity) * np.float32(0.5)
    while dz >= np.float32(0.01):
        r += 1
        dz = np.float32(dz / np.float32(10.0))
    return r


def eval_count(W, H, V, iters, max_disparity):
    """pmCost evaluations (plane x source view x window) of init + `iters` red/black iterations as the reference is
    written: every propagation candidate behind its border guard (gipuma.cu:889-1022) and every refinement round,
    over the rows the reference's checkerboard grid reaches (gipuma.cu:1721)."""
    yl = min(H, 32 * (((H // 2) + 15) // 16))
    sx = sum((x > 2) + (x < W - 3) + (x > 0) + (x < W - 1) for x in range(W))
    sy = sum((y > 2) + (y < H - 3) + (y > 0) + (y < H - 1) for y in range(yl))
    prop = yl * sx + W * sy
    refine = W * yl * refinement_rounds(max_disparity)
    return V * (W * H + iters * (prop + refine))
