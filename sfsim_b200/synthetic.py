"""Deterministic synthetic world rasters for tests, profiling runs and the bench (the NASA source rasters of
`clj -T:build map-tiles / elevation-tiles` are not part of the reference tree).  Plain numpy: no device code."""
import numpy as np


def synthetic_world(width, levels_elevation, levels_color, seed=1):
    """Deterministic rasters for tests and the bench: smooth terrain in [-500, 8000] m with sea (negative) regions plus
    per-pixel detail, random colours.  Each level is generated independently (not a pyramid): the path only reads."""
    rng = np.random.default_rng(seed)
    elevation, day, night = {}, {}, {}
    for level in levels_elevation:
        n = 1 << level
        rows, cols = 2 * n * width, 4 * n * width
        y = np.linspace(0, np.pi, rows, dtype=np.float32)[:, None]
        x = np.linspace(0, 2 * np.pi, cols, endpoint=False, dtype=np.float32)[None, :]
        h = 2500.0 * np.sin(3 * x + 1.0) * np.sin(2 * y) + 1500.0 * np.cos(7 * x) * np.sin(5 * y + 0.3) + 500.0
        h = h + rng.integers(-200, 200, size=(rows, cols)).astype(np.float32)
        h = np.clip(h, -500, 8000).astype(np.int16)
        elevation[level] = np.ascontiguousarray(h.reshape(2 * n, width, 4 * n, width).transpose(0, 2, 1, 3))
    for level in levels_color:
        n = 1 << level
        shape = (2 * n, 4 * n, width, width, 4)
        day[level] = rng.integers(0, 256, size=shape, dtype=np.uint8)
        night[level] = rng.integers(0, 256, size=shape, dtype=np.uint8)
    return elevation, day, night
