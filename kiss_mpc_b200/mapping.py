"""Occupancy map -> static obstacle circles (replaces obstacle_handling/static_obstacle.py:12-56).

The reference script thresholds the map, takes the distance transform of the occupied region and greedily packs it with the
largest inscribed circles -- and then only paints them.  ``map_to_circles`` returns them (libkmpc.so: kmpc_map_to_circles, host
code restating OpenCV's arithmetic to the bit), in pixels or, with a resolution / origin as in a ROS map_server YAML, in metres:
the candidate set ``BatchedMotionPlanner.select_obstacles`` / ``closed_loop(obstacle_centers=...)`` filter per agent."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def read_pgm(path: str) -> np.ndarray:
    """Binary (P5) 8-bit PGM -> uint8 array [h, w] (what cv2.imread(path, IMREAD_GRAYSCALE) returns for such a file)."""
    with open(path, "rb") as f:
        data = f.read()
    tokens, pos = [], 0
    while len(tokens) < 4:                       # magic, width, height, maxval; comments start with '#'
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            pos = data.index(b"\n", pos) + 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        tokens.append(data[pos:end]); pos = end
    if tokens[0] != b"P5" or int(tokens[3]) > 255:
        raise ValueError("only binary 8-bit PGM (P5) is supported")
    w, h = int(tokens[1]), int(tokens[2])
    return np.frombuffer(data, dtype=np.uint8, count=w * h, offset=pos + 1).reshape(h, w).copy()


def map_to_circles(image: np.ndarray, threshold: int = 127, min_radius: float = 1.0, resolution: float | None = None,
                   origin=(0.0, 0.0), max_circles: int | None = None):
    """image [h, w] uint8 (dark = occupied).  Returns (centers [M, 2], radii [M]) in the order the reference script finds them
    (largest first).  resolution None: pixels (x = column, y = row, int32).  resolution r (metres per pixel) and origin (x0, y0)
    of the lower-left pixel as in a map_server YAML: metres, float64, y up -- x = x0 + (col + 0.5) r, y = y0 + (h - row - 0.5) r."""
    L = _lib.load()
    img = np.ascontiguousarray(image, dtype=np.uint8)
    if img.ndim != 2:
        raise ValueError("image must be [h, w] uint8")
    h, w = img.shape
    cap = int(max_circles) if max_circles is not None else h * w // 2 + 1
    cen = np.empty((cap, 2), np.int32); rad = np.empty(cap, np.int32); n = C.c_int32()
    rc = L.kmpc_map_to_circles(img.ctypes.data, w, h, int(threshold), float(min_radius), cap, cen.ctypes.data, rad.ctypes.data, C.byref(n))
    if rc != 0:
        raise _lib.KmpcError(f"kmpc_map_to_circles failed (rc={rc})")
    m = min(int(n.value), cap)
    cen, rad = cen[:m].copy(), rad[:m].copy()
    if resolution is None:
        return cen, rad
    out = np.empty((m, 2))
    out[:, 0] = origin[0] + (cen[:, 0] + 0.5) * resolution
    out[:, 1] = origin[1] + (h - cen[:, 1] - 0.5) * resolution
    return out, rad * float(resolution)
