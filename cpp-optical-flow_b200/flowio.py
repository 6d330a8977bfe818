"""On-disk formats for the flow fields, the step after the hot path.

The reference dumps u and v as OpenCV FileStorage text (HornSchunckOF/main.cpp:99-102:
`%YAML:1.0`, `!!opencv-matrix`, dt: d) - about 9 MB of text per 1242x375 field, which is why its
own golden dumps are missing upstream.  This module writes that format for drop-in consumers and
two compact binary ones (Middlebury .flo, .npy).  Pure host code, no arithmetic.
"""
from __future__ import annotations

import numpy as np

FLO_MAGIC = 202021.25


def write_flo(path: str, u: np.ndarray, v: np.ndarray) -> None:
    """Middlebury .flo: float32 magic, int32 width, int32 height, then interleaved (u, v) float32."""
    u = np.asarray(u, np.float32)
    v = np.asarray(v, np.float32)
    if u.shape != v.shape or u.ndim != 2:
        raise ValueError("u and v must be 2-D arrays of the same shape")
    with open(path, "wb") as f:
        np.array([FLO_MAGIC], np.float32).tofile(f)
        np.array([u.shape[1], u.shape[0]], np.int32).tofile(f)
        np.stack([u, v], axis=-1).tofile(f)


def read_flo(path: str):
    with open(path, "rb") as f:
        if np.fromfile(f, np.float32, 1)[0] != np.float32(FLO_MAGIC):
            raise ValueError("not a .flo file")
        w, h = np.fromfile(f, np.int32, 2)
        d = np.fromfile(f, np.float32, 2 * w * h).reshape(h, w, 2)
    return d[..., 0].copy(), d[..., 1].copy()


def write_opencv_yaml(path: str, name: str, m: np.ndarray) -> None:
    """What `cv::FileStorage fs(path, WRITE); fs << name << m;` produces for a CV_64FC1 matrix
    (main.cpp:99-102 writes "u matrix" / "v matrix")."""
    m = np.asarray(m, np.float64)
    if m.ndim != 2:
        raise ValueError("expected a 2-D matrix")
    def fmt(x):
        if np.isnan(x):
            return ".Nan"
        if np.isinf(x):
            return ".Inf" if x > 0 else "-.Inf"
        t = repr(float(x))                      # shortest round-trip form, like OpenCV's
        return t[:-1] if t.endswith(".0") else t

    vals = [fmt(x) for x in m.ravel()]
    lines, cur = [], "   data: ["
    for i, t in enumerate(vals):
        piece = " " + t + ("," if i + 1 < len(vals) else " ]")
        if len(cur) + len(piece) > 76 and cur.strip() not in ("data: [",):
            lines.append(cur)
            cur = "      " + piece
        else:
            cur += piece
    lines.append(cur)
    key = name
    with open(path, "w") as f:
        f.write("%YAML:1.0\n---\n")
        f.write(f"{key}: !!opencv-matrix\n   rows: {m.shape[0]}\n   cols: {m.shape[1]}\n   dt: d\n")
        f.write("\n".join(lines) + "\n")
