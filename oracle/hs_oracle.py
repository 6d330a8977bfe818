"""CPU oracle for the Horn-Schunck hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product path (cpp-optical-flow_b200/, libhs_b200.so) never
touches it and fails loudly when the CUDA library is missing.

What it restates (all line numbers: /root/reference/HornSchunckOF/hornSchunck.cpp):
  * getGradients  :19-41   convertTo(CV_64F) :23-24, Sobel ksize 3 on the *previous* frame
                           only :27-28 (scale 1, BORDER_REFLECT_101), gradT = next - prev :39
  * getFlow       :43-75   zero init :49-50, kernel = ones(w,w)/w^2 :53, anchor = w-w/2-1 :54,
                           maxIterations Jacobi sweeps :56-74 (filter2D BORDER_CONSTANT :60-61,
                           update :63-73)
The arithmetic itself lives in OpenCV 4.4.0 core/imgproc (not vendored in the reference,
pinned by OpenCVx64d.props:6-11).  Two restatements are kept and tested against each other:
  cv_*  - through Python cv2 (same OpenCV entry points the reference calls)
  np_*  - plain NumPy, no OpenCV (this is the one that runs anywhere, in any dtype)
A plain-C twin of np_* lives in hs_oracle.c (faster, used for larger parity cases).

Pinning: the reference ships no tests.  The only known answers are the two flow plots
HornSchunckOF/img/resimage/0000{40,50}_10.pnghsbresenhamLineFlow.png (w=5, 100 iterations,
alpha=1: main.cpp:94-96,104); plot_bresenham() below restates plotFlow.cpp so the oracle can
be checked against them pixel for pixel (tests/test_oracle_golden.py).  Parity is therefore
"pinned weakly by the golden PNGs" (1197 grid samples per pair at 1/20 px), nothing stronger
exists upstream.
"""
from __future__ import annotations

import numpy as np

try:  # cv2 is optional: np_* needs nothing but NumPy
    import cv2  # type: ignore
except Exception:  # pragma: no cover
    cv2 = None


# --------------------------------------------------------------------------------------
# helpers shared by both restatements
# --------------------------------------------------------------------------------------
def anchor_of(w: int) -> int:
    """hornSchunck.cpp:54 - cv::Point anchor(w-(w/2)-1, w-(w/2)-1) with integer division."""
    return w - (w // 2) - 1


def bgr2gray(img_bgr: np.ndarray) -> np.ndarray:
    """main.cpp:11-26 preprocess(): cv::cvtColor(COLOR_BGR2GRAY) on 8UC3.

    OpenCV's 8-bit path is 15-bit fixed point: Y = (3735 B + 19235 G + 9798 R + 2^14) >> 15.
    Single-channel inputs are passed through (copyTo)."""
    if img_bgr.ndim == 2:
        return img_bgr.copy()
    b = img_bgr[..., 0].astype(np.uint32)
    g = img_bgr[..., 1].astype(np.uint32)
    r = img_bgr[..., 2].astype(np.uint32)
    return ((3735 * b + 19235 * g + 9798 * r + (1 << 14)) >> 15).astype(np.uint8)


# --------------------------------------------------------------------------------------
# restatement 1: through cv2 (OpenCV's own kernels)
# --------------------------------------------------------------------------------------
def cv_gradients(prev_u8: np.ndarray, next_u8: np.ndarray):
    """hornSchunck.cpp:19-41."""
    P = prev_u8.astype(np.float64)                      # :23
    N = next_u8.astype(np.float64)                      # :24
    gx = cv2.Sobel(P, -1, 1, 0, ksize=3)                # :27
    gy = cv2.Sobel(P, -1, 0, 1, ksize=3)                # :28
    gt = N - P                                          # :39
    return gx, gy, gt


def cv_flow(prev_u8, next_u8, w: int, iters: int, alpha: float):
    """hornSchunck.cpp:43-75 through cv2.  Returns (gx, gy, gt, u, v) in float64."""
    gx, gy, gt = cv_gradients(prev_u8, next_u8)         # :46
    u = np.zeros_like(gt)                               # :49
    v = np.zeros_like(gt)                               # :50
    kern = np.ones((w, w), np.float64) / float(w) ** 2  # :53
    a = anchor_of(w)                                    # :54
    for _ in range(iters):                              # :56
        ua = cv2.filter2D(u, -1, kern, anchor=(a, a), delta=0, borderType=cv2.BORDER_CONSTANT)  # :60
        va = cv2.filter2D(v, -1, kern, anchor=(a, a), delta=0, borderType=cv2.BORDER_CONSTANT)  # :61
        num = cv2.multiply(gx, ua) + cv2.multiply(gy, va) + gt                                  # :63-64,68
        den = alpha ** 2 + cv2.multiply(gx, gx) + cv2.multiply(gy, gy)                          # :65-66,68
        c = cv2.divide(num, den)                                                                # :68
        u = ua - cv2.multiply(gx, c)                                                            # :69,72
        v = va - cv2.multiply(gy, c)                                                            # :70,73
    return gx, gy, gt, u, v


# --------------------------------------------------------------------------------------
# restatement 2: NumPy only
# --------------------------------------------------------------------------------------
def _reflect101_pad1(P: np.ndarray) -> np.ndarray:
    """1-pixel BORDER_REFLECT_101 pad (index -1 -> 1, n -> n-2; a length-1 axis maps to 0)."""
    H, W = P.shape
    ys = np.array([1 if H > 1 else 0] + list(range(H)) + [H - 2 if H > 1 else 0])
    xs = np.array([1 if W > 1 else 0] + list(range(W)) + [W - 2 if W > 1 else 0])
    return P[np.ix_(ys, xs)]


def np_gradients(prev_u8, next_u8, dtype=np.float64):
    """hornSchunck.cpp:19-41 without OpenCV: 3x3 Sobel of prev (taps written out), next-prev."""
    P = prev_u8.astype(dtype)
    N = next_u8.astype(dtype)
    H, W = P.shape
    Pp = _reflect101_pad1(P)

    def s(dy, dx):
        return Pp[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]

    two = dtype(2)
    gx = (s(-1, 1) + two * s(0, 1) + s(1, 1)) - (s(-1, -1) + two * s(0, -1) + s(1, -1))
    gy = (s(1, -1) + two * s(1, 0) + s(1, 1)) - (s(-1, -1) + two * s(-1, 0) + s(-1, 1))
    return gx, gy, N - P


def np_box(f: np.ndarray, w: int) -> np.ndarray:
    """filter2D(f, ones(w,w)/w^2, anchor a, BORDER_CONSTANT) :53-54,60-61.

    Every tap is multiplied by fl(1/w^2) and accumulated row-major over the window, which is
    bit-identical to cv2.filter2D in float64 for w <= 7 (OpenCV changes algorithm at >= 50
    taps; then the difference is <= 1 ulp)."""
    dt = f.dtype.type
    a = anchor_of(w)
    H, W = f.shape
    fp = np.zeros((H + w - 1, W + w - 1), f.dtype)
    fp[a:a + H, a:a + W] = f
    acc = np.zeros((H, W), f.dtype)
    kf = dt(1.0) / dt(w * w)
    for dy in range(w):
        for dx in range(w):
            acc += kf * fp[dy:dy + H, dx:dx + W]
    return acc


def np_flow(prev_u8, next_u8, w: int, iters: int, alpha: float, dtype=np.float64):
    """hornSchunck.cpp:43-75 without OpenCV.  Returns (gx, gy, gt, u, v) in `dtype`."""
    dt = np.dtype(dtype).type
    gx, gy, gt = np_gradients(prev_u8, next_u8, dtype)
    u = np.zeros_like(gt)
    v = np.zeros_like(gt)
    den = dt(alpha) ** 2 + gx * gx + gy * gy
    with np.errstate(divide="ignore", invalid="ignore"):      # alpha = 0 is legal upstream (IEEE nan/inf)
        for _ in range(iters):
            ua = np_box(u, w)
            va = np_box(v, w)
            c = (gx * ua + gy * va + gt) / den
            u = ua - gx * c
            v = va - gy * c
    return gx, gy, gt, u, v


# --------------------------------------------------------------------------------------
# oracle of the NON-PARITY "textbook" mode (HS_FLAG_TEXTBOOK).  The reference does not compute this;
# it is the discretisation BASELINE.json's prose describes (Horn & Schunck 1981): gradients over
# the 2x2x2 cube (replicated at the right/bottom border), 1/6 - 1/12 weighted 3x3 average with
# zero padding.  The commented-out two-frame average at hornSchunck.cpp:30-36 hints at it.
# --------------------------------------------------------------------------------------
def np_gradients_textbook(prev_u8, next_u8, dtype=np.float64):
    P = np.pad(prev_u8.astype(dtype), ((0, 1), (0, 1)), mode="edge")
    N = np.pad(next_u8.astype(dtype), ((0, 1), (0, 1)), mode="edge")
    a00, a01, a10, a11 = P[:-1, :-1], P[:-1, 1:], P[1:, :-1], P[1:, 1:]
    c00, c01, c10, c11 = N[:-1, :-1], N[:-1, 1:], N[1:, :-1], N[1:, 1:]
    q = dtype(0.25)
    gx = q * ((a01 - a00) + (a11 - a10) + (c01 - c00) + (c11 - c10))
    gy = q * ((a10 - a00) + (a11 - a01) + (c10 - c00) + (c11 - c01))
    gt = q * ((c00 + c01 + c10 + c11) - (a00 + a01 + a10 + a11))
    return gx, gy, gt


def np_flow_textbook(prev_u8, next_u8, iters: int, alpha: float, dtype=np.float64):
    dt = np.dtype(dtype).type
    gx, gy, gt = np_gradients_textbook(prev_u8, next_u8, dtype)
    u = np.zeros_like(gt)
    v = np.zeros_like(gt)
    den = dt(alpha) ** 2 + gx * gx + gy * gy

    def bar(f):
        p = np.pad(f, 1)
        nsew = p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:]
        diag = p[:-2, :-2] + p[:-2, 2:] + p[2:, :-2] + p[2:, 2:]
        return nsew / dt(6) + diag / dt(12)

    with np.errstate(divide="ignore", invalid="ignore"):
        for _ in range(iters):
            ua, va = bar(u), bar(v)
            c = (gx * ua + gy * va + gt) / den
            u = ua - gx * c
            v = va - gy * c
    return gx, gy, gt, u, v


# --------------------------------------------------------------------------------------
# restatement of plotFlow.cpp - only needed to compare with the reference's golden PNGs
# --------------------------------------------------------------------------------------
def _ctrunc(x: float) -> int:
    return int(x)  # C (int) cast truncates toward zero, so does Python's int()


def plot_bresenham(image_bgr: np.ndarray, u: np.ndarray, v: np.ndarray,
                   delta: int = 20, scale: float = 20.0, outlier: int = 5) -> np.ndarray:
    """plotFlow.cpp:68-88 (call site main.cpp:104: delta 20, scale 20, outlier 5).

    Quirks kept on purpose: the first loop variable walks *rows* and is paired with u
    (:70-73); pixels are only written for 0 <= x < rows-1 and 0 <= y < cols-1 (:25-26);
    the residual starts at the integer half distance (:51)."""
    img = image_bgr.copy()
    rows, cols = img.shape[:2]

    def put(x, y, c0, c1, c2):                      # :24-32
        if 0 <= x < rows - 1 and 0 <= y < cols - 1:
            img[x, y, 0] = c0
            img[x, y, 1] = c1
            img[x, y, 2] = c2

    def sign(t):                                    # :18-22
        return -1 if t < 0 else (1 if t > 0 else 0)

    def line(xs, ys, xe, ye):                       # :43-66
        dx, dy = xe - xs, ye - ys
        sx, sy = sign(dx), sign(dy)
        dx, dy = abs(dx), abs(dy)
        dist = max(dx, dy)
        R = float(dist // 2)
        x, y = xs, ys
        if dx > dy:
            for _ in range(dist):
                put(x, y, 0, 255, 0)
                x += sx; R += dy                    # moveLateral :34-41
                if R >= dx:
                    y += sy; R -= dx
        else:
            for _ in range(dist):
                put(x, y, 0, 255, 0)
                y += sy; R += dx
                if R >= dy:
                    x += sx; R -= dy

    sc = float(np.float32(scale))                   # parameter type is float (:68)
    for x1 in range(0, rows, delta):                # :70
        for y1 in range(0, cols, delta):            # :71
            uu = float(u[x1, y1]); vv = float(v[x1, y1])
            x2 = _ctrunc(x1 + uu * sc)              # :72
            y2 = _ctrunc(y1 + vv * sc)              # :73
            if outlier > 0:                         # :74-78
                if uu < outlier and vv < outlier and uu > -outlier and vv > -outlier:
                    line(x1, y1, x2, y2)
            else:
                line(x1, y1, x2, y2)
            put(x2, y2, 0, 0, 255)                  # :82
    return img


# --------------------------------------------------------------------------------------
# crop oracle: exact interior of a huge solve from a small window
# --------------------------------------------------------------------------------------
def crop_margin(w: int, iters: int) -> int:
    """Influence radius of `iters` sweeps (+1 for the Sobel taps)."""
    return max(anchor_of(w), w // 2) * iters + 1
