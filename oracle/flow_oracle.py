"""CPU oracle for ``video_to_flow`` (lib/utils.py:94-129) -- TEST INFRASTRUCTURE ONLY (see vfd_oracle.py).

The reference computes dense optical flow on the host with OpenCV and encodes it as an RGB video:

    per frame index d: normalize(video[:, :, d]) over the whole batch            lib/utils.py:81-89,96
    per clip, per frame: cv2.cvtColor(RGB2GRAY)                                  :108
    cv2.calcOpticalFlowFarneback(prev, next, None, 0.5, 3, 15, 3, 5, 1.2, 0)     :115-116
    cv2.cartToPolar(angleInDegrees=True); H = ang / 2; S = 255;
    V = cv2.normalize(mag, None, 0, 255, NORM_MINMAX)                            :117-119
    cv2.cvtColor(HSV2RGB) on the float32 image, np.uint8(), last frame repeated  :121-126
    ClipToTensor (/255), * 2 - 1                                                 :127-129

Where the arithmetic lives: OpenCV (requirements.txt:24 pins opencv-python==4.2.0.32; absent from
/root/reference; 4.13 in this image). This file restates the published algorithms of the calls above in numpy:
Farneback's polynomial-expansion flow as implemented in modules/video/src/optflowgf.cpp (Gaussian pyramid,
``FarnebackPolyExp``, ``FarnebackUpdateMatrices``, ``FarnebackUpdateFlow_Blur``), ``getGaussianKernel``,
bilinear ``resize``, ``fastAtan2``, ``normalize(NORM_MINMAX)``, the float ``HSV2RGB`` conversion, and numpy's
float -> uint8 cast on x86 (truncate to int32, keep the low byte). Note what the reference's inputs do to the
algorithm: the frames are scaled to [0, 1], so the structure-tensor determinant is ~1e-8 against the 1e-3
regulariser in ``1 / (g11 * g22 - g12^2 + 1e-3)`` and the "flow" is a ~1e-6-pixel, heavily damped field; only its
direction and its per-image min-max normalised magnitude reach the output.

Pinning: against cv2.calcOpticalFlowFarneback itself (relative 3e-7 on the flow field) and against the reference's
own ``video_to_flow`` (> 99 % of the output bytes identical, the rest off by one grey level: float rounding at the
uint8 truncation) -- tests/test_flow_oracle.py, fixtures tests/golden/flow_small.pt.
"""
import numpy as np

F32 = np.float32


def cv_round(x):
    return int(np.rint(x))                      # cvRound: round half to even


def gaussian_kernel(ksize, sigma):
    """cv::getGaussianKernel: fixed table for sigma <= 0 and ksize <= 7, else normalised exp."""
    small = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
             7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125]}
    if sigma <= 0 and ksize in small:
        return np.array(small[ksize], F32)
    s = sigma if sigma > 0 else ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8
    x = np.arange(ksize) - (ksize - 1) * 0.5
    k = np.exp(-(x * x) / (2 * s * s))
    return (k / k.sum()).astype(F32)


def _reflect101(i, n):
    if n == 1:
        return 0
    while i < 0 or i >= n:
        i = -i if i < 0 else 2 * (n - 1) - i
    return i


def gaussian_blur(img, ksize, sigma):
    """cv::GaussianBlur, separable, BORDER_REFLECT_101, float accumulation: rows first, then columns."""
    k = gaussian_kernel(ksize, sigma)
    r = ksize // 2
    h, w = img.shape
    ix = np.array([[_reflect101(i + j - r, w) for j in range(ksize)] for i in range(w)])
    tmp = np.zeros_like(img)
    for j in range(ksize):
        tmp += k[j] * img[:, ix[:, j]]
    iy = np.array([[_reflect101(i + j - r, h) for j in range(ksize)] for i in range(h)])
    out = np.zeros_like(img)
    for j in range(ksize):
        out += k[j] * tmp[iy[:, j], :]
    return out


def resize_linear(img, w, h):
    """cv::resize(INTER_LINEAR): half-pixel centres, edge clamp (the exact 2x down-scale is a 2x2 mean)."""
    H, W = img.shape[:2]
    if (w, h) == (W, H):
        return img.copy()

    def coords(n_dst, n_src):
        f = (np.arange(n_dst) + 0.5) * (n_src / n_dst) - 0.5
        i0 = np.floor(f).astype(int)
        a = (f - i0).astype(F32)
        a = np.where((i0 < 0) | (i0 >= n_src - 1), 0, a).astype(F32)
        return np.clip(i0, 0, n_src - 1), np.clip(i0 + 1, 0, n_src - 1), a

    x0, x1, ax = coords(w, W)
    y0, y1, ay = coords(h, H)
    img = img.astype(F32)
    if img.ndim == 3:
        ax, ay = ax[None, :, None], ay[:, None, None]
    else:
        ax, ay = ax[None, :], ay[:, None]
    top = img[y0][:, x0] * (1 - ax) + img[y0][:, x1] * ax
    bot = img[y1][:, x0] * (1 - ax) + img[y1][:, x1] * ax
    return (top * (1 - ay) + bot * ay).astype(F32)


def prepare_gaussian(n, sigma):
    """FarnebackPrepareGaussian: the 1-D kernels g, x*g, x^2*g and four entries of inv(G)."""
    if sigma < 1.1920929e-07:
        sigma = n * 0.3
    x = np.arange(-n, n + 1)
    g = np.exp(-x * x / (2 * sigma * sigma)).astype(F32)
    g = (g * (1.0 / float(g.astype(np.float64).sum()))).astype(F32)
    xg, xxg = (x * g).astype(F32), (x * x * g).astype(F32)
    G = np.zeros((6, 6))
    for y in range(-n, n + 1):
        for xx in range(-n, n + 1):
            gg = float(g[y + n]) * float(g[xx + n])
            G[0, 0] += gg
            G[1, 1] += gg * xx * xx
            G[3, 3] += gg * xx ** 4
            G[5, 5] += gg * xx * xx * y * y
    G[2, 2] = G[0, 3] = G[0, 4] = G[3, 0] = G[4, 0] = G[1, 1]
    G[4, 4] = G[3, 3]
    G[3, 4] = G[4, 3] = G[5, 5]
    inv = np.linalg.inv(G)
    return g, xg, xxg, inv[1, 1], inv[0, 3], inv[3, 3], inv[5, 5]


def poly_exp(src, n=5, sigma=1.2):
    """FarnebackPolyExp: float vertical pass (rows clamped), double horizontal pass (columns clamped)."""
    g, xg, xxg, ig11, ig03, ig33, ig55 = prepare_gaussian(n, sigma)
    h, w = src.shape
    src = src.astype(F32)
    ys, xs = np.arange(h), np.arange(w)
    row0, row1, row2 = src * g[n], np.zeros_like(src), np.zeros_like(src)
    for k in range(1, n + 1):
        s0, s1 = src[np.maximum(ys - k, 0)], src[np.minimum(ys + k, h - 1)]
        p = s0 + s1
        row0 = row0 + g[n + k] * p
        row1 = row1 + xg[n + k] * (s1 - s0)
        row2 = row2 + xxg[n + k] * p

    def at(r, dx):
        return r[:, np.clip(xs + dx, 0, w - 1)]

    # C semantics of the reference loop: float (op) float is evaluated in float and only then widened to the
    # double accumulators; only ``tg`` (a double holding a float sum) multiplies in double
    d = np.float64
    b1, b3, b5 = (row0 * g[n]).astype(d), (row1 * g[n]).astype(d), (row2 * g[n]).astype(d)
    b2 = b4 = b6 = 0.0
    for k in range(1, n + 1):
        tg = (at(row0, k) + at(row0, -k)).astype(d)
        b1 = b1 + tg * d(g[n + k])
        b4 = b4 + tg * d(xxg[n + k])
        b2 = b2 + ((at(row0, k) - at(row0, -k)) * xg[n + k]).astype(d)
        b3 = b3 + ((at(row1, k) + at(row1, -k)) * g[n + k]).astype(d)
        b6 = b6 + ((at(row1, k) - at(row1, -k)) * xg[n + k]).astype(d)
        b5 = b5 + ((at(row2, k) + at(row2, -k)) * g[n + k]).astype(d)
    R = np.empty((h, w, 5), F32)
    R[..., 0], R[..., 1] = b3 * ig11, b2 * ig11
    R[..., 2], R[..., 3], R[..., 4] = b1 * ig03 + b5 * ig33, b1 * ig03 + b4 * ig33, b6 * ig55
    return R


BORDER = np.array([0.14, 0.14, 0.4472, 0.4472, 0.4472], F32)


def update_matrices(R0, R1, flow):
    """FarnebackUpdateMatrices: bilinear sample of R1 at x + flow, averaged with R0, border damping."""
    h, w = flow.shape[:2]
    ys, xs = np.mgrid[0:h, 0:w]
    dx, dy = flow[..., 0], flow[..., 1]
    fx, fy = (xs + dx).astype(F32), (ys + dy).astype(F32)
    x1, y1 = np.floor(fx).astype(int), np.floor(fy).astype(int)
    fx, fy = (fx - x1).astype(F32), (fy - y1).astype(F32)
    inside = (x1 >= 0) & (x1 < w - 1) & (y1 >= 0) & (y1 < h - 1)
    xc, yc = np.clip(x1, 0, w - 2), np.clip(y1, 0, h - 2)
    a00, a01, a10, a11 = (1 - fx) * (1 - fy), fx * (1 - fy), (1 - fx) * fy, fx * fy

    def interp(c):
        return a00 * R1[yc, xc, c] + a01 * R1[yc, xc + 1, c] + a10 * R1[yc + 1, xc, c] + a11 * R1[yc + 1, xc + 1, c]

    r2 = np.where(inside, interp(0), 0).astype(F32)
    r3 = np.where(inside, interp(1), 0).astype(F32)
    r4 = np.where(inside, (R0[..., 2] + interp(2)) * 0.5, R0[..., 2]).astype(F32)
    r5 = np.where(inside, (R0[..., 3] + interp(3)) * 0.5, R0[..., 3]).astype(F32)
    r6 = np.where(inside, (R0[..., 4] + interp(4)) * 0.25, R0[..., 4] * 0.5).astype(F32)
    r2, r3 = (R0[..., 0] - r2) * 0.5, (R0[..., 1] - r3) * 0.5
    r2, r3 = r2 + r4 * dy + r6 * dx, r3 + r6 * dy + r5 * dx
    sx, sy = np.ones(w, F32), np.ones(h, F32)
    for i in range(5):
        for s, n in ((sx, w), (sy, h)):
            if i < n:
                s[i] *= BORDER[i]
            if n - 1 - i >= 0:
                s[n - 1 - i] *= BORDER[i]
    scale = sx[None, :] * sy[:, None]       # (x < B ? border[x] : 1) * (x >= w - B ? ... : 1) * (y ...) * (y ...)
    r2, r3, r4, r5, r6 = [(r * scale).astype(F32) for r in (r2, r3, r4, r5, r6)]
    M = np.empty((h, w, 5), F32)
    M[..., 0], M[..., 1], M[..., 2] = r4 * r4 + r6 * r6, (r4 + r5) * r6, r5 * r5 + r6 * r6
    M[..., 3], M[..., 4] = r4 * r2 + r6 * r3, r6 * r2 + r5 * r3
    return M


def box_blur_solve(M, block=15):
    """FarnebackUpdateFlow_Blur without the in-place matrix update: (2m+1)^2 box sums with replicated borders
    (the reference keeps running sums in double) and the 2x2 solve with the 1e-3 regulariser."""
    m = block // 2
    h, w = M.shape[:2]
    ys, xs = np.arange(h), np.arange(w)
    Md = M.astype(np.float64)
    v = np.zeros_like(Md)
    for k in range(-m, m + 1):
        v += Md[np.clip(ys + k, 0, h - 1)]
    b = np.zeros_like(Md)
    for k in range(-m, m + 1):
        b += v[:, np.clip(xs + k, 0, w - 1)]
    b *= 1.0 / (block * block)
    g11, g12, g22, h1, h2 = [b[..., i] for i in range(5)]
    idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3)
    flow = np.empty((h, w, 2), F32)
    flow[..., 0] = (g11 * h2 - g12 * h1) * idet
    flow[..., 1] = (g22 * h1 - g12 * h2) * idet
    return flow


def pyramid_levels(H, W, levels=3, pyr_scale=0.5, min_size=32):
    k, scale = 0, 1.0
    while k < levels:
        scale *= pyr_scale
        if W * scale < min_size or H * scale < min_size:
            break
        k += 1
    return k


def farneback(prev, nxt, pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2):
    """cv2.calcOpticalFlowFarneback(prev, next, None, 0.5, 3, 15, 3, 5, 1.2, 0) for float32 single-channel images."""
    H, W = prev.shape
    prev_flow = None
    for k in range(pyramid_levels(H, W, levels, pyr_scale), -1, -1):
        scale = pyr_scale ** k
        sigma = (1.0 / scale - 1) * 0.5
        smooth = max(cv_round(sigma * 5) | 1, 3)
        w, h = cv_round(W * scale), cv_round(H * scale)
        if prev_flow is None:
            flow = np.zeros((h, w, 2), F32)
        else:
            flow = resize_linear(prev_flow, w, h) * F32(1.0 / pyr_scale)
        R = [poly_exp(resize_linear(gaussian_blur(img.astype(F32), smooth, sigma), w, h), poly_n, poly_sigma)
             for img in (prev, nxt)]
        M = update_matrices(R[0], R[1], flow)
        for i in range(iterations):
            flow = box_blur_solve(M, winsize)
            if i < iterations - 1:
                M = update_matrices(R[0], R[1], flow)
        prev_flow = flow
    return prev_flow


def fast_atan2_deg(y, x):
    """cv::fastAtan2 (the polynomial cartToPolar uses), degrees in [0, 360)."""
    k = 180.0 / np.pi
    p1, p3, p5, p7 = F32(0.9997878412794807 * k), F32(-0.3258083974640975 * k), F32(0.1555786518463281 * k), \
        F32(-0.04432655554792128 * k)
    ax, ay = np.abs(x), np.abs(y)
    eps = F32(2.220446049250313e-16)
    m = ax >= ay
    c = np.where(m, ay / (ax + eps), ax / (ay + eps)).astype(F32)
    c2 = c * c
    poly = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c
    a = np.where(m, poly, F32(90) - poly).astype(F32)
    a = np.where(x < 0, F32(180) - a, a)
    a = np.where(y < 0, F32(360) - a, a)
    return a.astype(F32)


def hsv2rgb_float(h, s, v):
    """cv::cvtColor(COLOR_HSV2RGB) on CV_32F (hue in degrees; no clamping of s or v)."""
    hh = h * F32(6.0 / 360.0)
    hh = np.where(hh < 0, hh + 6 * np.ceil(-hh / 6), hh)
    hh = np.where(hh >= 6, hh - 6 * np.floor(hh / 6), hh).astype(F32)
    sector = np.floor(hh).astype(int)
    f = (hh - sector).astype(F32)
    bad = (sector < 0) | (sector >= 6)
    sector, f = np.where(bad, 0, sector), np.where(bad, 0, f).astype(F32)
    tab = np.stack([v, v * (1 - s), v * (1 - s * f), v * (1 - s * (1 - f))], -1).astype(F32)
    sd = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])

    def take(idx):
        return np.take_along_axis(tab, idx[..., None], -1)[..., 0]

    b, g, r = take(sd[sector, 0]), take(sd[sector, 1]), take(sd[sector, 2])
    grey = s == 0
    return np.stack([np.where(grey, v, r), np.where(grey, v, g), np.where(grey, v, b)], -1).astype(F32)


def gray_frames(video):
    """(B, 3, D, H, W) in [-1, 1] -> (B, D, H, W) grey frames in [0, 1]: per-frame-index normalize over the batch
    (lib/utils.py:81-89,96) and cv2.cvtColor(RGB2GRAY) on float32."""
    v = np.asarray(video, dtype=F32)
    B, C, D, H, W = v.shape
    out = np.empty((B, D, H, W), F32)
    for d in range(D):
        fr = v[:, :, d]
        mn, mx = float(fr.min()), float(fr.max())
        n = (fr + F32(-mn)) / F32(mx - mn + 1e-5)
        out[:, d] = (n[:, 0] * F32(0.299) + n[:, 1] * F32(0.587) + n[:, 2] * F32(0.114)).astype(F32)
    return out


def encode_flow(flow):
    """(H, W, 2) flow -> (H, W, 3) uint8 RGB exactly as lib/utils.py:117-123 does it."""
    fx, fy = flow[..., 0], flow[..., 1]
    mag = np.sqrt(fx * fx + fy * fy).astype(F32)
    ang = fast_atan2_deg(fy, fx)
    mn, mx = mag.min(), mag.max()
    scale = F32(255.0) / (mx - mn) if mx > mn else F32(0)
    val = ((mag - mn) * scale).astype(F32)
    rgb = hsv2rgb_float(ang / 2, np.full_like(val, 255), val)
    return rgb.astype(np.int32).astype(np.uint8)          # np.uint8(float32) on x86: truncate, keep the low byte


def video_to_flow(video):
    """lib/utils.py:94-129 -> float32 (B, 3, D, H, W) in [-1, 1]; also returns the raw flow fields
    (B, D-1, H, W, 2) for tests."""
    grey = gray_frames(video)
    B, D, H, W = grey.shape
    out = np.empty((B, 3, D, H, W), F32)
    flows = np.empty((B, D - 1, H, W, 2), F32)
    for b in range(B):
        for i in range(1, D):
            flows[b, i - 1] = farneback(grey[b, i - 1], grey[b, i])
            rgb8 = encode_flow(flows[b, i - 1])
            out[b, :, i - 1] = rgb8.transpose(2, 0, 1).astype(F32) / F32(255)
        out[b, :, D - 1] = out[b, :, D - 2]
    return out * F32(2) - F32(1), flows
