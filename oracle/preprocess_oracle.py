"""TEST INFRASTRUCTURE - CPU restatement (numpy) of the image pre-processing in front of the detector.

Reference path: okvis_ros/src/Subscriber.cpp:123-157 (Subscriber::imageCallback):
    cv::resize(raw, resizeFactor) -> [cv::medianBlur(3)] -> [CLAHE | cv::equalizeHist] -> VioInterface::addImage
The arithmetic lives in OpenCV (third-party, absent from /root/reference): restated here from OpenCV 4.x
(modules/imgproc/src/{resize,median_blur,histogram,clahe}.cpp) and PINNED against the cv2 4.13 wheel of this image by
tests/golden/make_preprocess_golden.py -> tests/golden/preprocess_golden.npz (IPP disabled: the plain C++ paths).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np

F32 = np.float32


def _cv_round(x):
    """cvRound / saturate_cast from floating point: round half to even."""
    return np.rint(x)


def resize_linear(src: np.ndarray, factor: float) -> np.ndarray:
    """cv::resize(src, dst, Size(), factor, factor) with the default INTER_LINEAR, 8UC1 (resize.cpp)."""
    h, w = src.shape
    dw, dh = int(_cv_round(w * factor)), int(_cv_round(h * factor))
    scale_x, scale_y = 1.0 / factor, 1.0 / factor        # double inv_scale -> scale
    isx, isy = int(np.floor(scale_x + 0.5)), int(np.floor(scale_y + 0.5))   # saturate_cast<int>(scale)
    eps = np.finfo(np.float64).eps
    area_fast = abs(scale_x - isx) < eps and abs(scale_y - isy) < eps
    if area_fast and isx == 2 and isy == 2:
        # INTER_LINEAR is replaced by the INTER_AREA fast path for an exact 2x decimation (ResizeAreaFastVec)
        # resizeAreaFast_Invoker: full 2x2 blocks (a+b+c+d+2)>>2; blocks cut by the border (odd source sizes)
        # average the pixels that exist in float and round half to even
        p = np.zeros((2 * dh, 2 * dw), dtype=np.int32)
        m = np.zeros((2 * dh, 2 * dw), dtype=np.int32)
        hh, ww = min(h, 2 * dh), min(w, 2 * dw)
        p[:hh, :ww] = src[:hh, :ww]
        m[:hh, :ww] = 1
        ssum = p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2]
        cnt = m[0::2, 0::2] + m[0::2, 1::2] + m[1::2, 0::2] + m[1::2, 1::2]
        full = (ssum + 2) >> 2
        part = _cv_round(ssum.astype(F32) / np.maximum(cnt, 1).astype(F32)).astype(np.int32)
        return np.where(cnt == 4, full, np.where(cnt > 0, part, 0)).astype(np.uint8)

    def taps(dn, sn, scale, clamp_weights):
        d = np.arange(dn, dtype=np.float64)
        f = ((d + 0.5) * scale - 0.5).astype(F32)
        s = np.floor(f).astype(np.int64)
        f = (f - s.astype(F32)).astype(F32)
        if clamp_weights:   # columns: the tap is moved onto the border pixel (resize.cpp, xofs / alpha loop)
            lo = s < 0
            f[lo], s[lo] = 0, 0
            hi = s >= sn - 1
            f[hi], s[hi] = 0, sn - 1
        a0 = np.clip(_cv_round((F32(1.0) - f) * F32(2048)), -32768, 32767).astype(np.int64)
        a1 = np.clip(_cv_round(f * F32(2048)), -32768, 32767).astype(np.int64)
        return s, a0, a1

    sx, ax0, ax1 = taps(dw, w, scale_x, True)
    sy, by0, by1 = taps(dh, h, scale_y, False)   # rows: weights kept, the two row indices are clipped separately
    S = src.astype(np.int64)
    sx1 = np.minimum(sx + 1, w - 1)
    sy0 = np.clip(sy, 0, h - 1)
    sy1 = np.clip(sy + 1, 0, h - 1)
    H = S[:, sx] * ax0[None, :] + S[:, sx1] * ax1[None, :]        # HResizeLinear: rows in int, scale 2^11
    r0, r1 = H[sy0, :], H[sy1, :]
    out = (((by0[:, None] * (r0 >> 4)) >> 16) + ((by1[:, None] * (r1 >> 4)) >> 16) + 2) >> 2   # VResizeLinear<uchar>
    return np.clip(out, 0, 255).astype(np.uint8)


def median3(src: np.ndarray) -> np.ndarray:
    """cv::medianBlur(src, dst, 3): 3x3 median, BORDER_REPLICATE (median_blur.cpp)."""
    p = np.pad(src, 1, mode="edge")
    h, w = src.shape
    stack = np.stack([p[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)], axis=0)
    return np.sort(stack, axis=0)[4].astype(np.uint8)


def equalize_hist(src: np.ndarray) -> np.ndarray:
    """cv::equalizeHist (histogram.cpp:3386-3440)."""
    hist = np.bincount(src.ravel(), minlength=256).astype(np.int64)
    total = src.size
    i = int(np.nonzero(hist)[0][0])
    if hist[i] == total:
        return np.full_like(src, i)
    scale = F32(255.0) / F32(total - hist[i])
    lut = np.zeros(256, dtype=np.uint8)
    csum = np.cumsum(hist[i + 1:])
    lut[i + 1:] = np.clip(_cv_round(csum.astype(F32) * scale), 0, 255).astype(np.uint8)
    return lut[src]


def clahe_luts(src: np.ndarray, clip_limit: float, tiles: int):
    """Per-tile LUTs of CLAHE_CalcLut_Body<uchar,256,0> (clahe.cpp); returns (luts [ty][tx][256], padded image)."""
    h, w = src.shape
    if w % tiles == 0 and h % tiles == 0:
        img = src
    else:
        img = np.pad(src, ((0, tiles - h % tiles), (0, tiles - w % tiles)), mode="reflect")  # BORDER_REFLECT_101
    th, tw = img.shape[0] // tiles, img.shape[1] // tiles
    area = th * tw
    lut_scale = F32(255.0) / F32(area)
    clip = 0
    if clip_limit > 0.0:
        clip = max(int(clip_limit * area / 256), 1)
    luts = np.zeros((tiles, tiles, 256), dtype=np.uint8)
    for ty in range(tiles):
        for tx in range(tiles):
            t = img[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
            hist = np.bincount(t.ravel(), minlength=256).astype(np.int64)
            if clip > 0:
                clipped = int(np.maximum(hist - clip, 0).sum())
                hist = np.minimum(hist, clip)
                batch = clipped // 256
                residual = clipped - batch * 256
                hist += batch
                if residual != 0:
                    step = max(256 // residual, 1)
                    idx = np.arange(0, 256, step)[:residual]
                    hist[idx] += 1
            csum = np.cumsum(hist)
            luts[ty, tx] = np.clip(_cv_round(csum.astype(F32) * lut_scale), 0, 255).astype(np.uint8)
    return luts, img


def clahe(src: np.ndarray, clip_limit: float, tiles: int) -> np.ndarray:
    """cv::CLAHE::apply, 8UC1 (clahe.cpp: CLAHE_Interpolation_Body<uchar,0>)."""
    luts, img = clahe_luts(src, clip_limit, tiles)
    h, w = src.shape
    th, tw = img.shape[0] // tiles, img.shape[1] // tiles
    inv_tw, inv_th = F32(1.0) / F32(tw), F32(1.0) / F32(th)

    def table(n, inv, ntiles):
        f = (np.arange(n, dtype=F32) * inv - F32(0.5)).astype(F32)
        i1 = np.floor(f).astype(np.int64)
        a = (f - i1.astype(F32)).astype(F32)
        i2 = np.minimum(i1 + 1, ntiles - 1)
        i1 = np.maximum(i1, 0)
        return i1, i2, a, (F32(1.0) - a).astype(F32)

    tx1, tx2, xa, xa1 = table(w, inv_tw, tiles)
    ty1, ty2, ya, ya1 = table(h, inv_th, tiles)
    v = src.astype(np.int64)
    l11 = luts[ty1[:, None], tx1[None, :], v].astype(F32)
    l12 = luts[ty1[:, None], tx2[None, :], v].astype(F32)
    l21 = luts[ty2[:, None], tx1[None, :], v].astype(F32)
    l22 = luts[ty2[:, None], tx2[None, :], v].astype(F32)
    top = (l11 * xa1[None, :]).astype(F32) + (l12 * xa[None, :]).astype(F32)
    bot = (l21 * xa1[None, :]).astype(F32) + (l22 * xa[None, :]).astype(F32)
    res = (top.astype(F32) * ya1[:, None]).astype(F32) + (bot.astype(F32) * ya[:, None]).astype(F32)
    return np.clip(_cv_round(res.astype(F32)), 0, 255).astype(np.uint8)


HIST_NONE, HIST_EQUALIZE, HIST_CLAHE = 0, 1, 2


def preprocess(raw: np.ndarray, resize_factor=1.0, median=False, method=HIST_NONE, clip_limit=1.0, tiles=4):
    """Subscriber::imageCallback's image chain (Subscriber.cpp:123-147)."""
    img = resize_linear(raw, resize_factor) if resize_factor != 1.0 else raw.copy()
    if median:
        img = median3(img)
    if method == HIST_CLAHE:
        img = clahe(img, clip_limit, tiles)
    elif method == HIST_EQUALIZE:
        img = equalize_hist(img)
    return img
