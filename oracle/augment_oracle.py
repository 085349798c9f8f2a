"""ORACLE -- test infrastructure only (see oracle/vit_oracle.py header). numpy restatement of the reference's
training augmentation recipe, per sample, in pipeline order ($REF/TIC/ViT/ntrain.py:104-112 [a18]):

  RandomResizedCrop(224) -> RandomHorizontalFlip -> ColorJitter(.2,.2,.2,.1, random order) -> RandomGrayscale(.2)
  -> RandomErasing(.5, value 0) -> ToTensor -> Normalize(ImageNet) -> (here) 16x16 patch rows in bf16.

Parameter sampling follows torchvision 0.26 (`v2/_geometry.py:272-308`, `_transform.py:181`, `v2/_color.py:146-171`,
`v2/_augment.py:100-136`) but draws from a counter-based hash RNG, so the host sampler and this oracle consume the
same numbers. The pixel arithmetic follows torchvision's *tensor* kernels on uint8 images
(`v2/functional/_color.py:31-48,92-97,114-125,151-166,190-205,300-400`, `_misc.py:37-67,264-304`,
`_augment.py:35-42`); the resize is an antialiased (triangle filter) bilinear resample. On PIL inputs the reference
goes through PIL's own integer kernels, which differ by rounding: the bit-exactness contract of the CUDA kernel is
THIS restatement, every float32 operation separately rounded, no fused multiply-add.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)
OUT = 224
MAX_TAPS = 8

# ------------------------------------------------------------------------------------------------------------------
# counter-based RNG: 32 random bits from (seed, sample, draw) -- splitmix64 finaliser
# ------------------------------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def rng_u32(seed: int, sample: int, draw: int) -> int:
    z = (seed * 0x9E3779B97F4A7C15 + sample * 0xBF58476D1CE4E5B9 + draw * 0x94D049BB133111EB + 0x2545F4914F6CDD1D) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    z = z ^ (z >> 31)
    return (z >> 32) & 0xFFFFFFFF


class Stream:
    def __init__(self, seed, sample):
        self.seed, self.sample, self.draw = seed, sample, 0

    def uniform(self, lo=0.0, hi=1.0):
        u = (rng_u32(self.seed, self.sample, self.draw) >> 8) * (1.0 / (1 << 24))
        self.draw += 1
        return lo + (hi - lo) * u

    def randint(self, n):  # [0, n)
        v = rng_u32(self.seed, self.sample, self.draw) % n
        self.draw += 1
        return int(v)


def sample_params(seed: int, sample: int, height: int, width: int, size: int = OUT, recipe: str = "full"):
    """One sample's parameter record: (ints[16], floats[4]).

    ints  = [top, left, h, w, flip, op0, op1, op2, op3, gray, erase_i, erase_j, erase_h, erase_w, jitter_on, 0]
    floats = [brightness, contrast, saturation, hue]
    """
    st = Stream(seed, sample)
    do_crop = do_flip = do_erase = recipe in ("full", "generalization")
    do_jitter = recipe in ("full", "diversity")
    do_gray = recipe in ("full", "diversity", "grey")
    assert recipe in ("full", "generalization", "diversity", "grey", "none")
    area = height * width
    lr0, lr1 = math.log(3.0 / 4.0), math.log(4.0 / 3.0)
    top = left = 0
    h, w = height, width
    for _ in range(10 if do_crop else 0):  # v2/_geometry.py:277-292
        target_area = area * st.uniform(0.08, 1.0)
        aspect = math.exp(st.uniform(lr0, lr1))
        cw = int(round(math.sqrt(target_area * aspect)))
        chh = int(round(math.sqrt(target_area / aspect)))
        if 0 < cw <= width and 0 < chh <= height:
            top = st.randint(height - chh + 1)
            left = st.randint(width - cw + 1)
            h, w = chh, cw
            break
    else:  # central-crop fallback (:293-306); Resize-only recipes keep the whole frame
        in_ratio = float(width) / float(height)
        if not do_crop:
            pass
        elif in_ratio < 3.0 / 4.0:
            w = width
            h = int(round(w / (3.0 / 4.0)))
        elif in_ratio > 4.0 / 3.0:
            h = height
            w = int(round(h * (4.0 / 3.0)))
        else:
            w, h = width, height
        top, left = (height - h) // 2, (width - w) // 2
    flip = (1 if st.uniform() < 0.5 else 0) if do_flip else 0  # _transform.py:181
    perm = [0, 1, 2, 3]                                         # randperm(4): Fisher-Yates on the hash stream
    b = c = s = 1.0
    hue = 0.0
    if do_jitter:
        for i in range(3, 0, -1):
            j = st.randint(i + 1)
            perm[i], perm[j] = perm[j], perm[i]
        b = st.uniform(0.8, 1.2)
        c = st.uniform(0.8, 1.2)
        s = st.uniform(0.8, 1.2)
        hue = st.uniform(-0.1, 0.1)
    gray = (1 if st.uniform() < 0.2 else 0) if do_gray else 0
    ei = ej = eh = ew = 0
    if do_erase and st.uniform() < 0.5:                         # RandomErasing gate
        el0, el1 = math.log(0.3), math.log(3.3)
        for _ in range(10):                                     # v2/_augment.py:113-131
            erase_area = size * size * st.uniform(0.02, 0.33)
            aspect = math.exp(st.uniform(el0, el1))
            hh = int(round(math.sqrt(erase_area * aspect)))
            ww = int(round(math.sqrt(erase_area / aspect)))
            if not (hh < size and ww < size):
                continue
            ei = st.randint(size - hh + 1)
            ej = st.randint(size - ww + 1)
            eh, ew = hh, ww
            break
    jitter_on = 1 if do_jitter else 0
    ints = [top, left, h, w, flip, perm[0], perm[1], perm[2], perm[3], gray, ei, ej, eh, ew, jitter_on, 0]
    floats = [b, c, s, hue]
    return np.array(ints, dtype=np.int32), np.array(floats, dtype=np.float32)


# ------------------------------------------------------------------------------------------------------------------
# pixel arithmetic
# ------------------------------------------------------------------------------------------------------------------
def resample_weights(in_size: int, out_size: int):
    """Antialiased bilinear (triangle) filter taps per output coordinate: (lo[out], n[out], w[out, MAX_TAPS])."""
    scale = F32(in_size) / F32(out_size)
    fs = max(scale, F32(1.0))
    support = fs
    lo = np.zeros(out_size, np.int32)
    n = np.zeros(out_size, np.int32)
    w = np.zeros((out_size, MAX_TAPS), np.float32)
    for o in range(out_size):
        center = (F32(o) + F32(0.5)) * scale
        a = int(center - support + F32(0.5))
        b = int(center + support + F32(0.5))
        a = max(a, 0)
        b = min(b, in_size)
        cnt = b - a
        assert 0 < cnt <= MAX_TAPS, (in_size, out_size, cnt)
        total = F32(0.0)
        raw = np.zeros(MAX_TAPS, np.float32)
        for k in range(cnt):
            x = (F32(a + k) - center + F32(0.5)) / fs
            wk = max(F32(0.0), F32(1.0) - abs(x))
            raw[k] = wk
            total = F32(total + wk)
        for k in range(cnt):
            w[o, k] = F32(raw[k] / total)
        lo[o], n[o] = a, cnt
    return lo, n, w


def resized_crop(img: np.ndarray, top, left, h, w, size=OUT) -> np.ndarray:
    """uint8 HWC -> uint8 [size, size, 3]; horizontal pass then vertical pass in float32, round-half-up."""
    crop = img[top:top + h, left:left + w, :].astype(np.float32)
    lox, nx, wx = resample_weights(w, size)
    loy, ny, wy = resample_weights(h, size)
    tmp = np.zeros((h, size, 3), np.float32)          # horizontal: tmp[y, xo] = sum_t wx[xo,t] * crop[y, lox[xo]+t]
    for t in range(MAX_TAPS):
        idx = np.minimum(lox + t, w - 1)
        wt = np.where(t < nx, wx[:, t], F32(0.0)).astype(np.float32)
        tmp = tmp + crop[:, idx, :] * wt[None, :, None]
    out = np.zeros((size, size, 3), np.float32)       # vertical: out[yo, xo] = sum_t wy[yo,t] * tmp[loy[yo]+t, xo]
    for t in range(MAX_TAPS):
        idx = np.minimum(loy + t, h - 1)
        wt = np.where(t < ny, wy[:, t], F32(0.0)).astype(np.float32)
        out = out + tmp[idx, :, :] * wt[:, None, None]
    return np.clip(np.floor(out + F32(0.5)), 0, 255).astype(np.uint8)


def gray_floor(rgb_f32: np.ndarray) -> np.ndarray:
    """floor(0.2989 r + 0.587 g + 0.114 b) as float32 (integer valued), _color.py:31-48."""
    r, g, b = rgb_f32[..., 0], rgb_f32[..., 1], rgb_f32[..., 2]
    l_img = (r * F32(0.2989) + g * F32(0.587)) + b * F32(0.114)
    return np.floor(l_img).astype(np.float32)


def _to_u8(x: np.ndarray) -> np.ndarray:
    return np.clip(x, F32(0.0), F32(255.0)).astype(np.uint8)  # clamp then truncate (values are >= 0)


def adjust_brightness(img: np.ndarray, b: np.float32) -> np.ndarray:
    return _to_u8(img.astype(np.float32) * F32(b))


def blend(img: np.ndarray, other: np.ndarray, ratio: np.float32) -> np.ndarray:
    ratio = F32(ratio)
    return _to_u8(img.astype(np.float32) * ratio + other * F32(F32(1.0) - ratio))


def adjust_contrast(img: np.ndarray, c: np.float32) -> np.ndarray:
    g = gray_floor(img.astype(np.float32))
    total = int(g.astype(np.int64).sum())                  # exact: every term is an integer <= 255
    mean = F32(F32(total) / F32(g.size))
    return blend(img, np.full(img.shape, mean, np.float32), c)


def adjust_saturation(img: np.ndarray, s: np.float32) -> np.ndarray:
    g = gray_floor(img.astype(np.float32))
    return blend(img, np.repeat(g[..., None], 3, axis=-1), s)


def adjust_hue(img: np.ndarray, hue: np.float32) -> np.ndarray:
    x = img.astype(np.float32) * F32(1.0 / 255.0)
    r, g, b = x[..., 0], x[..., 1], x[..., 2]
    maxc = np.maximum(np.maximum(r, g), b)
    minc = np.minimum(np.minimum(r, g), b)
    eqc = maxc == minc
    cr = maxc - minc
    one = np.ones_like(maxc)
    s = cr / np.where(eqc, one, maxc)
    div = np.where(eqc, one, cr)
    rc, gc, bc = (maxc - r) / div, (maxc - g) / div, (maxc - b) / div
    neq_r = maxc != r
    eq_g = maxc == g
    hg = ((rc + F32(2.0)) - bc) * (eq_g & neq_r).astype(np.float32)
    hr = (bc - gc) * (~neq_r).astype(np.float32)
    hb = ((gc + F32(4.0)) - rc) * (neq_r & ~eq_g).astype(np.float32)
    h = (hr + hg) + hb
    h = np.fmod(h * F32(1.0 / 6.0) + F32(1.0), F32(1.0))
    h = np.fmod(h + F32(hue), F32(1.0))                    # torch.remainder: fmod, then + 1 when negative
    h = np.where(h < 0, h + F32(1.0), h).astype(np.float32)
    v = maxc
    h6 = h * F32(6.0)
    i = np.floor(h6)
    f = h6 - i
    i = i.astype(np.int32) % 6
    sxf = s * f
    oms = F32(1.0) - s
    q = np.clip((F32(1.0) - sxf) * v, F32(0.0), F32(1.0))
    t = np.clip((sxf + oms) * v, F32(0.0), F32(1.0))
    p = np.clip(oms * v, F32(0.0), F32(1.0))
    vpqt = np.stack([v, p, q, t], axis=0)
    select = np.array([[0, 2, 1, 1, 3, 0], [3, 0, 0, 2, 1, 1], [1, 1, 3, 0, 0, 2]])
    out = np.stack([np.take_along_axis(vpqt, select[c][i][None], axis=0)[0] for c in range(3)], axis=-1)
    return (out.astype(np.float32) * F32(255.999)).astype(np.uint8)   # _misc.py:296-300 (float -> uint8, truncating)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> bfloat16 bit patterns (uint16)."""
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    rounded = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return rounded.astype(np.uint16)


def augment_one(img: np.ndarray, ints: np.ndarray, floats: np.ndarray, size: int = OUT):
    """uint8 HWC image + parameter record -> (uint8 [size,size,3] after erase, bf16 bits [ (size/16)^2, 768 ])."""
    top, left, h, w, flip = (int(v) for v in ints[:5])
    out = resized_crop(img, top, left, h, w, size)
    if flip:
        out = out[:, ::-1, :].copy()
    if ints[14]:
        for op in ints[5:9]:
            if op == 0:
                out = adjust_brightness(out, floats[0])
            elif op == 1:
                out = adjust_contrast(out, floats[1])
            elif op == 2:
                out = adjust_saturation(out, floats[2])
            else:
                out = adjust_hue(out, floats[3])
    if ints[9]:
        g = gray_floor(out.astype(np.float32)).astype(np.uint8)
        out = np.repeat(g[..., None], 3, axis=-1)
    ei, ej, eh, ew = (int(v) for v in ints[10:14])
    if eh > 0 and ew > 0:
        out = out.copy()
        out[ei:ei + eh, ej:ej + ew, :] = 0
    x = out.astype(np.float32) / F32(255.0)                 # ToTensor
    x = (x - MEAN[None, None, :]) / STD[None, None, :]      # Normalize
    G = size // 16
    # patch rows: [gy, gx] -> row gy*G+gx ; columns ordered (c, py, px) like Conv2d weight.view(D, 768)
    patches = x.reshape(G, 16, G, 16, 3).transpose(0, 2, 4, 1, 3).reshape(G * G, 768)
    return out, f32_to_bf16_bits(patches)


def augment_batch(images: np.ndarray, seed: int, first_sample: int = 0, size: int = OUT, recipe: str = "full"):
    B, H, W, _ = images.shape
    G = size // 16
    pix = np.zeros((B, size, size, 3), np.uint8)
    tok = np.zeros((B * G * G, 768), np.uint16)
    for b in range(B):
        ints, floats = sample_params(seed, first_sample + b, H, W, size, recipe)
        pix[b], tok[b * G * G:(b + 1) * G * G] = augment_one(images[b], ints, floats, size)
    return pix, tok
