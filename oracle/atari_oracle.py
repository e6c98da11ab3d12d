"""CPU restatement of the reference's Atari observation pipeline -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module; the product path
(border_b200/csrc/atari.cu behind bb_atari_*) never does.

What it follows (border-atari-env/src/env.rs):
  skip_and_max        :126-147   element-wise max of the frames rendered at repeat 2 and 3 (RGB24, h x w x 3)
  clip_reward         :149-159   sign(r) in training (0 stays 0), r in evaluation
  warp_and_grayscale  :161-185   image::imageops::resize(.., 84, 84, FilterType::Triangle) on the RGB frame, then
                                 gray = (0.299*c2 + 0.587*c1 + 0.114*c0) as u8 (the source names the channels b, g, r in
                                 buffer order, so the 0.299 weight lands on the THIRD byte)
  stack_frame         :187-199   frames[1..4] = frames[0..3]; frames[0] = new  (newest first)
  reset               :263-300   all four frames = the warped first render

PARITY UNPINNED for the resize: `image` is a crates.io dependency (workspace Cargo.toml:50, image = "0.23.14") that is not
vendored under /root/reference, and the reference holds no golden frames.  `resize_triangle_u8` restates the published
algorithm of image 0.23.14's imageops/sample.rs -- resize() = vertical_sample (height) then horizontal_sample (width), each
pass: centre = (out + 0.5) * ratio, taps [floor(centre - support), ceil(centre + support)) clamped to the image, weights
triangle((i - (centre - 0.5)) / sratio) normalised by their sum, f32 accumulation in tap order, clamp to [0, 255], round half
away from zero (FloatNearest) back to u8 BETWEEN the two passes (the f32 intermediate image arrived in 0.24).  The device
kernel is tested bit-for-bit against this restatement; the restatement itself has no reference output to be checked against.
"""
import numpy as np

F = np.float32


def _taps(n_in, n_out):
    """(left, weights) per output index, exactly the f32 arithmetic of sample.rs."""
    ratio = F(n_in) / F(n_out)
    sratio = ratio if ratio >= F(1.0) else F(1.0)
    support = F(1.0) * sratio  # Triangle: support 1.0
    out = []
    for o in range(n_out):
        centre = (F(o) + F(0.5)) * ratio
        left = int(np.floor(centre - support))
        left = min(max(left, 0), n_in - 1)
        right = int(np.ceil(centre + support))
        right = min(max(right, left + 1), n_in)
        c = centre - F(0.5)
        ws = []
        s = F(0.0)
        for i in range(left, right):
            x = (F(i) - c) / sratio
            ax = np.abs(x)
            w = F(1.0) - ax if ax < F(1.0) else F(0.0)
            ws.append(F(w))
            s = F(s + w)
        ws = [F(w / s) for w in ws]
        out.append((left, ws))
    return out


def _round_u8(t):
    t = np.clip(t, F(0.0), F(255.0))
    # f32::round: half away from zero (all values are >= 0 here)
    return np.floor(t + F(0.5)).astype(np.uint8)


def _sample_axis0(img, n_out):
    """One pass along axis 0 of a [n][m][c] u8 image."""
    n_in = img.shape[0]
    out = np.empty((n_out,) + img.shape[1:], np.uint8)
    src = img.astype(np.float32)
    for o, (left, ws) in enumerate(_taps(n_in, n_out)):
        t = np.zeros(img.shape[1:], np.float32)
        for i, w in enumerate(ws):
            t = (t + src[left + i] * w).astype(np.float32)  # t += v * w, separate f32 multiply and add
        out[o] = _round_u8(t)
    return out


def resize_triangle_u8(img, new_w=84, new_h=84):
    """image::imageops::resize(img, new_w, new_h, Triangle) for an [h][w][c] u8 image (image 0.23.14)."""
    tmp = _sample_axis0(img, new_h)                                   # vertical_sample
    return _sample_axis0(tmp.transpose(1, 0, 2), new_w).transpose(1, 0, 2)  # horizontal_sample


def warp_and_grayscale(rgb):
    """env.rs:161-185 on an [h][w][3] u8 frame -> [84][84] u8."""
    small = resize_triangle_u8(rgb, 84, 84).astype(np.float32)
    g = (F(0.299) * small[..., 2] + F(0.587) * small[..., 1]).astype(np.float32)
    g = (g + F(0.114) * small[..., 0]).astype(np.float32)
    return np.clip(np.trunc(g), 0, 255).astype(np.uint8)  # `as u8`: truncation, saturating


def max_pool(a, b):
    return np.maximum(a, b)  # env.rs:140-145


def clip_reward(r, train):
    if not train:
        return float(r)
    return 0.0 if r == 0.0 else float(np.sign(r))


class FrameStack:
    """frames: [4][84][84] u8, newest first (env.rs:187-199, :288-297)."""

    def __init__(self):
        self.frames = np.zeros((4, 84, 84), np.uint8)

    def reset(self, rgb):
        self.frames[:] = warp_and_grayscale(rgb)[None]
        return self.frames.copy()

    def step(self, rgb_a, rgb_b):
        new = warp_and_grayscale(max_pool(rgb_a, rgb_b))
        self.frames[1:] = self.frames[:3].copy()
        self.frames[0] = new
        return self.frames.copy()
