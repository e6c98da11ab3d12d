"""Atari observation pipeline on the device (border-atari-env/src/env.rs:126-199, 263-300).

`AtariPreprocessor` mirrors what `BorderAtariEnv::reset / step` do to the emulator's RGB frames -- max of the last two
repeated frames, 84x84 Triangle resize, grey, newest-first stack of four, reward sign-clip -- and leaves the
[4][84][84] u8 observation in HBM, where `Agent.actor_step_dev` (policy + replay push) consumes it.  The emulator
(atari-env-sys) is out of scope: callers hand in its `render_rgb24` buffers.
"""
import ctypes as C

import numpy as np

from . import _lib as L


class AtariPreprocessor:
    def __init__(self, width=160, height=210, train=True, device=0):
        self._h = C.c_void_p()
        self.shape = (height, width, 3)
        L.check(L.lib().bb_atari_create(device, width, height, 1 if train else 0, C.byref(self._h)))

    def close(self):
        if self._h:
            L.lib().bb_atari_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        L.check(L.lib().bb_atari_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def _frame(self, rgb):
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        assert rgb.shape == self.shape, (rgb.shape, self.shape)
        return rgb

    def reset(self, rgb):
        """env.rs:263-300: all four frames = warp_and_grayscale(rgb)."""
        rgb = self._frame(rgb)
        L.check(L.lib().bb_atari_reset(self._h, rgb.ctypes.data_as(C.c_void_p)))

    def step(self, rgb_a, rgb_b, reward):
        """One env step: the frames rendered at repeats 2 and 3 and the summed reward; returns the clipped reward."""
        a, b = self._frame(rgb_a), self._frame(rgb_b)
        out = C.c_float()
        L.check(L.lib().bb_atari_step(self._h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), float(reward),
                                      C.byref(out)))
        return out.value

    def obs_device(self, consumer_stream):
        """Device pointer of the current [4][84][84] observation; `consumer_stream` (the CUDA stream of the agent / ring that
        will read it) is made to wait for the preprocessing kernel unless it is the preprocessor's own stream."""
        p = C.c_void_p()
        L.check(L.lib().bb_atari_obs_device(self._h, C.c_void_p(consumer_stream), C.byref(p)))
        return p.value

    def obs(self):
        out = np.empty((4, 84, 84), np.uint8)
        L.check(L.lib().bb_atari_obs_host(self._h, out.ctypes.data_as(C.c_void_p)))
        return out
