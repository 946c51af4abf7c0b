"""Synthetic input of the benchmark configs: 1/f "pink" noise frames.

seed = 1234 + frame_index; spectrum (N(0,1) + i N(0,1)) / f with f from fftfreq, DC = 0;
standardised; v = clip(128 + 48 a, 0, 255) as uint8; BGRA8 with B = G = R = v, A = 255.
"""
import numpy as np


def pink_noise_gray(width, height, frame_index=0):
    rng = np.random.Generator(np.random.PCG64(1234 + frame_index))
    spec = rng.standard_normal((height, width)) + 1j * rng.standard_normal((height, width))
    fx = np.fft.fftfreq(width)[None, :]
    fy = np.fft.fftfreq(height)[:, None]
    f = np.sqrt(fx * fx + fy * fy)
    f[0, 0] = np.inf
    a = np.real(np.fft.ifft2(spec / f))
    a = (a - a.mean()) / a.std()
    return np.clip(128 + 48 * a, 0, 255).astype(np.uint8)


def pink_noise_bgra(width, height, frame_index=0):
    v = pink_noise_gray(width, height, frame_index)
    out = np.empty((height, width, 4), dtype=np.uint8)
    out[..., 0] = v
    out[..., 1] = v
    out[..., 2] = v
    out[..., 3] = 255
    return out
