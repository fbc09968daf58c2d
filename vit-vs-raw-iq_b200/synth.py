"""Synthetic RadioML-shaped frames: random-symbol modulated IQ + AWGN across an SNR sweep.

There is no RadioML file and no h5py in this environment, so benchmarks and accuracy checks use this
generator.  The recipe follows the reference's own test-signal scripts:
  * unit-average-power constellations (QPSK /sqrt(2), 16QAM {+-1,+-3}/sqrt(10), ...):
    TT/test_dsp_functions.py:37-54, TT/test_sps_modes.py:15-17
  * SPS=1: one sample per symbol (L=1024).  SPS=2: zero-stuff x2 + root-raised-cosine (alpha 0.35,
    span 8) 'same' convolution (L=2048): TT/test_dsp_functions.py:58-72
  * AWGN: noise_power = signal_power / 10^(snr/10), noise = sqrt(noise_power/2) (randn + j randn):
    TT/test_sps_modes.py:20-24
  * SNR uniform over {-20,-18,...,+30} dB (the RadioML grid); every frame is then scaled to constant
    average power (per-channel std ~0.76 like the real dataset's statistics, V/main.ipynb:596) and
    given a random carrier phase.
Output layout is the dataset's: X [N, L, 2] float32 interleaved (I, Q), labels int64, snr float32.
"""
from __future__ import annotations

import numpy as np

CLASSES_11 = ["OOK", "4ASK", "8ASK", "BPSK", "QPSK", "8PSK", "16PSK", "32PSK", "16QAM", "64QAM", "256QAM"]
# the 19 modulations of R/training/train.py:61-81 (TARGET_MODULATIONS), same order
CLASSES_19 = ["OOK", "4ASK", "8ASK", "BPSK", "QPSK", "8PSK", "16PSK", "32PSK", "16APSK", "32APSK", "64APSK",
              "128APSK", "16QAM", "32QAM", "64QAM", "128QAM", "256QAM", "GMSK", "OQPSK"]
SNR_GRID = np.arange(-20, 32, 2, dtype=np.float32)


def _unit(c: np.ndarray) -> np.ndarray:
    c = np.asarray(c, dtype=np.complex64)
    return c / np.sqrt(np.mean(np.abs(c) ** 2))


def _ask(m):
    return _unit(np.arange(m, dtype=np.float32))


def _psk(m):
    return _unit(np.exp(2j * np.pi * np.arange(m) / m))


def _qam(m):
    k = int(round(np.sqrt(m)))
    if k * k == m:
        a = np.arange(-(k - 1), k, 2, dtype=np.float32)
        return _unit((a[:, None] + 1j * a[None, :]).ravel())
    # cross constellations (32, 128): square grid with the corners removed
    k = int(np.ceil(np.sqrt(m)))
    k += k % 2
    while True:
        a = np.arange(-(k - 1), k, 2, dtype=np.float32)
        pts = (a[:, None] + 1j * a[None, :]).ravel()
        pts = pts[np.argsort(np.abs(pts), kind="stable")]
        if len(pts) >= m:
            return _unit(pts[:m])
        k += 2


def _apsk(rings):
    pts = []
    for r, n in rings:
        pts.append(r * np.exp(2j * np.pi * (np.arange(n) + 0.5) / n))
    return _unit(np.concatenate(pts))


CONSTELLATIONS = {
    "OOK": _ask(2), "4ASK": _ask(4), "8ASK": _ask(8),
    "BPSK": _psk(2), "QPSK": _psk(4) * np.exp(1j * np.pi / 4), "8PSK": _psk(8), "16PSK": _psk(16), "32PSK": _psk(32),
    "16APSK": _apsk([(1.0, 4), (2.6, 12)]), "32APSK": _apsk([(1.0, 4), (2.6, 12), (4.3, 16)]),
    "64APSK": _apsk([(1.0, 4), (2.4, 12), (3.8, 20), (5.2, 28)]),
    "128APSK": _apsk([(1.0, 8), (2.0, 16), (3.0, 24), (4.0, 32), (5.0, 48)]),
    "16QAM": _qam(16), "32QAM": _qam(32), "64QAM": _qam(64), "128QAM": _qam(128), "256QAM": _qam(256),
}


def rrc_taps(alpha: float = 0.35, span: int = 8, sps: int = 2) -> np.ndarray:
    """Root-raised-cosine filter, unit energy (TT/test_dsp_functions.py:58-66)."""
    n = np.arange(-span * sps // 2, span * sps // 2 + 1, dtype=np.float64)
    t = n / sps
    h = np.zeros_like(t)
    for i, ti in enumerate(t):
        if abs(ti) < 1e-12:
            h[i] = 1.0 - alpha + 4 * alpha / np.pi
        elif abs(abs(ti) - 1 / (4 * alpha)) < 1e-9:
            h[i] = alpha / np.sqrt(2) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * alpha)) +
                                         (1 - 2 / np.pi) * np.cos(np.pi / (4 * alpha)))
        else:
            h[i] = (np.sin(np.pi * ti * (1 - alpha)) + 4 * alpha * ti * np.cos(np.pi * ti * (1 + alpha))) / \
                   (np.pi * ti * (1 - (4 * alpha * ti) ** 2))
    return (h / np.sqrt(np.sum(h ** 2))).astype(np.float32)


def _symbols(name: str, n: int, rng: np.random.Generator) -> np.ndarray:
    if name == "GMSK":      # continuous phase, +-pi/2 per symbol, Gaussian-smoothed frequency pulse
        bits = rng.integers(0, 2, n) * 2 - 1
        g = np.exp(-0.5 * (np.arange(-2, 3) / 0.85) ** 2)
        freq = np.convolve(bits, g / g.sum(), mode="same")
        return np.exp(1j * (np.pi / 2) * np.cumsum(freq)).astype(np.complex64)
    if name == "OQPSK":     # I and Q switch on alternate samples
        b = (rng.integers(0, 2, (2, n)) * 2 - 1).astype(np.float32)
        i = np.repeat(b[0, ::2], 2)[:n]
        q = np.roll(np.repeat(b[1, ::2], 2)[:n], 1)
        return ((i + 1j * q) / np.sqrt(2)).astype(np.complex64)
    c = CONSTELLATIONS[name]
    return c[rng.integers(0, len(c), n)]


def make_frames(n: int, classes=CLASSES_11, sps: int = 1, n_symbols: int = 1024, seed: int = 42,
                target_std: float = 0.7616):
    """-> (X [n, n_symbols*sps, 2] f32, labels [n] i64, snr [n] f32)."""
    rng = np.random.default_rng(seed)
    L = n_symbols * sps
    X = np.empty((n, L, 2), dtype=np.float32)
    labels = rng.integers(0, len(classes), n).astype(np.int64)
    snr = SNR_GRID[rng.integers(0, len(SNR_GRID), n)]
    taps = rrc_taps(sps=sps) if sps > 1 else None
    for k in range(n):
        s = _symbols(classes[labels[k]], n_symbols, rng)
        if sps > 1:
            up = np.zeros(L, dtype=np.complex64)
            up[::sps] = s
            s = np.convolve(up, taps, mode="same").astype(np.complex64)
        p = float(np.mean(np.abs(s) ** 2))
        npow = p / (10.0 ** (float(snr[k]) / 10.0))
        z = s + np.sqrt(npow / 2) * (rng.standard_normal(L) + 1j * rng.standard_normal(L))
        z = z * np.exp(2j * np.pi * rng.random())                 # random carrier phase
        z = z / np.sqrt(np.mean(np.abs(z) ** 2)) * (target_std * np.sqrt(2.0))
        X[k, :, 0] = z.real
        X[k, :, 1] = z.imag
    return X, labels, snr.astype(np.float32)


def normalization_stats(X: np.ndarray, max_frames: int = 5000, seed: int = 49) -> dict:
    """The dataset-level z-score statistics of R/dataloader/dataset.py:115-157 (<=5000 frames, seed 49)."""
    rng = np.random.default_rng(seed)
    idx = np.sort(rng.choice(len(X), min(max_frames, len(X)), replace=False))
    i = X[idx, :, 0].astype(np.float64).ravel()
    q = X[idx, :, 1].astype(np.float64).ravel()
    return {"i_mean": float(i.mean()), "i_std": max(float(i.std(ddof=1)), 1e-8),
            "q_mean": float(q.mean()), "q_std": max(float(q.std(ddof=1)), 1e-8)}
