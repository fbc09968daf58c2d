"""The synthetic RadioML-shaped generator (SURVEY §8d recipe; TT/test_dsp_functions.py:37-81, TT/test_sps_modes.py:15-24):
unit-power constellations, SPS-2 root-raised-cosine shaping, AWGN at the requested SNR, constant frame power, and the
dataset statistics of R/dataloader/dataset.py:115-157."""
import numpy as np

from oracle import amc_oracle as O
from vit_vs_raw_iq_b200 import synth


def test_constellations_have_unit_average_power_and_right_sizes():
    sizes = {"OOK": 2, "4ASK": 4, "8ASK": 8, "BPSK": 2, "QPSK": 4, "8PSK": 8, "16PSK": 16, "32PSK": 32, "16QAM": 16,
             "32QAM": 32, "64QAM": 64, "128QAM": 128, "256QAM": 256, "16APSK": 16, "32APSK": 32, "64APSK": 64, "128APSK": 128}
    for name, n in sizes.items():
        c = synth.CONSTELLATIONS[name]
        assert len(c) == n and len(np.unique(np.round(c, 5))) == n, name
        assert abs(np.mean(np.abs(c) ** 2) - 1.0) < 1e-5, name
    q = synth.CONSTELLATIONS["QPSK"]
    assert np.allclose(np.abs(q.real), 1 / np.sqrt(2), atol=1e-6) and np.allclose(np.abs(q.imag), 1 / np.sqrt(2), atol=1e-6)
    assert set(synth.CLASSES_11) <= set(synth.CLASSES_19) and len(synth.CLASSES_19) == 19


def test_rrc_filter_is_unit_energy_symmetric_and_nyquist_when_matched():
    h = synth.rrc_taps(alpha=0.35, span=8, sps=2).astype(np.float64)
    assert len(h) == 17 and abs(np.sum(h ** 2) - 1.0) < 1e-6 and np.allclose(h, h[::-1], atol=1e-7)
    rc = np.convolve(h, h)                       # matched pair = raised cosine: zero ISI at the symbol instants
    mid = len(rc) // 2
    isi = rc[mid % 2::2]
    isi = np.delete(isi, np.argmax(np.abs(isi)))
    assert np.abs(isi).max() < 0.02 * rc[mid]


def test_frames_layout_power_snr_and_determinism():
    X, y, snr = synth.make_frames(400, classes=synth.CLASSES_11, seed=5)
    X2, y2, _ = synth.make_frames(400, classes=synth.CLASSES_11, seed=5)
    assert X.shape == (400, 1024, 2) and X.dtype == np.float32 and y.dtype == np.int64 and snr.dtype == np.float32
    assert np.array_equal(X, X2) and np.array_equal(y, y2)
    assert set(np.unique(snr)) <= set(synth.SNR_GRID.tolist()) and snr.min() >= -20 and snr.max() <= 30
    p = (X.astype(np.float64) ** 2).sum(-1).mean(-1)            # every frame is scaled to the same average power
    assert np.allclose(p, 2 * 0.7616 ** 2, rtol=1e-4)
    Xs, _, _ = synth.make_frames(64, classes=synth.CLASSES_11, sps=2, seed=6)
    assert Xs.shape == (64, 2048, 2)
    # SPS-2 frames are band-limited by the RRC pulse at high SNR: little energy in the upper half band
    hi = [np.abs(np.fft.fft(Xs[k, :, 0] + 1j * Xs[k, :, 1]))[512:1536].mean() for k in range(64)]
    lo = [np.abs(np.fft.fft(Xs[k, :, 0] + 1j * Xs[k, :, 1]))[:512].mean() for k in range(64)]
    assert np.median(np.array(hi) / np.array(lo)) < 1.0


def test_snr_of_a_single_class_matches_the_request():
    rng = np.random.default_rng(0)
    s = synth._symbols("QPSK", 1024, rng)
    assert abs(np.mean(np.abs(s) ** 2) - 1.0) < 1e-6
    # reproduce the generator's noise law for one SNR and check it numerically (TT/test_sps_modes.py:20-24)
    snr_db = 8.0
    npow = 1.0 / 10 ** (snr_db / 10)
    n = np.sqrt(npow / 2) * (rng.standard_normal(200000) + 1j * rng.standard_normal(200000))
    assert abs(10 * np.log10(1.0 / np.mean(np.abs(n) ** 2)) - snr_db) < 0.05


def test_normalization_stats_agree_with_the_oracle_and_are_near_the_real_dataset():
    X, _, _ = synth.make_frames(600, classes=synth.CLASSES_19, seed=9)
    st = synth.normalization_stats(X)
    ref = O.normalization_stats(X)                 # all frames (<= 5000): same subset
    for k in st:
        assert abs(st[k] - ref[k]) < 2e-4, k
    assert 0.6 < st["i_std"] < 0.9 and 0.6 < st["q_std"] < 0.9 and abs(st["i_mean"]) < 0.05 and abs(st["q_mean"]) < 0.05
