"""The evaluation report must be byte-compatible with the reference's (sklearn text report inside the
header of R/training/utils.py:391-401) and parse with compare_models.py's regexes (TT/compare_models.py:39,44,49)."""
import re

import numpy as np
import pytest
import torch

from vit_vs_raw_iq_b200.evaluate import _report_from_confusion

CLASSES = ["OOK", "4ASK", "8ASK", "BPSK", "QPSK", "8PSK", "16PSK", "32PSK", "16APSK", "32APSK", "64APSK", "128APSK",
           "16QAM", "32QAM", "64QAM", "128QAM", "256QAM", "GMSK", "OQPSK"]


def test_report_text_equals_sklearn():
    from sklearn.metrics import classification_report, confusion_matrix
    rng = np.random.default_rng(0)
    y = rng.integers(0, len(CLASSES), 5000)
    p = np.where(rng.random(5000) < 0.6, y, rng.integers(0, len(CLASSES), 5000))
    cm = confusion_matrix(y, p, labels=np.arange(len(CLASSES)))
    ours = _report_from_confusion(cm, CLASSES, digits=4)
    ref = classification_report(y, p, target_names=CLASSES, digits=4)
    assert ours == ref


def test_report_parses_with_reference_regexes(tmp_path):
    rng = np.random.default_rng(1)
    C = len(CLASSES)
    cm = rng.integers(0, 50, (C, C)) + np.eye(C, dtype=np.int64) * 400
    text = "Classification Report - Test Set\n" + "=" * 80 + "\n\nOverall Accuracy: 63.44%\n\nAccuracy by SNR:\n"
    for snr, acc in ((-8, 0.1386), (0, 0.5708), (8, 0.9919)):
        text += f"  SNR {snr:+3d} dB: {acc*100:.2f}%\n"
    text += "\n" + "=" * 80 + "\n\n" + _report_from_confusion(cm, CLASSES)
    assert float(re.search(r'Overall Accuracy:\s+([\d.]+)%', text).group(1)) == 63.44
    snr = {int(a): float(b) for a, b in re.findall(r'SNR\s+([-+]\d+)\s+dB:\s+([\d.]+)%', text)}
    assert snr == {-8: 13.86, 0: 57.08, 8: 99.19}
    found = {}
    for line in text.split("\n"):
        m = re.match(r'^\s*(\w+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+(\d+)', line)
        if m and m.group(1) not in ("accuracy", "macro", "weighted"):
            found[m.group(1)] = int(m.group(5))
    assert set(found) == set(CLASSES)
    assert found["OOK"] == int(cm[0].sum())


@pytest.mark.gpu
def test_evaluate_model_end_to_end(tmp_path):
    import vit_vs_raw_iq_b200 as amc
    from vit_vs_raw_iq_b200 import synth
    from vit_vs_raw_iq_b200.evaluate import evaluate_model_with_confusion
    from vit_vs_raw_iq_b200.trainer import predict
    dev = "cuda:0"
    torch.manual_seed(0)
    X, y, snr = synth.make_frames(1000, classes=synth.CLASSES_11, seed=9)
    stats = synth.normalization_stats(X)
    model = amc.RawIQAMCTransformer(in_channels=2, seq_length=1024, num_classes=11, d_model=64, n_head=4, n_layers=2,
                                    ffn_hidden=128, drop_prob=0.1, device=dev, segment_size=16, compute_dtype="bf16")
    model.set_raw_input(stats)
    Xt, yt, st = torch.from_numpy(X), torch.from_numpy(y), torch.from_numpy(snr)
    batches = [(Xt[i:i + 256], yt[i:i + 256], st[i:i + 256]) for i in range(0, 1000, 256)]   # ragged last batch
    res = evaluate_model_with_confusion(model, batches, synth.CLASSES_11, tmp_path, prefix="test")
    ref_pred = predict(model, Xt.to(dev)).cpu().numpy()
    assert res["confusion_matrix"].sum() == 1000
    assert abs(res["overall_accuracy"] - float((ref_pred == y).mean())) < 1e-9
    for s in (-8, 0, 8):
        sel = np.abs(snr - s) <= 0.5
        assert abs(res["snr_accuracies"][s] - float((ref_pred[sel] == y[sel]).mean())) < 1e-9
    text = open(res["report_path"]).read()
    assert text.startswith("Classification Report - Test Set\n" + "=" * 80)
    assert re.search(r'Overall Accuracy:\s+([\d.]+)%', text)
