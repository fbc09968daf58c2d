"""``from models.transformer_rawIQ import AMCTransformer`` (R/training/train.py:29) -> the B200 module."""
from vit_vs_raw_iq_b200 import RawIQAMCTransformer as AMCTransformer  # noqa: F401

__all__ = ["AMCTransformer"]
