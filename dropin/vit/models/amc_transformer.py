"""``from models.amc_transformer import AMCTransformer`` (V/training/train.py:28) -> the B200 module."""
from vit_vs_raw_iq_b200 import ViTAMCTransformer as AMCTransformer  # noqa: F401

__all__ = ["AMCTransformer"]
