"""Drop-in for models/denoiser_h3d.py `MDM`: 256-d prompt style branch, learned null-prompt embedding,
audio/word masking (denoiser_h3d.py:116-146,160-181,199-200). Same native engine, variant 'h3d'."""
from .denoiser import MDM as _MDM


class MDM(_MDM):
    variant_default = "h3d"
