"""Module-path-compatible home of conv2d_resample (reference: lib/model_zoo/stylegan_utils/conv2d_resample.py:57)."""
from ...ops import conv2d_resample  # noqa: F401
