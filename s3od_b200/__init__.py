"""s3od_b200 - B200-native (sm_100a) implementation of S3OD's `remove_background` path.

Public surface mirrors `s3od/__init__.py` of the reference (/root/reference/src/s3od/__init__.py:1-4).
"""
__version__ = "0.1.0"

from .predictor import BackgroundRemoval, RemovalResult  # noqa: E402,F401

__all__ = ["BackgroundRemoval", "RemovalResult", "__version__"]
