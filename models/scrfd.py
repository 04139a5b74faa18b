"""Reference import path models/scrfd.py -> B200 engine implementation."""
from scrfd_arcface_facerecognition_b200.scrfd import SCRFD

__all__ = ["SCRFD"]
