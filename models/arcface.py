"""Reference import path models/arcface.py -> B200 engine implementation."""
from scrfd_arcface_facerecognition_b200.arcface import ArcFace

__all__ = ["ArcFace"]
