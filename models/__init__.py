"""Drop-in import path of the reference (`from models import SCRFD, ArcFace`, reference main.py:11,
models/__init__.py:1-2).  The classes live in scrfd_arcface_facerecognition_b200/."""
from .arcface import ArcFace
from .scrfd import SCRFD

__all__ = ["ArcFace", "SCRFD"]
