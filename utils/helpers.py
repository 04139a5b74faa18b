"""Reference import path utils/helpers.py -> B200 engine implementation
(`from utils.helpers import compute_similarity, draw_bbox_info, draw_bbox`, reference main.py:12)."""
from scrfd_arcface_facerecognition_b200.helpers import (  # noqa: F401
    compute_similarity, distance2bbox, distance2kps, draw_bbox, draw_bbox_info, estimate_norm,
    norm_crop_image, reference_alignment)
