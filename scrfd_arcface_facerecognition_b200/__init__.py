"""B200-native detect -> align -> embed -> match engine behind the reference's Python API.

Public surface (mirrors Kumar2421/scrfd_arcface_facerecognition):
    SCRFD(det_weight).detect(img, max_num) -> (bboxes, kpss)          reference models/scrfd.py
    ArcFace(rec_weight)(img, kps) / .get_embedding(img, kps) -> 512-d  reference models/arcface.py
    helpers.compute_similarity / estimate_norm / norm_crop_image ...   reference utils/helpers.py
    Gallery (cosine top-k, duplicate merge)                            reference main.py:136-142,
                                                                       qdrant_manager.py, duplicate.py:2726-2797
    QdrantManager (same method surface, GPU resident)                  reference qdrant_manager.py:17-300
    FaceAnalysis(name).prepare(...).get(img) -> [Face]                 reference duplicate.py:353-359, 1473-1496
    VideoRunner / FrameFeeder / FrameOverlay (batched video loop,      reference main.py:108-188,
        overlay painted on device-resident frames)                     utils/helpers.py:126-179
Importing the package does not touch CUDA; the heavy modules load lazily.
"""
__version__ = "0.1.0"

_LAZY = {"SCRFD": ".scrfd", "ArcFace": ".arcface", "Gallery": ".gallery", "FacePipeline": ".pipeline",
         "helpers": ".helpers", "FaceAnalysis": ".face_analysis", "Face": ".face_analysis",
         "QdrantManager": ".vector_store", "GalleryManager": ".vector_store",
         "VideoRunner": ".video", "FrameFeeder": ".video", "FrameOverlay": ".overlay",
         "PersonDatabase": ".result_store", "save_clustering_results": ".result_store"}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod = importlib.import_module(_LAZY[name], __name__)
        return mod if name == "helpers" else getattr(mod, name)
    raise AttributeError(name)
