"""B200-native track extraction + classifier-input preprocessing (hot path of
TheCacophonyProject/classifier-pipeline), imported as ``classifier_pipeline_b200``.

Sub-packages mirror the reference's module layout for the hot path only:
``track`` (ClipTrackExtractor, Clip, Track, Region), ``ml_tools`` (imageprocessing,
preprocess, rectangle, frame), ``piclassifier`` (motiondetector, cptvmotiondetector),
``config`` (tracking configuration data), ``cptv`` (CPTV v2 decoder) and ``native``
(ctypes binding of ``libcptrack.so``, the C-ABI over the sm_100a CUDA kernels).
"""
__version__ = "0.1.0"
