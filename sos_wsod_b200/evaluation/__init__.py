"""The rows either side of the hot path's output: the detection-results writers
the reference's evaluators implement, the image sharding of detection-result generation, and the PGF filter
(tools/pgf.py) that consumes the json."""
from .detection_results import (COCODetectionWriter, PascalVOCDetectionWriter, generate_detection_results,
                                inference_shard)
from .pgf import class_filter, pgf, pgf_voc_results

__all__ = ["COCODetectionWriter", "PascalVOCDetectionWriter", "generate_detection_results", "inference_shard",
           "class_filter", "pgf", "pgf_voc_results"]
