from .linear import linear_act
from .nms import batched_nms, nms
from .roi_pool import RoIPool, roi_pool

__all__ = ["RoIPool", "roi_pool", "nms", "batched_nms", "linear_act"]
