from .box_head import DiscriminativeAdaptionNeck, build_box_head
from .fast_rcnn_oicr import OICROutputLayers
from .fast_rcnn_wsddn import WSDDNOutputLayers
from .meta_arch import MultiInputRCNN, detector_postprocess
from .poolers import ROIPooler, convert_boxes_to_pooler_format
from .roi_heads_oicrplus import OICRPlusHeads, build_roi_heads, get_image_level_gt
from .test_time_augmentation_avg import DatasetMapperTTAAVG, GeneralizedRCNNWithTTAAVG, ViewSpec, resize_shortest_edge

__all__ = ["ROIPooler", "convert_boxes_to_pooler_format", "DiscriminativeAdaptionNeck", "build_box_head",
           "WSDDNOutputLayers", "OICROutputLayers", "OICRPlusHeads", "build_roi_heads", "get_image_level_gt",
           "MultiInputRCNN", "detector_postprocess", "DatasetMapperTTAAVG", "GeneralizedRCNNWithTTAAVG", "ViewSpec", "resize_shortest_edge"]
