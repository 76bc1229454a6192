"""yacs-style config shim (yacs is not installed here) carrying exactly the keys the OICR+ head reads
(SURVEY.md §5 "Config / flags"), with the values of the released configs
uwsod/projects/WSL/configs/Detection/code_release/voc07_oicr_plus.yaml (REFINE_NUM: 4, :56-58) +
Base-RCNN-DilatedC5.yaml and the defaults of uwsod/projects/WSL/wsl/config/defaults.py.  A real detectron2 CfgNode works in its place."""
from __future__ import annotations


class CfgNode(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self) -> "CfgNode":
        out = CfgNode()
        for k, v in self.items():
            out[k] = v.clone() if isinstance(v, CfgNode) else (list(v) if isinstance(v, list) else v)
        return out

    def merge_from_list(self, kv):
        assert len(kv) % 2 == 0
        for key, val in zip(kv[0::2], kv[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            assert parts[-1] in node, f"Non-existent config key: {key}"
            node[parts[-1]] = val
        return self


def get_cfg() -> CfgNode:
    C = CfgNode
    cfg = C()
    cfg.VIS_PERIOD = 0
    cfg.MODEL = C()
    cfg.MODEL.DEVICE = "cuda"
    cfg.MODEL.ROI_HEADS = C(
        NAME="OICRPlusHeads", NUM_CLASSES=20, IN_FEATURES=["plain5"], SCORE_THRESH_TEST=1e-6, NMS_THRESH_TEST=0.3,
        IOU_THRESHOLDS=[0.5, 0.6], IOU_LABELS=[0, -1, 1], PROPOSAL_APPEND_GT=False, BATCH_SIZE_PER_IMAGE=4096,
        POSITIVE_FRACTION=1.0)
    cfg.MODEL.ROI_BOX_HEAD = C(
        NAME="DiscriminativeAdaptionNeck", POOLER_TYPE="ROIPool", POOLER_RESOLUTION=7, POOLER_SAMPLING_RATIO=0,
        DAN_DIM=[4096, 4096], BBOX_REG_WEIGHTS=(10.0, 10.0, 5.0, 5.0), SMOOTH_L1_BETA=0.0,
        BBOX_REG_LOSS_TYPE="smooth_l1", BBOX_REG_LOSS_WEIGHT=1.0, CLS_AGNOSTIC_BBOX_REG=False, DROPOUT=0.5)
    cfg.WSL = C(REFINE_NUM=4, REFINE_REG=[True, True, True, True], REFINE_MIST=True, MIST_P=0.10, MIST_THRE=0.05,
                MIST_TYPE="nms", MEAN_LOSS=True)
    cfg.OICRPLUS = C(BBOX_UPDATE=False, PROPOSAL_NUM=2000)
    # detection_result_test.yaml:46-50 (shipped with ENABLED: False), Base-RCNN-DilatedC5.yaml:4-10
    cfg.TEST = C(DETECTIONS_PER_IMAGE=100, AUG=C(ENABLED=False, MIN_SIZES=(480, 576, 672, 768, 864, 960, 1056, 1152),
                                                  MAX_SIZE=4000, FLIP=True))
    cfg.INPUT = C(FORMAT="BGR")
    cfg.DATASETS = C(PRECOMPUTED_PROPOSAL_TOPK_TRAIN=4000, PRECOMPUTED_PROPOSAL_TOPK_TEST=4000)
    cfg.MODEL.LOAD_PROPOSALS = True
    cfg.MODEL.MASK_ON = False
    cfg.MODEL.KEYPOINT_ON = False
    # voc07_oicr_plus.yaml:36-45 + detectron2/config/defaults.py:509-547,655-656
    cfg.SOLVER = C(AMP=C(ENABLED=False), BASE_LR=0.001, MOMENTUM=0.9, NESTEROV=False, WEIGHT_DECAY=0.0005,
                   WEIGHT_DECAY_NORM=0.0, BIAS_LR_FACTOR=2.0, WEIGHT_DECAY_BIAS=0.0, REFINE_SCALE_ON=False,
                   REFINE_LR_SCALE=1.0, CLIP_GRADIENTS=C(ENABLED=False))
    return cfg
