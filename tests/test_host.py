"""CPU suite, part 2: the C-ABI library loads and exports every symbol include/soswsod_b200.h declares (no compute
calls without a GPU), the host-side plugin surface mirrors the reference's names, the product path refuses to run
on CPU tensors (no fallback), and the data-parallel gradient hook averages over ranks (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "soswsod_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(soswsod_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_and_exports_every_declared_symbol():
    from sos_wsod_b200 import _lib
    from sos_wsod_b200.build import build_library

    path = build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = _declared_symbols()
    assert len(syms) >= 21, syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/soswsod_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes prototypes and header must list the same entry points"
    loaded = _lib.load()
    assert loaded.soswsod_abi_version() == _lib.ABI_VERSION
    assert loaded.soswsod_nms_workspace_bytes(2000) > 2000 * 32 * 8
    assert loaded.soswsod_detect_workspace_bytes(2000, 20) > 20 * 2000 * 32 * 8
    assert loaded.soswsod_oicr_mine_workspace_bytes(200, 3, 3) >= 3 * 600 * 12


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    from sos_wsod_b200.build import LIB_PATH

    out = subprocess.run(["cuobjdump", "-sass", LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out, "tcgen05.mma must be present in the GEMM kernel"
    assert "UTMALDG" in out, "TMA loads must be present"
    assert "LDTM" in out, "tcgen05.ld (TMEM -> registers) must be present"
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout


def test_no_cpu_fallback():
    from sos_wsod_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.roi_pool_forward(torch.zeros(1, 4, 8, 8), torch.zeros(2, 5))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.nms(torch.zeros(3, 4), torch.zeros(3), 0.5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm_bf16(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from sos_wsod_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sos_wsod_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/_ref", ""), f"{f} must not reference the oracle"


def test_plugin_surface_names_match_reference():
    from sos_wsod_b200.config import get_cfg
    from sos_wsod_b200.modeling import OICRPlusHeads, build_roi_heads
    from sos_wsod_b200.registry import ROI_BOX_HEAD_REGISTRY, ROI_HEADS_REGISTRY
    from sos_wsod_b200.structures import ShapeSpec

    cfg = get_cfg()
    cfg.merge_from_list(["WSL.REFINE_NUM", 4])
    assert ROI_HEADS_REGISTRY.get("OICRPlusHeads") is OICRPlusHeads
    assert "DiscriminativeAdaptionNeck" in ROI_BOX_HEAD_REGISTRY
    heads = build_roi_heads(cfg, {"plain5": ShapeSpec(channels=512, stride=8)})
    names = dict(heads.named_parameters())
    # checkpoint keys of the reference: shapes [out, in] fp32
    assert names["box_head.fc1.weight"].shape == (4096, 25088) and names["box_head.fc2.weight"].shape == (4096, 4096)
    assert names["box_predictor.cls.weight"].shape == (20, 4096) and names["box_predictor.det.bias"].shape == (20,)
    for k in range(4):
        assert names[f"box_refinery_{k}.cls_score.weight"].shape == (21, 4096)
        assert names[f"box_refinery_{k}.bbox_pred.weight"].shape == (80, 4096)
    assert len(names) == 8 + 4 * 4
    # reference initialisers (box_head.py:64-67, fast_rcnn_oicr.py:465-468)
    assert abs(names["box_head.fc1.weight"].std().item() - 0.005) < 2e-4
    assert torch.all(names["box_head.fc1.bias"] == 0.1)
    assert abs(names["box_refinery_0.cls_score.weight"].std().item() - 0.01) < 1e-3
    hc = heads.head_config()
    assert (hc.in_dim, hc.fc_dim, hc.refine_k, hc.head_cols) == (25088, 4096, 4, 2 * 20 + 4 * 101)
    assert hc.top_k(2000) == 200 and hc.top_k(7) == 1
    sd = heads.state_dict()
    heads2 = build_roi_heads(cfg, {"plain5": ShapeSpec(channels=512, stride=8)})
    heads2.load_state_dict(sd)


def test_default_config_is_the_released_k4_and_loads_a_released_state_dict():
    """every code_release yaml ships WSL.REFINE_NUM: 4 (voc07_oicr_plus.yaml:56-58); the DEFAULT config and
    HeadConfig must build box_refinery_0..3 so that a released checkpoint loads strictly."""
    from sos_wsod_b200.config import get_cfg
    from sos_wsod_b200.engine import HeadConfig
    from sos_wsod_b200.modeling import build_roi_heads
    from sos_wsod_b200.structures import ShapeSpec

    cfg = get_cfg()
    assert cfg.WSL.REFINE_NUM == 4 and HeadConfig().refine_k == 4
    cfg.MODEL.ROI_BOX_HEAD.DAN_DIM = [32, 32]
    heads = build_roi_heads(cfg, {"plain5": ShapeSpec(channels=4, stride=8)})
    released = {"box_head.fc1.weight": (32, 196), "box_head.fc1.bias": (32,), "box_head.fc2.weight": (32, 32),
                "box_head.fc2.bias": (32,), "box_predictor.cls.weight": (20, 32), "box_predictor.cls.bias": (20,),
                "box_predictor.det.weight": (20, 32), "box_predictor.det.bias": (20,)}
    for k in range(4):
        released.update({f"box_refinery_{k}.cls_score.weight": (21, 32), f"box_refinery_{k}.cls_score.bias": (21,),
                         f"box_refinery_{k}.bbox_pred.weight": (80, 32), f"box_refinery_{k}.bbox_pred.bias": (80,)})
    heads.load_state_dict({k: torch.zeros(v) for k, v in released.items()}, strict=True)


@pytest.mark.parametrize("key,val", [("OICRPLUS.BBOX_UPDATE", True), ("MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_TYPE", "giou"),
                                     ("MODEL.ROI_BOX_HEAD.SMOOTH_L1_BETA", 0.5), ("WSL.MEAN_LOSS", False),
                                     ("MODEL.ROI_HEADS.IOU_LABELS", [0, 1, 1]), ("MODEL.ROI_HEADS.PROPOSAL_APPEND_GT", True),
                                     ("MODEL.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG", True), ("WSL.MIST_TYPE", "wetectron"),
                                     ("WSL.REFINE_REG", [True, False, True, True]), ("MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_WEIGHT", 2.0)])
def test_config_flags_that_change_the_maths_are_refused_at_construction(key, val):
    """a flag the fused path does not implement must raise in from_config, never train silently with the
    released configuration's arithmetic."""
    from sos_wsod_b200.config import get_cfg
    from sos_wsod_b200.modeling import build_roi_heads
    from sos_wsod_b200.structures import ShapeSpec

    cfg = get_cfg()
    cfg.MODEL.ROI_BOX_HEAD.DAN_DIM = [32, 32]
    cfg.merge_from_list([key, val])
    with pytest.raises(NotImplementedError):
        build_roi_heads(cfg, {"plain5": ShapeSpec(channels=4, stride=8)})


def test_synthetic_inputs_equal_the_fixture_generator():
    """sos_wsod_b200.synthetic (what bench.py feeds) and the test suite's generator (what the golden fixtures were made
    with) are independent implementations of SURVEY.md §8d: identical tensors for the same seed."""
    from oracle import oicr_plus_ref as ref
    from sos_wsod_b200 import synthetic as syn

    for seed, R, sizes, ch in [(1434, 2000, [(480, 640), (576, 768)], 4), (7, 200, [(240, 320), (288, 384)], 8)]:
        a = ref.synth_views(R, sizes, torch.Generator().manual_seed(seed), channels=ch)
        b = syn.synth_views(R, sizes, torch.Generator().manual_seed(seed), channels=ch)
        for x, y in zip(a, b):
            assert torch.equal(x.feat, y.feat) and torch.equal(x.boxes, y.boxes) and torch.equal(x.obj, y.obj)
            assert tuple(x.image_size) == tuple(y.image_size)
    views, gt = syn.training_image(0, 0, R=300, channels=4)
    assert len(views) == 4 and 1 <= gt.numel() <= 4 and bool((gt[1:] > gt[:-1]).all())
    feats, rois, obj = syn.pack_views(views)
    assert feats[0].shape[0] == 2 and rois[0].shape == (600, 5) and obj.shape == (1200,)
    assert rois[0][:300, 0].eq(0).all() and rois[0][300:, 0].eq(1).all()


def test_get_image_level_gt_and_structures():
    from sos_wsod_b200.modeling import convert_boxes_to_pooler_format, get_image_level_gt
    from sos_wsod_b200.structures import Boxes, Instances

    t = [Instances((10, 10), gt_classes=torch.tensor([7, 2, 7]), gt_boxes=Boxes(torch.zeros(3, 4)))]
    _, gt_int, oh = get_image_level_gt(t, 20)
    assert gt_int[0].tolist() == [2, 7] and oh.shape == (1, 20) and oh.sum() == 2
    r = convert_boxes_to_pooler_format([Boxes(torch.ones(2, 4)), Boxes(torch.zeros(1, 4))])
    assert r.shape == (3, 5) and r[:, 0].tolist() == [0.0, 0.0, 1.0]
    b = Boxes(torch.tensor([[-5.0, 2.0, 30.0, 50.0]]))
    b.clip((20, 25))
    assert b.tensor.tolist() == [[0.0, 2.0, 25.0, 20.0]]


_DIST_SCRIPT = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, world = dist.get_rank(), 2
from sos_wsod_b200.distributed import GradientExchange

# ---- the host logic of both exchange modes against plain DDP arithmetic (SGD + momentum, two steps) ----
torch.manual_seed(0)                       # same initial parameters on both ranks
rows, cols, panels = 16, 6, 2
init = {"fc1_w": torch.randn(rows, cols), "fc1_b": torch.randn(rows), "fc2_w": torch.randn(8, 8), "fc2_b": torch.randn(8),
        "head_w": torch.randn(4, 8), "head_b": torch.randn(4)}
lr, mom = 0.1, 0.9
def local_grads(step, r):
    g = torch.Generator().manual_seed(100 * step + r)
    return {k: torch.randn(v.shape, generator=g) for k, v in init.items()}
# DDP reference: every rank averages every gradient and updates everything
ref_p = {k: v.clone() for k, v in init.items()}
ref_b = {k: torch.zeros_like(v) for k, v in init.items()}
for step in range(2):
    gs = [local_grads(step, r) for r in range(world)]
    for k in ref_p:
        g = (gs[0][k] + gs[1][k]) / world
        ref_b[k] = mom * ref_b[k] + g
        ref_p[k] = ref_p[k] - lr * ref_b[k]
for mode in ("allreduce", "sharded"):
    master = {k: v.clone() for k, v in init.items()}
    bufs = {k: torch.zeros_like(v) for k, v in init.items()}
    opnd = {"fc1_w": master["fc1_w"].to(torch.bfloat16), "fc2_w": master["fc2_w"].to(torch.bfloat16)}
    ex = GradientExchange(master, mode=mode, min_shard_elems=1)
    assert ex.sharded == ({"fc1_w", "fc2_w"} if mode == "sharded" else set())
    for step in range(2):
        g = local_grads(step, rank)
        ex.begin_step()
        # the engine's hook order: head, fc2, fc1 in row panels, fc1 bias
        ex.hook("head_w", g["head_w"], 0); ex.hook("head_b", g["head_b"], 0)
        ex.hook("fc2_w", g["fc2_w"], 0); ex.hook("fc2_b", g["fc2_b"], 0)
        per = rows // panels
        for pi in range(panels):
            ex.hook("fc1_w", g["fc1_w"][pi * per:(pi + 1) * per], pi * per)
        ex.hook("fc1_b", g["fc1_b"], 0)
        ex.wait_gradients()
        for k in master:                   # the optimizer: only the owned rows of a sharded tensor
            for lo, hi in ex.owned_rows(k):
                bufs[k][lo:hi] = mom * bufs[k][lo:hi] + g[k][lo:hi]
                master[k][lo:hi] -= lr * bufs[k][lo:hi]
                if k in opnd:
                    opnd[k][lo:hi] = master[k][lo:hi].to(torch.bfloat16)
        ex.gather_operands(opnd)
        ex.operand_gate()
    if mode == "sharded":
        own = ex._panel_ranges("fc1_w")
        assert own == [(0, 8), (8, 16)], own
        assert ex.master_stale
        lo, hi = 4 * (1 - rank), 4 * (1 - rank) + 4          # rows of panel 0 owned by the OTHER rank: stale master
        assert not torch.allclose(master["fc1_w"][lo:hi], ref_p["fc1_w"][lo:hi])
        assert ex.bytes_last_step["reduce_scatter"] == (rows * cols + 64) * 4 and ex.bytes_last_step["all_gather"] == (rows * cols + 64) * 2
    # the bf16 operands every rank computes with equal DDP's result on EVERY row, in both modes
    for k in opnd:
        assert torch.equal(opnd[k], ref_p[k].to(torch.bfloat16)), (mode, k)
    ex.sync_master()
    for k in master:
        assert torch.allclose(master[k], ref_p[k], rtol=1e-6, atol=1e-7), (mode, k)
    if mode == "sharded":
        ex.gather_rows(bufs["fc1_w"], "fc1_w")
        assert torch.allclose(bufs["fc1_w"], ref_b["fc1_w"], rtol=1e-6, atol=1e-7)
# image sharding of detection-result generation: contiguous blocks, no collective (InferenceSampler)
n = 11
per = (n + 1) // 2
mine = list(range(rank * per, min(n, (rank + 1) * per)))
gathered = [None, None]
dist.all_gather_object(gathered, mine)
assert sorted(sum(gathered, [])) == list(range(n))
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_data_parallel_gradient_hook_gloo_world2(tmp_path):
    script = tmp_path / "dp.py"
    script.write_text(_DIST_SCRIPT)
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_tta_mapper_host_side_matches_reference_views():
    """DatasetMapperTTAAVG mirror: view order / sizes (ResizeShortestEdge arithmetic) and the PIL-resized, flipped
    pixels equal what the reference's own mapper produced (tests/golden/tta_golden.pt); the launch parameters carry
    the reference's Python-double scale factors."""
    import hashlib

    from sos_wsod_b200.config import get_cfg
    from sos_wsod_b200.modeling.test_time_augmentation_avg import DatasetMapperTTAAVG, ViewSpec, resize_shortest_edge

    gold = torch.load(os.path.join(ROOT, "tests", "golden", "tta_golden.pt"), weights_only=False)
    for c in gold["cases"]:
        cfg = get_cfg()
        cfg.MODEL.DEVICE = "cpu"
        cfg.TEST.AUG.MIN_SIZES, cfg.TEST.AUG.MAX_SIZE, cfg.TEST.AUG.FLIP = c["min_sizes"], c["max_size"], c["flip"]
        cfg.DATASETS.PRECOMPUTED_PROPOSAL_TOPK_TEST = c["topk"]
        mapper = DatasetMapperTTAAVG(cfg)
        h, w = c["stored_hw"]
        specs = mapper.view_specs(h, w)
        assert len(specs) == len(c["views"])
        img = c["image"].permute(1, 2, 0).numpy()
        for s, v in zip(specs, c["views"]):
            assert (3,) + s.image_size == v["image_shape"] and s.flip == ("HFlipTransform" in v["transforms"])
            out = s.apply_image(img.copy())
            chw = torch.from_numpy(out.transpose(2, 0, 1).copy())
            assert hashlib.sha1(chw.numpy().tobytes()).hexdigest() == v["image_sha1"], c["name"]
    assert resize_shortest_edge(375, 500, 480, 4000) == (480, 640)
    assert resize_shortest_edge(500, 375, 1152, 1200) == (1200, 900)
    p = ViewSpec(375, 500, 480, 640, True, orig_hw=(750, 1000)).params(batch_index=1.0)
    assert p == [640 / 500, 480 / 375, 1.0, 640.0, 480.0, 1.0, 500 / 640, 375 / 480, 2.0, 2.0]
    assert len(ViewSpec(10, 10, 20, 20, False).params()) == 10


def test_meta_arch_host_helpers():
    """ImageList.from_tensors and detector_postprocess of the MultiInputRCNN mirror (host logic, no GPU)."""
    from sos_wsod_b200.modeling import MultiInputRCNN, detector_postprocess
    from sos_wsod_b200.structures import Boxes, ImageList, Instances

    il = ImageList.from_tensors([torch.zeros(3, 10, 12), torch.ones(3, 8, 15)], 4)
    assert il.tensor.shape == (2, 3, 12, 16) and il.image_sizes == [(10, 12), (8, 15)]
    assert float(il.tensor[1, :, :8, :15].min()) == 1.0 and float(il.tensor[1, :, 8:, :].max()) == 0.0
    one = ImageList.from_tensors([torch.ones(3, 7, 9)])
    assert one.tensor.shape == (1, 3, 7, 9)
    r = Instances((10, 20), pred_boxes=Boxes(torch.tensor([[1.0, 2.0, 30.0, 8.0], [5.0, 5.0, 5.0, 9.0]])),
                  scores=torch.tensor([0.5, 0.4]))
    o = detector_postprocess(r, 20, 40)
    assert o.image_size == (20, 40) and o.pred_boxes.tensor.tolist() == [[2.0, 4.0, 40.0, 16.0]] and o.scores.tolist() == [0.5]
    m = MultiInputRCNN(backbone=torch.nn.Identity(), proposal_generator=None, roi_heads=torch.nn.Identity(),
                       pixel_mean=(1.0, 2.0, 3.0), pixel_std=(1.0, 1.0, 2.0))
    imgs = m.preprocess_image_inference([{"image": torch.full((3, 4, 5), 5.0)}])
    assert imgs.tensor.shape == (1, 3, 4, 5) and imgs.tensor[0, :, 0, 0].tolist() == [4.0, 3.0, 1.0]
    with pytest.raises(AssertionError, match="imgs_per_gpu=1"):
        m([{}, {}])


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path on the host cores): one JSON line with the keys the
    driver reads; under torchrun only rank 0 works (the other ranks exit 0 without output)."""
    import json

    env = dict(os.environ, SOSWSOD_REF_BUDGET_SECONDS="1.5")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    b = json.loads(lines[0])
    assert b["impl"] == "reference" and b["unit"] == "image-views/s" and b["higher_is_better"] is True
    assert b["metric"] == "OICR+ head fwd+bwd images/s" and b["steps"] == 2 and b["value"] > 0
    assert b["cpu_baseline"]["kind"] == "port" and b["cpu_baseline"]["cores"] >= 1 and b["cpu_baseline"]["value"] == b["value"]
    assert b["e2e"] == {"value": b["value"], "unit": b["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in b["config"] and "sample" in b["config"]
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2"],
                        capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=120)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_solver_param_groups_follow_the_reference_rules():
    """get_optimizer_param_groups (uwsod/detectron2/solver/build.py:143-218): bias lr x BIAS_LR_FACTOR with
    WEIGHT_DECAY_BIAS, norm layers WEIGHT_DECAY_NORM; with REFINE_SCALE_ON the name-keyed rules and REFINE_LR_SCALE."""
    from sos_wsod_b200.config import get_cfg
    from sos_wsod_b200.modeling import build_roi_heads
    from sos_wsod_b200.solver import B200SGD, build_optimizer, get_optimizer_param_groups
    from sos_wsod_b200.structures import ShapeSpec

    cfg = get_cfg()
    cfg.MODEL.ROI_BOX_HEAD.DAN_DIM = [64, 64]
    heads = build_roi_heads(cfg, {"plain5": ShapeSpec(channels=4, stride=8)})
    model = torch.nn.Sequential()
    model.add_module("backbone", torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4)))
    model.add_module("roi_heads", heads)
    names = {id(p): n for n, p in model.named_parameters()}
    groups = get_optimizer_param_groups(cfg, model)
    assert len(groups) == len(names)
    got = {names[id(g["params"][0])]: (g["lr"], g["weight_decay"]) for g in groups}
    for n, (lr, wd) in got.items():
        if n.startswith("backbone.1."):
            assert (lr, wd) == ((0.002 if n.endswith("bias") else 0.001), 0.0) or (lr, wd) == (0.001, 0.0), n   # norm layer
        elif n.endswith("bias"):
            assert (lr, wd) == (0.002, 0.0), n
        else:
            assert (lr, wd) == (0.001, 0.0005), n
    assert got["backbone.1.weight"] == (0.001, 0.0) and got["backbone.1.bias"] == (0.001, 0.0)
    cfg.SOLVER.REFINE_SCALE_ON, cfg.SOLVER.REFINE_LR_SCALE = True, 3.0
    got = {names[id(g["params"][0])]: (g["lr"], g["weight_decay"]) for g in get_optimizer_param_groups(cfg, model)}
    assert got["roi_heads.box_refinery_0.cls_score.bias"] == (0.001 * 2.0 * 3.0, 0.0)
    assert got["roi_heads.box_refinery_2.bbox_pred.weight"] == (0.001 * 3.0, 0.0005)
    assert got["roi_heads.box_head.fc1.bias"] == (0.002, 0.0) and got["roi_heads.box_predictor.cls.weight"] == (0.001, 0.0005)
    opt = build_optimizer(cfg, model)
    assert isinstance(opt, B200SGD) and opt.defaults["momentum"] == 0.9 and len(opt.param_groups) == len(names)
    model.backbone[0].weight.grad = torch.zeros_like(model.backbone[0].weight)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        opt.step()
    cfg.SOLVER.NESTEROV = True
    with pytest.raises(NotImplementedError):
        build_optimizer(cfg, model)


def test_half_table_cover_is_exact_for_every_bin_row():
    """The load list of roi_pool_fwd_half_kernel (csrc/roi_pool_fast.cu, struct BinCols), restated in Python and checked
    exhaustively: for every bin row [ws, we) of 4..16 cells and every uniform list length up to 6, the loads (single cells
    and 1 x 4 windows at EVEN starts) touch only cells of the bin, cover all of them, are visited left to right, and the
    running strict '>' over them returns the value and the FIRST position of the row's maximum -- with heavy ties."""
    import random

    def cols(ws, we, W, t_count):
        hd = ws & 1
        a = ws + hd
        bb = (we - 4) & ~1
        nw = ((bb - a + 3) >> 2) + 1 if bb >= a else 0
        cend = bb + 4 if nw else a
        count = hd + nw + (we - cend)
        out = []
        for t in range(t_count):
            if t < hd:
                out.append(ws)
                continue
            k = t - hd
            if k < nw:
                out.append(W + (min(a + 4 * k, bb) >> 1))
            else:
                out.append(min(cend + (k - nw), we - 1))
        return count, out

    rng = random.Random(7)
    W = 40
    for ws in range(0, 20):
        for bw in range(4, 17):
            we = ws + bw
            count, _ = cols(ws, we, W, 0)
            assert 1 <= count <= 6
            for ns in range(count, 7):           # a warp pads the lane's list to the longest list among its lanes
                _, cl = cols(ws, we, W, ns)
                covered = set()
                last_first = -1
                for c in cl:
                    cells = [c] if c < W else list(range(2 * (c - W), 2 * (c - W) + 4))
                    assert c < W or (2 * (c - W)) % 2 == 0
                    assert all(ws <= x < we for x in cells), (ws, we, cl)
                    assert cells[0] >= last_first, "loads must be visited left to right"
                    last_first = cells[0]
                    covered.update(cells)
                assert covered == set(range(ws, we)), (ws, we, cl)
                for _ in range(6):
                    row = [float(rng.randint(0, 3)) for _ in range(W)]      # many ties
                    best, pos = -1e30, -1
                    for c in cl:
                        if c < W:
                            v, p = row[c], c
                        else:
                            x0 = 2 * (c - W)
                            win = row[x0:x0 + 4]
                            v = max(win)
                            p = x0 + win.index(v)                            # the table stores the window's first maximum
                        if v > best:
                            best, pos = v, p
                    seg = row[ws:we]
                    assert best == max(seg) and pos == ws + seg.index(max(seg)), (ws, we, cl, row[ws:we])
