"""The committed fixtures are what the REFERENCE'S OWN CODE produces: where /root/reference is mounted (the build
container; never the GPU box) the generators under tests/golden/ are re-run into a scratch directory and must
reproduce the committed files value for value.  Skipped elsewhere."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/uwsod"), reason="the reference tree is not mounted here")


def _same(x, y):
    if isinstance(x, dict):
        return x.keys() == y.keys() and all(_same(x[k], y[k]) for k in x)
    if isinstance(x, (list, tuple)):
        return len(x) == len(y) and all(_same(a, b) for a, b in zip(x, y))
    if isinstance(x, torch.Tensor):
        return x.shape == y.shape and x.dtype == y.dtype and torch.equal(x, y)
    return x == y


@pytest.mark.parametrize("script,artefact", [("make_golden.py", "oicr_plus_golden.pt"), ("make_golden_eval.py", "eval_golden.json"),
                                             ("make_golden_tta.py", "tta_golden.pt"), ("make_golden_step.py", "step_golden.pt")])
def test_generators_reproduce_the_committed_fixtures(tmp_path, script, artefact):
    env = dict(os.environ, SOSWSOD_GOLDEN_OUT=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(GOLD, script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    new, old = os.path.join(str(tmp_path), artefact), os.path.join(GOLD, artefact)
    if artefact.endswith(".json"):
        assert json.load(open(new)) == json.load(open(old))
    else:
        a, b = torch.load(new, weights_only=False), torch.load(old, weights_only=False)
        for k in ("numpy", "torch"):      # library versions recorded by the generator, not results
            a.pop(k, None), b.pop(k, None)
        assert _same(a, b), f"{artefact}: the reference's code no longer reproduces the committed fixture"
