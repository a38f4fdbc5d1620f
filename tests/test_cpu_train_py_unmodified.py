"""BASELINE cfg1 (plumbing): the reference's OWN train.py, byte for byte, runs one synthetic LoveDA-shaped tile through the
`ever` / albumentations / skimage / timm work-alikes of compat/ (the real packages are neither in the reference tree nor in
this image).  CPU, the reference's own modules.  The `-m gpu` twin (tests/test_gpu_parity.py::test_train_py_unmodified_b200)
runs the same script with RSS_IMPL=b200, i.e. this repo's model behind the same registry name."""
import os
import subprocess
import sys
import zipfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TREE = "/root/reference/RSSFormer-TIP2023"
REF_ZIP = os.path.join(ROOT, "oracle", "_ref", "rssformer_reference.zip")


def reference_sources(tmp_path):
    """directory holding the unmodified RSSFormer-TIP2023 sources: the tree (authoring container) or the travelling archive"""
    if os.path.isdir(os.path.join(REF_TREE, "module", "baseline")):
        return REF_TREE
    if os.path.exists(REF_ZIP):
        dst = os.path.join(str(tmp_path), "ref_src")
        with zipfile.ZipFile(REF_ZIP) as z:
            z.extractall(dst)
        return dst
    return None


def run_train_py(tmp_path, impl, iters=1, batch=1, extra_env=None, timeout=900):
    src = reference_sources(tmp_path)
    if src is None:
        pytest.skip("reference sources not available (neither /root/reference nor oracle/_ref/rssformer_reference.zip)")
    work = os.path.join(str(tmp_path), "work")
    os.makedirs(work, exist_ok=True)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_synth_loveda
    make_synth_loveda.make(work, n=max(1, batch), size=1024)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), src, ROOT]), RSS_IMPL=impl)
    env.update(extra_env or {})
    cmd = [sys.executable, os.path.join(src, "train.py"), "--config_path=baseline.hrnetw32", "--model_dir=" + os.path.join(work, "log"),
           "train.num_iters", str(iters), "train.log_interval_step", "1", "train.eval_after_train", "False",
           "data.train.params.batch_size", str(batch), "data.train.params.num_workers", "0", "model.params.backbone.pretrained", "False"]
    r = subprocess.run(cmd, cwd=work, env=env, capture_output=True, text=True, timeout=timeout)
    return r, os.path.join(work, "log", "model-%d.pth" % iters)


def test_reference_train_py_runs_unmodified_cpu(tmp_path):
    r, ckpt = run_train_py(tmp_path, "reference", iters=1, batch=1, extra_env={"CUDA_VISIBLE_DEVICES": ""})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "step 1  loss" in r.stdout and os.path.exists(ckpt), r.stdout[-2000:]
    import json
    import torch
    sd = torch.load(ckpt, map_location="cpu")
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    assert [k[len("module."):] for k in sd] == list(keys)          # the checkpoint layout eval.py:36-38 expects
