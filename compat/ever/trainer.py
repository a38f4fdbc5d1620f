"""get_trainer('th_amp_ddp') (train.py:79-80): SGD(momentum, weight decay) + clip_grad_norm_ + poly LR + AMP + DDP, configured by
configs/base/loveda.py:68-113.  Command line: --config_path baseline.hrnetw32 --model_dir DIR [dotted.key value ...]."""
import argparse
import os
import time

import torch
import torch.distributed as dist
import torch.nn as nn

from .core import builder
from .core.config import apply_overrides, import_config
from .core.logger import get_logger


class _Single(nn.Module):
    """what DistributedDataParallel looks like to evaluate_cls_fn (`self.model.module.config`) when there is one process"""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **kw):
        return self.module(*a, **kw)


class _Checkpoint(object):
    def __init__(self):
        self.global_step = 0


def poly_lr(it, base_lr, power, max_iters):
    return base_lr * (1.0 - min(it, max_iters) / max_iters) ** power


class Launcher(object):
    def __init__(self, model, cfg, model_dir, logger, device):
        self.model, self.cfg, self._model_dir, self.logger, self.device = model, cfg, model_dir, logger, device
        self.checkpoint = _Checkpoint()
        self._evaluate = None

    def override_evaluate(self, fn):
        self._evaluate = fn

    def evaluate(self, test_loader):
        if self._evaluate is not None:
            self._evaluate(self, test_loader, self.cfg.get("test", {}))

    def save(self):
        os.makedirs(self._model_dir, exist_ok=True)
        path = os.path.join(self._model_dir, "model-%d.pth" % self.checkpoint.global_step)
        torch.save(self.model.state_dict(), path)        # keys carry the 'module.' prefix, as eval.py:38 expects
        return path

    def train_iters(self, train_loader, test_loader):
        tc, oc, lc = self.cfg["train"], self.cfg["optimizer"], self.cfg["learning_rate"]["params"]
        core = self.model.module
        b200 = type(core).__module__.startswith("representationlearning_b200")
        clip = oc.get("grad_clip", {}).get("max_norm", None)
        if b200:                      # this repo's model: fused clip + SGD + poly LR over one flat parameter buffer
            import representationlearning_b200 as P
            opt = P.FlatSGD(core, base_lr=lc["base_lr"], momentum=oc["params"]["momentum"], weight_decay=oc["params"]["weight_decay"],
                            max_norm=clip if clip is not None else 1e30, power=lc["power"], max_iters=lc["max_iters"])
        else:
            params = [p for p in self.model.parameters() if p.requires_grad]
            opt = torch.optim.SGD(params, lr=lc["base_lr"], **oc["params"])
            use_amp = self.device.type == "cuda"
            scaler = torch.amp.GradScaler("cuda", enabled=use_amp)
        it, t0 = 0, time.time()
        self.model.train()
        while it < tc["num_iters"]:
            for img, gt in train_loader:
                if it >= tc["num_iters"]:
                    break
                img = img.to(self.device, non_blocking=True)
                gt = {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in gt.items()}
                if b200:
                    loss = P.train_step(core, opt, img, gt["cls"])
                else:
                    lr = poly_lr(it, lc["base_lr"], lc["power"], lc["max_iters"])
                    for g in opt.param_groups:
                        g["lr"] = lr
                    opt.zero_grad(set_to_none=True)
                    for _ in range(tc.get("forward_times", 1)):
                        with torch.autocast(self.device.type, dtype=torch.float16, enabled=use_amp):
                            loss = sum(self.model(img, gt).values())
                        scaler.scale(loss).backward()
                    scaler.unscale_(opt)
                    if clip is not None:
                        nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], clip, oc["grad_clip"].get("norm_type", 2))
                    scaler.step(opt)
                    scaler.update()
                it += 1
                self.checkpoint.global_step = it
                if it % tc.get("log_interval_step", 50) == 0 or it == tc["num_iters"]:
                    self.logger.info("step %d  loss %.6f  %.2f s/step", it, float(loss), (time.time() - t0) / it)
        return it


class _Trainer(object):
    def run(self, after_construct_launcher_callbacks=None):
        ap = argparse.ArgumentParser()
        ap.add_argument("--config_path", type=str, required=True)
        ap.add_argument("--model_dir", type=str, default="./log")
        ap.add_argument("--local_rank", "--local-rank", type=int, default=int(os.environ.get("LOCAL_RANK", 0)))
        ap.add_argument("opts", nargs=argparse.REMAINDER)
        args = ap.parse_args()
        cfg = apply_overrides(import_config(args.config_path), args.opts)
        logger = get_logger("ever.trainer")
        cuda = torch.cuda.is_available()
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl" if cuda else "gloo")
        device = torch.device("cuda", args.local_rank) if cuda else torch.device("cpu")
        if cuda:
            torch.cuda.set_device(device)
        if os.environ.get("RSS_IMPL", "reference") == "b200":
            torch.backends.cudnn.enabled = True           # train.py:73 switches cuDNN off; the library convs of this repo's model use it
        model = builder.make_model(cfg["model"]).to(device)
        if cfg["train"].get("sync_bn", False) and world > 1 and cuda:
            model = nn.SyncBatchNorm.convert_sync_batchnorm(model)
        if world > 1 and not type(model).__module__.startswith("representationlearning_b200"):
            model = nn.parallel.DistributedDataParallel(model, device_ids=[args.local_rank] if cuda else None)
        else:
            model = _Single(model)                        # this repo's model all-reduces its flat gradient buffer itself
        launcher = Launcher(model, cfg, args.model_dir, logger, device)
        for cb in after_construct_launcher_callbacks or []:
            cb(launcher)
        train_loader = builder.make_dataloader(cfg["data"]["train"])
        n = launcher.train_iters(train_loader, None)
        path = launcher.save()
        logger.info("trained %d iterations, checkpoint %s", n, path)
        if cfg["train"].get("eval_after_train", False):
            launcher.evaluate(builder.make_dataloader(cfg["data"]["test"]))


_TRAINERS = {"th_amp_ddp": _Trainer, "th_ddp": _Trainer, "th_amp": _Trainer}


def get_trainer(name):
    return _TRAINERS[name]
