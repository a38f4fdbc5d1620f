"""Data-parallel training step for RSSFormer.

The reference delegates this to the un-vendored `ever` trainer 'th_amp_ddp' (RSSFormer-TIP2023/train.py:79-80)
configured by configs/base/loveda.py:68-113: SGD(momentum 0.9, weight decay 1e-4), clip_grad_norm_(35, L2),
poly LR (base 0.01, power 0.9, 30000 iters), AMP, DDP mean all-reduce, SyncBN.

B200-first layout: all trainable parameters live in ONE flat fp32 buffer (params/grads/momentum are views), so
  * zero_grad is folded into the optimiser kernel,
  * the gradient all-reduce is one NCCL call over NVLink/NVSwitch on the flat buffer (no bucket copies),
  * clip-norm + weight decay + momentum + update + bf16 shadow refresh are two kernels over 32.1 M elements
    (rss_grad_sumsq, rss_sgd_step) instead of ~10 passes and hundreds of launches.
Parameters that never receive a gradient in the reference (`headaux.*`: consumed under no_grad, CGFL.py:75-97)
are kept out of the flat buffer, exactly as torch.optim.SGD skips params whose .grad is None.
"""
import torch
import torch.distributed as dist

from . import conv, ops

NO_GRAD_PREFIXES = ("headaux.",)


def poly_lr(it, base_lr=0.01, power=0.9, max_iters=30000):
    """configs/base/loveda.py:87-93"""
    return base_lr * (1.0 - min(it, max_iters) / max_iters) ** power


class FlatSGD:
    def __init__(self, model, base_lr=0.01, momentum=0.9, weight_decay=1e-4, max_norm=35.0, power=0.9, max_iters=30000,
                 bf16_shadow=True):
        self.model = model
        self.hp = dict(base_lr=base_lr, momentum=momentum, weight_decay=weight_decay, max_norm=max_norm, power=power, max_iters=max_iters)
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad and not n.startswith(NO_GRAD_PREFIXES)]
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        dev = self.params[0].device
        # 16-byte align every view so that vector kernels can read weights in place
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.offsets, self.numel = offs, total
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(total, device=dev, dtype=torch.float32)
        self.shadow = torch.zeros(total, device=dev, dtype=torch.bfloat16) if bf16_shadow else None
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.lr_dev = torch.zeros(1, device=dev, dtype=torch.float32)     # poly LR lives on the device (graph-replay safe)
        # ring of pinned slots: with graph replay the host runs ahead of the device, so a slot must not be rewritten before the
        # copy that reads it has executed (8 steps of run-ahead are far more than the launch queue allows)
        self._lr_host = torch.zeros(8, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(8)
        self._lr_slot = 0
        for p, o in zip(self.params, offs):
            n = p.numel()
            self.flat_p[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + n].view(p.shape)
            p.grad = self.flat_g[o:o + n].view(p.shape)
            p._rss_flat = True
            if self.shadow is not None:
                conv.register_shadow(p, self.shadow[o:o + n].view(p.shape))
        if self.shadow is not None:
            self.shadow.copy_(self.flat_p)
        self._build_cl_shadow(dev)
        self._build_t_shadow(dev)
        self.iteration = 0
        from .modules import FusedBNAct
        self._bn_counters = []
        for m in model.modules():
            if isinstance(m, FusedBNAct):
                m.defer_counter = True                      # instance attribute: other models in the process keep counting themselves
                self._bn_counters.append(m.num_batches_tracked)
        self._tail_off, self._early_done, self._comm, self._group = None, False, None, None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and dev.type == "cuda":
            dist.broadcast(self.flat_p, src=0)              # replicas must start identical (DDP does this in its constructor)
            self.sync_shadows()
            # overlapped gradient all-reduce: parameters are laid out in module order, the backward pass produces gradients in reverse
            # order, so [first stage-4 parameter : end] (stage 4 + neck + head, ~2/3 of the bytes) is final when stage 4 has been
            # differentiated; HighResolutionNet._stage4_hook fires there and that tail is all-reduced on a side stream under stages 3..1
            import os
            from . import hrnet
            first = next((o for n, o in zip(self.names, self.offsets) if ".stage4." in n), None)
            if first is not None and os.environ.get("RSS_EARLY_ALLREDUCE", "1") != "0":
                nets = [m for m in model.modules() if isinstance(m, hrnet.HighResolutionNet)]
                if len(nets) == 1:
                    self._tail_off = first
                    self._comm = torch.cuda.Stream(dev)
                    nets[0]._stage4_hook = self.early_all_reduce

    def state_dict(self):
        """momentum buffer, schedule position and hyper-parameters (the parameters themselves are the model's state_dict)"""
        return {"momentum_buffer": self.flat_m.detach().clone(), "iteration": int(self.iteration), "hp": dict(self.hp),
                "names": list(self.names), "offsets": list(self.offsets)}

    def load_state_dict(self, state):
        """restores what state_dict() saved and refreshes the bf16 / channels-last / transposed weight shadows from the (already
        loaded) parameters: call AFTER model.load_state_dict()"""
        if list(state["names"]) != list(self.names) or list(state["offsets"]) != list(self.offsets):
            raise ValueError("optimizer state was saved for a different parameter layout")
        self.flat_m.copy_(state["momentum_buffer"])
        self.iteration = int(state["iteration"])
        self.hp.update(state["hp"])
        self.sync_shadows()

    def _build_cl_shadow(self, dev):
        """channels-last bf16 copies of every k x k (k > 1) convolution weight, refreshed by ONE kernel per step
        (rss_shadow_cl_refresh) instead of one permute kernel in front of each library convolution"""
        self.shadow_cl = None
        if self.shadow is None or dev.type != "cuda":
            return
        table, rows, total, nrow, max_row = [], [0], 0, 0, 1
        views = []
        for p, o in zip(self.params, self.offsets):
            if p.dim() != 4 or p.shape[2] * p.shape[3] == 1:
                continue
            cout, cin, kh, kw = p.shape
            table.append([o, total, cin, kh * kw])
            views.append((p, total, (cout, kh, kw, cin)))
            total += (p.numel() + 7) // 8 * 8                 # 16-byte aligned copies
            nrow += cout
            rows.append(nrow)
            max_row = max(max_row, cin * kh * kw)
        if not table:
            return
        self.shadow_cl = torch.zeros(total, device=dev, dtype=torch.bfloat16)
        self._cl_table = torch.tensor(table, dtype=torch.int64).to(dev)
        self._cl_rows = torch.tensor(rows, dtype=torch.int64).to(dev)
        self._cl_meta = (len(table), max_row)
        for p, o, shp in views:
            conv.register_shadow_cl(p, self.shadow_cl[o:o + p.numel()].view(shp).permute(0, 3, 1, 2))
        self.refresh_cl_shadow()

    def _build_t_shadow(self, dev):
        """transposed bf16 copies ([Cin][kh*kw][Cout]) of the 3x3 weights the fused conv (csrc/conv_cf.cu) differentiates: the
        data-gradient operand, refreshed by ONE kernel per step (rss_shadow_t_refresh) instead of a pack kernel per call"""
        self.shadow_t = None
        if self.shadow is None or dev.type != "cuda":
            return
        table, views, total = [], [], 0
        for p, o in zip(self.params, self.offsets):
            if p.dim() != 4 or tuple(p.shape[2:]) != (3, 3) or p.shape[0] != p.shape[1] or p.shape[0] not in conv.CF_SQUARE_3X3:
                continue
            cout, cin, kh, kw = p.shape
            table.append([o, total, cout, cin, kh * kw])
            views.append((p, total))
            total += p.numel()
        if not table:
            return
        self.shadow_t = torch.zeros(total, device=dev, dtype=torch.bfloat16)
        self._t_table = torch.tensor(table, dtype=torch.int64).to(dev)
        for p, o in views:
            p._rss_shadow_t = self.shadow_t[o:o + p.numel()]
        self.refresh_t_shadow()

    def refresh_t_shadow(self):
        if self.shadow_t is not None:
            ops.shadow_t_refresh(self.flat_p, self.shadow_t, self._t_table, self._t_table.shape[0])

    def sync_shadows(self):
        """call after writing parameters from outside the optimiser (e.g. load_state_dict on resume): refreshes the bf16 copies
        the convolution kernels read"""
        if self.shadow is not None:
            self.shadow.copy_(self.flat_p)
        self.refresh_cl_shadow()
        self.refresh_t_shadow()

    def refresh_cl_shadow(self):
        if self.shadow_cl is not None:
            ops.shadow_cl_refresh(self.flat_p, self.shadow_cl, self._cl_table, self._cl_rows, *self._cl_meta)

    def lr(self):
        return poly_lr(self.iteration, self.hp["base_lr"], self.hp["power"], self.hp["max_iters"])

    def rebind_grads(self):
        """autograd may replace .grad objects (e.g. after set_to_none); point them back at the flat buffer."""
        for p, o in zip(self.params, self.offsets):
            v = self.flat_g[o:o + p.numel()].view(p.shape)
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                if p.grad is not None:
                    v.copy_(p.grad)
                p.grad = v

    def early_all_reduce(self):
        """backward-pass hook (HighResolutionNet._stage4_hook): all-reduce the stage-4 / neck / head gradients on the communication stream.
        Every kernel that contributes to them has been ISSUED by now (autograd is past stage 4) on the chain streams or the
        weight-gradient streams: the communication stream waits for all of those, then runs the batched fp32 accumulation of the
        library gradients collected so far, then the collective."""
        if self._tail_off is None or self._early_done:
            return
        from . import hrnet
        dev = self.flat_g.device
        comm = self._comm
        comm.wait_stream(torch.cuda.current_stream(dev))
        if self._main_stream is not None:
            comm.wait_stream(self._main_stream)
        for s in hrnet._SIDE.get(dev, []) + hrnet._ROW.get(dev, []):      # branch streams and fuse-row streams
            comm.wait_stream(s)
        for s in conv.WGRAD["streams"].get(dev, []):
            comm.wait_stream(s)
        with torch.cuda.stream(comm):
            conv._flush_pending()
            dist.all_reduce(self.flat_g[self._tail_off:], group=self._group)
        self._early_done = True

    _main_stream = None

    def all_reduce_grads(self, group=None):
        conv.join_wgrad()                      # weight-gradient kernels run on a side stream: join before touching flat_g
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if self._early_done:               # the tail went out during the backward pass (early_all_reduce)
                dist.all_reduce(self.flat_g[:self._tail_off], group=group)
                torch.cuda.current_stream(self.flat_g.device).wait_stream(self._comm)
                self._early_done = False
            else:
                dist.all_reduce(self.flat_g, group=group)      # sum; the 1/world mean is folded into the step kernels
            return 1.0 / dist.get_world_size(group)
        return 1.0

    def push_lr(self):
        """host side of the schedule: write lr(iteration) into the device scalar (async, pinned)"""
        self._lr_slot = (self._lr_slot + 1) % self._lr_host.numel()
        self._lr_host[self._lr_slot] = self.lr()
        self.lr_dev.copy_(self._lr_host[self._lr_slot:self._lr_slot + 1], non_blocking=True)

    def device_step(self, grad_scale=1.0):
        """the two optimiser kernels only (capturable); the caller handles push_lr()/iteration"""
        hp = self.hp
        ops.grad_sumsq(self.flat_g, grad_scale, self.sumsq)
        ops.sgd_step(self.flat_p, self.flat_g, self.flat_m, self.sumsq, grad_scale, hp["max_norm"], self.lr_dev, hp["momentum"],
                     hp["weight_decay"], True, self.shadow)
        self.refresh_cl_shadow()
        self.refresh_t_shadow()
        if self._bn_counters and self.model.training:
            torch._foreach_add_(self._bn_counters, 1)       # num_batches_tracked of all 330 BN layers in one multi-tensor op

    def step(self, grad_scale=1.0):
        self.push_lr()
        self.device_step(grad_scale)
        self.iteration += 1

    def grad_norm(self):
        return float(self.sumsq.sqrt().item())


def _device_train_step(model, opt, img, labels, group=None):
    opt._group = group
    opt._main_stream = torch.cuda.current_stream(img.device) if img.is_cuda else None
    losses = model(img, {"cls": labels})
    loss = sum(losses.values())
    loss.backward()
    scale = opt.all_reduce_grads(group)
    opt.device_step(scale)
    return loss.detach()


def train_step(model, opt, img, labels, group=None):
    """forward + loss + backward + gradient all-reduce + clip + SGD.  Returns the loss tensor (device, no sync)."""
    opt.push_lr()
    loss = _device_train_step(model, opt, img, labels, group)
    opt.iteration += 1
    return loss


class GraphedTrainStep:
    """The whole training step captured ONCE in a CUDA graph and replayed: ~13 k kernel launches per step are
    launch-latency bound when issued from Python, and every shape in the step is static.  Inputs are copied into
    static device buffers; the poly LR is a device scalar refreshed by the host before each replay."""

    def __init__(self, model, opt, img, labels, group=None, warmup=3, restore_after_warmup=False):
        """warmup eager steps run first (one-time kernel attribute setup, allocator warm-up) and are REAL optimiser steps on the given
        batch.  restore_after_warmup=True snapshots parameters, momentum, the schedule position and every module buffer (BatchNorm
        running statistics, counters) before them and puts everything back afterwards, so the first replay is training step 0."""
        self.model, self.opt, self.group = model, opt, group
        self.img = img.clone()
        self.labels = labels.clone()
        snap = None
        if restore_after_warmup:
            snap = (opt.flat_p.clone(), opt.flat_m.clone(), opt.iteration, [b.clone() for b in model.buffers()])
        from .hrnet import CHAIN_PRIORITY
        side = self.stream = torch.cuda.Stream(img.device, priority=CHAIN_PRIORITY)     # chain streams outrank the wgrad streams
        side.wait_stream(torch.cuda.current_stream())
        self.warmup_losses = []
        with torch.cuda.stream(side):
            for _ in range(warmup):                     # eager warm-up (also performs one-time kernel attribute setup)
                self.warmup_losses.append(train_step(model, opt, self.img, self.labels, group))
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if snap is not None:
            opt.flat_p.copy_(snap[0]); opt.flat_m.copy_(snap[1]); opt.flat_g.zero_()
            opt.iteration = snap[2]
            for b, v in zip(model.buffers(), snap[3]):
                b.copy_(v)
            opt.sync_shadows()
            torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        opt.push_lr()
        # thread_local: the NCCL watchdog thread may touch CUDA while this thread captures (DDP all-reduce / SyncBN inside the graph)
        with torch.cuda.graph(self.graph, stream=self.stream, capture_error_mode="thread_local"):
            self.loss = _device_train_step(model, opt, self.img, self.labels, group)   # capture records, does not run

    def load(self, img, labels, non_blocking=True):
        self.img.copy_(img, non_blocking=non_blocking)
        self.labels.copy_(labels, non_blocking=non_blocking)

    def __call__(self):
        self.opt.push_lr()
        self.graph.replay()
        self.opt.iteration += 1
        return self.loss
