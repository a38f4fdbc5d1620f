"""samplers data/loveda.py:15 uses.  StepDistributedSampler = an endless, per-epoch reshuffled DistributedSampler."""
import torch
import torch.distributed as dist
from torch.utils.data import Sampler, SubsetRandomSampler


class StepDistributedSampler(Sampler):
    def __init__(self, dataset, num_replicas=None, rank=None, seed=2333):
        ok = dist.is_available() and dist.is_initialized()
        self.n = len(dataset)
        self.world = num_replicas if num_replicas is not None else (dist.get_world_size() if ok else 1)
        self.rank = rank if rank is not None else (dist.get_rank() if ok else 0)
        self.seed, self.epoch = seed, 0

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed + self.epoch)
        self.epoch += 1
        idx = torch.randperm(self.n, generator=g).tolist()
        idx += idx[: (-len(idx)) % self.world]
        return iter(idx[self.rank::self.world])

    def __len__(self):
        return (self.n + self.world - 1) // self.world


class CrossValSamplerGenerator(object):
    def __init__(self, dataset, distributed=True, seed=2333):
        self.n, self.seed = len(dataset), seed

    def k_fold(self, k):
        g = torch.Generator().manual_seed(self.seed)
        perm = torch.randperm(self.n, generator=g).tolist()
        folds = [perm[i::k] for i in range(k)]
        return [(SubsetRandomSampler([j for f in folds[:i] + folds[i + 1:] for j in f]), SubsetRandomSampler(folds[i])) for i in range(k)]
