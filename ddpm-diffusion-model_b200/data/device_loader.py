"""Device-resident batch feeder.

The reference feeds `train_one_epoch` from a `torch.utils.data.DataLoader` over PIL images: Resize / CenterCrop on
the CPU, then `ToTensor()` and `Normalize([0.5]*3, [0.5]*3)` (src/data/load_data_local.py:90-107,
src/data/celebraHQ.py:40-84, src/data/load_data_from_torch.py:34-58).  At ~10 k img/s per GPU a CPU loader cannot keep
up, and the resized dataset is small (CelebA 64x64: 202 599 x 12 KB = 2.5 GB of uint8), so it lives in HBM as uint8
NHWC and every batch is ONE kernel: gather by a permutation + ToTensor + Normalize -> fp32 NCHW in [-1, 1], bit-equal
to what the reference's transforms produce from the same uint8 pixels.

`DeviceLoader` iterates like the reference's loaders: `(images, labels)` pairs, `len()` = number of batches, a new
shuffle every epoch drawn exactly like `RandomSampler` (`torch.randperm(n, generator=g)` on the host), `drop_last`
honoured.  Under torch.distributed every rank takes its `rank::world` slice of the same permutation (what
`DistributedSampler` does), so the union over ranks is the single-process epoch.
"""
from typing import Iterator, Optional, Tuple

import torch

from .. import _lib
from .. import dist as _dist


def dataset_to_u8(dataset, limit: Optional[int] = None) -> torch.Tensor:
    """One pass over a reference-style dataset (items `(x, label)` or `x`, x = float CHW in [-1,1] as produced by
    ToTensor+Normalize(0.5, 0.5)) -> uint8 NHWC on the host.  Lossless: those floats are k/255 mapped to [-1,1]."""
    out = []
    for i in range(len(dataset) if limit is None else min(limit, len(dataset))):
        item = dataset[i]
        x = item[0] if isinstance(item, (tuple, list)) else item
        out.append(((x.float() * 0.5 + 0.5) * 255.0).round_().clamp_(0, 255).to(torch.uint8).permute(1, 2, 0))
    return torch.stack(out).contiguous()


class DeviceLoader:
    def __init__(self, images_u8: torch.Tensor, batch_size: int, *, shuffle: bool = True, drop_last: bool = False,
                 device="cuda", generator: Optional[torch.Generator] = None, labels: Optional[torch.Tensor] = None,
                 shard: bool = True):
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4:
            raise ValueError("DeviceLoader expects a uint8 tensor [N,H,W,3] (or [N,3,H,W])")
        if images_u8.shape[-1] != 3:
            if images_u8.shape[1] != 3:
                raise ValueError("DeviceLoader expects three colour channels")
            images_u8 = images_u8.permute(0, 2, 3, 1)
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("ddpm_b200.DeviceLoader is CUDA-only (the dataset lives in HBM); there is no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.data = images_u8.contiguous().to(dev)
        self.labels = labels
        self.N, self.H, self.W = self.data.shape[0], self.data.shape[1], self.data.shape[2]
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.generator = generator
        self.device = dev
        self.rank, self.world = _dist.world() if shard else (0, 1)

    @staticmethod
    def epoch_order(n: int, shuffle: bool, generator: Optional[torch.Generator]) -> torch.Tensor:
        """The index order `DataLoader(dataset, shuffle=shuffle, generator=generator)` walks in one epoch, drawn from the
        generator exactly as torch does it (so that seeded runs see the same batches): the loader iterator first takes
        its base seed from the generator (`_BaseDataLoaderIter.__init__`), then `RandomSampler` draws `randperm(n)` and
        a second, empty `randperm(n)[:0]`."""
        if not shuffle:
            return torch.arange(n)
        g = generator
        if g is None:                                        # RandomSampler.__iter__: a fresh seed from the global RNG
            torch.empty((), dtype=torch.int64).random_()     # (the iterator's base seed comes from the global RNG too)
            g = torch.Generator()
            g.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
        else:
            torch.empty((), dtype=torch.int64).random_(generator=g)
        order = torch.randperm(n, generator=g)
        torch.randperm(n, generator=g)
        return order

    @staticmethod
    def shard_order(order: torch.Tensor, rank: int, world: int) -> torch.Tensor:
        """Rank `rank`'s share of one epoch: every world-th index of the common permutation (what DistributedSampler
        does), equal counts on every rank (the tail that does not divide is dropped)."""
        if world <= 1:
            return order
        per = order.numel() // world
        return order[rank:per * world:world]

    def _order(self) -> torch.Tensor:
        order = self.epoch_order(self.N, self.shuffle, self.generator)
        if self.world > 1 and self.shuffle and self.generator is None:
            # every rank would otherwise draw its own permutation from its own global RNG and the rank::world slices
            # would overlap / miss samples: rank 0's permutation is the epoch's permutation for everybody
            import torch.distributed as dist
            o = order.to(self.device)
            dist.broadcast(o, src=0)
            order = o.cpu()
        return self.shard_order(order, self.rank, self.world)

    def __len__(self) -> int:
        n = self.N // self.world if self.world > 1 else self.N
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        order = self._order()
        idx_dev = order.to(self.device, non_blocking=False)  # 8 bytes per sample, once per epoch
        st = torch.cuda.current_stream(self.device).cuda_stream
        for b in range(len(self)):
            lo = b * self.batch_size
            hi = min(lo + self.batch_size, idx_dev.numel())
            B = hi - lo
            out = torch.empty((B, 3, self.H, self.W), dtype=torch.float32, device=self.device)
            _lib.call("ddpm_batch_from_u8", self.data.data_ptr(), self.N, idx_dev[lo:hi].data_ptr(), B, self.H, self.W,
                      out.data_ptr(), st)
            y = self.labels[order[lo:hi]] if self.labels is not None else torch.zeros(B, dtype=torch.long)
            yield out, y
