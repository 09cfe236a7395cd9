"""Device-resident copy of a VideoDataset + a loader that iterates it like the reference's DataLoader
(at > 100 frame-pairs/s the 4-worker file loader of depth_fine_tuning.py:205-218 is the
bottleneck; 50 frames + 276 flows + 276 masks at 224x384 are 340 MB -- nothing next to 180 GB of HBM).

Every frame is stored once (a pair batch gathers its two frames), flows / masks once per pair-direction.
`ResidentLoader` yields exactly what `DataLoader(VideoDataset, batch_size, shuffle, generator=)` yields -- same
collated layout (loaders/video_dataset.py:131-207 + default collate) and, for the same generator state, the same
order: it drives the very same `RandomSampler` / `BatchSampler` objects and consumes the generator the way
DataLoader's iterator does (one int64 draw for the worker base seed before sampling).
"""
import torch
from torch.utils.data import BatchSampler, RandomSampler, SequentialSampler


class ResidentVideoDataset:
    def __init__(self, dataset, device):
        self.device = torch.device(device)
        self.pairs = [tuple(p) for p in dataset.flow_indices]
        frames = sorted({k for p in self.pairs for k in p})
        self.slot = {k: i for i, k in enumerate(frames)}
        first = dataset[0]
        H, W = first[0].shape[-2:]
        n = len(self.pairs)
        dev = self.device
        self.images = torch.empty(len(frames), 3, H, W, device=dev)
        self.flows = [torch.empty(n, 2, H, W, device=dev) for _ in range(2)]
        self.masks = [torch.empty(n, 1, H, W, device=dev) for _ in range(2)]
        filled = set()
        for i in range(n):
            images, meta = dataset[i] if i else first
            for k, img in zip(self.pairs[i], images):
                if k not in filled:
                    self.images[self.slot[k]].copy_(img)
                    filled.add(k)
            g = meta["geometry_consistency"]
            for d in range(2):
                self.flows[d][i].copy_(g["flows"][d])
                self.masks[d][i].copy_(g["masks"][d])
        self.pair_index = torch.tensor(self.pairs, dtype=torch.long)                       # (P,2) frame ids (host)
        self.pair_slots = torch.tensor([[self.slot[a], self.slot[b]] for a, b in self.pairs], device=dev)
        self.extrinsics = dataset.extrinsics.to(dev)
        self.intrinsics = dataset.intrinsics.to(dev)
        self.frame_ids = torch.tensor(self.pairs, device=dev)

    def __len__(self):
        return len(self.pairs)

    def batch(self, ids):
        """Collated mini-batch of pair indices `ids` (list of ints), all tensors on the device."""
        idx = torch.as_tensor(ids, device=self.device)
        slots = self.pair_slots[idx]                                # (B,2)
        fid = self.frame_ids[idx]
        images = self.images[slots]                                 # (B,2,3,H,W)
        metadata = {
            "extrinsics": self.extrinsics[fid],
            "intrinsics": self.intrinsics[fid],
            "geometry_consistency": {
                "indices": self.pair_index[torch.as_tensor(ids)],
                "flows": [self.flows[d][idx] for d in range(2)],
                "masks": [self.masks[d][idx] for d in range(2)],
            },
        }
        return images, metadata


class ResidentLoader:
    """for images, metadata in ResidentLoader(resident, batch_size, shuffle, generator): ... (one epoch per iteration)."""

    def __init__(self, resident, batch_size, shuffle=False, generator=None):
        self.resident, self.generator = resident, generator
        sampler = RandomSampler(range(len(resident)), generator=generator) if shuffle else SequentialSampler(range(len(resident)))
        self.batch_sampler = BatchSampler(sampler, batch_size, drop_last=False)

    def __len__(self):
        return len(self.batch_sampler)

    def __iter__(self):
        # DataLoader's iterator draws the workers' base seed from the loader's generator before sampling starts
        torch.empty((), dtype=torch.int64).random_(generator=self.generator)
        for ids in self.batch_sampler:
            yield self.resident.batch(ids)
