"""Host-side helpers of the multi-GPU path (one process per GPU, frame pairs sharded across ranks).

Reference behaviour being replaced: nn.DataParallel frame-level scatter with the loss on GPU 0 and
`batch_size *= num_gpus` (depth_fine_tuning.py:155-159, pix2pix_model.py:108).  Here every rank runs the full
step on its contiguous slice of the global mini-batch; the only exchange is ONE all-reduce(sum) of the flat
buffer [parameter gradients | local loss] — cross-sample couplings are kept exact by normalising the local loss
with B_global and using the GLOBAL batch's mean focal length (consistency_loss.py:178,208).
"""
import torch


def shard_slice(n_pairs, rank, world):
    """Contiguous slice of a global mini-batch of n_pairs owned by `rank`.  A ragged batch (the last one of an epoch:
    138 pairs = 17 x 8 + 2) is split as evenly as possible, the first n_pairs % world ranks taking one extra pair; a
    rank may own no pair at all (empty slice) -- it still joins the step's all-reduce with a zero contribution."""
    per, extra = divmod(n_pairs, world)
    start = rank * per + min(rank, extra)
    return slice(start, start + per + (1 if rank < extra else 0))


def global_focal(intrinsics, n_pairs=None):
    """(f_dir0, f_dir1): mean of (fx, fy) over the GLOBAL batch for direction k, as torch.mean(focal_length(intrinsics_ref))
    does in the reference.  intrinsics: (B,2,4) host tensor of the whole global mini-batch."""
    n = intrinsics.shape[0] if n_pairs is None else n_pairs
    return float(intrinsics[:n, 0, :2].mean()), float(intrinsics[:n, 1, :2].mean())


def allreduce_flat(store, group=None):
    """In-place sum over ranks of the flat [gradients | loss] buffer (NCCL over NVLink on GPUs, gloo in CPU tests)."""
    import torch.distributed as dist
    dist.all_reduce(store, op=dist.ReduceOp.SUM, group=group)
    return store
