"""CPU, world_size 2 over gloo: the host logic of the multi-GPU path — pair sharding, global-batch scalars
(B_global normalisation, batch-mean focal length) and the single [gradients | loss] all-reduce — reproduces the
single-process result (checked with the oracle's closed form, which takes the same B_global / shard inputs)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import synth, consistency_oracle as co


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


PAIRS = [(0, 1), (1, 3), (2, 6), (4, 5)]


def _worker(rank, world, port, ret, B):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from consistent_depth_b200.distributed import allreduce_flat, global_focal, shard_slice
    H, W = 16, 24
    pairs = PAIRS[:B]
    batch = synth.make_pair_batch(5, pairs, H, W)
    batch["intrinsics"][-1, :, :2] *= 1.1           # make the per-pair focal lengths differ: f must be the GLOBAL mean
    depth = synth.synth_depth_pred(5, B, H, W)
    sl = shard_slice(B, rank, world)                # ragged batches: unequal shares, possibly none (FineTuneStep.step_empty)
    f_dir = global_focal(torch.tensor(batch["intrinsics"]))
    # local loss / gradient with the global scalars, via the oracle's closed form evaluated per direction
    loss_l, grad_l = 0.0, np.zeros((B, 2, H, W))
    if sl.stop > sl.start:
        l, r, d, g = co.closed_form(depth[sl], batch["extrinsics"][sl], batch["intrinsics"][sl],
                                    [f[sl] for f in batch["flows"]], [m[sl] for m in batch["masks"]], 1.0, 0.0, B_global=B)
        loss_l += l; grad_l[sl] += g
    store = torch.zeros(B * 2 * H * W + 4, dtype=torch.float64)
    store[:-4] = torch.tensor(grad_l).reshape(-1)
    store[-4] = loss_l
    allreduce_flat(store)
    if rank == 0:
        ret["store"] = store.numpy().copy()
        ret["f_dir"] = f_dir
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("B", [4, 3, 1])          # even, ragged (shares 2 + 1), and a rank without any pair (1 + 0)
def test_two_rank_sharding_matches_single_process(B):
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret, B), nprocs=world, join=True)
    H, W = 16, 24
    batch = synth.make_pair_batch(5, PAIRS[:B], H, W)
    batch["intrinsics"][-1, :, :2] *= 1.1
    depth = synth.synth_depth_pred(5, B, H, W)
    l, r, d, g = co.closed_form(depth, batch["extrinsics"], batch["intrinsics"], batch["flows"], batch["masks"], 1.0, 0.0)
    store = ret["store"]
    np.testing.assert_allclose(store[-4], l, rtol=1e-12)
    np.testing.assert_allclose(store[:-4].reshape(B, 2, H, W), g, rtol=1e-12, atol=1e-18)
    f0 = float(np.mean(batch["intrinsics"][:, 0, :2])); f1 = float(np.mean(batch["intrinsics"][:, 1, :2]))
    np.testing.assert_allclose(ret["f_dir"], (f0, f1), rtol=1e-6)


def test_shard_slices_partition_the_batch():
    """Even and ragged global batches (the last batch of 138 pairs at global batch 8 holds 2): contiguous, disjoint,
    complete, sizes differ by at most one, ranks beyond the batch get an empty slice."""
    from consistent_depth_b200.distributed import shard_slice
    for world in (1, 2, 4, 8):
        for n in (8, 2, 5, 1, 16, 7):
            seen, sizes = [], []
            for r in range(world):
                s = shard_slice(n, r, world)
                seen += list(range(n))[s]
                sizes.append(s.stop - s.start)
            assert seen == list(range(n)) and max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_pair_sampling_and_dataset_roundtrip(tmp_path):
    """loaders: hierarchical2 pair counts of the reference (138 @50 frames ...) and the .raw / VideoDataset contract."""
    import json
    from consistent_depth_b200.loaders.frame_sampling import hierarchical2_one_way
    from consistent_depth_b200.loaders.video_dataset import VideoDataset
    from consistent_depth_b200.utils import image_io
    assert [len(hierarchical2_one_way(n)) for n in (2, 50, 100, 200)] == [1, 138, 286, 584]
    H, W, pairs = 8, 12, [(0, 1), (0, 2)]
    root = str(tmp_path)
    for d in ("color_down", "flow", "mask"):
        os.makedirs(os.path.join(root, d))
    b = synth.make_pair_batch(3, pairs, H, W)
    frames = {0: b["images"][0, 0], 1: b["images"][0, 1], 2: b["images"][1, 1]}
    for i, im in frames.items():        # .raw colour files are stored BGR-swizzled (video.py:174): loader swaps back
        image_io.save_raw_float32_image(os.path.join(root, "color_down", f"frame_{i:06d}.raw"), im.transpose(1, 2, 0)[..., ::-1])
    try:
        import cv2
        wr = lambda p, a: cv2.imwrite(p, a)
    except ImportError:
        from PIL import Image
        wr = lambda p, a: Image.fromarray(a).save(p)
    for k, (i, j) in enumerate(pairs):
        for d, (r, t) in enumerate(((i, j), (j, i))):
            image_io.save_raw_float32_image(os.path.join(root, "flow", f"flow_{r:06d}_{t:06d}.raw"), b["flows"][d][k].transpose(1, 2, 0))
            wr(os.path.join(root, "mask", f"mask_{r:06d}_{t:06d}.png"), (b["masks"][d][k, 0] * 255).astype(np.uint8))
    json.dump([list(p) for p in pairs], open(os.path.join(root, "flow_list.json"), "w"))
    extr, intr = synth.camera_track(3, H, W)
    np.savez(os.path.join(root, "meta.npz"), extrinsics=extr, intrinsics=intr)
    ds = VideoDataset(root, os.path.join(root, "meta.npz"))
    assert len(ds) == 2
    images, meta = ds[1]
    np.testing.assert_allclose(images.numpy(), b["images"][1], rtol=0, atol=0)
    np.testing.assert_allclose(meta["geometry_consistency"]["flows"][1].numpy(), b["flows"][1][1])
    np.testing.assert_allclose(meta["geometry_consistency"]["masks"][0].numpy(), b["masks"][0][1])
    assert meta["geometry_consistency"]["indices"].tolist() == [0, 2]
