"""CPU: the device-resident dataset cache + loader yields exactly the batches, in exactly the order,
of the reference's DataLoader(VideoDataset, shuffle=True, generator=...) over the same files."""
import torch
from torch.utils.data import DataLoader


def _same(a, b):
    if isinstance(a, dict):
        assert a.keys() == b.keys()
        for k in a:
            _same(a[k], b[k])
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            _same(x, y)
    else:
        assert a.shape == b.shape and a.dtype == b.dtype and torch.equal(a.cpu(), b.cpu())


def test_resident_loader_matches_dataloader(tmp_path):
    from consistent_depth_b200.loaders.resident import ResidentLoader, ResidentVideoDataset
    from consistent_depth_b200.loaders.video_dataset import VideoDataset
    from consistent_depth_b200.synthetic_dataset import write_synthetic_dataset
    root, range_dir = str(tmp_path / "clip"), str(tmp_path / "clip" / "R")
    pairs = [(0, 1), (1, 2), (2, 3), (0, 2), (3, 5), (4, 5), (0, 4)]
    write_synthetic_dataset(root, range_dir, 6, 16, 24, pairs=pairs)
    ds = VideoDataset(root, range_dir + "/metadata_scaled.npz")
    res = ResidentVideoDataset(ds, "cpu")
    assert len(res) == len(ds) == 7 and res.images.shape[0] == 6           # every frame stored once
    for shuffle in (True, False):
        g1, g2 = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
        ref = DataLoader(ds, batch_size=3, shuffle=shuffle, num_workers=0, generator=g1)
        mine = ResidentLoader(res, 3, shuffle=shuffle, generator=g2)
        assert len(ref) == len(mine) == 3
        for _epoch in range(3):                                            # generator state carries across epochs
            batches = list(zip(ref, mine))
            assert len(batches) == 3 and batches[-1][0][0].shape[0] == 1   # ragged last batch
            for (ri, rm), (mi, mm) in batches:
                _same(ri, mi)
                _same(rm, mm)


def test_video_dataset_scales_and_flow_list_fallback(tmp_path):
    """metadata["scales"] follows video_dataset.py:196-204 (dict -> per frame, scalar -> both frames, absent -> no key) and,
    without flow_list.json, the pair list is parsed from the flow file names through to_in_range / to_one_way (:113-125)."""
    import os
    from consistent_depth_b200.loaders import frame_sampling
    from consistent_depth_b200.loaders.video_dataset import VideoDataset
    from consistent_depth_b200.synthetic_dataset import write_synthetic_dataset
    root, range_dir = str(tmp_path / "clip"), str(tmp_path / "clip" / "R")
    pairs = [(0, 1), (1, 2), (0, 2)]
    write_synthetic_dataset(root, range_dir, 3, 16, 24, pairs=pairs)
    ds = VideoDataset(root, range_dir + "/metadata_scaled.npz")
    _, md = ds[0]
    assert "scales" not in md
    ds.scales = {0: 2.0, 1: 0.5, 2: 4.0}
    pair = ds.flow_indices[1]
    _, md = ds[1]
    assert md["scales"].shape == (2, 1) and md["scales"].flatten().tolist() == [ds.scales[k] for k in pair]
    ds.scales = 3.0
    assert ds[2][1]["scales"].flatten().tolist() == [3.0, 3.0]
    with_list = list(ds.flow_indices)
    os.remove(os.path.join(root, "flow_list.json"))
    ds2 = VideoDataset(root, range_dir + "/metadata_scaled.npz")
    assert sorted(map(tuple, ds2.flow_indices)) == sorted(map(tuple, with_list))
    assert frame_sampling.to_in_range([(0, 1), (1, 5), (4, 2)], (0, 3)) == [(0, 1)]
    assert frame_sampling.to_in_range([(0, 9)]) == [(0, 9)]
