"""GPU: DepthFineTuner.fine_tune() launched as TWO ranks (the torchrun environment of INTEGRATION.md §C) -- the tuner itself
binds LOCAL_RANK's device and joins the process group, ranks shard every global mini-batch (including the ragged last
one: 3 pairs, global batch 2 -> the last batch gives rank 1 nothing), and all replicas end with identical weights.
With two GPUs the ranks use NCCL on separate devices; on a single-GPU box they share cuda:0 over gloo."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_rank_fine_tune(tmp_path):
    from consistent_depth_b200.synthetic_dataset import write_synthetic_dataset
    root, range_dir = str(tmp_path / "clip"), str(tmp_path / "clip" / "R0-4_hierarchical2_mc")
    write_synthetic_dataset(root, range_dir, 4, 32, 48, pairs=[(0, 1), (1, 2), (2, 3)])
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ngpu = torch.cuda.device_count()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank if ngpu >= 2 else 0),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        if ngpu < 2:
            env["CVD_DIST_BACKEND"] = "gloo"
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_two_rank_worker.py"), root, range_dir,
                                       str(tmp_path / "res")], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    r = [json.load(open(str(tmp_path / "res") + f".{k}")) for k in range(2)]
    assert r[0]["moved"] > 0
    assert r[0]["out_dir"] == r[1]["out_dir"] and "_BS1_" in r[0]["out_dir"]      # tagged before the batch scaling (:143)
    # replicas stay identical: same all-reduced gradient, same Adam update on every rank
    assert abs(r[0]["wsum"] - r[1]["wsum"]) <= 1e-9 * r[0]["wabs"], r
    if ngpu >= 2:
        assert r[0]["device"] != r[1]["device"]
    assert os.path.isfile(os.path.join(r[0]["out_dir"], "checkpoints", "0001.pth"))
    assert "Adjusting batch size to 2." in outs[0]
