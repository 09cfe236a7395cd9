"""Worker of tests/test_fine_tuner_2rank_gpu.py: one rank of a 2-process DepthFineTuner.fine_tune() run (launched with
RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT in the environment, as torchrun sets them)."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    root, range_dir, out = sys.argv[1:4]
    from consistent_depth_b200.depth_fine_tuning import DepthFineTuner
    p = types.SimpleNamespace(path=root, model_type="mc", batch_size=1, learning_rate=0, optimizer="Adam", num_epochs=1,
                              lambda_view_baseline=-1, lambda_reprojection=1.0, lambda_parameter=0, val_epoch_freq=1,
                              print_freq=1, display_freq=100, save_epoch_freq=1, log_dir=None)
    ft = DepthFineTuner(range_dir, list(range(4)), p)          # binds the device, joins the process group
    assert p.batch_size == 2                                    # batch_size *= num_gpus (:155-159)
    w0 = ft.model.P.flat.clone()
    ft.fine_tune(writer=None)
    torch.cuda.synchronize()
    rank = int(os.environ["RANK"])
    res = {"rank": rank, "out_dir": ft.out_dir, "moved": float((ft.model.P.flat - w0).abs().max()),
           "wsum": float(ft.model.P.flat.double().sum()), "wabs": float(ft.model.P.flat.double().abs().sum()),
           "device": torch.cuda.current_device()}
    json.dump(res, open(f"{out}.{rank}", "w"))
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
