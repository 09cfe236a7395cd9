#!/usr/bin/env python3
"""Per-entry-point CUDA-event times of ONE un-graphed fine-tune step (forward + loss + backward + Adam) of any of the
three backbones: every `cvd_*` call of the C-ABI is bracketed by events on the launching stream (CVD_MULTI_STREAM=0 so
kernels do not overlap).  Writes a JSON summary (total ms and launch count per entry point; convs also by shape).

    python tools/profile_engine.py --workload midas2 --out gpurun_out/midas_ops.json
"""
import argparse
import collections
import json
import os
import sys

os.environ["CVD_MULTI_STREAM"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mc", choices=["mc", "monodepth2", "midas2"])
    ap.add_argument("--out", default="gpurun_out/engine_ops.json")
    args = ap.parse_args()
    import __graft_entry__ as graft
    graft.build()
    from consistent_depth_b200 import _lib
    from consistent_depth_b200.fine_tune_step import FineTuneStep
    from consistent_depth_b200.synthetic import SyntheticVideo
    dev = torch.device("cuda", 0)
    if args.workload == "mc":
        from consistent_depth_b200.monodepth.mannequin_challenge_model import MannequinChallengeModel as M
        H, W, B, model = 224, 384, 4, None
        model = M()
    elif args.workload == "monodepth2":
        from consistent_depth_b200.monodepth.monodepth2_model import Monodepth2Model as M
        H, W, B = 192, 640, 4
        model = M()
    else:
        from consistent_depth_b200.monodepth.midas_v2_model import MidasV2Model as M
        H, W, B = 384, 672, 1
        model = M(pretrained=False)
    video = SyntheticVideo(50, H, W, dev, seed=1236)
    step = FineTuneStep(model, B, H, W, lr=model.learning_rate, use_graph=False)
    b = video.batch(list(range(B)))
    step.load_batch(b["images"], b["flows"], b["masks"], b["extrinsics"], b["intrinsics"])
    step.step(); step.step()
    torch.cuda.synchronize()

    real = _lib.lib()
    recs = []

    class Proxy:
        def __getattr__(self, name):
            fn = getattr(real, name)
            if not name.startswith("cvd_") or name in ("cvd_last_error", "cvd_launch_count", "cvd_version"):
                return fn

            def timed(*a):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn(*a)
                e1.record()
                key = name
                if name in ("cvd_conv_fwd", "cvd_conv_fwd_bn", "cvd_conv_wgrad", "cvd_conv_wgrad_grouped"):
                    ints = [x for x in a if type(x) is int]          # N, H, W, cin, cout, k, ... (pointers are ctypes objects)
                    key = f"{name} {ints[:6]}"
                elif name == "cvd_conv2_fwd":                        # zc8, zc8_off, N, H, W, cin, cout, k, flags
                    ints = [x for x in a if type(x) is int]
                    key = f"{name} {ints[2:8]}"
                recs.append((key, name, e0, e1))
                return rc
            return timed
    proxy = Proxy()
    _lib.lib = lambda: proxy
    step.step()
    torch.cuda.synchronize()
    _lib.lib = lambda: real
    by_fn, by_key = collections.defaultdict(lambda: [0.0, 0]), collections.defaultdict(lambda: [0.0, 0])
    for key, name, e0, e1 in recs:
        ms = e0.elapsed_time(e1)
        by_fn[name][0] += ms; by_fn[name][1] += 1
        by_key[key][0] += ms; by_key[key][1] += 1
    total = sum(v[0] for v in by_fn.values())
    out = {"workload": args.workload, "total_ms_serial": total, "calls": len(recs),
           "by_entry_point": {k: {"ms": v[0], "calls": v[1]} for k, v in sorted(by_fn.items(), key=lambda kv: -kv[1][0])},
           "top_shapes": {k: {"ms": v[0], "calls": v[1]} for k, v in sorted(by_key.items(), key=lambda kv: -kv[1][0])[:40]}}
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("workload", "total_ms_serial", "calls")}))
    for k, v in list(out["by_entry_point"].items())[:14]:
        print(f"{k:32s} {v['calls']:5d} {v['ms']:9.3f} ms")


if __name__ == "__main__":
    main()
