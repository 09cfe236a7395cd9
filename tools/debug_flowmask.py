import os, sys, subprocess, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import flowmask_oracle as fo
from oracle.make_golden import FLOWMASK_CASES
from consistent_depth_b200.utils.consistency import consistent_flow_masks_batched
subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "consistent_depth_b200", "csrc"), "-o", "/tmp/fm.so",
                       os.path.join(ROOT, "tests", "host_emul", "flow_mask_host.c"), "-lm"])
HL = C.CDLL("/tmp/fm.so")
dev = torch.device("cuda:0")
for name, (seed, H, W, ft, ct) in FLOWMASK_CASES.items():
    flows, colors = fo.synthetic_pair(seed, H, W)
    want = fo.consistent_flow_masks(flows, colors, ft, ct)
    fn = np.ascontiguousarray(np.stack([x.transpose(2, 0, 1) for x in flows])[None], np.float32)
    cn = np.ascontiguousarray(np.stack([x.transpose(2, 0, 1) for x in colors])[None], np.float32)
    f, c = torch.tensor(fn, device=dev), torch.tensor(cn, device=dev)
    print(name, "roundtrip equal", np.array_equal(f.cpu().numpy(), fn), np.array_equal(c.cpu().numpy(), cn), f.is_contiguous(), c.is_contiguous(),
          hex(f.data_ptr()), hex(c.data_ptr()))
    for (a, b) in ((ft, ct), (1e4, 1e4), (ft, 1e4), (1e4, ct), (0.0, 0.0)):
        mm = consistent_flow_masks_batched(f, c, a, b)[0]
        torch.cuda.synchronize()
        got = mm.cpu().numpy() > 0.5
        m = np.zeros((1, 2, H, W), np.float32); p = lambda x: x.ctypes.data_as(C.c_void_p)
        HL.flow_mask_host(p(fn), p(cn), p(m), 1, H, W, C.c_float(a), C.c_float(b))
        hm = m[0] > 0.5
        print(f"  thresholds ({a}, {b}): gpu ones {[float(got[d].mean()) for d in range(2)]} host ones {[float(hm[d].mean()) for d in range(2)]} "
              f"mismatch {[float((got[d] != hm[d]).mean()) for d in range(2)]}")
    # where do they differ?
    mm = consistent_flow_masks_batched(f, c, ft, ct)[0].cpu().numpy() > 0.5
    m = np.zeros((1, 2, H, W), np.float32)
    HL.flow_mask_host(p(fn), p(cn), p(m), 1, H, W, C.c_float(ft), C.c_float(ct))
    diff = mm[0] != (m[0, 0] > 0.5)
    ys, xs = np.nonzero(diff)
    print("  diff rows hist", np.bincount(ys, minlength=H).tolist())
    print("  diff cols hist", np.bincount(xs, minlength=W).tolist())
