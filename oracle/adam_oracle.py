"""CPU restatement of torch.optim.Adam.step as the reference configures it.

optimizer/__init__.py:5-17 maps "Adam" -> torch.optim.Adam; depth_fine_tuning.py:231-236
creates it with lr, betas=(0.9, 0.999) and defaults eps=1e-8, weight_decay=0,
amsgrad=False.  Parameters whose grad is None are skipped.  The NaN guard of
depth_fine_tuning.py:278-280 (`continue` before backward/step) is `skip`.
TEST INFRASTRUCTURE ONLY. Pinned by tests/golden/adam.npz (real torch.optim.Adam).
"""
import numpy as np


class AdamOracle:
    def __init__(self, lr, betas=(0.9, 0.999), eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, betas[0], betas[1], eps
        self.t = 0
        self.m = None
        self.v = None

    def step(self, p, g, skip=False):
        """p, g: float32 numpy arrays (flat). Updates p in place; returns p."""
        if skip:
            return p
        if self.m is None:
            self.m = np.zeros_like(p); self.v = np.zeros_like(p)
        self.t += 1
        f32 = np.float32
        self.m += f32(1 - self.b1) * (g - self.m)                     # exp_avg.lerp_(grad, 1-beta1)
        self.v *= f32(self.b2); self.v += f32(1 - self.b2) * g * g   # mul_(beta2).addcmul_(g, g, 1-beta2)
        bc1 = 1 - self.b1 ** self.t
        bc2_sqrt = (1 - self.b2 ** self.t) ** 0.5
        denom = np.sqrt(self.v) / f32(bc2_sqrt) + f32(self.eps)
        p -= f32(self.lr / bc1) * (self.m / denom)
        return p
