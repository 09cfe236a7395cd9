"""One fused fine-tuning step: the body of the reference's inner loop (depth_fine_tuning.py:264-283)

    depth = model(stacked_img, metadata); opt.zero_grad(); loss, meta = criterion(depth, metadata, ...)
    if isnan(loss): continue; loss.backward(); opt.step()

as a static sequence of C-ABI launches captured ONCE in a CUDA graph and replayed per mini-batch:
    zero flat grads -> hourglass forward (tcgen05 convs) -> mask sums -> fused reproject+consistency
    loss fwd+bwd -> hourglass backward -> [NCCL all-reduce of the flat gradient + loss] -> fused Adam
The NaN guard runs on the device (the Adam kernel skips the update when the loss is NaN), so the host
never has to synchronise inside the loop; scalars for logging are read back asynchronously.

Multi-GPU (one process per GPU, frame pairs sharded, reference semantics of nn.DataParallel:
batch_size *= num_gpus, BatchNorm statistics per replica): each rank runs the graph on its B_local pairs
with the loss normalised by B_global and the batch-mean focal length of the GLOBAL batch, then ONE
all-reduce(sum) of [flat gradient | local loss] over NVLink, then Adam on every rank.
"""
import ctypes as C

import torch

from . import _lib, ops
from .distributed import allreduce_flat
from .utils.geometry import _Workspace


class FineTuneStep:
    def __init__(self, model, B_local, H, W, lr, lambda_reprojection=1.0, lambda_view_baseline=None,
                 betas=(0.9, 0.999), eps=1e-8, world_size=1, process_group=None, use_graph=True, B_global=None,
                 lambda_parameter=0.0, parameters_init=None, rank=0):
        self.model, self.B, self.H, self.W = model, B_local, H, W
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        self.lam_r = float(lambda_reprojection)
        self.lam_b = float(model.lambda_view_baseline if lambda_view_baseline is None else lambda_view_baseline)
        self.world, self.pg, self.rank = world_size, process_group, rank
        # lambda_parameter regulariser (loss/parameter_loss.py:13-19, joint_loss.py:34-39): lambda * sum |p - p0| added to the
        # loss and lambda * sign(p - p0) to the flat gradient by ONE launch over the flat parameter buffer.  The reference
        # evaluates it once per step (on GPU 0 under DataParallel), so with several ranks only rank 0 contributes it.
        self.lam_p = float(lambda_parameter)
        self.p_init = None
        if self.lam_p > 0:
            self.p_init = (parameters_init if parameters_init is not None else model.P.flat).detach().clone()
            assert self.p_init.numel() == model.P.n_flat, "parameters_init must be the flat parameter buffer"
        # pairs of the GLOBAL mini-batch (the 1/B of consistency_loss.py:208); ranks may hold unequal shares of a ragged batch
        self.B_global = B_local * world_size if B_global is None else int(B_global)
        dev = model.device_
        self.dev = dev
        self.engine = model.engine(2 * B_local, H, W)
        P = model.P
        # static inputs (the graph reads these addresses)
        z = lambda *s: torch.zeros(*s, device=dev)
        self.images = z(B_local, 2, 3, H, W)
        self.flows = [z(B_local, 2, H, W), z(B_local, 2, H, W)]
        self.masks = [z(B_local, 1, H, W), z(B_local, 1, H, W)]
        self.extr, self.intr = z(B_local, 2, 3, 4), z(B_local, 2, 4)
        # global-batch mean focal length (consistency_loss.py:178) for world > 1: a DEVICE buffer the kernels read at run
        # time, so the captured graph stays valid when the intrinsics change from batch to batch
        self.f_dir = torch.zeros(2, device=dev) if world_size > 1 else None
        self._f_host = torch.zeros(64, 2).pin_memory() if world_size > 1 else None    # ring of pinned staging slots
        self._f_slot = 0
        # the loss lives in the tail slot of the flat gradient storage so a single all-reduce carries both
        self.loss = P.loss_slot
        self.pair_losses = z(2, B_local)
        self.grad_depth = z(B_local, 2, H, W)
        self.ws = _Workspace.get(dev, B_local)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(P.flat), torch.zeros_like(P.flat)
        self.adam_state = torch.zeros(4, dtype=torch.int32, device=dev)
        self.graph = None
        self.use_graph = use_graph
        self.launches_per_step = 0

    # ------------------------------------------------------------------ pieces
    def _fwd_bwd(self):
        L = _lib.lib()
        P, eng, B, H, W = self.model.P, self.engine, self.B, self.H, self.W
        st = _lib.stream()
        P.grad_flat.zero_()
        eng.train_mode = True
        depth = eng.forward(self.images.view(2 * B, 3, H, W))          # (2B,H,W) == (B,2,H,W)
        _lib.check(L.cvd_mask_sums(_lib.ptr(self.masks[0]), _lib.ptr(self.masks[1]), B, H, W, _lib.ptr(self.ws["msum"]), st),
                   "cvd_mask_sums")
        _lib.check(L.cvd_consistency_fwd_bwd(
            _lib.ptr(depth), _lib.ptr(self.flows[0]), _lib.ptr(self.flows[1]), _lib.ptr(self.masks[0]), _lib.ptr(self.masks[1]),
            _lib.ptr(self.extr), _lib.ptr(self.intr), _lib.ptr(self.ws["msum"]), None, _lib.ptr(self.f_dir),
            C.c_float(self.lam_r), C.c_float(self.lam_b), B, self.B_global, H, W, _lib.ptr(self.ws["acc"]),
            _lib.ptr(self.pair_losses), _lib.ptr(self.loss), _lib.ptr(self.grad_depth), st), "cvd_consistency_fwd_bwd")
        eng.backward(self.grad_depth.view(2 * B, H, W))
        if self.lam_p > 0 and self.rank == 0:
            _lib.check(L.cvd_param_l1(_lib.ptr(P.flat), _lib.ptr(self.p_init), C.c_longlong(P.n_flat), C.c_float(self.lam_p),
                                      _lib.ptr(P.grad_flat), _lib.ptr(self.loss), st), "cvd_param_l1")

    def _adam(self, flag):
        P = self.model.P
        _lib.check(_lib.lib().cvd_adam_flat(
            _lib.ptr(P.flat), _lib.ptr(P.grad_flat), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
            C.c_longlong(P.n_flat), C.c_float(self.lr), C.c_float(self.betas[0]), C.c_float(self.betas[1]),
            C.c_float(self.eps), C.c_float(1.0), _lib.ptr(self.adam_state), _lib.ptr(flag), _lib.stream()), "cvd_adam_flat")

    def _body_single(self):
        self._fwd_bwd()
        self._adam(self.loss)

    # ------------------------------------------------------------------ public
    def load_batch(self, images, flows, masks, extrinsics, intrinsics, f_dir=None):
        """Copy one collated mini-batch (host pinned or device tensors, reference layout) into the static buffers."""
        self.images.copy_(images, non_blocking=True)
        self.flows[0].copy_(flows[0], non_blocking=True); self.flows[1].copy_(flows[1], non_blocking=True)
        self.masks[0].copy_(masks[0], non_blocking=True); self.masks[1].copy_(masks[1], non_blocking=True)
        self.extr.copy_(extrinsics, non_blocking=True); self.intr.copy_(intrinsics, non_blocking=True)
        if self.f_dir is not None:
            if f_dir is None:
                raise _lib.CvdError("world_size > 1: load_batch needs the global batch's mean focal length (f_dir)")
            if torch.is_tensor(f_dir):
                self.f_dir.copy_(f_dir.reshape(2), non_blocking=True)
            else:
                slot = self._f_host[self._f_slot % 64]           # a slot is reused only 64 batches later
                self._f_slot += 1
                slot[0], slot[1] = float(f_dir[0]), float(f_dir[1])
                self.f_dir.copy_(slot, non_blocking=True)

    def step(self):
        """Run one fine-tuning step on the loaded batch. Returns the device loss tensor (shape (1,), no sync)."""
        if self.world > 1:
            return self._step_distributed()
        if not self.use_graph:
            self._body_single()
            return self.loss
        if self.graph is None:
            n0 = _lib.launch_count()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):                    # warm-up outside capture (lazy attribute setup)
                self._snapshot_and_restore(self._body_single)
            torch.cuda.current_stream().wait_stream(s)
            self.launches_per_step = _lib.launch_count() - n0
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._body_single()
        self.graph.replay()
        self.model.P.num_batches_tracked += 1          # the replayed forward ran in train mode (BN momentum update)
        return self.loss

    def _snapshot_and_restore(self, fn):
        """Run fn once without leaving a trace in the optimisation state (used for the pre-capture warm-up)."""
        P = self.model.P
        keep = (P.flat.clone(), P.buf_flat.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.adam_state.clone(),
                P.num_batches_tracked)
        fn()
        P.flat.copy_(keep[0]); P.buf_flat.copy_(keep[1]); self.exp_avg.copy_(keep[2]); self.exp_avg_sq.copy_(keep[3])
        self.adam_state.copy_(keep[4]); P.num_batches_tracked = keep[5]

    def _step_distributed(self):
        P = self.model.P
        if self.graph is None and self.use_graph:
            n0 = _lib.launch_count()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._snapshot_and_restore(self._fwd_bwd)
            torch.cuda.current_stream().wait_stream(s)
            self.launches_per_step = _lib.launch_count() - n0 + 2
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._fwd_bwd()
        if self.use_graph:
            self.graph.replay()
            P.num_batches_tracked += 1
        else:
            self._fwd_bwd()
        # ONE all-reduce over NVLink: [flat gradient | local loss (already divided by B_global)], in place
        allreduce_flat(P.grad_store, self.pg)
        self._adam(self.loss)
        return self.loss

    def loss_meta(self):
        """The reference's per-sample sub-loss dict of the LAST step (joint_loss.py:41-46, consistency_loss.py:195-207):
        device clones, no synchronisation.  `parameter_loss` is the total minus the consistency terms."""
        pl = self.pair_losses.clone()
        meta = {"reprojection": pl[0], "disparity": pl[1]}
        if self.lam_p > 0:
            meta["parameter_loss"] = (self.loss.clone() - pl.sum() / self.B_global).reshape(1, 1)
        return meta

    def depth(self):
        """(B,2,H,W) depth of the last step (engine-owned buffer)."""
        return self.engine.depth.view(self.B, 2, self.H, self.W)

    def step_empty(self):
        """This rank owns no pair of the (ragged) global mini-batch: contribute zeros to the all-reduce, apply the same
        Adam update as every other rank (the reduced loss drives the same NaN guard)."""
        assert self.world > 1
        P = self.model.P
        P.grad_store.zero_()
        allreduce_flat(P.grad_store, self.pg)
        self._adam(self.loss)
        return self.loss
