"""DepthFineTuner — drop-in for the reference's depth_fine_tuning.py:28-406.

Same constructor, attributes (.out_dir, .model), `save_depth`, `fine_tune(writer=None)`, CLI flags
(`DepthFineTuningParams.add_arguments`) and on-disk products (checkpoints/%04d.pth, eval/*, depth/*.raw).
The inner loop (:261-293) is replaced by `FineTuneStep`: the whole step is one CUDA-graph replay, the NaN
guard runs on the device and the loss scalar is read back one iteration late, so the host never stalls the
GPU (the reference synchronises three times per iteration, :277,278,288).
"""
import json
import os
import time
from os.path import join as pjoin

import torch
from torch.utils.data import DataLoader

from . import optimizer
from .distributed import global_focal, shard_slice
from .fine_tune_step import FineTuneStep
from .loaders.resident import ResidentLoader, ResidentVideoDataset
from .loaders.video_dataset import VideoDataset, VideoFrameDataset
from .loss.joint_loss import JointLoss
from .loss.loss_params import LossParams
from .monodepth.depth_model_registry import get_depth_model
from .utils import image_io
from .utils.torch_helpers import to_device


class DepthFineTuningParams:
    @staticmethod
    def add_arguments(parser):
        parser = LossParams.add_arguments(parser)
        parser.add_argument("--optimizer", default="Adam", choices=optimizer.OPTIMIZER_NAMES)
        parser.add_argument("--val_epoch_freq", type=int, default=1)
        parser.add_argument("--learning_rate", type=float, default=0)
        parser.add_argument("--batch_size", type=int, default=4)
        parser.add_argument("--num_epochs", type=int, default=20)
        parser.add_argument("--log_dir")
        parser.add_argument("--display_freq", type=int, default=100)
        parser.add_argument("--print_freq", type=int, default=1)
        parser.add_argument("--save_epoch_freq", type=int, default=1)
        return parser


def log_loss_stats(writer, name_prefix, loss_meta, n, log_histogram=False):
    """max / min / mean of every sub-loss (reference :66-91)."""
    for sub, value in loss_meta.items():
        full = name_prefix + "/" + sub
        writer.add_scalar(full + "/max", value.max(), n)
        writer.add_scalar(full + "/min", value.min(), n)
        writer.add_scalar(full + "/mean", value.mean(), n)
        if log_histogram:
            writer.add_histogram(full, value, n)


def write_summary(writer, mode_name, input_images, depth, masks, n_iter):
    """Image grids of the inputs, the predicted depth and the masks (reference :94-117)."""
    import torchvision.utils as vutils
    B = depth.shape[0]
    pred = depth.unsqueeze(-3)
    mask = torch.stack(masks, dim=1)

    def to_vis(x):
        return x[:8].transpose(0, 1).reshape((-1,) + x.shape[-3:])

    writer.add_image(mode_name + "/image", vutils.make_grid(to_vis(input_images), nrow=B, normalize=True), n_iter)
    writer.add_image(mode_name + "/pred_full", vutils.make_grid(to_vis(1.0 / pred), nrow=B, normalize=True), n_iter)
    writer.add_image(mode_name + "/mask", vutils.make_grid(to_vis(mask), nrow=B, normalize=True), n_iter)


def log_loss(writer, mode_name, loss, loss_meta, niters):
    """Main loss scalar + sub-loss statistics (reference :120-126)."""
    main = mode_name + "/loss"
    writer.add_scalar(main, loss, niters)
    log_loss_stats(writer, main, loss_meta, niters)


def make_tag(params):
    return (LossParams.make_str(params) + f"_LR{params.learning_rate}" + f"_BS{params.batch_size}"
            + f"_O{params.optimizer.lower()}")


class DepthFineTuner:
    def __init__(self, range_dir, frames, params):
        self.frames, self.params = frames, params
        self.base_dir, self.range_dir = params.path, range_dir
        model_cls = get_depth_model(params.model_type)
        # params.py:110-119 resolves these from the model class when the CLI left them at their sentinels
        if getattr(params, "learning_rate", 0) <= 0:
            params.learning_rate = model_cls.learning_rate
        if getattr(params, "lambda_view_baseline", -1) < 0:
            params.lambda_view_baseline = model_cls.lambda_view_baseline
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        # the run directory is tagged with the batch size BEFORE the multi-GPU scaling, as the reference does (:143 vs :155-159)
        self.out_dir = pjoin(self.range_dir, make_tag(params))
        os.makedirs(self.out_dir, exist_ok=True)
        print(f"Fine-tuning directory: '{self.out_dir}'")
        self.checkpoints_dir = pjoin(self.out_dir, "checkpoints")
        os.makedirs(self.checkpoints_dir, exist_ok=True)
        if self.world > 1:
            # one process per GPU (torchrun): bind this rank's device BEFORE the model allocates, and join the NCCL group
            # that FineTuneStep's gradient all-reduce uses (gloo when no GPU is present: the CPU tests)
            import torch.distributed as dist
            if torch.cuda.is_available():
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", self.rank)) % torch.cuda.device_count())
            if not dist.is_initialized():
                if "MASTER_ADDR" not in os.environ:
                    raise RuntimeError("WORLD_SIZE > 1 but no rendezvous: launch with torchrun (MASTER_ADDR/MASTER_PORT/RANK) "
                                       "or call torch.distributed.init_process_group before constructing DepthFineTuner")
                # NCCL over NVLink; CVD_DIST_BACKEND=gloo lets two ranks share ONE GPU (single-GPU test boxes)
                dist.init_process_group(os.environ.get("CVD_DIST_BACKEND", "nccl" if torch.cuda.is_available() else "gloo"))
        self.model = model_cls()
        print(f"Using {self.world} GPUs.")
        if self.world > 1:                    # reference: batch_size *= num_gpus (:155-159)
            self.params.batch_size *= self.world
            print(f"Adjusting batch size to {self.params.batch_size}.")
        self.vis_depth_scale = None

    def save_depth(self, dir=None, frames=None):
        dir = self.out_dir if dir is None else dir
        frames = self.frames if frames is None else frames
        color_fmt = pjoin(self.base_dir, "color_down", "frame_{:06d}.raw")
        depth_dir = pjoin(dir, "depth")
        os.makedirs(depth_dir, exist_ok=True)
        loader = DataLoader(VideoFrameDataset(color_fmt, frames), batch_size=1, shuffle=False, num_workers=0)
        self.model.eval()                                  # BN running statistics (:182)
        for images, meta in loader:
            frame_id = int(meta["frame_id"][0])
            with torch.no_grad():
                depth = self.model.forward(to_device(images), None)
            inv_depth = 1.0 / depth.detach().cpu().numpy().squeeze()
            image_io.save_raw_float32_image(pjoin(depth_dir, f"frame_{frame_id:06d}.raw"), inv_depth)

    def fine_tune(self, writer=None):
        P = self.params
        dataset = VideoDataset(self.base_dir, pjoin(self.range_dir, "metadata_scaled.npz"))
        pin = torch.cuda.is_available()
        B_global = P.batch_size
        B_local = B_global // self.world
        gen = torch.Generator().manual_seed(0)              # same shuffle on every rank; ranks take disjoint slices
        if getattr(P, "resident_dataset", True):
            # the whole clip lives in HBM (340 MB at 224x384 / 50 frames); same batches in the same order
            # as the DataLoader below (tests/test_resident_loader_cpu.py), no per-step file reads, PNG decodes or H2D copies
            resident = ResidentVideoDataset(dataset, self.model.device_)
            train_loader = ResidentLoader(resident, B_global, shuffle=True, generator=gen)
            val_loader = ResidentLoader(resident, B_local, shuffle=False)
        else:
            train_loader = DataLoader(dataset, batch_size=B_global, shuffle=True, num_workers=4, pin_memory=pin, generator=gen)
            val_loader = DataLoader(dataset, batch_size=B_local, shuffle=False, num_workers=4, pin_memory=pin)
        criterion = JointLoss(P)
        eval_dir = pjoin(self.out_dir, "eval")
        os.makedirs(eval_dir, exist_ok=True)
        self.model.train()
        images0, _ = dataset[0]
        H, W = images0.shape[-2:]
        steps = {}
        # parameters_init (reference :223-224): the weights before fine-tuning, for the lambda_parameter regulariser
        p_init = self.model.P.flat.detach().clone() if getattr(P, "lambda_parameter", 0.0) > 0 else None

        def get_step(b, b_global):
            key = (b, b_global)
            if key not in steps:
                first = next(iter(steps.values())) if steps else None
                steps[key] = FineTuneStep(self.model, b, H, W, lr=P.learning_rate, lambda_reprojection=P.lambda_reprojection,
                                          lambda_view_baseline=P.lambda_view_baseline, world_size=self.world, B_global=b_global,
                                          lambda_parameter=getattr(P, "lambda_parameter", 0.0), parameters_init=p_init, rank=self.rank)
                if first is not None:                            # all batch shapes share ONE Adam state
                    s = steps[key]
                    s.exp_avg, s.exp_avg_sq, s.adam_state = first.exp_avg, first.exp_avg_sq, first.adam_state
                    s.p_init = first.p_init
            return steps[key]

        def validate(epoch, niters):
            meta = self.eval_and_save(criterion, val_loader, f"_e{epoch:04d}_iter{niters:06d}")
            if writer is not None:
                for name, v in meta.items():
                    writer.add_scalar(f"validation/{name}/mean", v.mean(), epoch)
            print(f"Done Validation for epoch {epoch} ({niters} iterations)")

        self.vis_depth_scale = None
        validate(0, 0)
        total_iters = 0
        # Logging (reference :277-293): the loss scalar, the per-sub-loss statistics (log_loss_stats :66-91) and the image
        # summary (:94-117) of iteration i are read back while iteration i+1 runs, so the host never stalls the GPU.
        pending = None

        def flush(p):
            if p is None:
                return 0
            e, pr, l, meta, n_before, nb_, imgs = p
            lv = float(l)
            print(f"Epoch = {e}, pairs = {pr}, loss = {lv}")
            if lv != lv:                                          # :278-280 (the update was skipped on the device)
                print("Loss is NaN. Skipping.")
                return 0
            n_after = n_before + nb_
            if writer is not None and self.rank == 0:
                if n_after % P.print_freq == 0:
                    log_loss(writer, "Train", l, meta, n_after)
                if imgs is not None:
                    write_summary(writer, "Train", *imgs, n_after)
            return nb_

        for epoch in range(P.num_epochs):
            t0 = time.perf_counter()
            for images, metadata in train_loader:
                nb = images.shape[0]
                sl = shard_slice(nb, self.rank, self.world)      # ragged last batch: uneven shares, possibly none
                bl = sl.stop - sl.start
                geom = metadata["geometry_consistency"]
                pairs = geom["indices"][:nb].tolist()
                if bl == 0:
                    loss = next(iter(steps.values())).step_empty()
                    total_iters += flush(pending)
                    pending = (epoch, pairs, loss.clone(), {}, total_iters, nb, None)
                    continue
                step = get_step(bl, nb)
                f_dir = None
                if self.world > 1:
                    f_dir = global_focal(metadata["intrinsics"], nb)
                step.load_batch(images[sl], [f[sl] for f in geom["flows"]], [m[sl] for m in geom["masks"]],
                                metadata["extrinsics"][sl], metadata["intrinsics"][sl], f_dir)
                loss = step.step()
                n_prev = flush(pending)                          # the previous step, while this one runs
                total_iters += n_prev
                meta = step.loss_meta()                          # device clones, no sync
                imgs = None
                if writer is not None and self.rank == 0 and (total_iters + nb) % P.display_freq == 0:
                    imgs = (step.images.clone(), step.depth().clone(), [m.clone() for m in step.masks])
                pending = (epoch, pairs, loss.clone(), meta, total_iters, nb, imgs)
            total_iters += flush(pending)
            pending = None
            print(f"Epoch {epoch} took {time.perf_counter() - t0:.2f}s.")
            if (epoch + 1) % P.val_epoch_freq == 0:
                validate(epoch + 1, total_iters)
            if (epoch + 1) % P.save_epoch_freq == 0 and self.rank == 0:
                self.model.save(pjoin(self.checkpoints_dir, f"{epoch + 1:04d}.pth"))
        if P.num_epochs % P.val_epoch_freq != 0:
            validate(P.num_epochs, total_iters)
        print("Finished Training")

    def eval_and_save(self, criterion, data_loader, suf):
        """Forward + loss over all pairs under no_grad but in TRAIN mode (BN batch statistics, running stats
        keep updating) exactly as the reference does (:312-406); writes eval/depth_*.raw and loss*.json."""
        loss_dict, saved, all_pairs = {}, set(), []
        for images, metadata in data_loader:
            metadata = to_device(metadata)
            with torch.no_grad():
                depth = self.model(to_device(images), metadata)
                _, loss_meta = criterion(depth, metadata, parameters=None)
            idx = metadata["geometry_consistency"]["indices"].cpu().numpy().tolist()
            all_pairs += idx
            for name, losses in loss_meta.items():
                for pr, l in zip(idx, losses):
                    loss_dict.setdefault(name, {})[str(pr)] = float(l)
            inv = 1.0 / depth.cpu().numpy()
            for invs, pr in zip(inv, idx):
                for inv_depth, index in zip(invs, pr):
                    if index in saved or self.rank != 0:
                        continue
                    saved.add(index)
                    image_io.save_raw_float32_image(pjoin(self.out_dir, "eval", f"depth_{index:06d}{suf}.raw"), inv_depth)
        loss_meta = {name: torch.tensor(tuple(v.values())) for name, v in loss_dict.items()}
        loss_dict["mean"] = {k: float(v.mean()) for k, v in loss_meta.items()}
        if self.rank == 0:
            with open(pjoin(self.out_dir, "eval", f"loss{suf}.json"), "w") as f:
                json.dump(loss_dict, f)
        return loss_meta
