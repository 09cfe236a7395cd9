"""CPU restatement of the geometric-consistency loss (fp32 or fp64, torch CPU).

Follows, op for op:
  utils/geometry.py:9-19    pixel_grid
  utils/geometry.py:38-61   pixels_to_rays      ((x-cx)/fx, -(y-cy)/fy, -1)
  utils/geometry.py:86-100  pixels_to_points
  utils/geometry.py:103-128 reproject_points    (t_r + R_r p ; R_t^T (p - t_t))
  utils/geometry.py:64-83   project
  utils/geometry.py:201-208 sample              (grid_sample bilinear/border/align_corners=False)
  loss/consistency_loss.py:73-89   weighted_mean_loss
  loss/consistency_loss.py:98-208  geometry_consistency_loss
  loss/consistency_loss.py:210-253 ConsistencyLoss.__call__
  loss/joint_loss.py:26-47         JointLoss.__call__ (shape-(1,) loss)
Gradients come from torch.autograd on this restatement (what loss.backward()
does in depth_fine_tuning.py:282); `closed_form` is an independent numpy
re-derivation used to cross-check the backward.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned by
tests/golden/consistency_*.npz generated from the real reference.
"""
import numpy as np
import torch
import torch.nn.functional as F


def pixel_grid(batch_size, shape, dtype, device=None):
    # the reference creates these on utils.torch_helpers._device; here: wherever the depths live
    H, W = shape
    x = torch.linspace(0, W - 1, W, dtype=dtype, device=device)
    y = torch.linspace(0, H - 1, H, dtype=dtype, device=device)
    Y, X = torch.meshgrid(y, x, indexing="ij")
    return torch.stack((X, Y), dim=0)[None].expand(batch_size, -1, -1, -1)


def pixels_to_rays(pixels, intrinsics):
    B, _, H, W = pixels.shape
    uvs = pixels - intrinsics[:, 2:].view(-1, 2, 1, 1)
    uvs = torch.stack((uvs[:, 0], -uvs[:, 1]), dim=1)
    fxys = intrinsics[:, :2].view(-1, 2, 1, 1)
    return torch.cat((uvs / fxys, -torch.ones((B, 1, H, W), dtype=uvs.dtype, device=uvs.device)), dim=1)


def project(points, intrinsics):
    rays = points / -points[:, -1:]
    uvs = rays[:, :2] * intrinsics[:, :2].view(-1, 2, 1, 1)
    uvs = torch.stack((uvs[:, 0], -uvs[:, 1]), dim=1)
    return uvs + intrinsics[:, 2:].view(-1, 2, 1, 1)


def reproject_points(points_cam_ref, extrinsics_ref, extrinsics_tgt):
    B, _, H, W = points_cam_ref.shape
    R_ref, t_ref = extrinsics_ref[..., :3], extrinsics_ref[..., -1:]
    points_world = torch.baddbmm(t_ref, R_ref, points_cam_ref.reshape(B, 3, -1))
    R_tgt, t_tgt = extrinsics_tgt[..., :3], extrinsics_tgt[..., -1:]
    return torch.bmm(R_tgt.transpose(1, 2), points_world - t_tgt).view(B, 3, H, W)


def sample(data, uv):
    H, W = data.shape[2:]
    size = torch.tensor((W - 1, H - 1), dtype=uv.dtype, device=uv.device).view(1, -1, 1, 1)
    grid = (2 * uv / size - 1).permute(0, 2, 3, 1)
    return F.grid_sample(data, grid, padding_mode="border", align_corners=False)


def weighted_mean_loss(x, weights, eps=1e-6):
    B = weights.shape[0]
    weights_sum = torch.clamp(torch.sum(weights.view(B, -1), dim=-1).view(B, 1, 1, 1), min=eps)
    return torch.sum((weights / weights_sum * x).reshape(B, -1), dim=1)


def consistency_loss(depths, extrinsics, intrinsics, flows, masks,
                     lambda_reprojection=1.0, lambda_view_baseline=0.1):
    """depths (B,2,H,W) -> (loss shape (1,), {"reprojection": (B,), "disparity": (B,)})."""
    B, N, H, W = depths.shape
    dtype = depths.dtype
    pixels = pixel_grid(B * N, (H, W), dtype, depths.device)
    rays = pixels_to_rays(pixels, intrinsics.reshape(B * N, 4))
    points_cam = (rays * depths.reshape(B * N, 1, H, W)).reshape(B, N, 3, H, W)
    pixels = pixels.reshape(B, N, 2, H, W)
    reproj_losses, disp_losses = [], []
    for k in range(2):
        t = 1 - k
        points_cam_tgt = reproject_points(points_cam[:, k], extrinsics[:, k], extrinsics[:, t])
        matched = pixels[:, k] + flows[k]
        pixels_tgt = project(points_cam_tgt, intrinsics[:, t])
        if lambda_reprojection > 0:
            dist = torch.norm(pixels_tgt - matched, dim=1, keepdim=True)
            reproj_losses.append(weighted_mean_loss(torch.abs(dist), masks[k]))
        if lambda_view_baseline > 0:
            f = torch.mean(intrinsics[:, k, :2])
            warped = sample(points_cam[:, t], matched)
            disp_diff = 1.0 / points_cam_tgt[:, -1:] - 1.0 / warped[:, -1:]
            disp_losses.append(f * weighted_mean_loss(torch.abs(disp_diff), masks[k]))
    reproj = (lambda_reprojection * torch.mean(torch.stack(reproj_losses, -1), -1)
              if reproj_losses else torch.zeros(B, dtype=dtype, device=depths.device))
    disp = (lambda_view_baseline * torch.mean(torch.stack(disp_losses, -1), -1)
            if disp_losses else torch.zeros(B, dtype=dtype, device=depths.device))
    loss = torch.zeros(1, dtype=dtype, device=depths.device) + torch.mean(reproj + disp)
    return loss, {"reprojection": reproj, "disparity": disp}


def consistency_loss_and_grad(depths_np, batch, lambda_r=1.0, lambda_b=0.1, dtype=torch.float32):
    """numpy in / numpy out convenience wrapper: loss, per-pair losses, dL/d depth."""
    t = lambda a: torch.tensor(np.asarray(a), dtype=dtype)
    d = t(depths_np).requires_grad_(True)
    loss, meta = consistency_loss(d, t(batch["extrinsics"]), t(batch["intrinsics"]),
                                  [t(f) for f in batch["flows"]], [t(m) for m in batch["masks"]],
                                  lambda_r, lambda_b)
    loss.backward()
    return (loss.detach().numpy(), {k: v.detach().numpy() for k, v in meta.items()}, d.grad.numpy())


def closed_form(depths, extrinsics, intrinsics, flows, masks, lambda_r=1.0, lambda_b=0.1, B_global=None):
    """Independent float64 numpy evaluation of loss and dL/d depth (SURVEY.md §8(a) closed form)."""
    depths = np.asarray(depths, np.float64)
    B, _, H, W = depths.shape
    Bg = B_global or B
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    grad = np.zeros_like(depths)
    reproj = np.zeros(B); disp = np.zeros(B)
    for k in range(2):
        t = 1 - k
        f = float(np.mean(np.asarray(intrinsics, np.float64)[:, k, :2]))
        for b in range(B):
            fx, fy, cx, cy = np.asarray(intrinsics[b, k], np.float64)
            fxt, fyt, cxt, cyt = np.asarray(intrinsics[b, t], np.float64)
            Er, Et = np.asarray(extrinsics[b, k], np.float64), np.asarray(extrinsics[b, t], np.float64)
            M = Et[:, :3].T @ Er[:, :3]
            c = Et[:, :3].T @ (Er[:, 3] - Et[:, 3])
            ray = np.stack([(xx - cx) / fx, -(yy - cy) / fy, -np.ones_like(xx)], 0)
            m = np.einsum("ab,bhw->ahw", M, ray)
            d = depths[b, k]
            Q = d[None] * m + c[:, None, None]
            u = -fxt * Q[0] / Q[2] + cxt
            v = fyt * Q[1] / Q[2] + cyt
            mx = xx + np.asarray(flows[k][b, 0], np.float64)
            my = yy + np.asarray(flows[k][b, 1], np.float64)
            mk = np.asarray(masks[k][b, 0], np.float64)
            w = mk / max(mk.sum(), 1e-6)
            if lambda_r > 0:
                ex, ey = u - mx, v - my
                dist = np.sqrt(ex * ex + ey * ey)
                reproj[b] += 0.5 * lambda_r * np.sum(w * dist)
                du = -fxt * (m[0] * Q[2] - Q[0] * m[2]) / Q[2] ** 2
                dv = fyt * (m[1] * Q[2] - Q[1] * m[2]) / Q[2] ** 2
                safe = np.where(dist > 0, dist, 1.0)
                grad[b, k] += lambda_r / (2 * Bg) * w * np.where(dist > 0, (ex * du + ey * dv) / safe, 0.0)
            if lambda_b > 0:
                sx = np.clip(mx * W / (W - 1) - 0.5, 0, W - 1)
                sy = np.clip(my * H / (H - 1) - 0.5, 0, H - 1)
                x0 = np.floor(sx).astype(int); y0 = np.floor(sy).astype(int)
                wx1 = sx - x0; wy1 = sy - y0
                taps = [(x0, y0, (1 - wx1) * (1 - wy1)), (x0 + 1, y0, wx1 * (1 - wy1)),
                        (x0, y0 + 1, (1 - wx1) * wy1), (x0 + 1, y0 + 1, wx1 * wy1)]
                zw = np.zeros_like(sx)
                for tx, ty, wt in taps:
                    ok = (tx < W) & (ty < H)
                    zw += np.where(ok, wt * -depths[b, t][np.minimum(ty, H - 1), np.minimum(tx, W - 1)], 0.0)
                s = 1.0 / Q[2] - 1.0 / zw
                disp[b] += 0.5 * lambda_b * f * np.sum(w * np.abs(s))
                coef = lambda_b * f / (2 * Bg) * w * np.sign(s)
                grad[b, k] += coef * (-m[2] / Q[2] ** 2)
                for tx, ty, wt in taps:
                    ok = (tx < W) & (ty < H)
                    np.add.at(grad[b, t], (np.minimum(ty, H - 1)[ok], np.minimum(tx, W - 1)[ok]),
                              (-coef * wt / zw ** 2)[ok])
    loss = float(np.sum(reproj + disp) / Bg)
    return loss, reproj, disp, grad
