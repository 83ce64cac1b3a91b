"""Validation / test epochs of the Voice2Pose pipeline on the device.

Reference: ``Trainer.validate`` / ``Trainer.test`` (core/pipelines/trainer.py:407-457) loop ``Voice2Pose.test_step``
(voice2pose.py:333-384) over the test loader -- eval-mode forward, ``dataset.get_final_results`` x2, ``evaluate_step`` --
move EVERY result to the host with ``.cpu().numpy()``, concatenate the FGD features of the whole epoch there and finally
fit two Gaussians and take their Frechet distance (``evaluate_epoch`` voice2pose.py:432-446, core/utils/fgd.py).

Here the per-step work stays on the GPU: the eval forward runs through the same kernels (BatchNorm layers from running
statistics), the f64 final results + metrics come from ``sdt_pose_final_results`` / ``sdt_pose_metrics``, the loss sums of the
epoch are accumulated in a device vector, and the FGD features are reduced on the fly to the sufficient statistics of a
Gaussian (count, sum, sum of outer products, f64) -- 2 x (64 + 64 x 64) doubles instead of the epoch's feature matrices.  Only
``finish()`` touches the host: one small D2H copy and the 64 x 64 matrix square root (scipy, as in the reference).
"""
from collections import OrderedDict

import numpy as np
import torch

from . import ops


def mutiply_batch(batch, multiple):
    """trainer.py:343-353 (the reference's spelling): every sample repeated `multiple` times, whole batch tiled."""
    if isinstance(batch, dict):
        return {k: mutiply_batch(v, multiple) for k, v in batch.items() if not callable(v)}   # (drops data.py's "_release" hook)
    if isinstance(batch, list):
        return batch * multiple
    if isinstance(batch, torch.Tensor):
        return batch.unsqueeze(0).repeat_interleave(multiple, dim=0).reshape(multiple * batch.shape[0], *batch.shape[1:])
    raise NotImplementedError


class GaussianStats:
    """Streaming sufficient statistics of a D-dimensional sample on the device, f64: n, sum x, sum x x^T."""

    def __init__(self, dim, device):
        self.dim = dim
        self.n = 0
        self.s1 = torch.zeros(dim, dtype=torch.float64, device=device)
        self.s2 = torch.zeros(dim, dim, dtype=torch.float64, device=device)

    def add(self, x):
        x = x.detach().reshape(-1, self.dim).double()
        self.n += x.shape[0]
        self.s1 += x.sum(0)
        self.s2.addmm_(x.t(), x)

    def mean_cov(self):
        """mean and the UNBIASED covariance (np.cov default, fgd.py:63) as numpy f64."""
        assert self.n >= 2, "need at least two samples for a covariance"
        s1, s2 = self.s1.cpu().numpy(), self.s2.cpu().numpy()
        mean = s1 / self.n
        cov = (s2 - self.n * np.outer(mean, mean)) / (self.n - 1)
        return mean, cov


def _sqrtm(m):
    """scipy.linalg.sqrtm across scipy versions (the reference calls it with disp=False, removed in recent releases)."""
    from scipy import linalg
    try:
        out = linalg.sqrtm(m, disp=False)
        return np.asarray(out[0] if isinstance(out, tuple) else out)
    except TypeError:
        return np.asarray(linalg.sqrtm(m))


def frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """d^2 = |mu1 - mu2|^2 + Tr(C1 + C2 - 2 sqrt(C1 C2))   (core/utils/fgd.py:6-58, incl. its singular-product fallback)."""
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    assert mu1.shape == mu2.shape and sigma1.shape == sigma2.shape
    diff = mu1 - mu2
    covmean = _sqrtm(sigma1.dot(sigma2))
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = _sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))


class Voice2PoseEvaluator:
    """``validate`` / ``test`` for a drop-in ``Voice2PoseModel`` (or a ``Voice2PoseTrainer``'s model).

        ev = Voice2PoseEvaluator(model, test_batch_size=cfg.TEST.BATCH_SIZE, multiple=cfg.TEST.MULTIPLE)
        for batch in test_loader: ev.step(batch)
        metrics = ev.finish(num_test_samples)        # losses / L2_dist / lip_sync_error_n epoch means, FGD_mu, FGD_mu_logvar
    """

    LOSS_KEYS = ("G_reg_loss", "G_loss", "L2_dist", "lip_sync_error_n")

    def __init__(self, model, test_batch_size, multiple=1, dataset=None):
        assert isinstance(multiple, int) and multiple >= 1, "TEST.MULTIPLE should be an integer >= 1 (voice2pose.py:338-340)"
        self.model, self.bs, self.multiple, self.dataset = model, int(test_batch_size), multiple, dataset
        self.device = next(model.parameters()).device
        self.reset()

    def reset(self):
        self.sums = torch.zeros(len(self.LOSS_KEYS), dtype=torch.float64, device=self.device)
        self.code_dim = None
        self.pred = self.gt = None
        self.steps = 0

    @torch.no_grad()
    def step(self, batch):
        """One ``test_step`` (voice2pose.py:333-384) without the host round trips.  Returns the step's (losses, results) with
        device tensors, in case the caller wants to log or save them like the reference does."""
        was_training = self.model.training
        self.model.eval()                                   # trainer.py:412
        try:
            if self.multiple > 1:
                batch = mutiply_batch(batch, self.multiple)
            losses, results = self.model(batch, self.dataset)
        finally:
            self.model.train(was_training)
        st = batch["speaker_stat"]
        dev = self.device
        mean, std = torch.as_tensor(st["mean"]).to(dev).double().contiguous(), torch.as_tensor(st["std"]).to(dev).double().contiguous()
        scale = torch.as_tensor(st["scale_factor"]).to(dev).double().contiguous()
        hier = bool(self.model.cfg.DATASET.HIERARCHICAL_POSE)
        pred = results["poses_pred_batch"].detach().float().contiguous()
        gt = results["poses_gt_batch"].detach().float().contiguous()
        results["poses_pred_batch"] = ops.pose_final_results(pred, mean, std, scale, hier)     # get_final_results, f64
        results["poses_gt_batch"] = ops.pose_final_results(gt, mean, std, scale, hier)
        met = ops.pose_metrics(results["poses_pred_batch"], results["poses_gt_batch"])          # evaluate_step
        losses = OrderedDict(losses)
        losses["L2_dist"], losses["lip_sync_error_n"] = met[0], met[1]
        # batch_losses = mean * TEST.BATCH_SIZE (voice2pose.py:378), summed over the epoch on the device
        vals = torch.stack([losses[k].detach().double().reshape(()) for k in self.LOSS_KEYS])
        self.sums += vals * self.bs
        if "mu_pred" in results:
            d = results["mu_pred"].shape[-1]
            if self.pred is None:
                self.code_dim = d
                self.pred, self.gt = GaussianStats(2 * d, dev), GaussianStats(2 * d, dev)
            self.pred.add(torch.cat([results["mu_pred"], results["logvar_pred"]], 1))
            self.gt.add(torch.cat([results["mu_gt"], results["logvar_gt"]], 1))
        self.steps += 1
        return losses, results

    def finish(self, num_test_samples):
        """Epoch means (trainer.py:423) + ``evaluate_epoch`` (voice2pose.py:432-446).  The only host synchronisation."""
        out = OrderedDict((k, v / num_test_samples) for k, v in zip(self.LOSS_KEYS, self.sums.cpu().tolist()))
        if self.pred is not None:
            d = self.code_dim
            mp, cp = self.pred.mean_cov()
            mg, cg = self.gt.mean_cov()
            out["FGD_mu"] = frechet_distance(mp[:d], cp[:d, :d], mg[:d], cg[:d, :d])      # the mu block of [mu | logvar]
            out["FGD_mu_logvar"] = frechet_distance(mp, cp, mg, cg)
        return out
