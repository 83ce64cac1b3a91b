"""Validation epoch on the device (speechdrivestemplates_b200/evaluation.py) against the reference's host-side recipe:
per-step eval forward with the oracle, features concatenated on the host, np.cov + scipy sqrtm (core/utils/fgd.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from util import oliver_stat  # noqa: E402


def _fgd_host(a, b):
    from scipy import linalg
    mu_a, mu_b = a.mean(0), b.mean(0)
    sa, sb = np.cov(a, rowvar=False), np.cov(b, rowvar=False)
    covmean = np.asarray(linalg.sqrtm(sa.dot(sb)))
    return float((mu_a - mu_b).dot(mu_a - mu_b) + np.trace(sa) + np.trace(sb) - 2 * np.trace(covmean.real))


def test_gaussian_stats_and_frechet_distance_match_numpy_scipy():
    from speechdrivestemplates_b200 import evaluation as E
    g = torch.Generator().manual_seed(5)
    a = torch.randn(300, 64, generator=g, dtype=torch.float64) * torch.linspace(0.5, 2.0, 64, dtype=torch.float64) + 0.3
    b = torch.randn(280, 64, generator=g, dtype=torch.float64) @ torch.randn(64, 64, generator=g, dtype=torch.float64) * 0.2
    sa, sb = E.GaussianStats(64, "cuda:0"), E.GaussianStats(64, "cuda:0")
    for chunk in a.split(37):
        sa.add(chunk.cuda().float())                       # the features are f32 on the device
    for chunk in b.split(41):
        sb.add(chunk.cuda().float())
    ma, ca = sa.mean_cov()
    mb, cb = sb.mean_cov()
    a32, b32 = a.float().double().numpy(), b.float().double().numpy()
    assert np.allclose(ma, a32.mean(0), rtol=0, atol=1e-12) and np.allclose(ca, np.cov(a32, rowvar=False), rtol=1e-10, atol=1e-12)
    ref = _fgd_host(a32, b32)
    assert abs(E.frechet_distance(ma, ca, mb, cb) - ref) < 1e-8 * max(1.0, abs(ref))
    assert abs(E.frechet_distance(ma[:32], ca[:32, :32], mb[:32], cb[:32, :32]) - _fgd_host(a32[:, :32], b32[:, :32])) < 1e-8 * max(1.0, ref)


def test_validation_epoch_matches_the_host_side_recipe():
    from oracle import sdt_oracle as O
    from speechdrivestemplates_b200 import config, evaluation as E, pipeline
    st = oliver_stat()
    n_train, bs, steps = 32, 6, 4
    cfg = config.get_cfg("voice2pose_sdt_bp")
    torch.manual_seed(0)
    model = pipeline.Voice2PoseModel(cfg, num_train_samples=n_train).cuda()
    model.clips_code.data.copy_(0.1 * torch.randn(n_train, 32, generator=torch.Generator().manual_seed(11)))
    # give the BatchNorm layers of the FGD extractor non-trivial running statistics
    for name, buf in model.pose_encoder.named_buffers():
        if name.endswith("running_mean"):
            buf.copy_(0.05 * torch.randn(buf.shape, generator=torch.Generator().manual_seed(len(name))))
        if name.endswith("running_var"):
            buf.copy_(1.0 + 0.2 * torch.rand(buf.shape, generator=torch.Generator().manual_seed(len(name) + 1)))
    ev = E.Voice2PoseEvaluator(model, test_batch_size=bs, multiple=1)
    feats = {k: [] for k in ("mu_pred", "logvar_pred", "mu_gt", "logvar_gt")}
    sums = np.zeros(4)
    for i in range(steps):
        b = O.synthetic_batch(bs, n_train, st, seed=900 + i)
        hb = dict(b)
        hb["speaker_stat"] = {k: torch.from_numpy(np.asarray(v)) for k, v in b["speaker_stat"].items()}
        losses, results = ev.step(hb)
        # host-side recipe on the SAME forward outputs: .cpu().numpy(), mean * batch size, concatenate
        for k in feats:
            feats[k].append(results[k].detach().cpu().numpy())
        sums += np.array([float(losses[k]) for k in E.Voice2PoseEvaluator.LOSS_KEYS]) * bs
        # the step's final results equal the oracle's get_final_results on the same prediction, bit for bit
        if i == 0:
            pred_n = None
    out = ev.finish(bs * steps)
    for j, k in enumerate(E.Voice2PoseEvaluator.LOSS_KEYS):
        assert abs(out[k] - sums[j] / (bs * steps)) < 1e-9 * max(1.0, abs(out[k])), k
    cat = {k: np.concatenate(v, 0).astype(np.float64) for k, v in feats.items()}
    ref_mu = _fgd_host(cat["mu_pred"], cat["mu_gt"])
    ref_all = _fgd_host(np.concatenate([cat["mu_pred"], cat["logvar_pred"]], 1), np.concatenate([cat["mu_gt"], cat["logvar_gt"]], 1))
    assert abs(out["FGD_mu"] - ref_mu) < 1e-6 * max(1.0, abs(ref_mu)), (out["FGD_mu"], ref_mu)
    assert abs(out["FGD_mu_logvar"] - ref_all) < 1e-6 * max(1.0, abs(ref_all)), (out["FGD_mu_logvar"], ref_all)
    assert model.training                                  # the evaluator restores the mode


def test_multiple_replicates_the_batch_like_the_reference():
    from speechdrivestemplates_b200 import evaluation as E
    b = {"audio": torch.arange(6.).view(3, 2), "speaker": ["a", "b", "c"], "speaker_stat": {"mean": torch.arange(3.).view(3, 1)}}
    m = E.mutiply_batch(b, 2)
    assert m["audio"].shape == (6, 2) and torch.equal(m["audio"][:3], b["audio"]) and torch.equal(m["audio"][3:], b["audio"])
    assert m["speaker"] == ["a", "b", "c", "a", "b", "c"] and m["speaker_stat"]["mean"].shape == (6, 1)
