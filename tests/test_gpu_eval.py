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
    ev_pred_inputs = []
    hook = model.netG.register_forward_hook(lambda mod, args, out: ev_pred_inputs.append(out.detach().float().cpu().numpy()))
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
        # the step's f64 final results equal the oracle's get_final_results on the same fp32 prediction, bit for bit
        if i == 0:
            pred32 = ev_pred_inputs[-1]
            ref_fin = O.get_final_results(pred32, b["speaker_stat"]["mean"], b["speaker_stat"]["std"], b["speaker_stat"]["scale_factor"], True)
            assert np.array_equal(results["poses_pred_batch"].cpu().numpy(), ref_fin)
    hook.remove()
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


def _eval_model(tag, name, opts):
    """Drop-in Voice2PoseModel built like the fixture's reference model: torch.manual_seed(0) + constructor, then the stored
    BatchNorm buffers."""
    from speechdrivestemplates_b200 import config, pipeline
    from util import golden
    g = golden("eval_golden")
    n_train = int(g[tag + "/n_train"])
    torch.manual_seed(0)
    model = pipeline.Voice2PoseModel(config.get_cfg(name, opts), num_train_samples=n_train)
    sd = model.state_dict()
    for k in g.files:
        if k.startswith(tag + "/buf/"):
            sd[k[len(tag) + 5:]] = torch.from_numpy(g[k])
    model.load_state_dict(sd)
    return g, model.cuda().eval(), n_train, int(g[tag + "/batch_size"])


@pytest.mark.parametrize("tag,name,opts", [("s2g", "voice2pose_s2g", []),
                                           ("gtcode", "voice2pose_sdt_bp", ["VOICE2POSE.GENERATOR.CLIP_CODE.TEST_WITH_GT_CODE", True])])
@pytest.mark.parametrize("mode", [0, 3])
def test_eval_forward_vs_reference_fixture(tag, name, opts, mode):
    """The eval path of the drop-in model (what Voice2PoseEvaluator / the reference's test_step drive) against the reference's
    own recorded validation forward: BatchNorm from running statistics, parted->global FGD input (s2g), ground-truth code + KL
    (TEST_WITH_GT_CODE), LSGAN terms.  fp32 mode 1e-4 / TF32 mode 5e-3 (stated)."""
    from oracle import sdt_oracle as O
    g, model, n_train, bs = _eval_model(tag, name, opts)
    model.set_conv_math(mode)
    tol = 1e-4 if mode == 0 else 5e-3
    hier = bool(model.cfg.DATASET.HIERARCHICAL_POSE)
    b = O.synthetic_batch(bs, n_train, oliver_stat(hier), seed=410, stat_parted=oliver_stat(True), stat_global=oliver_stat(False))
    hb = dict(b)
    hb["speaker_stat"] = {k: torch.from_numpy(np.asarray(v)) for k, v in b["speaker_stat"].items()}
    with torch.no_grad():
        losses, results = model(hb, None)
    ref_keys = {k.split("/")[-1] for k in g.files if k.startswith(tag + "/loss/")}
    assert set(losses) == ref_keys, (set(losses), ref_keys)
    for k, v in losses.items():
        ref = float(g["%s/loss/%s" % (tag, k)])
        assert abs(float(v) - ref) <= tol * max(1.0, abs(ref)), (k, float(v), ref)
    rel = lambda a, r: float(np.abs(a - r).max() / (np.abs(r).max() + 1e-30))
    assert rel(results["poses_pred_batch"].cpu().numpy(), g[tag + "/pred"]) < tol
    for k in ("mu_pred", "mu_gt", "logvar_pred", "logvar_gt"):
        assert rel(results[k].cpu().numpy(), g["%s/%s" % (tag, k)]) < 10 * tol, k
    if tag == "gtcode":
        assert rel(results["condition_code"].cpu().numpy(), g[tag + "/condition_code"]) < 10 * tol


def test_evaluator_epoch_on_s2g_matches_oracle_eval_forward():
    """One validation epoch of voice2pose_s2g through Voice2PoseEvaluator against the CPU oracle's eval forward + the host-side
    recipe (losses * batch size summed / num samples; np.cov + sqrtm on the concatenated FGD features)."""
    from oracle import sdt_oracle as O
    from speechdrivestemplates_b200 import evaluation as E
    g, model, n_train, bs = _eval_model("s2g", "voice2pose_s2g", [])
    model.set_conv_math(0)
    cfgo = O.make_cfg("voice2pose_s2g")
    orc = O.Voice2PoseOracle(cfgo, n_train, seed=0)
    for k in g.files:
        if k.startswith("s2g/buf/"):
            orc.sd[k[len("s2g/buf/"):]] = torch.from_numpy(g[k])
    ev = E.Voice2PoseEvaluator(model, test_batch_size=bs, multiple=1)
    steps, feats, sums = 30, {k: [] for k in ("mu_pred", "logvar_pred", "mu_gt", "logvar_gt")}, {}
    for i in range(steps):
        b = O.synthetic_batch(bs, n_train, oliver_stat(False), seed=420 + i, stat_parted=oliver_stat(True), stat_global=oliver_stat(False))
        hb = dict(b)
        hb["speaker_stat"] = {k: torch.from_numpy(np.asarray(v)) for k, v in b["speaker_stat"].items()}
        ev.step(hb)
        with torch.no_grad():
            losses, results = orc.forward(b, training=False)
        st = b["speaker_stat"]
        fp = O.get_final_results(results["poses_pred_batch"].float().numpy(), st["mean"], st["std"], st["scale_factor"], False)
        fg = O.get_final_results(results["poses_gt_batch"].float().numpy(), st["mean"], st["std"], st["scale_factor"], False)
        losses = dict(losses, **O.evaluate_step(fp, fg))
        for k in E.Voice2PoseEvaluator.LOSS_KEYS:
            sums[k] = sums.get(k, 0.0) + float(losses[k]) * bs
        for k in feats:
            feats[k].append(results[k].numpy())
    out = ev.finish(bs * steps)
    for k in E.Voice2PoseEvaluator.LOSS_KEYS:
        ref = sums[k] / (bs * steps)
        assert abs(out[k] - ref) <= 1e-4 * max(1.0, abs(ref)), (k, out[k], ref)
    cat = {k: np.concatenate(v, 0).astype(np.float64) for k, v in feats.items()}
    ref_mu = _fgd_host(cat["mu_pred"], cat["mu_gt"])
    # 90 samples of a 32-d feature: sqrtm of a poorly conditioned covariance product amplifies fp32-level feature differences
    assert abs(out["FGD_mu"] - ref_mu) <= 2e-2 * max(1.0, abs(ref_mu)), (out["FGD_mu"], ref_mu)
