"""The pipeline-registry seam (speechdrivestemplates_b200/pipelines.py) driven the way the reference's Trainer drives it.

/root/reference does not exist on the GPU box, so the parent classes here are stand-ins that replay the parts of
core/pipelines/trainer.py the subclasses touch: the epoch / step loop (:375-398), ``logger_writer_step`` (:246-262: reads
``optimizer.param_groups[i]['lr']`` and the loss tensors), ``save_checkpoint`` (:305-321) and the resume order of
``setup_experiment`` (:171-186: ``setup_model(cfg, state_dict)`` then ``setup_optimizer(checkpoint, last_epoch)``).  The CPU test
``test_plugin_registers_into_reference_registries`` checks the same factory against the REAL parent classes (registry entry,
method set, signatures)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from util import oliver_stat  # noqa: E402


class _RefTrainerStandIn:
    def __init__(self, cfg):
        self.cfg, self.model, self.optimizers, self.schedulers = cfg, None, {}, {}
        self.logged, self.base_path = [], None
        self.result_saving_interval_train = 1

    def get_rank(self):
        return 0

    def is_master_process(self):
        return True

    def logger_writer_step(self, tag, losses, step, epoch=None, global_step=None):
        lrs = {k: [g["lr"] for g in v.param_groups] for k, v in self.optimizers.items()}
        self.logged.append({"tag": tag, "step": step, "epoch": epoch, "global_step": global_step, "lr": lrs,
                            "losses": {k: float(v.detach().cpu().numpy()) for k, v in losses.items()}})

    def checkpoint_dict(self, epoch, global_step):
        d = {"epoch": epoch, "step": global_step, "model_state_dict": self.model.state_dict()}
        for k, v in self.optimizers.items():
            d["%s_state_dict" % k] = v.state_dict()
        return d

    def run_epochs(self, batches, first_epoch, n_epochs, global_step=0):
        epoch = first_epoch
        for _ in range(n_epochs):
            epoch += 1
            self.model.train()
            for t_step, batch in enumerate(batches):
                global_step += 1
                self.train_step(batch, t_step + 1, global_step, epoch)
            if self.cfg.TRAIN.LR_SCHEDULER:
                for v in self.schedulers.values():
                    v.step()
        return epoch, global_step


class _RefVoice2PoseStandIn(_RefTrainerStandIn):
    pass


class _RefPose2PoseStandIn(_RefTrainerStandIn):
    pass


def _batches(n, bs, n_train, seed0):
    from oracle import sdt_oracle as O
    st = oliver_stat()
    out = []
    for i in range(n):
        b = O.synthetic_batch(bs, n_train, st, seed=seed0 + i)
        b["speaker_stat"] = {k: torch.from_numpy(np.asarray(v)) for k, v in b["speaker_stat"].items()}
        out.append(b)
    return out


def _cfg(name, epochs=12):
    from speechdrivestemplates_b200 import config
    return config.get_cfg(name, ["SYS.LOG_INTERVAL", 1, "TRAIN.NUM_EPOCHS", epochs, "TRAIN.SAVE_VIDEO", False, "TRAIN.SAVE_NPZ", False])


@pytest.mark.parametrize("math", ["0", "3"])
def test_voice2pose_pipeline_runs_the_fused_step_through_the_reference_loop(monkeypatch, math):
    from speechdrivestemplates_b200 import pipeline, pipelines
    monkeypatch.setenv("SDT_CONV_MATH", math)
    V2P, _ = pipelines.make_pipelines(_RefVoice2PoseStandIn, _RefPose2PoseStandIn)
    n_train, bs = 16, 4
    cfg = _cfg("voice2pose_sdt_bp")
    batches = _batches(2, bs, n_train, 300)
    code0 = 0.1 * torch.randn(n_train, 32, generator=torch.Generator().manual_seed(11))

    p = V2P(cfg)
    p.num_train_samples = n_train
    p.setup_model(cfg)
    assert p.fused.conv_math == int(math)
    assert p.model.module is p.fused.model and all(k.startswith("module.") for k in p.model.state_dict())
    p.model.module.clips_code.data.copy_(code0)
    p.setup_optimizer()
    assert set(p.optimizers) == {"optimizerG", "optimizerClipCode"} and set(p.schedulers) == {"schedulerG", "schedulerClipCode"}
    epoch, gstep = p.run_epochs(batches, 0, 2)
    assert [r["global_step"] for r in p.logged] == [1, 2, 3, 4] and p.logged[0]["lr"]["optimizerG"] == [pytest.approx(1e-4)]

    # the same four steps straight on the fused trainer: identical scalars
    tr = pipeline.Voice2PoseTrainer(cfg, n_train, "cuda:0", seed=0, conv_math=int(math))
    tr.model.clips_code.data.copy_(code0)
    for k in range(4):
        host = tr.losses_to_host(tr.train_step(batches[k % 2]))
        assert host == pytest.approx(p.logged[k]["losses"], rel=1e-6, abs=1e-9), k
    assert torch.equal(tr.flat_p, p.fused.flat_p)
    tr.close()

    # milestones [NUM_EPOCHS - 10, NUM_EPOCHS - 2] = [2, 10]: two scheduler steps -> first decay, all three rates follow
    assert p.fused.lr == pytest.approx(1e-5) and p.optimizers["optimizerClipCode"].param_groups[0]["lr"] == pytest.approx(1e-5)

    # checkpoint -> resume in a fresh pipeline (setup_experiment's order) -> continues bit-identically
    ck = p.checkpoint_dict(epoch, gstep)
    ck = {k: ({kk: vv.detach().clone().cpu() for kk, vv in v.items()} if k == "model_state_dict" else v) for k, v in ck.items()}
    adam = torch.optim.Adam(p.model.module.netG.parameters(), lr=1e-4)
    adam.load_state_dict(ck["optimizerG_state_dict"])               # torch's own Adam accepts what the handle writes
    assert len(adam.state_dict()["state"]) == len(list(p.model.module.netG.parameters()))
    q = V2P(cfg)
    q.num_train_samples = n_train
    q.setup_model(cfg, state_dict=ck["model_state_dict"])
    q.setup_optimizer(checkpoint=ck, last_epoch=epoch)
    assert q.fused.lr == pytest.approx(1e-5)
    assert torch.equal(q.fused.flat_p[:q.fused.n_g], p.fused.flat_p[:p.fused.n_g])
    p.run_epochs(batches, epoch, 1, gstep)
    q.run_epochs(batches, epoch, 1, gstep)
    assert q.logged[-1]["losses"] == pytest.approx(p.logged[-1]["losses"], rel=1e-6)
    assert torch.equal(q.fused.flat_p, p.fused.flat_p)

    # test / demo construction (no dataset -> no fused trainer): the drop-in model behind the same handle, loaded from the checkpoint
    t = V2P(cfg)
    t.num_train_samples = None
    t.setup_model(cfg, state_dict=ck["model_state_dict"])
    assert t.fused is None
    t.model.eval()
    with torch.no_grad():
        losses, results = t.model(batches[0], None)
    assert np.isfinite(float(losses["G_loss"])) and tuple(results["poses_pred_batch"].shape) == (bs, 64, 2, 121)
    p.fused.close()
    q.fused.close()


def test_pose2pose_pipeline_runs_the_fused_step_through_the_reference_loop():
    from speechdrivestemplates_b200 import pipeline, pipelines
    _, P2P = pipelines.make_pipelines(_RefVoice2PoseStandIn, _RefPose2PoseStandIn)
    n_train, bs = 16, 4
    cfg = _cfg("pose2pose")
    batches = _batches(2, bs, n_train, 400)
    p = P2P(cfg)
    p.num_train_samples = n_train
    p.setup_model(cfg)
    p.setup_optimizer()
    assert set(p.optimizers) == {"optimizer"} and set(p.schedulers) == {"scheduler"}
    torch.manual_seed(5)
    epoch, gstep = p.run_epochs(batches, 0, 1)
    tr = pipeline.Pose2PoseTrainer(cfg, n_train, "cuda:0", seed=0)
    torch.manual_seed(5)
    for k in range(2):
        host = tr.losses_to_host(tr.train_step(batches[k]))
        assert host == pytest.approx(p.logged[k]["losses"], rel=1e-6, abs=1e-9), k
    assert torch.equal(tr.flat_p, p.fused.flat_p)
    ck = p.checkpoint_dict(epoch, gstep)
    assert "optimizer_state_dict" in ck and "module.clip_code_mu" in ck["model_state_dict"]
    q = P2P(cfg)
    q.num_train_samples = n_train
    q.setup_model(cfg, state_dict={k: v.detach().clone() for k, v in ck["model_state_dict"].items()})
    q.setup_optimizer(checkpoint=ck, last_epoch=epoch)
    assert torch.equal(q.fused.flat_p, p.fused.flat_p) and torch.equal(q.fused.exp_avg, p.fused.exp_avg)
    tr.close()
    p.fused.close()
    q.fused.close()
