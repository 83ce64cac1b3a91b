"""Pipeline-registry seam: ``Voice2Pose`` / ``Pose2Pose`` subclasses of the reference's own pipeline classes whose training
step is the fused device program (``pipeline.Voice2PoseTrainer`` / ``pipeline.Pose2PoseTrainer``).

The reference resolves ``cfg.PIPELINE_TYPE`` through ``core.pipelines.module_dict`` (core/pipelines/__init__.py:5-16, main.py:
``get_pipeline(cfg.PIPELINE_TYPE)(cfg)``).  ``plugin.register()`` installs the classes built here under the same two names, so the
reference's ``main.py``, YAML configs, data loaders, logging, checkpoint files and validation / test / demo loops
(core/pipelines/trainer.py) run unchanged while

* ``setup_model`` (voice2pose.py:216-242, pose2pose.py:94-107) builds the fused trainer and exposes its drop-in step model as
  ``self.model`` behind a handle with the surface the reference uses on its DDP / DataParallel wrapper: ``.module``, ``__call__``,
  ``train()`` / ``eval()``, ``state_dict()`` / ``load_state_dict()`` with the ``module.`` key prefix;
* ``setup_optimizer`` (voice2pose.py:244-279, pose2pose.py:109-122) registers handles over the fused flat Adam under the
  reference's optimizer names (so ``save_checkpoint`` trainer.py:305-321 and ``logger_writer_step`` :246-262 work as they are) and
  MultiStepLR-equivalent schedule handles that drive ``trainer.set_lr``;
* ``train_step`` (voice2pose.py:281-331, pose2pose.py:124-166) runs ONE fused step and then the reference's own logging /
  result-saving code on the step's outputs.

``test_step`` / ``demo_step`` / ``evaluate_step`` / ``evaluate_epoch`` are inherited: they call ``self.model(batch, dataset)``, which
the handle routes to the drop-in ``Voice2PoseModel`` / ``Pose2PoseModel`` over the same parameters.

The classes are produced by a factory taking the parent classes, so the module itself imports nothing from the reference (the
GPU box has no /root/reference; tests/test_gpu_pipelines.py drives the same code with a stand-in parent that replays the
reference's training loop).
"""
import bisect
import os
from collections import OrderedDict

import torch

from . import checkpoint as ckpt_io
from . import pipeline


def default_conv_math(cfg=None):
    """Math mode of the fused trainers behind the registry: SDT_CONV_MATH in the environment, else cfg.SYS.SDT_CONV_MATH when the
    config carries it, else 3 (tcgen05 TF32 with operand reuse: TF32 is the reference's own GPU default through cuDNN)."""
    env = os.environ.get("SDT_CONV_MATH")
    if env is not None and env != "":
        return int(env)
    sys_node = getattr(cfg, "SYS", None) if cfg is not None else None
    if sys_node is not None and getattr(sys_node, "get", None) is not None and sys_node.get("SDT_CONV_MATH") is not None:
        return int(sys_node.get("SDT_CONV_MATH"))
    return 3


class ModuleHandle:
    """What the reference touches on ``self.model`` (a DDP / DataParallel wrapper there): ``.module``, call, train / eval,
    ``state_dict`` with ``module.``-prefixed keys (trainer.py:316, voice2pose.py:226-229), ``parameters``."""

    def __init__(self, module, trainer=None):
        self.module = module
        self.trainer = trainer

    def __call__(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def train(self, mode=True):
        self.module.train(mode)
        return self

    def eval(self):
        self.module.eval()
        return self

    def parameters(self):
        return self.module.parameters()

    def state_dict(self):
        return OrderedDict(("module." + k, v) for k, v in self.module.state_dict().items())

    def load_state_dict(self, state_dict, strict=True):
        sd = OrderedDict((k[len("module."):] if k.startswith("module.") else k, v) for k, v in state_dict.items())
        own = self.module.state_dict()
        missing = [k for k in own if k not in sd]
        unexpected = [k for k in sd if k not in own]
        if strict and (missing or unexpected):
            raise RuntimeError("Error(s) in loading state_dict: missing %s, unexpected %s" % (missing, unexpected))
        with torch.no_grad():
            for k, v in sd.items():
                if k in own:
                    own[k].copy_(v)           # in place: the parameters stay views of the trainer's flat buffer
        if self.trainer is not None:
            self.trainer._graphs = None       # captured graphs hold derived tables (mel bands): re-capture
            mel = getattr(self.module, "mel_transfm", None)
            if mel is not None and hasattr(mel, "_tables_key"):
                mel._tables_key = None
        return missing, unexpected


class FusedAdamHandle:
    """One of the reference's ``torch.optim.Adam`` objects, backed by a slice of the fused trainer's flat Adam state.
    ``state_dict()`` / ``load_state_dict()`` speak torch.optim.Adam's format (checkpoint.py); ``param_groups[0]['lr']`` is what
    ``logger_writer_step`` prints; ``zero_grad`` / ``step`` are no-ops (the fused step owns them)."""

    def __init__(self, trainer, names, params, offset, scalars, lr_of, weight_decay=0.0):
        self.trainer, self.names, self.params, self.offset, self.scalars = trainer, names, params, offset, scalars
        self._lr_of, self.weight_decay = lr_of, weight_decay

    @property
    def param_groups(self):
        return [{"lr": self._lr_of(), "weight_decay": self.weight_decay, "params": self.params}]

    def state_dict(self):
        return ckpt_io._optimizer_state(self.trainer, self.names, self.params, self.offset, self.scalars, self._lr_of(), self.weight_decay)

    def load_state_dict(self, sd):
        return ckpt_io._load_optimizer_state(self.trainer, sd, self.names, self.params, self.offset, self.scalars)

    def zero_grad(self, set_to_none=True):
        pass

    def step(self):
        pass


class MultiStepHandle:
    """torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones, gamma, last_epoch) over the fused trainer: the reference builds
    one scheduler per optimizer with identical milestones (voice2pose.py:251-279); ``step()`` of the FIRST handle moves the base
    rate of the trainer (``set_lr`` derives the clip-code and discriminator rates), the others only keep count."""

    def __init__(self, trainer, base_lr, milestones, gamma=0.1, last_epoch=-1, drives=True):
        self.trainer, self.base_lr, self.milestones, self.gamma = trainer, float(base_lr), sorted(milestones), gamma
        self.last_epoch = last_epoch + 1          # MultiStepLR's constructor performs the initial step
        self.drives = drives
        self._apply()

    def get_last_lr(self):
        return [self.base_lr * self.gamma ** bisect.bisect_right(self.milestones, self.last_epoch)]

    def _apply(self):
        if self.drives:
            self.trainer.set_lr(self.get_last_lr()[0])

    def step(self):
        self.last_epoch += 1
        self._apply()


def _dist_group(cfg):
    if getattr(cfg.SYS, "DISTRIBUTED", False) and torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.group.WORLD
    return None


def _as_tensor_dict(host_losses, device):
    return OrderedDict((k, torch.tensor(v, device=device)) for k, v in host_losses.items())


def make_pipelines(RefVoice2Pose, RefPose2Pose):
    """-> (Voice2Pose, Pose2Pose): subclasses of the given reference pipeline classes running the fused train step."""

    class Voice2Pose(RefVoice2Pose):
        """core/pipelines/voice2pose.py:211-331 on the fused trainer (SDT configs and voice2pose_s2g alike)."""
        fused = None

        def setup_model(self, cfg, state_dict=None):
            rank = self.get_rank()
            if self.is_master_process():
                print(torch.cuda.device_count(), "GPUs are available.")
            print("Setting up models on rank", rank, "(speechdrivestemplates_b200 fused pipeline)")
            if getattr(self, "num_train_samples", None) is None:
                # test / demo: no optimizer, no fused step -- the drop-in step model alone (voice2pose.py:221)
                model = pipeline.Voice2PoseModel(cfg, state_dict, None, rank).cuda()
                model.set_conv_math(default_conv_math(cfg))
                self.model = ModuleHandle(model)
            else:
                try:
                    self.fused = pipeline.Voice2PoseTrainer(cfg, self.num_train_samples, torch.device("cuda", rank), process_group=_dist_group(cfg),
                                                            seed=int(getattr(cfg.SYS, "SEED", 0) or 0), conv_math=default_conv_math(cfg))
                except NotImplementedError as e:
                    # a configuration the fused step does not cover (e.g. POSE_DISCRIMINATOR.WHITE_LIST): the reference's own
                    # setup_model / setup_optimizer / train_step over the drop-in step model and networks (autograd path)
                    print("speechdrivestemplates_b200: fused trainer unavailable for this config (%s); using the drop-in modules" % e)
                    self.fused = None
                    return super().setup_model(cfg, state_dict)
                self.model = ModuleHandle(self.fused.model, self.fused)
            if state_dict is not None:
                self.model.load_state_dict(state_dict, strict=bool(cfg.VOICE2POSE.STRICT_LOADING))
            # AE_CHECKPOINT -> pose encoder (voice2pose.py:231-242): done by Voice2PoseTrainer / repeated here for the test path
            if self.fused is None and cfg.VOICE2POSE.POSE_ENCODER.NAME is not None and cfg.VOICE2POSE.POSE_ENCODER.AE_CHECKPOINT is not None:
                ck = torch.load(cfg.VOICE2POSE.POSE_ENCODER.AE_CHECKPOINT, map_location="cpu")
                enc = OrderedDict((k.replace("module.ae.encoder.", ""), v) for k, v in ck["model_state_dict"].items() if "encoder" in k)
                self.model.module.pose_encoder.load_state_dict(enc)

        def setup_optimizer(self, checkpoint=None, last_epoch=-1):
            if self.fused is None:
                return super().setup_optimizer(checkpoint, last_epoch)
            tr, cfg, m = self.fused, self.cfg, self.fused.model
            g_params = [p for _, p in m.netG.named_parameters()]
            self.optimizers["optimizerG"] = FusedAdamHandle(tr, tr.g_names, g_params, 0, tr.adam_g, lambda: tr.lr, float(cfg.TRAIN.WD))
            if tr.has_d:
                d_params = [p for _, p in m.netD_pose.named_parameters()]
                self.optimizers["optimizerD_pose"] = FusedAdamHandle(tr, tr.d_names, d_params, tr.off_d, tr.adam_d, lambda: tr.lr)
            if tr.train_code:
                self.optimizers["optimizerClipCode"] = FusedAdamHandle(tr, ["clips_code"], [m.clips_code], tr.n_g_pad, tr.adam_c, lambda: tr.code_lr)
            if checkpoint is not None:
                lr = None
                for k, h in self.optimizers.items():
                    got = h.load_state_dict(checkpoint["%s_state_dict" % k])
                    lr = got if k == "optimizerG" else lr
                if lr is not None:
                    tr.set_lr(lr)
            if cfg.TRAIN.LR_SCHEDULER:
                ms = [cfg.TRAIN.NUM_EPOCHS - 10, cfg.TRAIN.NUM_EPOCHS - 2]
                for i, k in enumerate(self.optimizers):
                    self.schedulers[k.replace("optimizer", "scheduler")] = MultiStepHandle(tr, float(cfg.TRAIN.LR), ms, 0.1, last_epoch, drives=(i == 0))

        def train_step(self, batch, t_step, global_step, epoch):
            if self.fused is None:
                return super().train_step(batch, t_step, global_step, epoch)
            tag = "TRAIN"
            tr = self.fused
            out = tr.train_step(batch)
            if not self.is_master_process():
                return
            log_now = t_step % self.cfg.SYS.LOG_INTERVAL == 0
            save_now = t_step % self.result_saving_interval_train == 0
            if log_now:
                # one 96-byte D2H read; with several ranks these are already the means over ranks (reduce_tensor_dict)
                self.logger_writer_step(tag, _as_tensor_dict(tr.losses_to_host(out), tr.device), t_step, epoch, global_step)
            if save_now and (self.cfg.TRAIN.SAVE_NPZ or self.cfg.TRAIN.SAVE_VIDEO):
                results = {"poses_pred_batch": out["final_pred"], "poses_gt_batch": out["final_gt"]}
                for k in ("mu_pred", "logvar_pred", "mu_gt", "logvar_gt", "condition_code"):
                    if out.get(k) is not None:
                        results[k] = out[k]
                results = {k: v.detach().cpu().numpy() for k, v in results.items()}
                if self.cfg.TRAIN.SAVE_NPZ:
                    self.save_results(tag, t_step, epoch, self.base_path, results)
                if self.cfg.TRAIN.SAVE_VIDEO:
                    vid = self.generate_video_pair(results["poses_pred_batch"][0], results["poses_gt_batch"][0])
                    self.video_writer.save_video(self.cfg, tag, vid, t_step, epoch, global_step, audio=batch["audio"][0].numpy(),
                                                 writer=self.tb_writer, base_path=self.base_path)

    class Pose2Pose(RefPose2Pose):
        """core/pipelines/pose2pose.py:90-166 on the fused trainer."""
        fused = None

        def setup_model(self, cfg, state_dict=None):
            rank = self.get_rank()
            if self.is_master_process():
                print(torch.cuda.device_count(), "GPUs are available.")
            print("Setting up models on rank", rank, "(speechdrivestemplates_b200 fused pipeline)")
            if getattr(self, "num_train_samples", None) is None:
                model = pipeline.Pose2PoseModel(cfg, state_dict, None, rank).cuda()
                model.ae.set_conv_math(default_conv_math(cfg))
                self.model = ModuleHandle(model)
            else:
                self.fused = pipeline.Pose2PoseTrainer(cfg, self.num_train_samples, torch.device("cuda", rank), process_group=_dist_group(cfg),
                                                       seed=int(getattr(cfg.SYS, "SEED", 0) or 0), conv_math=default_conv_math(cfg))
                self.model = ModuleHandle(self.fused.model, self.fused)
            if state_dict is not None:
                self.model.load_state_dict(state_dict)

        def setup_optimizer(self, checkpoint=None, last_epoch=-1):
            tr, cfg = self.fused, self.cfg
            params = [p for _, p in tr.model.ae.named_parameters()]
            self.optimizers["optimizer"] = FusedAdamHandle(tr, tr.names, params, 0, tr.adam, lambda: tr.lr, float(cfg.TRAIN.WD))
            if checkpoint is not None:
                tr.set_lr(self.optimizers["optimizer"].load_state_dict(checkpoint["optimizer_state_dict"]))
            if cfg.TRAIN.LR_SCHEDULER:
                self.schedulers["scheduler"] = MultiStepHandle(tr, float(cfg.TRAIN.LR), [cfg.TRAIN.NUM_EPOCHS - 10, cfg.TRAIN.NUM_EPOCHS - 2],
                                                               0.1, last_epoch)

        def train_step(self, batch, t_step, global_step, epoch):
            tag = "TRAIN"
            tr = self.fused
            out = tr.train_step(batch)
            if not self.is_master_process():
                return
            if t_step % self.cfg.SYS.LOG_INTERVAL == 0:
                self.logger_writer_step(tag, _as_tensor_dict(tr.losses_to_host(out), tr.device), t_step, epoch, global_step)
            if t_step % self.result_saving_interval_train == 0 and (self.cfg.TRAIN.SAVE_NPZ or self.cfg.TRAIN.SAVE_VIDEO):
                results = {"poses_pred_batch": out["final_pred"], "poses_gt_batch": out["final_gt"],
                           "clip_code_mu": out["clip_code_mu"], "clip_code_logvar": out["clip_code_logvar"]}
                results = {k: v.detach().cpu().numpy() for k, v in results.items()}
                if self.cfg.TRAIN.SAVE_NPZ:
                    self.save_results(tag, t_step, epoch, self.base_path, results)
                if self.cfg.TRAIN.SAVE_VIDEO:
                    vid = self.generate_video_pair(results["poses_pred_batch"][0], results["poses_gt_batch"][0])
                    self.video_writer.save_video(self.cfg, tag, vid, t_step, epoch, global_step, audio=batch["audio"][0].numpy(),
                                                 writer=self.tb_writer, base_path=self.base_path)

    Voice2Pose.__qualname__, Pose2Pose.__qualname__ = "Voice2Pose", "Pose2Pose"
    return Voice2Pose, Pose2Pose
