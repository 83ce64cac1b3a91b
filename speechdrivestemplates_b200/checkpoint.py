"""Checkpoints of the fused trainers in the reference's exact ``.pth`` layout.

``Trainer.save_checkpoint`` (core/pipelines/trainer.py:305-321) writes::

    {'epoch', 'step', 'model_state_dict': DataParallel/DDP-wrapped model => every key prefixed 'module.',
     'optimizerG_state_dict', ['optimizerClipCode_state_dict'], ['optimizerD_pose_state_dict']: torch.optim.Adam.state_dict()}

and ``Voice2Pose.__init__`` / ``setup_optimizer`` (voice2pose.py:214-279) read it back on resume; a pose2pose checkpoint
additionally feeds ``VOICE2POSE.POSE_ENCODER.AE_CHECKPOINT`` and the external clip codes of voice2pose_sdt_vae
(voice2pose.py:40-55,234-242).  The fused trainers keep parameters, ``exp_avg`` and ``exp_avg_sq`` in flat buffers and the step
count in a device scalar block; this module converts both ways, so a run can move between the reference and this
implementation at any epoch boundary and ``torch.optim.Adam.load_state_dict`` accepts what is written here.
"""
from collections import OrderedDict

import torch


def _adam_group(lr, n_params, weight_decay=0.0):
    """param_groups entry of torch.optim.Adam (all defaults of the constructor the reference calls)."""
    return {"lr": lr, "betas": (0.9, 0.999), "eps": 1e-08, "weight_decay": weight_decay, "amsgrad": False, "maximize": False,
            "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
            "params": list(range(n_params))}


def _step_of(scalars):
    return int(scalars[4:6].view(torch.int64).item())


def _set_step(scalars, t):
    scalars[4:6].view(torch.int64).fill_(int(t))
    scalars[2] = float(t)


def _views(flat, names, params, offset):
    out, off = [], offset
    for n, p in zip(names, params):
        out.append(flat[off:off + p.numel()].view(p.shape))
        off += p.numel()
    return out


def _optimizer_state(trainer, names, params, offset, scalars, lr, wd=0.0):
    t = _step_of(scalars)
    state = {}
    if t > 0:                       # torch creates the per-parameter state lazily at the first step
        for i, (m, v) in enumerate(zip(_views(trainer.exp_avg, names, params, offset), _views(trainer.exp_avg_sq, names, params, offset))):
            state[i] = {"step": torch.tensor(float(t)), "exp_avg": m.detach().clone().cpu(), "exp_avg_sq": v.detach().clone().cpu()}
    return {"state": state, "param_groups": [_adam_group(lr, len(params), wd)]}


def _load_optimizer_state(trainer, sd, names, params, offset, scalars):
    steps = set()
    ms, vs = _views(trainer.exp_avg, names, params, offset), _views(trainer.exp_avg_sq, names, params, offset)
    for i in range(len(params)):
        st = sd["state"].get(i)
        if st is None:
            ms[i].zero_()
            vs[i].zero_()
            continue
        ms[i].copy_(st["exp_avg"])
        vs[i].copy_(st["exp_avg_sq"])
        steps.add(int(float(st["step"])))
    assert len(steps) <= 1, "parameters of one optimizer carry different step counts: %s" % sorted(steps)
    _set_step(scalars, steps.pop() if steps else 0)
    return float(sd["param_groups"][0]["lr"])


def voice2pose_checkpoint(trainer, epoch, step):
    """The dict ``Trainer.save_checkpoint`` would ``torch.save`` for a ``pipeline.Voice2PoseTrainer``."""
    m = trainer.model
    ckpt = {"epoch": int(epoch), "step": int(step),
            "model_state_dict": OrderedDict(("module." + k, v.detach().clone().cpu()) for k, v in m.state_dict().items())}
    g_params = [p for _, p in m.netG.named_parameters()]
    ckpt["optimizerG_state_dict"] = _optimizer_state(trainer, trainer.g_names, g_params, 0, trainer.adam_g, trainer.lr,
                                                     float(trainer.cfg.TRAIN.WD))
    if trainer.train_code:
        ckpt["optimizerClipCode_state_dict"] = _optimizer_state(trainer, ["clips_code"], [m.clips_code], trainer.n_g_pad,
                                                                trainer.adam_c, trainer.code_lr)
    if trainer.has_d:
        d_params = [p for _, p in m.netD_pose.named_parameters()]
        ckpt["optimizerD_pose_state_dict"] = _optimizer_state(trainer, trainer.d_names, d_params, trainer.off_d, trainer.adam_d,
                                                              trainer.lr)
    return ckpt


def save_voice2pose(trainer, path, epoch, step):
    assert str(path).split(".")[-1] == "pth", "file type not supported: %s" % path        # trainer.py:173
    torch.save(voice2pose_checkpoint(trainer, epoch, step), path)


def load_voice2pose(trainer, ckpt, strict=True):
    """Resume a ``pipeline.Voice2PoseTrainer`` from a checkpoint dict / path written by the reference or by save_voice2pose.
    Returns (epoch, step).  Parameters are copied INTO the flat buffers (the nn.Parameters stay views of them)."""
    if not isinstance(ckpt, dict):
        ckpt = torch.load(ckpt, map_location="cpu")
    m = trainer.model
    sd = OrderedDict((k[len("module."):] if k.startswith("module.") else k, v) for k, v in ckpt["model_state_dict"].items())
    own = m.state_dict()
    missing = [k for k in own if k not in sd]
    unexpected = [k for k in sd if k not in own]
    if strict and (missing or unexpected):                  # VOICE2POSE.STRICT_LOADING (voice2pose.py:226-229)
        raise RuntimeError("Error(s) in loading state_dict: missing %s, unexpected %s" % (missing, unexpected))
    with torch.no_grad():
        for k, v in sd.items():
            if k in own:
                own[k].copy_(v)                             # in place: keeps the views into the flat parameter buffer
    g_params = [p for _, p in m.netG.named_parameters()]
    lr = _load_optimizer_state(trainer, ckpt["optimizerG_state_dict"], trainer.g_names, g_params, 0, trainer.adam_g)
    if trainer.train_code and "optimizerClipCode_state_dict" in ckpt:
        _load_optimizer_state(trainer, ckpt["optimizerClipCode_state_dict"], ["clips_code"], [m.clips_code], trainer.n_g_pad, trainer.adam_c)
    if trainer.has_d and "optimizerD_pose_state_dict" in ckpt:
        d_params = [p for _, p in m.netD_pose.named_parameters()]
        _load_optimizer_state(trainer, ckpt["optimizerD_pose_state_dict"], trainer.d_names, d_params, trainer.off_d, trainer.adam_d)
    trainer.set_lr(lr)
    trainer._graphs = None          # captured graphs hold the old mel band tables / weight-operand tables: re-capture
    if hasattr(m.mel_transfm, "_tables_key"):
        m.mel_transfm._tables_key = None
    return int(ckpt["epoch"]), int(ckpt["step"])


def pose2pose_checkpoint(trainer, epoch, step):
    """The dict ``Trainer.save_checkpoint`` would ``torch.save`` for a ``pipeline.Pose2PoseTrainer`` (trainer.py:305-321 with
    the single optimizer named 'optimizer', pose2pose.py:114): ``module.clip_code_mu`` / ``module.clip_code_logvar`` /
    ``module.mel_transfm.*`` / ``module.ae.*`` + ``optimizer_state_dict``.  This is the file VOICE2POSE.POSE_ENCODER.AE_CHECKPOINT
    and CLIP_CODE.EXTERNAL_CODE of voice2pose_sdt_vae read (voice2pose.py:40-55,234-242)."""
    m = trainer.model
    params = [p for _, p in m.ae.named_parameters()]
    return {"epoch": int(epoch), "step": int(step),
            "model_state_dict": OrderedDict(("module." + k, v.detach().clone().cpu()) for k, v in m.state_dict().items()),
            "optimizer_state_dict": _optimizer_state(trainer, trainer.names, params, 0, trainer.adam, trainer.lr,
                                                     float(trainer.cfg.TRAIN.WD))}


def save_pose2pose(trainer, path, epoch, step):
    assert str(path).split(".")[-1] == "pth", "file type not supported: %s" % path        # trainer.py:173
    torch.save(pose2pose_checkpoint(trainer, epoch, step), path)


def load_pose2pose(trainer, ckpt):
    """Resume a ``pipeline.Pose2PoseTrainer`` (pose2pose.py:104-119: strict state-dict load + optimizer state).  Returns (epoch, step)."""
    if not isinstance(ckpt, dict):
        ckpt = torch.load(ckpt, map_location="cpu")
    m = trainer.model
    sd = OrderedDict((k[len("module."):] if k.startswith("module.") else k, v) for k, v in ckpt["model_state_dict"].items())
    own = m.state_dict()
    missing = [k for k in own if k not in sd]
    unexpected = [k for k in sd if k not in own]
    if missing or unexpected:
        raise RuntimeError("Error(s) in loading state_dict: missing %s, unexpected %s" % (missing, unexpected))
    with torch.no_grad():
        for k, v in sd.items():
            own[k].copy_(v)
    params = [p for _, p in m.ae.named_parameters()]
    trainer.set_lr(_load_optimizer_state(trainer, ckpt["optimizer_state_dict"], trainer.names, params, 0, trainer.adam))
    trainer._graphs = None
    return int(ckpt["epoch"]), int(ckpt["step"])
