"""Registration into the reference's own registries (its only plugin API):

    import speechdrivestemplates_b200.plugin as sdt; sdt.register()

overwrites the entries of ``core.networks.module_dict`` (core/networks/__init__.py:6-11) whose B200 implementation
exists, so ``get_model(cfg.VOICE2POSE.GENERATOR.NAME)(cfg)`` (voice2pose.py:33,76) builds the CUDA-backed modules
with the reference's main.py, YAML configs and checkpoints unchanged.  Optionally swaps the step model class used by
the Voice2Pose pipeline so the mel front end and the loss graph run on the fused kernels too.
"""


def register(swap_step_model=True):
    """Returns the list of registry names that now resolve to B200 implementations."""
    import core.networks as ref_networks            # the reference package must be importable (its repo root on sys.path)
    from . import networks
    done = []
    for name, cls in networks.module_dict().items():
        ref_networks.module_dict[name] = cls
        done.append(name)
    if swap_step_model:
        import core.pipelines.voice2pose as ref_v2p
        from . import pipeline
        ref_v2p.Voice2PoseModel = pipeline.Voice2PoseModel      # constructed at voice2pose.py:221
        done.append("Voice2PoseModel")
        import core.pipelines.pose2pose as ref_p2p
        ref_p2p.Pose2PoseModel = pipeline.Pose2PoseModel        # constructed at pose2pose.py:100
        done.append("Pose2PoseModel")
    return done
