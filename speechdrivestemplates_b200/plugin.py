"""Registration into the reference's own registries (its only plugin API):

    import speechdrivestemplates_b200.plugin as sdt; sdt.register()

overwrites the entries of ``core.networks.module_dict`` (core/networks/__init__.py:6-11) whose B200 implementation
exists, so ``get_model(cfg.VOICE2POSE.GENERATOR.NAME)(cfg)`` (voice2pose.py:33,76) builds the CUDA-backed modules
with the reference's main.py, YAML configs and checkpoints unchanged; swaps the step model classes used by the reference
pipelines so the mel front end and the loss graph run on the fused kernels too; and installs ``Voice2Pose`` / ``Pose2Pose``
subclasses into ``core.pipelines.module_dict`` (core/pipelines/__init__.py:5-16) whose ``setup_model`` / ``setup_optimizer`` /
``train_step`` run the fused trainers (pipelines.py) -- so ``get_pipeline(cfg.PIPELINE_TYPE)(cfg).train(...)`` in the reference's
main.py reaches the benchmarked path.  Convolution math of that path: 3 (tcgen05 TF32) unless SDT_CONV_MATH / cfg.SYS.SDT_CONV_MATH
say otherwise (per trainer, not process-global).
"""


def register(swap_step_model=True, swap_pipelines=True):
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
    if swap_pipelines:
        import core.pipelines as ref_pipelines
        from . import pipelines
        base_v2p = getattr(ref_pipelines, "_sdt_ref_Voice2Pose", ref_pipelines.module_dict["Voice2Pose"])
        base_p2p = getattr(ref_pipelines, "_sdt_ref_Pose2Pose", ref_pipelines.module_dict["Pose2Pose"])
        ref_pipelines._sdt_ref_Voice2Pose, ref_pipelines._sdt_ref_Pose2Pose = base_v2p, base_p2p      # register() twice stays one level deep
        v2p, p2p = pipelines.make_pipelines(base_v2p, base_p2p)
        ref_pipelines.module_dict["Voice2Pose"] = v2p
        ref_pipelines.module_dict["Pose2Pose"] = p2p
        done += ["pipeline:Voice2Pose", "pipeline:Pose2Pose"]
    return done
