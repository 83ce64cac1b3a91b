"""Configuration tree with the reference's key names (configs/default.py:4-97 is the contract: the YAML overlays
and ``KEY VALUE`` command-line overrides of main.py:22-33 address these keys).

The drop-in modules accept EITHER the reference's own frozen yacs ``CfgNode`` (when driven by the reference's
main.py) OR the node built here (stand-alone use, tests, bench); both expose attribute access.
"""
import ast
import copy

_DEFAULTS = {
    "PIPELINE_TYPE": None,
    "VOICE2POSE": {
        "STRICT_LOADING": True,
        "GENERATOR": {
            "NAME": None, "LEAKY_RELU": True, "NORM": "IN", "LAMBDA_REG": 1.0, "LAMBDA_CLIP_KL": 0.1,
            "CLIP_CODE": {
                "DIMENSION": None, "LR_SCALING": 1.0, "TRAIN": True, "FRAME_VARIANT": False, "SAMPLE_FROM_NORMAL": False,
                "TEST_WITH_GT_CODE": False, "EXTERNAL_CODE": False, "EXTERNAL_CODE_PTH": None,
            },
        },
        "POSE_ENCODER": {"NAME": "PoseSeqEncoder", "AE_CHECKPOINT": None},
        "POSE_DISCRIMINATOR": {"NAME": None, "LEAKY_RELU": False, "LAMBDA_GAN": 1.0, "MOTION": True, "WHITE_LIST": None},
    },
    "POSE2POSE": {
        "AUTOENCODER": {"NAME": None, "LEAKY_RELU": True, "NORM": "BN", "CODE_DIM": 32},
        "LAMBDA_REG": 1.0, "LAMBDA_KL": 0.1,
    },
    "DATASET": {
        "NAME": "GestureDataset", "ROOT_DIR": "datasets/speakers", "SUBSET": None, "NUM_LANDMARKS": 121,
        "HIERARCHICAL_POSE": True, "SPEAKER": None, "NUM_FRAMES": 64, "AUDIO_LENGTH": 68267, "MAX_DEMO_LENGTH": 24,
        "AUDIO_SR": 16000, "FPS": 15, "CACHING": False,
    },
    "TRAIN": {
        "NUM_EPOCHS": 100, "BATCH_SIZE": 32, "SAVE_VIDEO": True, "SAVE_NPZ": False, "LR": 1e-4, "WD": 0,
        "LR_SCHEDULER": True, "PRETRAIN_FROM": None, "VALIDATE": True, "NUM_RESULT_SAMPLE": 2, "CHECKPOINT_INTERVAL": 1,
    },
    "TEST": {"BATCH_SIZE": 32, "NUM_RESULT_SAMPLE": 8, "SAVE_VIDEO": True, "SAVE_NPZ": True, "MULTIPLE": 1},
    "DEMO": {"MULTIPLE": 1, "NUM_SAMPLES": 1, "CODE_INDEX": None, "CODE_INDEX_B": None, "CODE_PATH": None},
    "SYS": {
        "OUTPUT_DIR": "output/", "CANVAS_SIZE": (720, 1280), "VISUALIZATION_SCALING": 0.85, "VIDEO_FORMAT": ["mp4", "img"],
        "ASYNC_VIDEO_SAVING": False, "LOG_INTERVAL": 100, "NUM_WORKERS": 8, "DISTRIBUTED": False, "WORLD_SIZE": 1,
        "MASTER_ADDR": "localhost", "MASTER_PORT": 21379,
    },
}

# the reference's four YAML overlays (configs/*.yaml), as override lists
OVERLAYS = {
    "voice2pose_sdt_bp": ["PIPELINE_TYPE", "Voice2Pose", "VOICE2POSE.GENERATOR.NAME", "SequenceGeneratorCNN",
                          "VOICE2POSE.GENERATOR.CLIP_CODE.DIMENSION", 32, "VOICE2POSE.GENERATOR.CLIP_CODE.EXTERNAL_CODE", False],
    "voice2pose_sdt_vae": ["PIPELINE_TYPE", "Voice2Pose", "VOICE2POSE.GENERATOR.NAME", "SequenceGeneratorCNN",
                           "VOICE2POSE.GENERATOR.CLIP_CODE.DIMENSION", 32, "VOICE2POSE.GENERATOR.CLIP_CODE.EXTERNAL_CODE", True],
    "voice2pose_s2g": ["PIPELINE_TYPE", "Voice2Pose", "VOICE2POSE.GENERATOR.NAME", "SequenceGeneratorCNN",
                       "VOICE2POSE.GENERATOR.NORM", "BN", "VOICE2POSE.POSE_DISCRIMINATOR.NAME", "PoseSequenceDiscriminator",
                       "VOICE2POSE.POSE_DISCRIMINATOR.LAMBDA_GAN", 0.1, "VOICE2POSE.POSE_DISCRIMINATOR.LEAKY_RELU", True,
                       "DATASET.HIERARCHICAL_POSE", False, "SYS.NUM_WORKERS", 16],
    "pose2pose": ["PIPELINE_TYPE", "Pose2Pose", "POSE2POSE.AUTOENCODER.NAME", "Autoencoder", "TRAIN.NUM_EPOCHS", 100,
                  "TRAIN.LR", 1e-4],
}


class Node(dict):
    """Attribute-access config node (the subset of yacs.CfgNode behaviour the reference relies on)."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = Node(v) if isinstance(v, dict) else copy.deepcopy(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return Node(self)

    def merge_from_list(self, opts):
        opts = list(opts)
        if len(opts) % 2:
            raise ValueError("override list must hold KEY VALUE pairs")
        for key, val in zip(opts[0::2], opts[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            if parts[-1] not in node:
                raise KeyError("Non-existent config key: %s" % key)
            if isinstance(val, str):
                try:
                    val = ast.literal_eval(val)
                except (ValueError, SyntaxError):
                    pass
            node[parts[-1]] = val
        return self

    def freeze(self):
        return self


def get_cfg_defaults():
    return Node(_DEFAULTS)


def get_cfg(name, opts=()):
    """Defaults + one of the reference's overlays (by YAML stem) + KEY VALUE overrides."""
    if name not in OVERLAYS:
        raise KeyError("Unknown config: %s" % name)
    return get_cfg_defaults().merge_from_list(OVERLAYS[name]).merge_from_list(opts)
