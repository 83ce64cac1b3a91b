"""speechdrivestemplates_b200 -- B200-native (sm_100a) kernels + host glue for the Voice2Pose / Pose2Pose
training-step hot path of ShenhanQian/SpeechDrivesTemplates.  See DESIGN.md and include/sdt_b200.h."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
