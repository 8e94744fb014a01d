"""dbot_ros_b200 -- B200 (sm_100a) depth observation model for the dbot / dbot_ros particle tracker.

Only the hot path of SURVEY.md section 8 lives here: the CUDA library (csrc/ -> libdbot_b200.so,
C-ABI in include/dbot_b200.h) and the host-side mirror of the reference's sensor interface
(`RbSensorBuilder`, `RbSensor.reset/set_observation/loglikes`).  There is no CPU fallback: importing
works anywhere, but building a sensor needs libdbot_b200.so and a CUDA device.
"""
from .abi import LIB_PATH, DbotB200Error, load_library  # noqa: F401
from . import io, replay, tracker  # noqa: F401
from .rb_sensor import (  # noqa: F401
    B200KinectImageModel,
    CameraData,
    ObjectModel,
    Parameters,
    RbSensor,
    RbSensorBuilder,
)
