"""libbagpu: B200-native sparse bundle adjustment behind ORB-SLAM3's Optimizer BA entry points."""
from .problem import (BAProblem, BAResult, PoseBatch, PoseResult, Round, Schedule, schedule_global_ba,  # noqa: F401
                      schedule_local_ba, schedule_merge_ba)
