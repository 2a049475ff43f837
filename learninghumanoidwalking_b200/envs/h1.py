"""H1Env — single-environment view of the Unitree H1 standing environment with the reference protocol
(envs/h1/h1_env.py + envs/h1/h1_base.py + tasks/standing_task.py): numpy in, (obs f64, float, bool, dict) out.
A 1-env BatchedHumanoidEnv(model="h1"): observation noise, per-episode dynamics randomisation (joint damping /
frictionloss, body mass / CoM), random pushes and initial-pose noise all run inside the same CUDA kernel."""
from __future__ import annotations

from .jvrc_walk import JvrcWalkEnv


class H1Env(JvrcWalkEnv):
    MODEL = "h1"
