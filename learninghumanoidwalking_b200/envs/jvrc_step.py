"""JvrcStepEnv — single-environment view of the JVRC-1 stepping environment with the reference protocol
(envs/jvrc/jvrc_step.py + tasks/stepping_task.py): numpy in, (obs f64[39], float, bool, dict of 6 reward terms) out.
A 1-env BatchedHumanoidEnv(model="jvrc_step"): the footstep sequence, the 20 stepping-stone slabs it places, the floor that
drops away in FORWARD mode, target tracking and the goal-step observation all live inside the same CUDA kernel.
`env.robot.iteration_count = itr` (rl/workers/rollout_worker.py:95) drives the height curriculum."""
from __future__ import annotations

from .jvrc_walk import JvrcWalkEnv


class JvrcStepEnv(JvrcWalkEnv):
    MODEL = "jvrc_step"
