"""BaseTask — the task protocol of the reference (tasks/base_task.py:13-83): reset(iter_count), step(), calc_reward(prev_torque,
prev_action, action) -> dict, done() -> bool, substep().

On this path the five hooks run INSIDE the step kernel (csrc/sim_core.h: env_step / env_reset, one compile-time task policy
per Cfg<NJ, TK>).  The Python subclasses in tasks/descriptors.py are the host-side descriptors SURVEY.md §8b asks for: they
carry what the reference's task objects carry (weights, durations, body names, mass, neutral pose) and their hooks report
what the kernel computed for one environment of the batch — they do not recompute anything on the CPU."""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np


class BaseTask(ABC):
    @abstractmethod
    def reset(self, iter_count: int = 0) -> None: ...

    @abstractmethod
    def step(self) -> None: ...

    @abstractmethod
    def calc_reward(self, prev_torque: np.ndarray, prev_action: np.ndarray, action: np.ndarray) -> dict[str, float]: ...

    @abstractmethod
    def done(self) -> bool: ...

    def substep(self) -> None:
        pass
