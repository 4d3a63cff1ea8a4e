"""Stand-in for the subset of dm_env 1.5 used by PushWorldEnv."""

import enum
from typing import Any, NamedTuple

from . import specs  # noqa: F401


class StepType(enum.IntEnum):
    FIRST = 0
    MID = 1
    LAST = 2

    def first(self):
        return self is StepType.FIRST

    def mid(self):
        return self is StepType.MID

    def last(self):
        return self is StepType.LAST


class TimeStep(NamedTuple):
    step_type: Any
    reward: Any
    discount: Any
    observation: Any

    def first(self):
        return self.step_type == StepType.FIRST

    def mid(self):
        return self.step_type == StepType.MID

    def last(self):
        return self.step_type == StepType.LAST


class Environment:
    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def observation_spec(self):
        raise NotImplementedError

    def action_spec(self):
        raise NotImplementedError

    def close(self):
        pass


def restart(observation):
    return TimeStep(StepType.FIRST, None, None, observation)


def transition(reward, observation, discount=1.0):
    return TimeStep(StepType.MID, reward, discount, observation)


def termination(reward, observation):
    return TimeStep(StepType.LAST, reward, 0.0, observation)


def truncation(reward, observation, discount=1.0):
    return TimeStep(StepType.LAST, reward, discount, observation)
