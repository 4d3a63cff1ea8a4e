"""Stand-in for the subset of gym 0.26 used by PushWorldEnv."""

from . import spaces  # noqa: F401

__version__ = "0.26.1-shim"


class Env:
    metadata = {}
    render_mode = None

    def reset(self, *, seed=None, options=None):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def render(self):
        raise NotImplementedError

    def close(self):
        pass
