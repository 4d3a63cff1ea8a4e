"""Minimal stand-ins for ``gym`` (0.26 API) and ``dm_env`` (1.5 API).

The reference pins ``gym==0.26.1`` and ``dm-env==1.5`` (python3/requirements.txt);
neither is installable offline.  ``ensure_gym()`` / ``ensure_dm_env()`` import the
real package when present and otherwise put ``compat/shims`` on ``sys.path``.  The
shims provide exactly the symbols the wrappers touch (gym_env.py:18,32,114-123,195;
dm_env.py:19-20,35,117-129,183,197,232-234).
"""

import importlib
import os
import sys

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def _ensure(name):
    try:
        return importlib.import_module(name)
    except ImportError:
        if _SHIMS not in sys.path:
            sys.path.append(_SHIMS)
        return importlib.import_module(name)


def ensure_gym():
    return _ensure("gym")


def ensure_dm_env():
    return _ensure("dm_env")
