"""Top-level 'neural_style_transfer' module for drop-in use: put this directory first on PYTHONPATH and the reference's
lab.py / tlbot.py / start_nn.py / task_executor.py import the B200 path unchanged."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from artstyletransfer_b200.neural_style_transfer import *  # noqa: F401,F403,E402
from artstyletransfer_b200 import neural_style_transfer as _impl  # noqa: E402

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
