"""Stand-ins for the third-party packages the reference's scripts import but this image does not ship
(``pytorch_lightning``, ``timm``): just the surface ``scripts/training/pretrain_mae.py:6-13,116-128``,
``scripts/training/train_mae.py:6-11,205-217`` and ``scripts/utils.py:3-7`` touch, backed by this package's trainer and
modules.  ``install()`` appends this directory to ``sys.path`` -- at the END, so a real installation always wins -- and is
called by ``ssrl_vit_mae_jepa_b200`` on import; scripts that import ``pytorch_lightning`` before anything of ours are run
with ``PYTHONPATH=<repo>/ssrl-vit-mae-jepa_b200/shims``."""
import os
import sys

SHIM_DIR = os.path.dirname(os.path.abspath(__file__))


def install() -> None:
    if SHIM_DIR not in sys.path:
        sys.path.append(SHIM_DIR)
