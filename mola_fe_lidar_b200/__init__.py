"""Importable alias of the `mola-fe-lidar_b200/` package directory (a hyphen
is not a valid Python identifier).  All modules live there."""
import os as _os

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
PACKAGE_DIR = _os.path.join(_ROOT, "mola-fe-lidar_b200")
__path__.append(PACKAGE_DIR)  # noqa: F821
