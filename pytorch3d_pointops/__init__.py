"""Drop-in alias: `import pytorch3d_pointops` -> pytorch3d_pointops_b200.

Every module path of the reference package (pytorch3d_pointops.functions.knn,
pytorch3d_pointops.structures.point_structure, pytorch3d_pointops._C, ...) resolves to the
Blackwell-native implementation, so user code written against the reference runs unchanged.
"""
import importlib as _importlib
import sys as _sys

import pytorch3d_pointops_b200 as _impl

_SUBMODULES = [
    "_C",
    "functions",
    "functions.knn",
    "functions.ball_query",
    "functions.chamfer",
    "functions.sample_farthest_points",
    "functions.sample_pdf",
    "functions.packed_to_padded",
    "functions.utils",
    "structures",
    "structures.point_structure",
    "structures.utils",
]
for _name in _SUBMODULES:
    _mod = _importlib.import_module("pytorch3d_pointops_b200." + _name)
    _sys.modules[__name__ + "." + _name] = _mod
    if "." not in _name:
        globals()[_name] = _mod

__version__ = _impl.__version__
