"""Import shim: the product package lives in the directory ``imagecaptioning.pytorch_b200/`` (a literal dot in the
name), which Python's default finder cannot resolve as ``imagecaptioning.pytorch_b200``.  Importing this package
registers that directory under the dotted module name, so ``import imagecaptioning.pytorch_b200`` works."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(os.path.dirname(_here), 'imagecaptioning.pytorch_b200')
_name = 'imagecaptioning.pytorch_b200'
if _name not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(_pkg_dir, '__init__.py'), submodule_search_locations=[_pkg_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
pytorch_b200 = sys.modules[_name]
