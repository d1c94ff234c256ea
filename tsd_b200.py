"""Import alias: the product package lives in the directory `opencv-traffic-sign-detector_b200/`, whose name is
not a Python identifier.  `import tsd_b200` loads that directory as the package `tsd_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "opencv-traffic-sign-detector_b200")
_spec = importlib.util.spec_from_file_location("tsd_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["tsd_b200"] = _mod
_spec.loader.exec_module(_mod)
