"""Import alias: loads the package in ./amv-codec-tools_b200/ (whose directory name is not a
Python identifier) under the name ``amv_codec_tools_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "amv-codec-tools_b200")
_spec = importlib.util.spec_from_file_location("amv_codec_tools_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["amv_codec_tools_b200"] = _mod
_spec.loader.exec_module(_mod)
