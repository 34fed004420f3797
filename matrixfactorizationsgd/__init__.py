"""Import shim: makes ``import matrixfactorizationsgd.java_b200`` resolve to the sibling directory
``matrixfactorizationsgd.java_b200/`` (a directory name with a dot cannot be imported directly)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "matrixfactorizationsgd.java_b200")
_spec = importlib.util.spec_from_file_location(
    "matrixfactorizationsgd.java_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
java_b200 = importlib.util.module_from_spec(_spec)
sys.modules["matrixfactorizationsgd.java_b200"] = java_b200
_spec.loader.exec_module(java_b200)
