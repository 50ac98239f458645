"""Imports fruid_b200/scene.py WITHOUT importing the fruid_b200 package (whose __init__
loads the CUDA library): the scene driver is pure host logic and the CPU-only tests and
the golden generator drive the oracle twins with it."""
import importlib.util
import os

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fruid_b200", "scene.py")
_spec = importlib.util.spec_from_file_location("_fruid_scene", _p)
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
Scene = _mod.Scene
