"""snowmocap_b200 -- B200-native multi-camera triangulation behind the SnowMocap API.

Drop-in names (reference ``snowvision.camera`` / ``snowvision.triangulation``):
``Camera``, ``CameraGroup``, ``Skew_Ray_Solver``, ``Human_Triangulation``,
``Human_Triangulation_Condense``, ``Human_Triangulation_Smooth``; from ``snowvision.blender``:
``Human_Triangulation_Blender``, ``Human_Triangulation_Blender_Smooth``, ``Human_Triangulation_To_Blender_Result``,
``save_blender_result``.  Batch API: ``TriangulationEngine``, ``triangulate_batch``, ``SmoothState``,
``BlenderControl``, ``BlenderSmoothState``, ``clip_to_blender_result_list``.
Importing the package does not need a GPU; calling into it does (no CPU fallback).
"""
from .camera import Camera, CameraGroup  # noqa: F401

_LAZY = {"TriangulationEngine": "engine", "triangulate_batch": "engine", "Skew_Ray_Solver": "triangulation",
         "Human_Triangulation": "triangulation", "Human_Triangulation_Condense": "triangulation",
         "Human_Triangulation_Smooth": "triangulation", "SmoothState": "engine",
         "Human_Triangulation_Blender": "blender", "Human_Triangulation_Blender_Smooth": "blender",
         "Human_Triangulation_To_Blender_Result": "blender", "save_blender_result": "blender",
         "BlenderControl": "blender", "BlenderSmoothState": "blender", "clip_to_blender_result_list": "blender"}


__all__ = ["Camera", "CameraGroup"] + sorted(_LAZY)   # `from snowmocap_b200 import *` overrides the reference's names


def __getattr__(name):
    import importlib
    if name in _LAZY:
        return getattr(importlib.import_module("." + _LAZY[name], __name__), name)
    if name in ("triangulation", "engine", "dist", "synth", "blender"):   # submodules load on first use (they import torch)
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
