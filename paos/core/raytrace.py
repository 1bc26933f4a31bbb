"""``paos.core.raytrace`` (reference ``paos/core/raytrace.py``)."""
import importlib

raytrace = importlib.import_module("paos_b200.raytrace").raytrace
