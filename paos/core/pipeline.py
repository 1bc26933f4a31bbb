"""``paos.core.pipeline`` (reference ``paos/core/pipeline.py``)."""
import importlib

pipeline = importlib.import_module("paos_b200.pipeline").pipeline
