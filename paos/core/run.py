"""``paos.core.run`` (reference ``paos/core/run.py``): ``run`` and ``push_results`` of the device path."""
import importlib

_impl = importlib.import_module("paos_b200.run")  # the module (the package attribute of that name is the function)
run = _impl.run
push_results = _impl.push_results
