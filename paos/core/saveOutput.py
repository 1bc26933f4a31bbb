"""``paos.core.saveOutput`` (reference ``paos/core/saveOutput.py``)."""
from paos_b200.save_output import load_output, remove_keys, save_datacube, save_output  # noqa: F401
