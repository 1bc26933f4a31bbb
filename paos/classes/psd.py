"""``paos.classes.psd`` (reference ``paos/classes/psd.py``)."""
from paos_b200.psd import PSD  # noqa: F401
