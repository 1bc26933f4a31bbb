"""``paos.core.coordinateBreak`` (reference ``paos/core/coordinateBreak.py``)."""
from paos_b200.coordinate_break import coordinate_break  # noqa: F401
