"""``paos.classes.abcd`` (reference ``paos/classes/abcd.py``)."""
from paos_b200.abcd import ABCD  # noqa: F401
