"""``paos.classes.wfo`` (reference ``paos/classes/wfo.py``): the device-resident ``WFO``."""
from paos_b200.wfo import WFO  # noqa: F401
