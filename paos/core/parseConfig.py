"""``paos.core.parseConfig`` (reference ``paos/core/parseConfig.py``)."""
from paos_b200.parse_config import parse_config  # noqa: F401
