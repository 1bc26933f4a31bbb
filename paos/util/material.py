"""``paos.util.material`` (reference ``paos/util/material.py``)."""
from paos_b200.material import Material  # noqa: F401
