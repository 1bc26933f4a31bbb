"""``paos.core``: module paths of the reference (``paos/core/``) mapped onto ``paos_b200``."""
from paos.core import coordinateBreak, parseConfig, pipeline, plot, raytrace, run, saveOutput  # noqa: F401
