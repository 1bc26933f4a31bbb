"""Print the fused pass plan (PAOS_DEBUG_PLAN=1) and per-kind pass timings of one AIRS-CH0 2048^2 job."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paos_b200
from paos_b200 import configs
from paos_b200.sweep import Sweep
jobs = configs.airs_ch0(grid=2048, n_wl=256)[100:101]
sw = Sweep(2048, slots=1)
sw.enable_timing(True)
sw.run(jobs)
sw.run(jobs)
print(sw.timing_detail())
