"""Per-kind timings of the batched pass launches of the headline sweep (one slot, so that launches do not overlap)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paos_b200 import configs
from paos_b200.sweep import Sweep

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
jobs = configs.airs_ch0(grid=2048, n_wl=256)[96:96 + 4 * batch]
sw = Sweep(2048, slots=1, batch=batch)
sw.run(jobs)
sw.enable_timing(True)
sw.run(jobs)
det = sw.timing_detail()
out = {("col" if k[0] else "row") + "x%d" % k[1]: round(1e3 * v[0] / max(v[1], 1), 1) for k, v in sorted(det.items())}
print(json.dumps({"batch": batch, "us_per_launch": out, "totals": sw.timing_totals()}))
