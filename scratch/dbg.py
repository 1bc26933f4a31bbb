import sys; sys.path.insert(0,'/root/repo')
import numpy as np
import paos_b200
from oracle import paos_np
def mk():
    d = paos_b200.WFO(1.0, 3e-6, 128, 4); o = paos_np.WFO(1.0, 3e-6, 128, 4)
    for w in (d,o):
        w.aperture(0,0,r=0.5,shape='circular'); w.lens(2.0)
    return d,o
d,o = mk()
f = d.wfo
print('A: wfo first', np.abs(f-o._wfo).max(), f[64,70], o._wfo[64,70])
d,o = mk()
a = d.amplitude; f = d.wfo
print('B: amp first', np.abs(f-o._wfo).max(), f[64,70], o._wfo[64,70])
d,o = mk()
ph = d.phase
print('C: phase first', np.abs(ph-o.phase).max(), ph[64,70], o.phase[64,70])
d,o = mk()
d.flush(); ph = d.phase
print('D: flush, phase', np.abs(ph-o.phase).max(), ph[64,70], o.phase[64,70])
