import sys, os, time, subprocess, threading
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import paos_b200
from paos_b200 import _lib
n=2048; K=int(sys.argv[1]) if len(sys.argv)>1 else 4
zero = len(sys.argv)>2 and sys.argv[2]=='zero'
rng=np.random.default_rng(0); x=(rng.standard_normal((n,n))+1j*rng.standard_normal((n,n)))
if zero: x[:, :] = 0; x[900:1100, 900:1100] = 1.0
ws=[]
for _ in range(4):
    w=paos_b200.WFO(1.0,1e-6,n,4); w.wfo=x; ws.append(w)
p=subprocess.Popen(["nvidia-smi","--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active","--format=csv,noheader","-lms","50"],stdout=subprocess.PIPE,text=True)
lines=[]
threading.Thread(target=lambda:[lines.append(l.strip()) for l in p.stdout],daemon=True).start()
time.sleep(0.3)
for phase in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
    reps=200
    for r in range(reps):
        w=ws[r%4]
        for i in range(K): w._fft2(inverse=bool(i&1))
        w.flush()
        if zero and r%8==7: w.wfo_device  # noop
    e1.record(); torch.cuda.synchronize()
    print(f"phase {phase}: {e0.elapsed_time(e1)*1e3/reps:.1f} us per (row x{K} + col x{K}) pair, host {1e6*(time.perf_counter()-t0)/reps:.1f} us")
p.terminate()
print(lines[::4][:40])
