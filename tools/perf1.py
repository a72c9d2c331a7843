import sys, time, json
sys.path.insert(0,'.')
import numpy as np
from vapor_b200 import synth
from vapor_b200.engine import Engine
n_sv = int(sys.argv[1]) if len(sys.argv)>1 else 400
t=time.time(); w = synth.make_workload(n_sv, seed=1); print('gen', time.time()-t)
eng = Engine(0)
for which,name in ((0,'isetp'),(1,'lop3'),(2,'iadd3')):
    print(name, '%.3e lane-ops/s'%eng.int_peak(which))
for it in range(3):
    t=time.time(); res = eng.score(w.batch); dt=time.time()-t
    tm = eng.timings()
    print('e2e %.1f ms'%(dt*1e3), {k:(round(v,2) if isinstance(v,float) else v) for k,v in tm.items()})
    print('cells/s (tile) %.3e  reads/s e2e %.1f'%(tm['cells']/(tm['tile_ms']*1e-3), w.batch.n_task/dt))
