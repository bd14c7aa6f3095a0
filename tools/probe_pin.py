import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np
from opengate_b200 import core
core.load_library()
import ctypes as C
# a CUDA context first
eng = core.Engine(0, 1)
def pin_once(n=78_643_200):
    t0=time.perf_counter(); a=np.empty(n, dtype=np.float64); t1=time.perf_counter()
    ok=core.host_register(a); t2=time.perf_counter()
    core.host_unregister(a); t3=time.perf_counter()
    return a, (t1-t0, t2-t1, t3-t2, ok)
keep=[]
for it in range(6):
    a,t=pin_once(); print(it, 'alloc %.3f pin %.3f unpin %.3f'%t[:3], t[3], flush=True)
    if it%2==0: keep.append(a)   # keep some alive like the actors do
import threading
def worker(i, out): out[i]=pin_once()[1]
for rep in range(2):
    out=[None]*3; th=[threading.Thread(target=worker,args=(i,out)) for i in range(3)]
    t0=time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]; print('3 parallel', time.perf_counter()-t0, [('%.3f'%o[1]) for o in out], flush=True)
