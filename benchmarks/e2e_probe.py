import sys, time
sys.path.insert(0,'pynbody-extras_b200'); sys.path.insert(0,'.')
import numpy as np
from benchmarks.synthetic import nfw_disc
from pynbodyext.gravity import Gravity, KernelKind
import pynbodyext._rust as r
n=10_000_000
pos,mass,h=nfw_disc(n,seed=3)
def step(tag):
    t0=time.perf_counter(); g=Gravity(pos,mass,softening=h,kernel=KernelKind.Spline); t1=time.perf_counter()
    tree=g.tree; t2=time.perf_counter()
    out=g.tree_potentials(theta=0.7); t3=time.perf_counter()
    del g, tree; t4=time.perf_counter()
    print(tag, 'init %.1f build %.1f eval %.1f del %.1f ms'%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3), flush=True)
for i in range(4): step(i)
import os
os.environ['X']='1'
