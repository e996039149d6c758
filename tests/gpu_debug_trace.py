import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from common import *
from hybridmc_b200.engine import *
name=sys.argv[1]; nsteps=int(sys.argv[2])
p = params_from_json(load_config(name)) if name=='8bead_1bond' else transition_params(name,0)
o=OracleRun(p); pos=o.init_pos(); o.sim.init_s(); o.sim.reset_clocks(); o.sim.enable_trace(200000)
normals=o.normals(nsteps)
for s in range(nsteps): assert o.sim.md_step(0,s,normals[s])==0
otr,on=o.sim.get_trace()
e=Engine(p,1); e.set_positions(pos[None]); e.init_s(); e.reset_clocks(); e.enable_trace(0,200000)
t0=time.time(); e.md_steps_injected(0,0,normals[None]); t1=time.time()
gtr,gn=e.get_trace()
print(name,'oracle events',on,'gpu events',gn,'gpu time',t1-t0, 'err', e.get_counts()[3], 'stats', e.launch_stats())
k=first_trace_mismatch(otr,gtr)
print('first mismatch', k)
if k is not None:
    for q in range(max(0,k-3), min(k+4, len(otr), len(gtr))):
        print(q, 'O', otr[q], 'G', gtr[q], otr[q]['t'].hex() if hasattr(otr[q]['t'],'hex') else '', float(gtr[q]['t']).hex())
