import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import common, parity_cases as pc
from backends import CudaBackend as EmuBackend
name='rodent'
m,cfg,clip,tables = common.setup(name)
b = EmuBackend(tables)
o,_ = common.oracles(name)
st, ctrl = pc.random_states(m, 16)
full, cdist, niter = b.forward_debug(st, ctrl, 0)
at5,_,_ = b.forward_debug(st, ctrl, 5)
for e in range(16):
    o.set_state(st["qpos"][e].astype(np.float64), st["qvel"][e].astype(np.float64), st["act"][e].astype(np.float64), st["qacc_warmstart"][e].astype(np.float64), ctrl[e].astype(np.float64))
    o.forward(); d=o.d
    sa=np.abs(d.qacc_smooth).max()
    qs = common.region(tables, at5[e], "qacc_smooth", m.nv); qa = common.region(tables, full[e], "qacc", m.nv)
    print(e, 'qacc_smooth rel', np.abs(qs-d.qacc_smooth).max()/sa, 'qacc rel', np.abs(qa-d.qacc).max()/sa, 'niter', niter[e], d.solver_niter[0])
