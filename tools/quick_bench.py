import time, numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_qem_b200 import backends, engine, families as F, noise
eng=engine.Engine(0)
# cfg2-like
be=backends.synthetic_chain(16, seed=2); nm=noise.from_backend(be); eng.set_noise(nm)
tw,base,obs=F.config_brick10_twirl(n_base=4,n_twirls=50)
b=engine.encode_batch(tw,[obs]*len(tw))
PF=lambda d: d<<8
for kq,low,chunk,fl in ((6,2,0,0),(6,2,0,PF(0xffff)),(6,2,0,PF(148)),(6,2,0,PF(296)),(6,2,0,PF(888)),(6,2,0,3|PF(0xffff)),(6,1,0,0),(7,2,0,0)):
    eng.set_options(tile_qubits=kq,low_qubits=low,chunk_circuits=chunk,flags=fl)
    for _ in range(4):
        t=time.time(); v,s=eng.run_dm(b); dt=time.time()-t
    st=eng.stats()
    print('brick10',kq,low,chunk,'flags',fl,'circ/s',len(tw)/dt,'kernel_ms',st['kernel_ms'],'lower_ms',st['lower_ms'],'h2d',st['h2d_ms'],'GB/s',st['state_bytes_swept']/st['kernel_ms']/1e6,'sweeps/circ',st['n_state_sweeps']/len(tw),'passes/circ',st['n_passes']/len(tw))
# cfg3-like n=12,13
for n in (12,13):
    be=backends.synthetic_chain(n, seed=n); eng.set_noise(noise.from_backend(be))
    circs,obs=F.config_tfim_dm(n=n,n_circuits=4,max_steps=4)
    b=engine.encode_batch(circs,[obs]*len(circs))
    for kq,low,fl in ((6,2,0),(6,2,PF(0xffff)),(6,2,PF(148)),(6,2,PF(296)),(6,2,PF(888)),(6,2,3|PF(0xffff)),(6,1,0),(7,2,0)):
        eng.set_options(tile_qubits=kq,low_qubits=low,flags=fl)
        for _ in range(2):
            t=time.time(); v,s=eng.run_dm(b); dt=time.time()-t
        st=eng.stats()
        print('tfim',n,kq,low,'flags',fl,'circ/s',len(circs)/dt,'kernel_ms',st['kernel_ms'],'GB/s',st['state_bytes_swept']/st['kernel_ms']/1e6,'sweeps',st['n_state_sweeps'],'passes',st['n_passes'])
