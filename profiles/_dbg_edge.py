import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from test_gpu_domain import _setup, _slabs
from sde_burst_coast_b200.domain import step_loopback, gather_global, attach_loopback_peers
N, L = 6000, 400.0
sp, st = _setup(N, L, 21)
st['vproj'] = st['vproj'] * 4; st['vx'], st['vy'] = st['vx'] * 4, st['vy'] * 4
def run(peer, steps=2):
    A = _slabs(sp, st, N, 2, seed=5)
    if peer: attach_loopback_peers(A)
    step_loopback(A, 0, steps, peer=peer)
    g,_ = gather_global(A, N)
    for s in A: s.close()
    return g
a1, a2, b1, b2 = run(False), run(False), run(True), run(True)
print('nonpeer vs nonpeer', {k: int((a1[k] != a2[k]).sum()) for k in a1 if (a1[k] != a2[k]).any()})
print('peer vs peer', {k: int((b1[k] != b2[k]).sum()) for k in a1 if (b1[k] != b2[k]).any()})
print('nonpeer vs peer', {k: int((a1[k] != b1[k]).sum()) for k in a1 if (a1[k] != b1[k]).any()})
i = np.where(a1['fx'] != b1['fx'])[0]
print(i[:10], a1['fx'][i[:5]], b1['fx'][i[:5]], a1['x'][i[:10]], st['bin_step'][i[:10]])
