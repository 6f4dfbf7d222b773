import sys, torch
sys.path.insert(0,'.')
from stroke_prediction_b200 import ops
from stroke_prediction_b200.common import data, metrics
b = data.synthetic_cae_batch(8, seed=4)
lab = b[data.KEY_LABELS].cuda()
core, penu, les = lab[:,0:1].contiguous(), lab[:,1:2].contiguous(), lab[:,2:3].contiguous()
rec = torch.rand_like(core)*0.3 + 0.5*penu
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
print('surface_distances core-vs-penu   %.3f ms' % t(lambda: ops.surface_distances(core, penu)))
print('surface_distances rec-vs-lesion  %.3f ms' % t(lambda: ops.surface_distances(rec, les)))
print('binary_counts                    %.3f ms' % t(lambda: ops.binary_counts(rec, les, 0.5)))
print('binary_measures_many (3 pairs, incl. D2H)  %.3f ms' % t(lambda: metrics.binary_measures_many([(rec,les),(core,core),(penu,penu)])))
print('  without surface distances                %.3f ms' % t(lambda: metrics.binary_measures_many([(rec,les),(core,core),(penu,penu)], surface_distances=False)))
ops.start_profile(); ops.surface_distances(rec, les); print(ops.stop_profile())
