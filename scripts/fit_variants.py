import sys, time, json, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from doubletdetection_b200 import _capi
from doubletdetection_b200.classifier import _pca_plan
wl = bench.WORKLOADS['c3']; counts = bench.make_counts(wl)
n_cells, n_genes = counts.shape; n_aug = n_cells + n_cells//4
omega, npi = _pca_plan(n_aug, n_genes, 30, 0)
h = _capi.Handle(0); h.upload_counts(counts)
rng = np.random.default_rng(0)
kw = dict(pseudocount=0.1, standard_scaling=False, n_comp=30, n_power_iter=npi, n_host_threads=int(sys.argv[1]) if len(sys.argv)>1 else 16)
def run(tag, timing=False, sampler=False, reps=2):
    h.set_kernel_timing(timing)
    s = bench.ClockSampler(0)
    if sampler: s.start()
    for r in range(reps):
        par = bench.draw_parents(rng, n_cells, 25)
        t0 = time.perf_counter(); out = h.fit_iterations(par, omega, **kw); dt = time.perf_counter()-t0
        print(tag, 'wall %.0f ms'%(dt*1e3), {k: round(v) for k,v in out['stage_ms'].items()}, flush=True)
    if sampler: s.stop()
    h.set_kernel_timing(False)
run('warm'); run('plain'); run('timing', timing=True); run('sampler', sampler=True); run('both', True, True); run('plain2')
