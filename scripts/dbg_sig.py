import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'tests'))
import numpy as np, mpmath as mp
mp.mp.prec=200
from test_gpu_math import run
rng = np.random.default_rng(2)
rng.uniform(-40, 40, 4000); rng.uniform(-745, 709, 2000)
xs = np.concatenate([rng.uniform(-100, 100, 4000), rng.standard_normal(2000) * 1e3, [0.0, 1e4, -1e4, 1e308, -1e308]])
sg,_ = run(2, xs)
t = -0.4*xs
e_gpu,_ = run(1, t)
e_cr = np.array([float(mp.e ** mp.mpf(float(v))) if v < 709.78 else np.inf for v in t])
with np.errstate(over='ignore'):
    ref_cr = 1/(e_cr+1)
sel = np.isfinite(e_cr) & (ref_cr > 1e-300)
err = np.abs(sg - ref_cr)/np.array([math.ulp(v) if math.isfinite(v) and v>0 else 1.0 for v in ref_cr])
idx = np.argsort(-np.where(sel, err, 0))[:8]
for i in idx:
    print(f"x={xs[i]!r} t={t[i]!r} e_cr={e_cr[i]!r} e_gpu={e_gpu[i]!r} ref={ref_cr[i]!r} got={sg[i]!r} err={err[i]:.2f} ulps; exp ulp err={(e_gpu[i]-e_cr[i])/math.ulp(e_cr[i]) if math.isfinite(e_cr[i]) else 0:.2f}")
