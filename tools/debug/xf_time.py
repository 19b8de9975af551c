"""Times ree_cbree_step (split path, Philox) alone.  Env REE_XF_VARIANT / REE_XF_DEBUG select the kernel variant."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rareeventestimation_b200 import _lib
from rareeventestimation_b200.engine import DeviceLsf, Engine
n, d = int(os.environ.get("N", 10_000_000)), int(os.environ.get("D", 100))
eng = Engine.get()
X = eng.philox_normal(n, d, 1, 1)
Y = eng.alloc_ensemble(n, d)
lsf = DeviceLsf(_lib.LSF_AFFINE, d, None, 3.5)
m_noise = eng.zeros(d)
C = torch.eye(d, dtype=torch.float64, device=eng.device) * 0.5 + 0.01
F = eng.empty(d * d); st = eng.empty(8)
eng.chol_scaled(C.reshape(-1).contiguous(), d, 1.0, F, st)
fn = lambda: eng.cbree_step(X, Y, 0.6, m_noise, F, lsf, None, None, seed=7, draw=3)
fn(); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(json.dumps({"variant": os.environ.get("REE_XF_VARIANT", "0"), "debug": os.environ.get("REE_XF_DEBUG", "0"), "n": n,
                  "ms": round(float(np.median(ts)), 3)}))
