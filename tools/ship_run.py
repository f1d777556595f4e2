import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from sfsim_b200 import atmosphere_lut as al, _lib
b = al.AtmosphereLutBuilder()
for i in range(3):
    t=time.time(); b.run(); b.sync(); print('run', i, time.time()-t)
for n, ms in b.stage_times(): print('%-40s %8.3f ms' % (n, ms))
print(b.work())
out = b.download()
for n, o in zip(al.FILE_NAMES, out): print(n, o.shape, float(o.min()), float(o.max()), bool(np.isfinite(o).all()))
t=time.time(); al.generate_tables(); print('generate e2e', time.time()-t)
t=time.time(); al.generate_tables(); print('generate e2e', time.time()-t)
