import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import perturb_b200 as pb
from oracle import port
n = 6000
b1, b2 = pb.synth_catalog(3, n)
L1, L2 = pb.block_to_lines(b1, n), pb.block_to_lines(b2, n)
i = 812
jd, fr = pb.synth_time_grid(10080)
sel = np.linspace(0, 10079, 500).astype(int)
b = pb.SatelliteBatch.from_tles([L1[i]], [L2[i]])
o = port.OracleSatellites.from_tle_lines([L1[i]], [L2[i]], 1, "i", flavor=1)
pv, err = b.propagate(jd[sel], fr[sel])
eo, ro, vo, _ = o.propagate_grid(jd[sel], fr[sel], use_jd=True)
np.set_printoptions(linewidth=200, precision=6)
for k in range(0, 34):
    print(k, err[0, k], eo[0, k], pv[0:3, 0, k], ro[0, k], np.abs(pv[0:3, 0, k] - ro[0, k]).max())
names = b.record_field_names()
for nm in ("cc1", "cc4", "t2cof", "nodecf", "mdot", "no_unkozai", "dedt", "didt", "con41"):
    print(nm, b.record_field(nm)[0], o.field(nm)[0])
