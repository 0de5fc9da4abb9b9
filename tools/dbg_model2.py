import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import youth_pkg, oracle_py as O
pkg = youth_pkg.load()
from slam_rgbd_b200 import binding as B
from test_model import SMALL, SMALL_T, IDENT, pose_of
cfg = pkg.default_config(**SMALL, batch=4, traj_capacity=16)
tcfg = pkg.tsdf_config(**SMALL_T)
trk = B.Tracker(cfg); trk.enable_model(tcfg)
ocfg = O.config_from(cfg); otcfg = O.tsdf_config_from(tcfg)
frames = pkg.synth_sequence(3, 160, 120)
got = trk.track_batch([frames])[0]
want, st = O.track_sequence_model(ocfg, otcfg, frames)
print("traj equal", np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.abs(got - want).max())
vol = trk.read_volume()
ovol = vol.copy()
trk.debug_integrate(0, IDENT)
O.tsdf_integrate(ocfg, otcfg, ovol, O.OFrame(ocfg, frames[0]).depth(0), IDENT)
v1 = trk.read_volume()
diff = (v1 != ovol).any(-1)
print("differing", diff.sum(), "changed dev", (v1 != vol).any(-1).sum(), "changed orc", (ovol != vol).any(-1).sum())
for z, y, x in np.argwhere(diff)[:12]:
    print((x, y, z), "before", vol[z, y, x], "dev", v1[z, y, x], "orc", ovol[z, y, x])
