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
poses = trk.track_batch([frames])[0]
print("poses", poses[:, [3, 7, 11]])
want, st = O.track_sequence_model(ocfg, otcfg, frames)
print("oracle", want[:, [3, 7, 11]], st)
print("traj equal", np.array_equal(poses.view(np.uint32), want.view(np.uint32)))
vol = trk.read_volume()
# oracle volume after frame 0 only
ov = O.tsdf_new(otcfg)
of0 = O.OFrame(ocfg, frames[0])
d0 = trk.debug_read(B.DBG_DEPTH, 0, 0)
print("depth0 equal", np.array_equal(d0, of0.depth(0)))
trk.reset()
trk.track_batch([frames[:1]])
v1 = trk.read_volume()
O.tsdf_integrate(ocfg, otcfg, ov, of0.depth(0), IDENT)
diff = (v1 != ov).any(-1)
print("after frame0: differing voxels", diff.sum(), "observed dev", (v1[..., 1] > 0).sum(), "oracle", (ov[..., 1] > 0).sum())
idx = np.argwhere(diff)[:10]
for z, y, x in idx:
    print((x, y, z), v1[z, y, x], ov[z, y, x])
for lvl in range(3):
    vm, nm = O.tsdf_raycast(ocfg, otcfg, ov, IDENT, lvl)
    dv, dn = trk.read_model(B.DBG_VERTEX, lvl), trk.read_model(B.DBG_NORMAL, lvl)
    print("level", lvl, "vertex equal", np.array_equal(dv.view(np.uint32), vm.view(np.uint32)), (dv != vm).any(-1).sum(),
          "normal equal", np.array_equal(dn.view(np.uint32), nm.view(np.uint32)), (dn != nm).any(-1).sum(), "valid", (vm[..., 3] > 0).mean())
vm, nm = O.tsdf_raycast(ocfg, otcfg, ov, IDENT, 0)
dv, dn = trk.read_model(B.DBG_VERTEX, 0), trk.read_model(B.DBG_NORMAL, 0)
d = np.argwhere((dv != vm).any(-1))
print("dev valid", (dv[..., 3] > 0).mean(), "oracle valid", (vm[..., 3] > 0).mean())
for (r, c) in d[:: max(1, len(d) // 12)][:12]:
    print((r, c), "dev", dv[r, c], "orc", vm[r, c])
trk.debug_raycast(IDENT)
dv2 = trk.read_model(B.DBG_VERTEX, 0)
print("debug_raycast == tracked raycast", np.array_equal(dv2, dv))
