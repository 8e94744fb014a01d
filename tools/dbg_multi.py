import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import oracle_lib
from dbot_ros_b200 import CameraData, ObjectModel, Parameters, RbSensorBuilder, synth
PRM = synth.default_parameters()
mesh, ds, n = "three20k", 1, 24
rows, cols = synth.resolution(ds); K = synth.camera_matrix(ds)
V,T = synth.mesh_set(mesh)
for cull in (True, False):
    orc = oracle_lib.Oracle(rows,cols,K,V,T,PRM,cull=1 if cull else 0)
    sensor = RbSensorBuilder(ObjectModel(V, T), CameraData(K, rows, cols), Parameters(sample_count=n, cull_backfaces=cull)).build()
    rng = synth.rng_for(2)
    gt = synth.ground_truth_poses(len(V), rng)
    frame = synth.make_frame(orc.render, gt, rows, cols, rng)
    poses = synth.particle_poses(gt, n, rng)
    orc.set_observation(frame); sensor.set_observation(frame)
    want = orc.loglikes_poses(poses, np.zeros(n,np.int32), False)
    for rep in range(3):
        got = sensor.loglikes_poses(poses, np.zeros(n,np.int32), False)
        bad = np.nonzero(np.abs(got-want)/np.maximum(1,np.abs(want))>1e-4)[0]
        print('cull',cull,'rep',rep,'bad particles',bad.tolist(), (got-want)[bad])
    for i in range(n):
        a = sensor.render(poses[i]); b = orc.render(poses[i])
        d = a.view(np.uint32)!=b.view(np.uint32)
        if d.any():
            rr,cc = np.nonzero(d)
            fin = np.isfinite(b); wr,wc = np.nonzero(fin)
            print('  render differs particle',i,'npix',d.sum(),'rows',rr.min(),rr.max(),'cols',cc.min(),cc.max(),'window rows',wr.min(),wr.max(),'cols',wc.min(),wc.max(), 'missing', int((np.isinf(a)&fin).sum()), 'extra', int((np.isfinite(a)&~fin).sum()))
