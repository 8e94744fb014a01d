"""Work of alternative span-raster schedules on BASELINE config 2's geometry, counted on the CPU (numpy projection of the
synthetic mesh; no GPU, no oracle): warp rounds of the shipped balanced (triangle, row) units against one triangle per
lane (box scan / row walk, with and without sorting).  DESIGN.md section 4 quotes the result.  usage: python tools/sim_raster.py"""
import numpy as np, sys
sys.path.insert(0,'.')
from dbot_ros_b200 import synth
V,T = synth.mesh_set("icosphere5k"); V=V[0].astype(np.float64); T=T[0]
rng = synth.rng_for(2)
gt = synth.ground_truth_poses(1, rng)
poses = synth.particle_poses(gt, 8, rng)
K = synth.K_VGA
print("gt t", gt[0][1])
for pi in range(3):
    P = poses[pi,0].astype(np.float64); R=P[:9].reshape(3,3); t=P[9:]
    v = V@R.T + t
    u = (v[:,0]/v[:,2])*K[0,0]+K[0,2]; w=(v[:,1]/v[:,2])*K[1,1]+K[1,2]
    X = np.rint(u*256)/256; Y=np.rint(w*256)/256
    a,b,c = T[:,0],T[:,1],T[:,2]
    area2 = (X[b]-X[a])*(Y[c]-Y[a])-(Y[b]-Y[a])*(X[c]-X[a])
    sgn = np.sign(area2)
    # which sign faces camera: majority weighted... use cull: front = sign with smaller z-mean? just take sign count that gives ~half
    front = sgn>0 if (sgn>0).sum()< (sgn<0).sum()*1.2 and True else sgn>0
    # determine via normal·view
    n = np.cross(v[b]-v[a], v[c]-v[a]); facing = (n*v[a]).sum(1) < 0
    front = facing & (sgn!=0)
    # check consistent sign
    cmin=np.ceil(np.minimum(np.minimum(X[a],X[b]),X[c])); cmax=np.floor(np.maximum(np.maximum(X[a],X[b]),X[c]))
    rmin=np.ceil(np.minimum(np.minimum(Y[a],Y[b]),Y[c])); rmax=np.floor(np.maximum(np.maximum(Y[a],Y[b]),Y[c]))
    ok = front & (cmin<=cmax)&(rmin<=rmax)
    idx = np.nonzero(ok)[0]
    bw = (cmax-cmin+1)[idx].astype(int); bh=(rmax-rmin+1)[idx].astype(int)
    # exact fragments per tri and per row via edge functions
    frags=np.zeros(len(idx),int); maxspan=np.zeros(len(idx),int); rows_nonempty=np.zeros(len(idx),int)
    spans=[]
    for k,ti in enumerate(idx):
        xs=np.arange(cmin[ti],cmax[ti]+1); ys=np.arange(rmin[ti],rmax[ti]+1)
        gx,gy=np.meshgrid(xs,ys)
        def E(p,q): return (X[q]-X[p])*(gy-Y[p])-(Y[q]-Y[p])*(gx-X[p])
        s=sgn[ti]
        cov=(E(a[ti],b[ti])*s>=0)&(E(b[ti],c[ti])*s>=0)&(E(c[ti],a[ti])*s>=0)
        rs=cov.sum(1); frags[k]=cov.sum(); maxspan[k]=rs.max() if rs.size else 0; rows_nonempty[k]=(rs>0).sum()
        spans.append(rs)
    nt=len(idx)
    print(f"particle {pi}: survivors {nt} frags {frags.sum()} rows {bh.sum()} bbox px {(bw*bh).sum()} mean bw {bw.mean():.2f} bh {bh.mean():.2f}")
    # scheme A (current): batches of 32; unit rounds = ceil(H/32); per round pixel iterations = max span among 32 units
    nb=(nt+31)//32
    roundsA=0; pixitA=0
    for bi in range(nb):
        sl=slice(32*bi,32*bi+32)
        allspans=np.concatenate(spans[sl]) if len(spans[sl]) else np.zeros(0,int)
        H=len(allspans)
        for q in range(0,H,32):
            roundsA+=1; pixitA+=allspans[q:q+32].max()
    print(f"  A: batches {nb} unit rounds {roundsA} pixel iters {pixitA}  (ideal pixel iters {frags.sum()/32:.0f})")
    # scheme B: per-lane bbox scan: iters per batch = max bbox area
    itB=sum((bw[32*bi:32*bi+32]*bh[32*bi:32*bi+32]).max() for bi in range(nb))
    print(f"  B: bbox-scan iters {itB}  (ideal {(bw*bh).sum()/32:.0f})")
    # scheme C: per-lane row walk flattened: iters = frags + rows per lane, max per batch
    itC=sum((frags[32*bi:32*bi+32]+bh[32*bi:32*bi+32]).max() for bi in range(nb))
    print(f"  C: row-walk iters {itC}  (ideal {(frags+bh).sum()/32:.0f})")
    # scheme D: per-lane rows, each row handles up to 4 px in one step (span<=4 common?) 
    import collections
    allsp=np.concatenate(spans); print("  span hist", np.bincount(allsp)[:12], "mean", allsp.mean())
    # scheme E: sort survivors by bbox area before batching (B sorted)
    order=np.argsort(bw*bh); ar=(bw*bh)[order]
    itE=sum(ar[32*bi:32*bi+32].max() for bi in range(nb))
    print(f"  E: bbox-scan sorted iters {itE}")
    # scheme F: per-lane row per iteration with up to 8-wide mask? iterations = rows; per row cost fixed handles <=W px
    itF=sum(bh[32*bi:32*bi+32].max() for bi in range(nb)); print(f"  F: per-lane rows iters {itF} (ideal {bh.sum()/32:.0f}); spans>4: {(allsp>4).mean():.3f} >6: {(allsp>6).mean():.3f} >8 {(allsp>8).mean():.3f}")
