"""CPU only (needs oracle/_ref): the C restatement against the reference binary end to end on random small noisy
scenes over all options (early-termination parameters, bundle tolerances, losses, PINHOLE / SIMPLE_PINHOLE cameras,
float32 input).  Prints the mismatch classification quoted in DESIGN.md §2.

    python tools/oracle_vs_wheel_scan.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
from mdrp_b200 import synth
from oracle import port, ref_wheel
pl=ref_wheel.poselib()
rng=np.random.default_rng(4242)
bad=0; tot=0; fails=[]; kinds={}; detail={}
LOSS=["TRIVIAL","TRUNCATED","HUBER","CAUCHY","TRUNCATED_CAUCHY"]
for variant in (0,1,2,3):
    for _ in range(400):
        n=int(rng.choice([30,80,200]))
        kw=dict(outlier_ratio=float(rng.choice([0.2,0.5])), sigma_px=float(rng.choice([0.5,2.0])), depth_noise=float(rng.choice([0.01,0.1])))
        if variant==1: kw.update(shift1=0.3,shift2=-0.2)
        if variant==3: kw.update(f1=700.,f2=900.)
        sc=synth.make_scene(int(rng.integers(0,10**6)), n, **kw)
        mi=int(rng.choice([1,20,100])); mx=int(rng.choice([100,500]))
        mult=float(rng.choice([1.0,3.0,5.0])); sp=float(rng.choice([0.9,0.99,0.9999]))
        te=float(rng.choice([1.0,2.0,4.0])); tr=float(rng.choice([8.0,16.0,32.0])); seed=int(rng.integers(0,2**31))
        b=dict(max_iterations=int(rng.choice([5,100])), loss_type=str(rng.choice(LOSS)), loss_scale=float(rng.choice([0.3,1.0,2.0])),
               gradient_tol=float(rng.choice([1e-10,1e-6])), step_tol=float(rng.choice([1e-8,1e-5])), initial_lambda=float(rng.choice([1e-3,1e-1])),
               min_lambda=float(rng.choice([1e-10,1e-6])), max_lambda=float(rng.choice([1e10,1e2])))
        rd={"max_iterations":mx,"min_iterations":min(mi,mx),"max_epipolar_error":te,"max_reproj_error":tr,"seed":seed,"dyn_num_trials_mult":mult,"success_prob":sp,"monodepth_estimate_shift":variant==1}
        ro=port.ransac_opt(max_iterations=mx,min_iterations=min(mi,mx),dyn_num_trials_mult=mult,success_prob=sp,max_reproj_error=tr,max_epipolar_error=te,seed=seed,estimate_shift=variant==1)
        bo=port.bundle_opt(**b)
        if variant<2:
            fx1,fy1,cx1,cy1=sc.f1*float(rng.uniform(.95,1.05)),sc.f1,640.+float(rng.uniform(-20,20)),480.
            cam={"model":"PINHOLE","width":-1,"height":-1,"params":[fx1,fy1,cx1,cy1]}; cam2={"model":"SIMPLE_PINHOLE","width":-1,"height":-1,"params":[sc.f2,640.,480.]}
            x1=sc.x1.astype(np.float32).astype(np.float64) if rng.integers(0,2) else sc.x1
            g,info=pl.estimate_monodepth_relative_pose(x1.astype(np.float32) if (x1.astype(np.float32)==x1).all() else x1,sc.x2,sc.d1,sc.d2,cam,cam2,rd,b)
            m,st,mask=port.estimate(variant,x1,sc.x2,sc.d1,sc.d2,[fx1,fy1,cx1,cy1],[sc.f2,sc.f2,640,480],ro,bo)
            ref=np.r_[np.array(g.pose.q),np.array(g.pose.t).ravel(),g.scale,g.shift1,g.shift2]
            got=np.r_[np.array(m.q),np.array(m.t),m.scale,m.shift1,m.shift2]
        else:
            a1,a2=sc.x1-[640.,480.],sc.x2-[640.,480.]
            f=pl.estimate_monodepth_shared_focal_relative_pose if variant==2 else pl.estimate_monodepth_varying_focal_relative_pose
            ip,info=f(a1,a2,sc.d1,sc.d2,rd,b); g=ip.geometry
            m,st,mask=port.estimate(variant,a1,a2,sc.d1,sc.d2,None,None,ro,bo)
            ref=np.r_[np.array(g.pose.q),np.array(g.pose.t).ravel(),g.scale,ip.camera1.focal(),ip.camera2.focal()]
            got=np.r_[np.array(m.q),np.array(m.t),m.scale,m.f1,m.f2]
        tot+=1
        same=(st.refinements,st.iterations,st.num_inliers)==(info["refinements"],info["iterations"],info["num_inliers"])
        close=np.allclose(got,ref,rtol=1e-6,atol=1e-8) or st.num_inliers<10
        if not (same and close):
            bad+=1; kinds[(variant,"stats" if not same else "model")]=kinds.get((variant,"stats" if not same else "model"),0)+1; detail[(variant,(st.refinements!=info["refinements"]),(st.iterations!=info["iterations"]),(st.num_inliers!=info["num_inliers"]))]=detail.get((variant,(st.refinements!=info["refinements"]),(st.iterations!=info["iterations"]),(st.num_inliers!=info["num_inliers"])),0)+1; fails.append((variant,sc,rd,b)); (lambda *a: None)(variant,n,"oracle",(st.refinements,st.iterations,st.num_inliers),"wheel",(info["refinements"],info["iterations"],info["num_inliers"]),"close",close, {k:b[k] for k in ("loss_type","max_iterations","loss_scale")}, "mult",mult,"sp",sp,"mi",mi,"mx",mx,"te",te,"tr",tr)
print(tot,bad); print(kinds); print("variant, refinements differ, iterations differ, inliers differ:", detail)

