# first GPU contact: stage parity + e2e on a few pairs, verbose
import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from mdrp_b200 import _native as nv, synth
from oracle import port
ctx = nv.Context(0)
print('pipes TF/s fp64,fp32:', ctx.measure_pipes())
# sampler
for n,it in ((2000,1000),(5,500),(3,64),(10000,2000)):
    s=ctx.sample(n,0,it); st=0; bad=0
    for k in range(it):
        a,st=port.draw_sample(3,n,st)
        if not (a==s[k]).all(): bad+=1
    print('sampler n',n,'iters',it,'bad',bad)
rng=np.random.default_rng(0)
cfgs=(('cfg1_calib_scale',0,port.solve_calib_scale),('cfg2_calib_shift',1,port.solve_calib_shift),('cfg3_shared_focal',2,port.solve_shared_focal),('cfg4_varying_focal',3,port.solve_varying_focal))
for cfg,var,fn in cfgs:
    sc=synth.scene_for(cfg,0)
    if var<2: x1=(sc.x1-synth.PP)/sc.f1; x2=(sc.x2-synth.PP)/sc.f2; ns=800.
    else:
        x1,x2=sc.centred(); ns=(np.linalg.norm(x1,axis=1)+np.linalg.norm(x2,axis=1)).sum()/(np.sqrt(2)*len(x1)); x1=x1/ns; x2=x2/ns
    n=2000
    idx=np.stack([rng.choice(len(x1),3,replace=False) for _ in range(n)])
    x1h=np.concatenate([x1[idx],np.ones((n,3,1))],axis=2); x2h=np.concatenate([x2[idx],np.ones((n,3,1))],axis=2)
    d1=sc.d1[idx]; d2=sc.d2[idx]
    models,counts=ctx.solve(var,x1h,x2h,d1,d2)
    nexact=0;nclose=0;ncnt=0; allm=[]
    for i in range(n):
        ref=fn(x1h[i],x2h[i],d1[i],d2[i])
        if len(ref)!=counts[i]: ncnt+=1; continue
        ok=True; ex=True
        for k,m in enumerate(ref):
            g=models[i,k]
            a=np.r_[m[0],m[1],m[2],m[3],m[4],m[5],m[6]]; b=np.r_[g['q'],g['t'],g['scale'],g['shift1'],g['shift2'],g['f1'],g['f2']]
            if not np.array_equal(a,b,equal_nan=True): ex=False
            if not np.allclose(a,b,rtol=1e-9,atol=1e-12,equal_nan=True): ok=False
            allm.append(g)
        nexact+=ex; nclose+=ok
    print(cfg,'solve: count mismatch',ncnt,'exact',nexact,'close',nclose,'of',n)
    # score
    allm=np.array(allm[:600],dtype=nv.MODEL_DTYPE)
    thr2=(2/ns)**2
    t0=time.perf_counter(); scores,cnts,masks=ctx.score(var,allm,x1,x2,thr2,want_masks=True); t1=time.perf_counter()
    bad=0; badm=0; maxrel=0
    for i,g in enumerate(allm):
        if var<2:
            s,c=port.msac_score_pose(g['q'],g['t'],x1,x2,thr2); mk=port.get_inliers_pose(g['q'],g['t'],x1,x2,thr2)
        else:
            mm=port.make_model(g['q'],g['t'],g['scale'],0,0,g['f1'],g['f2']); F=port.fundamental_from_model(mm)
            s,c=port.msac_score_F(F,x1,x2,thr2); mk=port.get_inliers_F(F,x1,x2,thr2)
        if c!=cnts[i]: bad+=1
        if not np.array_equal(mk,masks[i].astype(bool)): badm+=1
        if s==s: maxrel=max(maxrel,abs(s-scores[i])/abs(s))
    print(cfg,'score: count mismatches',bad,'mask mismatches',badm,'max rel score diff',maxrel,'of',len(allm))
    # refine
    good=[g for g,c in zip(allm,cnts) if c>0.3*len(x1)][:8]
    if good:
        good=np.array(good,dtype=nv.MODEL_DTYPE)
        for loss,iters in (('TRUNCATED',25),('TRUNCATED_CAUCHY',100)):
            bo=nv.bundle_options(max_iterations=iters,loss_type=loss,loss_scale=2/ns)
            out,st=ctx.refine(var,good,x1,x2,sc.d1,sc.d2,(2/16.)**2,1.0,bo)
            for g,o,s_ in zip(good,out,st):
                mm=port.make_model(g['q'],g['t'],g['scale'],g['shift1'],g['shift2'],g['f1'],g['f2'])
                mo,so=port.refine(var,x1,x2,sc.d1,sc.d2,mm,(2/16.)**2,1.0,port.bundle_opt(max_iterations=iters,loss_type=loss,loss_scale=2/ns))
                dq=min(np.abs(np.array(mo.q)-o['q']).max(),np.abs(np.array(mo.q)+o['q']).max())
                print('  refine',loss,'it',s_['iterations'],so.iterations,'cost',s_['cost'],so.cost,'dq %.1e dt %.1e ds %.1e'%(dq,np.abs(np.array(mo.t)-o['t']).max(),abs(mo.scale-o['scale'])))
# e2e
def e2e(cfg,npairs,iters):
    c=synth.CONFIGS[cfg]; var={'calib':1 if c['shift'] else 0,'shared':2,'varying':3}[c['variant']]
    scs=[synth.scene_for(cfg,i) for i in range(npairs)]
    offs=np.r_[0,np.cumsum([len(s.d1) for s in scs])]
    if var<2:
        x1=np.concatenate([s.x1 for s in scs]); x2=np.concatenate([s.x2 for s in scs])
        cams=np.array([[s.f1,s.f1,640,480,s.f2,s.f2,640,480] for s in scs])
    else:
        x1=np.concatenate([s.centred()[0] for s in scs]); x2=np.concatenate([s.centred()[1] for s in scs]); cams=None
    d1=np.concatenate([s.d1 for s in scs]); d2=np.concatenate([s.d2 for s in scs])
    o=nv.default_options(); o.max_iterations=iters;o.min_iterations=iters;o.max_epipolar_error=2.0;o.max_reproj_error=16.0;o.seed=0
    o.estimate_shift=int(c['shift']); o.loss_type=nv.LOSS['TRUNCATED_CAUCHY']; o.loss_scale=1.0
    t0=time.perf_counter(); models,stats,masks=ctx.estimate_batch_host(var,offs,x1,x2,d1,d2,cams,o); t1=time.perf_counter()
    tm,cn=ctx.last_timing()
    print(cfg,'pairs',npairs,'iters',iters,'wall %.3f s'%(t1-t0),{k:round(v,3) for k,v in tm.items()},cn)
    for i,s in enumerate(scs[:6]):
        rop=port.ransac_opt(max_iterations=iters,min_iterations=iters,max_epipolar_error=2.0,max_reproj_error=16.0,seed=0,estimate_shift=c['shift'])
        bop=port.bundle_opt(loss_type='TRUNCATED_CAUCHY',loss_scale=1.0)
        if var<2: m,st,mk=port.estimate(var,s.x1,s.x2,s.d1,s.d2,[s.f1,s.f1,640,480],[s.f2,s.f2,640,480],rop,bop)
        else: m,st,mk=port.estimate(var,*s.centred(),s.d1,s.d2,None,None,rop,bop)
        g=models[i]; ss=stats[i]
        dq=min(np.abs(np.array(m.q)-g['q']).max(),np.abs(np.array(m.q)+g['q']).max())
        print('  pair',i,'ref',st.refinements,st.num_inliers,st.model_score,'gpu',ss['refinements'],ss['num_inliers'],ss['model_score'],'dq %.1e dt %.1e ds %.1e df %.1e'%(dq,np.abs(np.array(m.t)-g['t']).max(),abs(m.scale-g['scale']),abs(m.f1-g['f1'])),'mask diff',(mk!=masks[offs[i]:offs[i+1]].astype(bool)).sum())
for cfg in ('cfg1_calib_scale','cfg2_calib_shift','cfg3_shared_focal','cfg4_varying_focal'):
    e2e(cfg,8,1000)
e2e('cfg2_calib_shift',64,10000)
e2e('cfg2_calib_shift',256,10000)
print('launches',ctx.launch_count)
