"""Latency of the per-pair Python call (the reference's own calling convention) on one GPU:
    gpurun -- python tools/single_pair_latency.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mdrp_b200 import api, synth, _native as nv
for cfg,iters in (('cfg1_calib_scale',1000),('cfg2_calib_shift',10000),('cfg5_roma_calib',1000),('cfg5_roma_calib',10000)):
    c=synth.CONFIGS[cfg]; sc=synth.scene_for(cfg,0)
    c1,c2=sc.camera_dicts()
    ro={'max_iterations':iters,'min_iterations':iters,'max_epipolar_error':2.0,'max_reproj_error':16.0,'monodepth_estimate_shift':c['shift']}
    bo={'loss_type':'TRUNCATED_CAUCHY'}
    for _ in range(3): api.estimate_monodepth_relative_pose(sc.x1,sc.x2,sc.d1,sc.d2,c1,c2,ro,bo)
    t0=time.perf_counter(); n=20
    for _ in range(n): g,info=api.estimate_monodepth_relative_pose(sc.x1,sc.x2,sc.d1,sc.d2,c1,c2,ro,bo)
    dt=(time.perf_counter()-t0)/n
    tm,cn=api.context(0).last_timing()
    print(cfg,iters,'single-pair latency %.3f ms'%(dt*1e3),'device %.3f ms'%tm['device_total'],{k:round(v,3) for k,v in tm.items() if v>0.02})
