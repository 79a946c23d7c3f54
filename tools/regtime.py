import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from warpsense_b200 import api, fixedpoint as fp
from warpsense_b200.synth import ScanStream
res=50; s=ScanStream(128,1024,512,res)
size=(513,)*3
class V:
    size_=np.array(size,np.int32); offset_=np.array([256]*3,np.int32); pos_=np.zeros(3,np.int32); data_=None
t=api.TSDFCuda(V(),1000,640,res,upload=False); r=api.RegistrationCuda(t)
for k in range(2):
    f=s.frame(k); pos,up=fp.convert_pose_to_gpu(f["pose"],res); t.update_tsdf(f["points_map"],pos,up)
cloud=s.frame(2,prior_pose=s.pose(1))["points_prior"].copy()
for rep in range(3):
    c=cloud.copy(); T,it=r.register_cloud(c,np.eye(4,dtype=np.float32),20,0.1,0.0,res)
tr=r.trace()[:20,:4].astype(np.int64)
d=np.diff(tr,axis=1)
print("per-iteration us: accumulate+reduce %.2f  sum/barrier %.2f  solve %.2f ; iteration period %.2f" % (d[:,0].mean()/1e3, d[:,1].mean()/1e3, d[:,2].mean()/1e3, np.diff(tr[:,0]).mean()/1e3))
print(d[5:10]/1e3)
