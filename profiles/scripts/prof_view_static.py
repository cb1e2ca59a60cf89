import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import torch
from dynibar_b200 import synthetic, render_ray as rr
DEV='cuda:0'
R,S=512,128
batch, feat_c, feat_f, frame, t, offs = synthetic.make_scene(rays=R, seed=0)
model, args = synthetic.make_model(64,64)
d=lambda x: synthetic.to_device(x, DEV)
b, ff = d(batch), d(feat_f)
m=synthetic.model_to(model, DEV)
pts, z, s = rr.sample_along_camera_ray(b['ray_o'], b['ray_d'], b['depth_range'], S, True, True)
fcl = rr.featmaps_channels_last(ff[2])
for i in range(3):
  raw, mask = rr.net_static_fused(m.net_fine_st, pts, b['ray_o'], b['ray_d'], b['camera'], b['static_src_rgbs'], b['static_src_cameras'], fcl)
torch.cuda.synchronize()
print('ok', raw.shape)
