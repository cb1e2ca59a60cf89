import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import torch, scenes
from dynibar_b200 import synthetic, render_ray as rr
from dynibar_b200.projection import Projector
from oracle import dynibar_oracle as orc
torch.set_printoptions(linewidth=200, precision=5, sci_mode=False)
DEV='cuda:0'
fx=torch.load(os.path.join(ROOT,'tests/golden/mv_small.pt'),weights_only=False)
cfg, st = fx['cfg'], fx['stages']
batch, feat_c, feat_f, frame, t, offs, model, args = scenes.build(cfg)
d=lambda x: synthetic.to_device(x, DEV)
b, fc = d(batch), d(feat_c)
wst = orc._sd(model.net_coarse_st)
m = synthetic.model_to(model, DEV)
pts=d(st['pts'])
for k in ('rgb_feat_st','ray_diff_st','mask_st','src_plucker','ref_plucker','pts'):
  print(k, st[k].shape, st[k].stride(), st[k].is_contiguous())
raw = rr.net_static_forward(m.net_coarse_st, pts, d(st['ref_plucker']), d(st['src_plucker']), d(st['rgb_feat_st']), d(st['ray_diff_st']), d(st['mask_st'])).cpu()
want = st['raw_st']
with torch.no_grad():
  orc_raw = orc.net_static(wst, st['pts'], st['ref_plucker'], st['src_plucker'], st['rgb_feat_st'], st['ray_diff_st'], st['mask_st'], True, False)
print('oracle vs golden', (orc_raw-want).abs()[st['mask_st'].sum(2).expand(-1,-1,4)>0].max())
err=(raw-want).abs()
ms=st['mask_st'][...,0].sum(-1)
for nv in range(0,6):
  sel=ms==nv
  if sel.any(): print('nvalid',nv,'count',sel.sum().item(),'max err rgb',err[...,:3][sel].max().item(),'sigma',err[...,3][sel].max().item())
print('err per ray (sigma):', err[...,3].clamp(max=1).amax(1))
r=int(err[...,3].clamp(max=1).amax(1).argmax())
print('ray',r,'nvalid',ms[r]); print('got',raw[r,:,3]); print('want',want[r,:,3])
# now with contiguous copies
c=lambda x: d(x.contiguous())
raw2 = rr.net_static_forward(m.net_coarse_st, pts, c(st['ref_plucker']), c(st['src_plucker']), c(st['rgb_feat_st']), c(st['ray_diff_st']), c(st['mask_st'])).cpu()
print('contig-first err', (raw2-want).abs()[...,3].clamp(max=1).max().item(), 'same as before?', torch.equal(raw,raw2))
