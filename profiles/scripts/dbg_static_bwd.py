import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, torch.nn.functional as F
import test_train_gpu as T
from dynibar_b200 import synthetic, autograd as ag, mlp_network as nets
from oracle import dynibar_oracle as orc
DEV='cuda:0'
def run(R,S,V,aa,mrgb, dark=True):
  torch.manual_seed(R*S+V)
  args = synthetic.make_args(aa, mrgb)
  mod = nets.DynibarStatic(args, 32, S)
  g, pts, feat, mask, ray_dir = T._net_inputs(R,S,V,11+V)
  if mrgb and dark: feat[1,2,1,:3]=0.0
  ref_rays=torch.randn(R,6,generator=g); src_rays=torch.randn(R,S,V,6,generator=g)
  ray_diff=torch.cat([F.normalize(torch.randn(R,S,V,3,generator=g),dim=-1), torch.rand(R,S,V,1,generator=g)*0.3+0.7],-1)
  gen=torch.randn(R,S,4,generator=g)
  w=T._leaves(mod); fo=feat.clone().requires_grad_(True)
  want=orc.net_static(w,pts,ref_rays,src_rays,fo,ray_diff,mask,anti_alias_pooling=bool(aa),mask_rgb=bool(mrgb))
  meff = mask*(feat[..., :3].sum(-1,keepdim=True)>1e-3).float() if mrgb else mask
  live=(meff.sum(2)>=1).float(); scale=torch.cat([torch.ones(R,S,3),live],-1)
  (want*gen*scale).sum().backward()
  mod=mod.to(DEV).requires_grad_(True); fd=feat.to(DEV).requires_grad_(True); d=lambda x:x.to(DEV)
  got=ag.net_static(mod,d(pts),d(ref_rays),d(src_rays),fd,d(ray_diff),d(mask))
  print((R,S,V,aa,mrgb,dark),'fwd maxerr',(got.detach().cpu()-want.detach()).abs().max().item())
  (got*d(gen*scale)).sum().backward()
  for k,p in list(mod.named_parameters())+[('rgb_feat',fd)]:
    ref = w[k].grad if k!='rgb_feat' else fo.grad
    e=(p.grad.cpu()-ref).norm().item()/(ref.norm().item()+1e-12)
    if e>1e-3: print('   BAD',k,e, ref.norm().item())
run(4,16,11,1,1); run(4,16,11,1,0); run(4,16,11,0,1); run(6,16,5,1,1); run(6,16,5,1,1,False); run(4,16,8,1,0); run(4,16,12,1,0)
