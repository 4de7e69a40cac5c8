import torch, sys, numpy as np
sys.path.insert(0,'/root/repo')
import multimodal_vqvae_compression_audio_tactile_b200 as pkg
from oracle import cases, proposed, training as otr
dev=torch.device('cuda',0)
case=otr.TRAIN_CASE
ref=cases.build_reference_style_model(proposed.ProposedEval, case)
net=pkg.build_proposed(case["books"], case["K"]); net.load_state_dict(ref.state_dict()); net=net.to(dev).eval(); net.precision='f32'
a,t=cases.codec_inputs(case)
g=np.load('tests/golden/train_step.npz')
out=net.forward_step(a.to(dev), t.to(dev))
with torch.no_grad():
    y2=net.forward_eval(a.to(dev), t.to(dev))
    idx_eval=net.last_indices.clone()
print('train vs eval y', float((out['y_hat'].detach()-y2[..., :out['y_hat'].shape[-1]]).abs().max()))
print('train vs golden', float((out['y_hat'].detach().cpu()-torch.from_numpy(g['y_hat'])).abs().max()))
print('eval vs golden', float((y2.cpu()[..., :7992]-torch.from_numpy(g['y_hat'])).abs().max()))
tr={}
yo=ref.forward_eval(a,t,None,trace=tr)
print('oracle eval vs golden(train)', float((yo[..., :7992]-torch.from_numpy(g['y_hat'])).abs().max()))
print('idx eval vs oracle', int((idx_eval.cpu()!=tr['idx']).sum()), 'of', idx_eval.numel())
d=(out['y_hat'].detach().cpu()-torch.from_numpy(g['y_hat'])).abs()
print('where', d.argmax().item(), d.mean().item())
# z_run compare
keep={}
o2=otr.forward_step(ref,a,t,keep)
zr=keep['z_run'].detach()
dz=(out['z_run'].detach().cpu()-zr).abs()
print('z_run diff max', float(dz.max()), 'at token', int(dz.amax(dim=(0,1)).argmax()), dz.amax(dim=(0,1)))
