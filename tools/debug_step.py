import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import representationlearning_b200 as P
from oracle import rssformer_ref as R
img, lbl = R.synth_batch(2, 64)
sd = R.synth_state_dict(2333)
keys = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k and not k.startswith("headaux.")]
params = {k: sd[k].clone().requires_grad_(True) for k in keys}
mom = [None] * len(keys)
cur = dict(sd); cur.update(params)
out, stats = R.model_forward(cur, img, lbl, training=True)
grads = torch.autograd.grad(out["fc_loss"], [params[k] for k in keys], allow_unused=True)
gref = {k: g.clone() for k, g in zip(keys, grads)}
tot = R.sgd_step([params[k] for k in keys], list(grads), mom, R.poly_lr(0))
print("oracle loss", out["fc_loss"].item(), "grad norm", float(tot))
m = P.build_rssformer(compute_dtype=torch.float32); m.load_state_dict(sd); m.train()
opt = P.FlatSGD(m, bf16_shadow=False)
named = dict(m.named_parameters())
# gradient check before the step
losses = m(img.cuda(), {"cls": lbl.cuda()}); loss = sum(losses.values()); loss.backward()
from representationlearning_b200 import conv
conv.join_wgrad(); torch.cuda.synchronize()
print("cuda loss", loss.item(), "grad norm", float(opt.flat_g.double().norm()))
rows = []
for k in keys:
    g = named[k].grad.detach().cpu()
    r = gref[k]
    if r is None: continue
    rows.append(((g - r).abs().max().item() / (r.abs().max().item() + 1e-12), r.abs().max().item(), k))
rows.sort(reverse=True)
for r in rows[:12]: print("grad relerr %.3g (max|ref| %.3g) %s" % r)
opt.step(1.0); torch.cuda.synchronize()
rows = []
for k in keys:
    a = named[k].detach().cpu(); b = params[k].detach()
    d0 = (b - sd[k]).abs().max().item()
    rows.append(((a - b).abs().max().item() / (d0 + 1e-12), d0, k))
rows.sort(reverse=True)
for r in rows[:8]: print("update relerr %.3g (|dp| %.3g) %s" % r)
