"""Diagnostic (GPU): meta-gradient of a C2 meta-batch vs the oracle for each layer implementation."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmeta_b200 import _lib
from gmeta_b200.meta import Meta
from gmeta_b200.synthetic import make_dataset
from oracle import gmeta_oracle as O
from tests import helpers as H

ds = make_dataset('C2')
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ds.update_step = K
mb = ds.sample_meta_batch(np.random.default_rng(41), 3)
xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
torch.manual_seed(222)
m0 = Meta(ds.args(), ds.config()).to('cuda')
params = [p.detach().cpu().clone().requires_grad_(True) for p in m0.net.parameters()]
om = O.OracleMeta(ds.args(), ds.config(), params=[p.detach().clone().requires_grad_(True) for p in params])
om.forward([H.to_ograph(x) for x in xs], ys, [H.to_ograph(x) for x in xq], yq, cs, cq, ns, nq, gs, gq, ds.feats)
om64 = None
for impl, pruned in ((_lib.IMPL_SIMT, True), (_lib.IMPL_SIMT, False), (_lib.IMPL_AUTO, True), (_lib.IMPL_AUTO, False)):
    args = ds.args(); args.impl = impl; args.pruned_forward = pruned
    torch.manual_seed(222)
    m = Meta(args, ds.config()).to('cuda')
    m.return_meta_grad = True
    accs = m(*mb, ds.feats)
    out = []
    for k, (g, r) in enumerate(zip(m.last["meta_grad"], om.last_grads)):
        err = (g.cpu().double() - r.double()).abs()
        out.append("g%d %.2e/%.2e" % (k, float(err.max()), float(r.abs().max())))
    print("impl", impl, "pruned", pruned, "K", K, "loss", m.last["loss_q"], om.last_loss_q, " ".join(out))
    if impl == _lib.IMPL_SIMT and pruned:
        g3 = m.last["meta_grad"][3].cpu().double(); r3 = om.last_grads[3].double()
        d = (g3 - r3)
        idx = torch.argsort(d.abs(), descending=True)[:8]
        print("  worst b2 entries", [(int(i), float(g3[i]), float(r3[i])) for i in idx])
