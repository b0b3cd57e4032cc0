import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from tests.test_train_gpu import _setup, _oracle_grads
from timbre_trap_b200.framework.train import TrainStep
R, model, sd, c, audio, gt = _setup()
want, _, _ = _oracle_grads(R, sd, c, audio, gt)
ts = TrainStep(model)
out = ts.losses(audio.cuda(), gt.cuda()); ts.backward(out['total'])
tot = sum(float(v.norm())**2 for v in want.values())**0.5
rows = []
for k, p in model.named_parameters():
    g = p.grad.cpu(); w = want[k]
    rows.append((float((g-w).norm())/tot, float((g-w).norm())/max(float(w.norm()),1e-12), float(w.norm()), float(g.norm()), k))
for r in sorted(rows, reverse=True)[:14]: print('abs/total %.4f rel %.3f |want| %.4f |got| %.4f %s' % r)
