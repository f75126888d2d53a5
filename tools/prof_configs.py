"""One step of C4 (conv net) or a short C5 (char-rnn, second order) for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autodiff_b200 as ad
from autodiff_b200 import workloads as wl
from autodiff_b200.core.engine import evaluate_many

ad.init(0)
which = sys.argv[1] if len(sys.argv) > 1 else "c4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rng = np.random.default_rng(1337)
if which == "c4":
    batch = 256
    xd = ad.to_device(rng.standard_normal((batch, 64, 64, 3)))
    yd = ad.to_device(np.eye(10)[rng.integers(0, 10, batch)])
    pd = [ad.to_device(p) for p in wl.conv_net_params(rng)]
    for _ in range(reps):
        X, Y = ad.Variable(xd, name="x"), ad.Variable(yd, name="y")
        P = [ad.Variable(p, name="p%d" % i) for i, p in enumerate(pd)]
        loss = wl.conv_net(ad, X, Y, *P)
        l0 = ad.launch_count()
        evaluate_many(ad.grad(loss, P) + [loss])
        torch.cuda.synchronize()
        print("launches", ad.launch_count() - l0)
else:
    T = 8
    prm = [ad.to_device(p) for p in wl.char_rnn_params(rng, 1024, 71)]
    oh = wl._SlicedDevice(ad.to_device(np.eye(71)[rng.integers(0, 71, (512, T))]))
    for _ in range(reps):
        P = [ad.Variable(p, name="p%d" % i) for i, p in enumerate(prm)]
        loss = wl.char_rnn(ad, oh, *P)
        top = ad.grad(ad.grad(loss, [P[0]])[0], [P[0]])[0]
        l0 = ad.launch_count()
        top.eval_device()
        torch.cuda.synchronize()
        print("launches", ad.launch_count() - l0)
