"""Three steps of the 1024-row shard of configs[2] (one rank of an 8-GPU job, no communication) for an ncu launch list."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autodiff_b200 as ad
from autodiff_b200 import dp
ad.init(0)
np.random.seed(1337)
local = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sizes, gbatch = [4096] * 9, 8192
nn = ad.NN(sizes)
for w in nn.w:
    w.value = ad.to_device(w.value)
opt = ad.Adam(len(nn.w), lr=1e-3)
rng = np.random.default_rng(0)
x = ad.to_device(rng.standard_normal((local, 4096)))
y = np.zeros((local, 4096)); y[np.arange(local), rng.integers(0, 4096, local)] = 1.0
y = ad.to_device(y)
for _ in range(3):
    l0 = ad.launch_count()
    dp.train_step(nn, opt, x, y, gbatch, None, True)
    torch.cuda.synchronize()
    print("launches", ad.launch_count() - l0)
