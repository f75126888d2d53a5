import sys; sys.path.insert(0, ".")
import numpy as np
import autodiff_b200 as ad
ad.init(0)
rng = np.random.default_rng(0)
n = (512, 512)
x, y = rng.standard_normal(n), rng.random(n) + 0.5
X, Y = ad.Variable(x), ad.Variable(y)
lab = (rng.random(n) > 0.5).astype(float)
for _ in range(2):
    ad.SigmoidCEWithLogits(labels=ad.Variable(lab), logits=ad.Variable(x))()
    g = ad.grad(ad.Einsum("ab->", ad.SigmoidCEWithLogits(labels=ad.Variable(lab), logits=X)), [X])[0](); X.cached = None
    ad.NormalDistribution(ad.Variable(x), 0.3, 1.7)()
    ad.SquaredDifference(ad.Variable(x), ad.Variable(y))()
    ad.grad(ad.Einsum("ab->", ad.SquaredDifference(X, Y)), [X])[0]()
    ad.Pow(ad.Variable(y), 3)()
    ad.Log(ad.Variable(y))(); ad.Sigmoid(ad.Variable(x))()
    (ad.Variable(x) / ad.Variable(y))(); (ad.Variable(x) - ad.Variable(y))(); (-ad.Variable(x))()
    sm = ad.Softmax(ad.Variable(x)); sm()
    ad.grad(ad.Einsum("ab->", ad.Softmax(X) * Y), [X])[0]()
    for opt in (ad.SGDOptimizer(0.1), ad.Momentum(1, 0.1), ad.Adagrad(1, 0.1), ad.Adam(1, 0.1)):
        p = [ad.to_device(x)]
        for _ in range(3):
            p = opt(p, [ad.to_device(y)])
open("gpurun_out/ewise_stats_extra.txt", "w").write(ad.ewise_stats())
