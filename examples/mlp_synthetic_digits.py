"""Three-layer MLP trained with Adam on a synthetic 10-class "digits" problem, written exactly the way a user of
bgavran/autodiff writes it (cf. the reference's examples/feedforward_mnist.py) - only the import changes.
There is no network here, so instead of MNIST the data are 28x28 class prototypes plus noise.

    python examples/mlp_synthetic_digits.py [--steps 300] [--batch 100] [--device-weights]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autodiff_b200 as ad

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--batch", type=int, default=100)
ap.add_argument("--hidden", type=int, default=32)
ap.add_argument("--device-weights", action="store_true", help="keep weights / gradients / Adam state in HBM (no host round trip)")
args = ap.parse_args()

rng = np.random.default_rng(0)
prototypes = rng.random((10, 784))


def batch(n):
    labels = rng.integers(0, 10, n)
    return np.clip(prototypes[labels] + 0.35 * rng.standard_normal((n, 784)), 0, 1), labels


np.random.seed(0)
nn = ad.NN([784, args.hidden, 10])
optimizer = ad.Adam(len(nn.w), lr=0.001)
if args.device_weights:
    for w in nn.w:
        w.value = ad.to_device(w.value)

start = time.time()
for step in range(args.steps):
    images, labels = batch(args.batch)
    x = ad.Variable(images, name="images")
    y_true = ad.Variable(np.eye(10)[labels], name="labels")
    cost = ad.Einsum("i->", ad.SoftmaxCEWithLogits(labels=y_true, logits=nn(x))) / args.batch
    grads = ad.grad(cost, nn.w)
    if args.device_weights:
        new_w = optimizer(list(nn.w), [g.eval_device() for g in grads])
    else:
        new_w = optimizer([w() for w in nn.w], [g() for g in grads])          # the reference's idiom
    optimizer.apply_new_weights(nn.w, new_w)
    if step % 50 == 0:
        print("step %4d  cost %.3f  grad norm %.3f  %.2f s" % (step, float(cost()), float(ad.FrobeniusNorm(*grads)()), time.time() - start))
        start = time.time()

images, labels = batch(2000)
probs = ad.Softmax(nn(ad.Variable(images, name="images")))()
print("accuracy on 2000 fresh samples: %.1f %%" % (100.0 * np.mean(np.argmax(probs, -1) == labels)))
