"""A small classifier on Iris with vulkpy.nn (Dense-ReLU-Dense-Softmax, cross-entropy), the workflow of
the reference's example/02-nn.py with the shuffle and the arg-max done on the device.

    python examples/02_nn_iris.py [--optimizer adam|sgd] [--nepoch 100]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sklearn.datasets import load_iris

import vulkpy_b200 as vk
from vulkpy_b200 import nn

ap = argparse.ArgumentParser()
ap.add_argument("--optimizer", choices=["adam", "sgd"], default="adam")
ap.add_argument("--nepoch", type=int, default=100)
ap.add_argument("--batch", type=int, default=32)
args = ap.parse_args()

gpu = vk.GPU()
iris = load_iris()
x_all = ((iris.data - iris.data.mean(axis=0)) / iris.data.std(axis=0)).astype(np.float32)
y_all = iris.target.astype(np.uint32)
order = np.random.default_rng(0).permutation(len(x_all))
test, train = order[:30], order[30:]

opt = (lambda: nn.Adam(gpu, lr=0.01)) if args.optimizer == "adam" else (lambda: nn.SGD(0.05))
net = nn.Sequence([nn.Dense(gpu, 4, 32, w_opt=opt(), b_opt=opt(), w_init=nn.HeNormal(gpu, 4, seed=1)), nn.ReLU(),
                   nn.Dense(gpu, 32, 3, w_opt=opt(), b_opt=opt(), w_init=nn.HeNormal(gpu, 32, seed=2)), nn.Softmax()],
                  nn.CrossEntropyLoss())

X = vk.Array(gpu, data=x_all[train])
Y = vk.U32Array(gpu, data=y_all[train]).to_onehot(3)
Xt = vk.Array(gpu, data=x_all[test])
rng = vk.random.Xoshiro128pp(gpu, seed=3)

t0 = time.perf_counter()
for epoch in range(args.nepoch):
    perm = rng.permutation(len(train))                    # device shuffle (README.md:77 lists it as missing upstream)
    xs, ys = X.gather(perm, axis=0), Y.gather(perm, axis=0)
    total = 0.0
    for lo in range(0, len(train) - args.batch + 1, args.batch):
        idx = vk.U32Array(gpu, data=np.arange(lo, lo + args.batch, dtype=np.uint32))
        _, loss = net.train(xs.gather(idx, axis=0), ys.gather(idx, axis=0))
        total += float(np.asarray(loss).reshape(-1)[0])    # reduce="mean" leaves a 0-d array, as in the reference
    if epoch % max(1, args.nepoch // 5) == 0 or epoch == args.nepoch - 1:
        pred = np.asarray(net.predict(Xt).argmax(axis=1))
        print(f"epoch {epoch:4d}  train loss {total / (len(train) // args.batch):.4f}  test accuracy {(pred == y_all[test]).mean():.3f}")
print(f"{args.nepoch} epochs in {time.perf_counter() - t0:.2f} s")
assert (pred == y_all[test]).mean() >= 0.6   # chance is 0.33; the reference's diagonal-only softmax gradient learns slowly
