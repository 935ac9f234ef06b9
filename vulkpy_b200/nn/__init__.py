"""
Neural Network Module (:mod:`vulkpy_b200.nn`; reference: vulkpy/nn/__init__.py)

>>> import vulkpy_b200 as vk
>>> from vulkpy_b200 import nn
>>> gpu = vk.GPU()
>>> opt = nn.Adam(gpu, lr=1e-4)
>>> net = nn.Sequence([nn.Dense(gpu, 3, 32, w_opt=opt, b_opt=opt), nn.ReLU(),
...                    nn.Dense(gpu, 32, 4, w_opt=opt, b_opt=opt), nn.Softmax()],
...                   nn.CrossEntropyLoss())
>>> pred_y, loss = net.train(x, y)
"""
from .core import Optimizer, OptimizerState, Loss, Regularizer, Module
from .initializers import Constant, HeNormal
from .optimizers import SGD, SGDState, Adam, AdamState, AdaGrad, AdaGradState
from .layers import Dense, ReLU, Sigmoid, Softmax
from .losses import CrossEntropyLoss, SoftmaxCrossEntropyLoss, MSELoss, HuberLoss, MixLoss
from .regularizers import Lasso, Ridge, Elastic
from .models import Sequence
from . import parameters
