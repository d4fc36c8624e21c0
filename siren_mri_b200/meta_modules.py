"""Producer of per-sample hypo-network parameters (the `params` operand of the native path).

Mirrors meta_modules.HyperNetwork (/root/reference/meta_modules.py:11-54) and its initialisers
hyper_weight_init / hyper_bias_init (meta_modules.py:303-321): one ReLU FCBlock per parameter of
the hypo module, whose output is reshaped to ``[B, *param_shape]``.  The ReLU MLPs themselves are
ordinary PyTorch (out of the hot path); what matters here is the output contract -- names in
``meta_named_parameters()`` order, leading task axis -- which feeds
``hypo_net(model_input, params=...)`` and receives the per-task weight gradients.
"""
import math
from collections import OrderedDict

import torch
from torch import nn

from . import modules


def _hyper_last_layer_init(m, bias_bound=None):
    """He-normal weights scaled by 1e-2; bias ~ U(+-bias_bound) (default 1/fan_in of this layer)."""
    if hasattr(m, "weight"):
        nn.init.kaiming_normal_(m.weight, a=0.0, nonlinearity="relu", mode="fan_in")
        with torch.no_grad():
            m.weight.mul_(1.0e-2)
    if hasattr(m, "bias"):
        bound = bias_bound
        if bound is None:
            bound = 1.0 / nn.init._calculate_fan_in_and_fan_out(m.weight)[0]
        with torch.no_grad():
            m.bias.uniform_(-bound, bound)


def hyper_weight_init(m, in_features_main_net):
    """Last layer of a hypernetwork head predicting a weight matrix (meta_modules.py:303-310)."""
    _hyper_last_layer_init(m, 1.0 / in_features_main_net)


def hyper_bias_init(m):
    """Last layer of a hypernetwork head predicting a bias vector (meta_modules.py:313-321)."""
    _hyper_last_layer_init(m, None)


class HyperNetwork(nn.Module):
    def __init__(self, hyper_in_features, hyper_hidden_layers, hyper_hidden_features, hypo_module):
        super().__init__()
        self.names = []
        self.nets = nn.ModuleList()
        self.param_shapes = []
        for name, param in hypo_module.meta_named_parameters():
            self.names.append(name)
            self.param_shapes.append(param.size())
            hn = modules.FCBlock(in_features=hyper_in_features, out_features=int(math.prod(param.size())),
                                 num_hidden_layers=hyper_hidden_layers, hidden_features=hyper_hidden_features,
                                 outermost_linear=True, nonlinearity="relu")
            self.nets.append(hn)
            if "weight" in name:
                fan = param.size()[-1]
                self.nets[-1].net[-1].apply(lambda m, fan=fan: hyper_weight_init(m, fan))
            elif "bias" in name:
                self.nets[-1].net[-1].apply(lambda m: hyper_bias_init(m))

    def forward(self, z):
        """z: [B, hyper_in_features] -> OrderedDict name -> [B, *param_shape]."""
        params = OrderedDict()
        for name, net, shape in zip(self.names, self.nets, self.param_shapes):
            params[name] = net(z).reshape((-1,) + tuple(shape))
        return params
