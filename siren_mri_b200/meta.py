"""Parameter-dict plumbing of the reference's vendored torchmeta subset.

Mirrors (behaviour, not code) torchmeta/modules/module.py:6-27 (MetaModule),
torchmeta/modules/container.py:6-19 (MetaSequential) and torchmeta/modules/utils.py:4-11
(get_subdict): modules whose forward accepts ``params``, an OrderedDict of tensors keyed by the
dotted parameter names, peeled one prefix per nesting level.
"""
from collections import OrderedDict

import torch.nn as nn


def get_subdict(dictionary, key=None):
    """Entries of ``dictionary`` below the prefix ``key + '.'`` with the prefix removed."""
    if dictionary is None:
        return None
    if key is None or key == "":
        return dictionary
    prefix = key + "."
    n = len(prefix)
    return OrderedDict((k[n:], v) for k, v in dictionary.items() if k.startswith(prefix) and len(k) > n)


class MetaModule(nn.Module):
    """nn.Module whose forward takes an optional ``params`` dict (full autograd support)."""

    def meta_named_parameters(self, prefix="", recurse=True):
        # only parameters owned by MetaModules are "meta" parameters
        def members(module):
            return module._parameters.items() if isinstance(module, MetaModule) else []
        for elem in self._named_members(members, prefix=prefix, recurse=recurse):
            yield elem

    def meta_parameters(self, recurse=True):
        for _, p in self.meta_named_parameters(recurse=recurse):
            yield p


class MetaSequential(nn.Sequential, MetaModule):
    """nn.Sequential that threads ``params`` to the MetaModules among its children."""

    def forward(self, input, params=None):
        for name, module in self._modules.items():
            if isinstance(module, MetaModule):
                input = module(input, params=get_subdict(params, name))
            elif isinstance(module, nn.Module):
                input = module(input)
            else:
                raise TypeError("expected nn.Module or MetaModule, got %r" % type(module))
        return input
