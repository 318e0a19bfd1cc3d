"""Minimal `torchmeta.modules` surface used by the reference: MetaModule
(nn.Module + get_subdict) and MetaSequential (nn.Sequential that forwards an
optional `params` sub-dict to MetaModule children)."""
import re
from collections import OrderedDict

from torch import nn


class MetaModule(nn.Module):
    def meta_named_parameters(self, prefix='', recurse=True):
        return self.named_parameters(prefix=prefix, recurse=recurse)

    def get_subdict(self, params, key=None):
        if params is None:
            return None
        if key is None or key == '':
            return params
        key_re = re.compile(r'^{0}\.(.+)'.format(re.escape(key)))
        out = OrderedDict()
        for k, v in params.items():
            m = key_re.match(k)
            if m is not None:
                out[m.group(1)] = v
        return out


class MetaSequential(nn.Sequential, MetaModule):
    def forward(self, input, params=None):
        for name, module in self._modules.items():
            if isinstance(module, MetaModule):
                input = module(input, params=self.get_subdict(params, name))
            else:
                input = module(input)
        return input
