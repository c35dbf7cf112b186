"""Module containers of the sparseconvnet API (reference: sparseconvnet/sequential.py:9-17,
tables.py:13-55, identity.py:10-12, utils.py:46-66)."""
import torch

from . import native
from .tensor import SparseConvNetTensor


def _like(x, features):
    out = SparseConvNetTensor()
    out.metadata = x.metadata
    out.spatial_size = x.spatial_size
    out.features = features
    return out


def _sum_features(tensors):
    """Elementwise sum of feature matrices sharing one Metadata (our add kernel when no autograd
    graph is being recorded, torch's `+` otherwise so gradients flow)."""
    feats = [t.features for t in tensors]
    acc = feats[0]
    for f in feats[1:]:
        if torch.is_grad_enabled() and (acc.requires_grad or f.requires_grad):
            a, acc = acc, acc + f
            tr = native._tr()
            if tr is not None:  # a training forward is being recorded (program.Trace(training=True))
                tr.add(5, [tr.reg(a), tr.reg(f), tr.new_reg(acc)])
        elif acc.is_cuda and acc.dtype == torch.float32 and acc.is_contiguous() and f.is_contiguous():
            acc = native.add_features(acc, f)
        else:
            acc = acc + f
    return acc


class Sequential(torch.nn.Sequential):
    def input_spatial_size(self, out_size):
        for name in reversed(self._modules):
            out_size = self._modules[name].input_spatial_size(out_size)
        return out_size

    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    append = add

    def insert(self, index, module):
        for i in range(len(self._modules), index, -1):
            self._modules[str(i)] = self._modules[str(i - 1)]
        self._modules[str(index)] = module


class Identity(torch.nn.Module):
    def forward(self, input):
        return input

    def input_spatial_size(self, out_size):
        return out_size


class ConcatTable(torch.nn.Sequential):
    """Feeds one input to every child, returns the list of outputs."""

    def forward(self, input):
        return [m(input) for m in self._modules.values()]

    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def input_spatial_size(self, out_size):
        return self._modules['0'].input_spatial_size(out_size)


class AddTable(torch.nn.Sequential):
    def forward(self, input):
        return _like(input[0], _sum_features(input))

    def input_spatial_size(self, out_size):
        return out_size


class JoinTable(torch.nn.Sequential):
    def forward(self, input):
        f = torch.cat([i.features for i in input], 1) if input[0].features.numel() else input[0].features
        return _like(input[0], f)

    def input_spatial_size(self, out_size):
        return out_size


def add_feature_planes(input):
    return _like(input[0], _sum_features(input))


def concatenate_feature_planes(input):
    return _like(input[0], torch.cat([i.features for i in input], 1))
