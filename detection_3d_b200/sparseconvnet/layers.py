"""Layer modules and autograd Functions of the sparseconvnet API used by the Detection_3D backbone.

Same class names, constructor signatures, attribute names and state_dict keys / shapes as the
reference (sparseconvnet/{ioLayers,submanifoldConvolution,convolution,deconvolution,
batchNormalization,metadata,utils}.py); every native call goes through `native` (the C-ABI).
"""
import torch
from torch.autograd import Function
from torch.nn import Module, Parameter

from . import native
from .tensor import SparseConvNetTensor


def toLongTensor(dimension, x):
    if isinstance(x, torch.Tensor) and x.dtype == torch.int64 and not x.is_cuda:
        return x
    if isinstance(x, (list, tuple)):
        assert len(x) == dimension
        return torch.LongTensor([int(v) for v in x])
    return torch.LongTensor(dimension).fill_(int(x))


def optionalTensor(module, name):
    return getattr(module, name) if hasattr(module, name) else torch.Tensor()


def optionalTensorReturn(t):
    return t if t.numel() else None


def Metadata(dim):
    """reference: sparseconvnet/metadata.py:16-17"""
    if dim != 3:
        raise RuntimeError("detection_3d_b200 is built for dimension 3 only (the Detection_3D backbone)")
    return native.Metadata_3()


import os as _os
_PREFETCH = _os.environ.get("SCN_PREFETCH", "1") != "0"  # developer switch: build rulebooks lazily on the calling thread only


class _NoCtx(object):
    """Stand-in for the autograd context when no graph is being recorded."""

    def save_for_backward(self, *tensors):
        pass


def _run(fn, *args):
    """fn.apply(*args) when autograd is recording, else the forward alone (Function.apply costs
    ~15 us per call and the backbone has ~100 calls per forward)."""
    if torch.is_grad_enabled():
        return fn.apply(*args)
    return fn.forward(_NoCtx(), *args)


def _counters():
    import detection_3d_b200.sparseconvnet as pkg
    return pkg


def _conv_weight(volume, groups, n_in, n_out):
    std = (2.0 * groups / n_in / volume) ** 0.5
    return Parameter(torch.Tensor(volume, groups, n_in // groups, n_out // groups).normal_(0, std))


# ------------------------------------------------------------------------------- InputLayer
class InputLayer(Module):
    """(coords LongTensor [N, 3|4], features [N, C][, batch_size]) -> SparseConvNetTensor.
    mode: 0 unique, 1 first, 2 last (as implemented by the reference's rules, IOLayersRules.h:100-110),
    3 sum, 4 mean.  reference: sparseconvnet/ioLayers.py:15-65."""

    def __init__(self, dimension, spatial_size, mode=3):
        Module.__init__(self)
        self.dimension = dimension
        self.spatial_size = toLongTensor(dimension, spatial_size)
        self.mode = mode
        self.device = None
        self.prefetch_ops = None  # rulebook requests of an earlier forward of the owning network (a hint)

    def to(self, device):
        self.device = device
        return self

    def forward(self, input):
        out = SparseConvNetTensor(metadata=Metadata(self.dimension), spatial_size=self.spatial_size)
        coords = input[0]
        if not coords.is_cuda:  # the reference keeps coordinates on the host; a CUDA tensor is accepted as is
            coords = coords.long()
        feats = input[1].to(self.device) if self.device else input[1]
        out.features = _run(InputLayerFunction, self.dimension, out.metadata, self.spatial_size, coords.long(), feats,
                                                0 if len(input) == 2 else input[2], self.mode)
        if self.prefetch_ops and _PREFETCH:
            out.metadata.prefetch(self.prefetch_ops)
        return out


class InputLayerFunction(Function):
    @staticmethod
    def forward(ctx, dimension, metadata, spatial_size, coords, input_features, batch_size, mode):
        ctx.metadata_ = metadata
        out = input_features.new()
        native.InputLayer_updateOutput(metadata, spatial_size, coords, input_features.contiguous(), out, batch_size, mode)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        grad_input = grad_output.new()
        native.InputLayer_updateGradInput(ctx.metadata_, grad_input, grad_output.contiguous())
        return None, None, None, None, grad_input, None, None


class OutputLayer(Module):
    """Parameter-free inverse of InputLayer: SparseConvNetTensor -> float [N input rows, planes], every input row receives
    the feature row of its voxel.  reference: sparseconvnet/ioLayers.py:68-90,198-223."""

    def __init__(self, dimension):
        Module.__init__(self)
        self.dimension = dimension

    def forward(self, input):
        return _run(OutputLayerFunction, self.dimension, input.metadata, input.features)


class OutputLayerFunction(Function):
    @staticmethod
    def forward(ctx, dimension, metadata, input_features):
        out = input_features.new()
        ctx.metadata_ = metadata
        native.OutputLayer_updateOutput(metadata, input_features.contiguous(), out)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        grad_input = grad_output.new()
        native.OutputLayer_updateGradInput(ctx.metadata_, grad_input, grad_output.contiguous())
        return None, None, grad_input


class SparseToDense(Module):
    """SparseConvNetTensor -> dense [batch, nPlanes, X, Y, Z] (zeros at inactive sites).
    reference: sparseconvnet/sparseToDense.py:25-78; used by layers/roi_align_rotated_3d.py:81 and tools_3d_2d.py."""

    def __init__(self, dimension, nPlanes):
        Module.__init__(self)
        self.dimension = dimension
        self.nPlanes = nPlanes

    def forward(self, input):
        return _run(SparseToDenseFunction, input.features, input.metadata, input.spatial_size, self.dimension, self.nPlanes)

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        return 'SparseToDense(' + str(self.dimension) + ',' + str(self.nPlanes) + ')'


class SparseToDenseFunction(Function):
    @staticmethod
    def forward(ctx, input_features, input_metadata, spatial_size, dimension, nPlanes):
        ctx.input_metadata = input_metadata
        ctx.dimension = dimension
        ctx.save_for_backward(input_features, spatial_size)
        out = input_features.new()
        native.SparseToDense_updateOutput(spatial_size, input_metadata, input_features.contiguous(), out, nPlanes)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input_features, spatial_size = ctx.saved_tensors
        grad_input = grad_output.new()
        native.SparseToDense_updateGradInput(spatial_size, ctx.input_metadata, input_features, grad_input, grad_output.contiguous())
        return grad_input, None, None, None, None


def sparse_3d_to_dense_2d(feat_s3d):
    """reference: sparseconvnet/tools_3d_2d.py:7-48 -- densify and crop to the occupied extent [0:x_size, 0:y_size, 0:z_size]."""
    loc = feat_s3d.get_spatial_locations()
    x_size, y_size, z_size, _ = (loc.max(0)[0] + 1).tolist()
    dense = SparseToDense(dimension=4, nPlanes=feat_s3d.features.shape[1])(feat_s3d)
    return dense[:, :, 0:x_size, 0:y_size, 0:z_size]


# ------------------------------------------------------------------------------- convolutions
class SubmanifoldConvolution(Module):
    """reference: sparseconvnet/submanifoldConvolution.py:14-59"""

    def __init__(self, dimension, nIn, nOut, filter_size, bias, groups=1):
        Module.__init__(self)
        self.dimension, self.groups, self.nIn, self.nOut = dimension, groups, nIn, nOut
        self.filter_size = toLongTensor(dimension, filter_size)
        self.filter_volume = self.filter_size.prod().item()
        self.weight = _conv_weight(self.filter_volume, groups, nIn, nOut)
        if bias:
            self.bias = Parameter(torch.Tensor(nOut).zero_())

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn, (self.nIn, self.nOut, input)
        out = SparseConvNetTensor(metadata=input.metadata, spatial_size=input.spatial_size)
        out.features = _run(SubmanifoldConvolutionFunction, input.features, self.weight, optionalTensor(self, 'bias'), input.metadata,
                                                            input.spatial_size, self.dimension, self.filter_size)
        return out

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        f = self.filter_size.tolist()
        fs = str(f[0]) if min(f) == max(f) else '(' + ','.join(str(v) for v in f) + ')'
        return 'SubmanifoldConvolution %d->%d C%s' % (self.nIn, self.nOut, fs)


class ValidConvolution(SubmanifoldConvolution):
    pass


class SubmanifoldConvolutionFunction(Function):
    @staticmethod
    def forward(ctx, input_features, weight, bias, input_metadata, spatial_size, dimension, filter_size):
        ctx.input_metadata = input_metadata
        out = input_features.new()
        ctx.save_for_backward(input_features, spatial_size, weight, bias, filter_size)
        pkg = _counters()
        pkg.forward_pass_multiplyAdd_count += native.SubmanifoldConvolution_updateOutput(
            spatial_size, filter_size, input_metadata, input_features.contiguous(), out, weight, bias)
        pkg.forward_pass_hidden_states += out.nelement()
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input_features, spatial_size, weight, bias, filter_size = ctx.saved_tensors
        # the network's first convolution: its input (the InputLayer output of leaf features) needs no gradient
        grad_input = grad_output.new() if ctx.needs_input_grad[0] else None
        grad_weight = torch.zeros_like(weight)
        grad_bias = torch.zeros_like(bias)
        native.SubmanifoldConvolution_backward(spatial_size, filter_size, ctx.input_metadata, input_features.contiguous(), grad_input,
                                               grad_output.contiguous(), weight, grad_weight, grad_bias)
        return grad_input, grad_weight, optionalTensorReturn(grad_bias), None, None, None, None


def _strided_repr(name, m):
    f, s = m.filter_size.tolist(), m.filter_stride.tolist()
    if min(f) == max(f) and min(s) == max(s):
        tail = '%d/%d' % (f[0], s[0])
    else:
        tail = '(' + ','.join(map(str, f)) + ')/(' + ','.join(map(str, s)) + ')'
    return '%s %d->%d C%s' % (name, m.nIn, m.nOut, tail)


class Convolution(Module):
    """reference: sparseconvnet/convolution.py:13-70.  Output size = (in - filter) // stride + 1
    (the reference's `/` at :36 is integer division under the torch it was written for)."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        Module.__init__(self)
        self.dimension, self.groups, self.nIn, self.nOut = dimension, groups, nIn, nOut
        self.filter_size = toLongTensor(dimension, filter_size)
        self.filter_volume = self.filter_size.prod().item()
        self.filter_stride = toLongTensor(dimension, filter_stride)
        self.weight = _conv_weight(self.filter_volume, groups, nIn, nOut)
        if bias:
            self.bias = Parameter(torch.Tensor(nOut).zero_())

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn
        out = SparseConvNetTensor(metadata=input.metadata)
        out.spatial_size = (input.spatial_size - self.filter_size) // self.filter_stride + 1
        assert ((out.spatial_size - 1) * self.filter_stride + self.filter_size == input.spatial_size).all(), \
            (input.spatial_size, out.spatial_size, self.filter_size, self.filter_stride)
        out.features = _run(ConvolutionFunction, input.features, self.weight, optionalTensor(self, 'bias'), input.metadata, input.spatial_size,
                                                 out.spatial_size, self.dimension, self.filter_size, self.filter_stride)
        return out

    def input_spatial_size(self, out_size):
        return (out_size - 1) * self.filter_stride + self.filter_size

    def __repr__(self):
        return _strided_repr('Convolution', self)


class ConvolutionFunction(Function):
    @staticmethod
    def forward(ctx, input_features, weight, bias, input_metadata, input_spatial_size, output_spatial_size, dimension, filter_size, filter_stride):
        ctx.input_metadata = input_metadata
        out = input_features.new()
        ctx.save_for_backward(input_features, input_spatial_size, weight, bias, output_spatial_size, filter_size, filter_stride)
        pkg = _counters()
        pkg.forward_pass_multiplyAdd_count += native.Convolution_updateOutput(
            input_spatial_size, output_spatial_size, filter_size, filter_stride, input_metadata, input_features.contiguous(), out, weight, bias)
        pkg.forward_pass_hidden_states += out.nelement()
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input_features, in_size, weight, bias, out_size, filter_size, filter_stride = ctx.saved_tensors
        grad_input = grad_output.new()
        grad_weight = torch.zeros_like(weight)
        grad_bias = torch.zeros_like(bias)
        native.Convolution_backward(in_size, out_size, filter_size, filter_stride, ctx.input_metadata, input_features.contiguous(), grad_input,
                                    grad_output.contiguous(), weight, grad_weight, grad_bias)
        return grad_input, grad_weight, optionalTensorReturn(grad_bias), None, None, None, None, None, None


class Deconvolution(Module):
    """reference: sparseconvnet/deconvolution.py:13-90.  Output size = (in - 1) * stride + filter."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        Module.__init__(self)
        self.dimension, self.groups, self.nIn, self.nOut = dimension, groups, nIn, nOut
        self.filter_size = toLongTensor(dimension, filter_size)
        self.filter_volume = self.filter_size.prod().item()
        self.filter_stride = toLongTensor(dimension, filter_stride)
        self.weight = _conv_weight(self.filter_volume, groups, nIn, nOut)
        if bias:
            self.bias = Parameter(torch.Tensor(nOut).zero_())

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn
        out = SparseConvNetTensor(metadata=input.metadata)
        out.spatial_size = (input.spatial_size - 1) * self.filter_stride + self.filter_size
        out.features = _run(DeconvolutionFunction, input.features, self.weight, optionalTensor(self, 'bias'), input.metadata, input.spatial_size,
                                                   out.spatial_size, self.dimension, self.filter_size, self.filter_stride)
        return out

    def input_spatial_size(self, out_size):
        in_size = (out_size - self.filter_size) // self.filter_stride + 1
        assert ((in_size - 1) * self.filter_stride + self.filter_size == out_size).all()
        return in_size

    def __repr__(self):
        return _strided_repr('Deconvolution', self)


class DeconvolutionFunction(Function):
    @staticmethod
    def forward(ctx, input_features, weight, bias, input_metadata, input_spatial_size, output_spatial_size, dimension, filter_size, filter_stride):
        ctx.input_metadata = input_metadata
        out = input_features.new()
        pkg = _counters()
        pkg.forward_pass_multiplyAdd_count += native.Deconvolution_updateOutput(
            input_spatial_size, output_spatial_size, filter_size, filter_stride, input_metadata, input_features.contiguous(), out, weight, bias)
        pkg.forward_pass_hidden_states += out.nelement()
        ctx.save_for_backward(input_features, input_spatial_size, weight, bias, output_spatial_size, filter_size, filter_stride)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input_features, in_size, weight, bias, out_size, filter_size, filter_stride = ctx.saved_tensors
        grad_input = grad_output.new()
        grad_weight = torch.zeros_like(weight)
        grad_bias = torch.zeros_like(bias)
        native.Deconvolution_backward(in_size, out_size, filter_size, filter_stride, ctx.input_metadata, input_features.contiguous(), grad_input,
                                      grad_output.contiguous(), weight, grad_weight, grad_bias)
        return grad_input, grad_weight, optionalTensorReturn(grad_bias), None, None, None, None, None, None


# ------------------------------------------------------------------------------- batch norm
class BatchNormalization(Module):
    """BatchNorm + in-place leaky ReLU (leakiness 0 = ReLU, 1 = none).
    reference: sparseconvnet/batchNormalization.py:13-78.  In eval mode with
    track_running_stats=False the statistics of the current input are used (mean, UNBIASED
    variance) -- the reference computes them with torch ops (:55-56); here the kernel does."""

    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1, track_running_stats=True):
        Module.__init__(self)
        self.nPlanes, self.eps, self.momentum, self.affine, self.leakiness = nPlanes, eps, momentum, affine, leakiness
        self.register_buffer("running_mean", torch.Tensor(nPlanes).fill_(0))
        self.register_buffer("running_var", torch.Tensor(nPlanes).fill_(1))
        if affine:
            self.weight = Parameter(torch.Tensor(nPlanes).fill_(1))
            self.bias = Parameter(torch.Tensor(nPlanes).fill_(0))
        self.track_running_stats = track_running_stats

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nPlanes, (self.nPlanes, input.features.shape)
        out = SparseConvNetTensor(metadata=input.metadata, spatial_size=input.spatial_size)
        instance = not (self.training or self.track_running_stats)
        out.features = _run(BatchNormalizationFunction, input.features, optionalTensor(self, 'weight'), optionalTensor(self, 'bias'),
                                                        self.running_mean, self.running_var, self.eps, self.momentum, self.training,
                                                        self.leakiness, instance)
        return out

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        s = 'BatchNorm(%d,eps=%s,momentum=%s,affine=%s' % (self.nPlanes, self.eps, self.momentum, self.affine)
        if self.leakiness > 0:
            s += ',leakiness=' + str(self.leakiness)
        return s + ')'


class BatchNormReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, track_running_stats=True):
        BatchNormalization.__init__(self, nPlanes, eps, momentum, True, 0, track_running_stats)

    def __repr__(self):
        return 'BatchNormReLU(%d,eps=%s,momentum=%s,affine=%s)' % (self.nPlanes, self.eps, self.momentum, self.affine)


class BatchNormLeakyReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, leakiness=0.333, track_running_stats=True):
        BatchNormalization.__init__(self, nPlanes, eps, momentum, True, leakiness, track_running_stats)

    def __repr__(self):
        return 'BatchNormLeakyReLU(%d,eps=%s,momentum=%s,affine=%s,leakiness=%s)' % (self.nPlanes, self.eps, self.momentum, self.affine, self.leakiness)


class BatchNormalizationFunction(Function):
    @staticmethod
    def forward(ctx, input_features, weight, bias, running_mean, running_var, eps, momentum, train, leakiness, instance_stats=False):
        ctx.train, ctx.leakiness = train, leakiness
        out = input_features.new()
        save_mean, save_invstd = input_features.new(), input_features.new()
        x = input_features.contiguous()
        native.BatchNormalization_updateOutput(x, out, save_mean, save_invstd, running_mean, running_var, weight, bias, eps, momentum, train,
                                               leakiness, instance_stats)
        ctx.save_for_backward(x, out, weight, bias, running_mean, running_var, save_mean, save_invstd)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        x, out, weight, bias, running_mean, running_var, save_mean, save_invstd = ctx.saved_tensors
        assert ctx.train
        grad_input = grad_output.new()
        grad_weight = torch.zeros_like(weight)
        grad_bias = torch.zeros_like(bias)
        # the reference rewrites grad_output in place (CPU/BatchNormalization.cpp:79-82); work on a private copy
        g = grad_output.contiguous().clone()
        native.BatchNormalization_backward(x, grad_input, out, g, save_mean, save_invstd, running_mean, running_var, weight, bias,
                                           grad_weight, grad_bias, ctx.leakiness)
        return grad_input, optionalTensorReturn(grad_weight), optionalTensorReturn(grad_bias), None, None, None, None, None, None, None


class NetworkInNetworkFunction(Function):
    """reference: sparseconvnet/networkInNetwork.py:14-56"""

    @staticmethod
    def forward(ctx, input_features, weight, bias):
        out = input_features.new()
        x = input_features.contiguous()
        ctx.save_for_backward(x, weight, bias)
        pkg = _counters()
        pkg.forward_pass_multiplyAdd_count += native.NetworkInNetwork_updateOutput(x, out, weight, bias)
        pkg.forward_pass_hidden_states += out.nelement()
        return out

    @staticmethod
    def backward(ctx, grad_output):
        x, weight, bias = ctx.saved_tensors
        g = grad_output.contiguous()
        grad_input = grad_output.new()
        grad_weight = torch.zeros_like(weight)
        grad_bias = torch.zeros_like(bias)
        native.NetworkInNetwork_updateGradInput(grad_input, g, weight)
        native.NetworkInNetwork_accGradParameters(x, g, grad_weight, grad_bias)
        return grad_input, grad_weight, optionalTensorReturn(grad_bias)


class NetworkInNetwork(Module):
    """1x1 'convolution' on feature rows (reference: sparseconvnet/networkInNetwork.py:59-92; SCN/CPU/NetworkInNetwork.cpp:7-46);
    built by FPN_Net when a residual block changes width (fpn_net.py:63)."""

    def __init__(self, nIn, nOut, bias):
        Module.__init__(self)
        self.nIn, self.nOut = nIn, nOut
        std = (2.0 / nIn) ** 0.5
        self.weight = Parameter(torch.Tensor(nIn, nOut).normal_(0, std))
        if bias:
            self.bias = Parameter(torch.Tensor(nOut).zero_())

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn, (self.nIn, input.features.shape)
        out = SparseConvNetTensor(metadata=input.metadata, spatial_size=input.spatial_size)
        out.features = _run(NetworkInNetworkFunction, input.features, self.weight, optionalTensor(self, 'bias'))
        return out

    def __repr__(self):
        return 'NetworkInNetwork' + str(self.nIn) + '->' + str(self.nOut)

    def input_spatial_size(self, out_size):
        return out_size
