"""Drop-in `sparseconvnet` API for the Detection_3D backbone, backed by hand-written sm_100a kernels.

    import detection_3d_b200.sparseconvnet as scn          # or detection_3d_b200.install_as_sparseconvnet()

exports the names Detection_3D uses from the reference package (SparseConvNet/sparseconvnet/__init__.py:13-42):
InputLayer, SubmanifoldConvolution, Convolution, Deconvolution, BatchNormalization / BatchNormReLU /
BatchNormLeakyReLU, SparseConvNetTensor, Metadata, Sequential, ConcatTable, AddTable, JoinTable, Identity,
NetworkInNetwork, OutputLayer, SparseToDense (+ tools_3d_2d.sparse_3d_to_dense_2d), add_feature_planes,
concatenate_feature_planes, FPN_Net and the two global counters.
"""
forward_pass_multiplyAdd_count = 0
forward_pass_hidden_states = 0

from . import native as SCN  # noqa: E402  (the reference exposes its extension as sparseconvnet.SCN)
from .containers import (AddTable, ConcatTable, Identity, JoinTable, Sequential, add_feature_planes,  # noqa: E402
                         concatenate_feature_planes)
from .fpn import FPN_Net, OutputLayer, c6_fpn4321_config, sw4c_fpn432_config  # noqa: E402
from .layers import (BatchNormalization, BatchNormLeakyReLU, BatchNormReLU, Convolution, Deconvolution, InputLayer,  # noqa: E402
                     Metadata, NetworkInNetwork, SparseToDense, SubmanifoldConvolution, ValidConvolution, optionalTensor,
                     optionalTensorReturn, sparse_3d_to_dense_2d, toLongTensor)
from . import layers as sparseToDense  # noqa: E402  (the reference reaches the class as scn.sparseToDense.SparseToDense, tools_3d_2d.py:26)
from .native import empty_cache, kernel_launch_count, set_math_mode  # noqa: E402
from .tensor import SparseConvNetTensor  # noqa: E402
