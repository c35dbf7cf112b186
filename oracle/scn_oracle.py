"""ORACLE / TEST INFRASTRUCTURE ONLY.

ctypes front-ends for the two CPU checkers:

* ``OracleMetadata`` + ``o_*`` functions  -> oracle/_build/liboracle.so, our plain-C
  restatement (oracle/scn_oracle.c).
* ``RefMetadata``                         -> oracle/_ref/libscn_ref_rules.so, the
  reference's own ``Metadata<3>`` compiled unmodified from /root/reference
  (oracle/ref_rules.cpp is the C glue).

Both expose the same numpy interface so the parity tests can run either as the
checker.  Nothing under detection_3d_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
_REF_RULES_SO = os.path.join(HERE, "_ref", "libscn_ref_rules.so")
_REF_SCN_SO = os.path.join(HERE, "_ref", "SCN.so")

_L3 = C.c_long * 3


def _l3(v):
    v = [int(x) for x in (v.tolist() if hasattr(v, "tolist") else v)]
    assert len(v) == 3
    return _L3(*v)


def build_oracle(force=False):
    """Compile oracle/scn_oracle.c -> oracle/_build/liboracle.so (gcc, seconds)."""
    src = os.path.join(HERE, "scn_oracle.c")
    if force or not os.path.exists(_ORACLE_SO) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return _ORACLE_SO


def build_ref():
    """Compile the reference's CPU extension from /root/reference (build container only)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def have_ref():
    return os.path.exists(_REF_RULES_SO) and os.path.exists(_REF_SCN_SO)


_olib = None


def olib():
    global _olib
    if _olib is None:
        lib = C.CDLL(build_oracle())
        lib.oracle_point_hash.restype = C.c_uint64
        lib.oracle_point_hash.argtypes = [C.c_int32] * 3
        lib.omd_create.restype = C.c_void_p
        lib.omd_destroy.argtypes = [C.c_void_p]
        lib.omd_nactive.restype = C.c_long
        lib.omd_nactive.argtypes = [C.c_void_p, _L3]
        lib.omd_batch_size.restype = C.c_long
        lib.omd_batch_size.argtypes = [C.c_void_p, _L3]
        lib.omd_input_layer.restype = C.c_long
        lib.omd_input_layer.argtypes = [C.c_void_p, _L3, C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_long]
        lib.omd_spatial_locations.argtypes = [C.c_void_p, _L3, C.c_void_p]
        lib.omd_iteration_order.restype = C.c_long
        lib.omd_iteration_order.argtypes = [C.c_void_p, _L3, C.c_long, C.c_void_p]
        lib.omd_submanifold_rules.restype = C.c_void_p
        lib.omd_submanifold_rules.argtypes = [C.c_void_p, _L3, _L3]
        lib.omd_conv_rules.restype = C.c_void_p
        lib.omd_conv_rules.argtypes = [C.c_void_p, _L3, _L3, _L3, _L3]
        lib.omd_sparse_to_dense_rules.restype = C.c_void_p
        lib.omd_sparse_to_dense_rules.argtypes = [C.c_void_p, _L3]
        lib.o_sparse_to_dense_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_long]
        lib.o_sparse_to_dense_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_long]
        lib.o_roi_align_rotated_3d.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_long, C.c_void_p, C.c_long, C.c_float, C.c_long, C.c_long,
                                               C.c_long, C.c_long, C.c_void_p, C.c_int]
        lib.omd_input_rules.restype = C.c_void_p
        lib.omd_input_rules.argtypes = [C.c_void_p]
        lib.orb_nlists.restype = C.c_long
        lib.orb_nlists.argtypes = [C.c_void_p]
        lib.orb_list_size.restype = C.c_long
        lib.orb_list_size.argtypes = [C.c_void_p, C.c_long]
        lib.orb_list_copy.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        vp, l, f, i = C.c_void_p, C.c_long, C.c_float, C.c_int
        lib.o_input_layer_forward.argtypes = [vp, vp, l, l, l, vp, i]
        lib.o_input_layer_backward.argtypes = [vp, vp, l, l, l, l, vp, i]
        lib.o_conv_list_forward.argtypes = [vp, vp, vp, l, l, vp, l, i, i]
        lib.o_conv_list_backward.argtypes = [vp, vp, vp, vp, vp, l, l, vp, l, i, i]
        lib.o_bn_forward.argtypes = [vp, vp, l, l, vp, vp, vp, vp, vp, vp, f, f, i, f]
        lib.o_bn_backward.argtypes = [vp, vp, vp, vp, l, l, vp, vp, vp, vp, vp, f]
        _olib = lib
    return _olib


_rlib = None


def rlib():
    global _rlib
    if _rlib is None:
        import torch  # noqa: F401  (libtorch must be loaded first)

        lib = C.CDLL(_REF_RULES_SO)
        lib.ref_md_create.restype = C.c_void_p
        lib.ref_md_destroy.argtypes = [C.c_void_p]
        lib.ref_md_input_layer.restype = C.c_long
        lib.ref_md_input_layer.argtypes = [C.c_void_p, _L3, C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_long]
        lib.ref_md_nactive.restype = C.c_long
        lib.ref_md_nactive.argtypes = [C.c_void_p, _L3]
        lib.ref_md_batch_size.restype = C.c_long
        lib.ref_md_batch_size.argtypes = [C.c_void_p, _L3]
        lib.ref_md_spatial_locations.argtypes = [C.c_void_p, _L3, C.c_void_p]
        lib.ref_md_iteration_order.restype = C.c_long
        lib.ref_md_iteration_order.argtypes = [C.c_void_p, _L3, C.c_long, C.c_void_p]
        lib.ref_md_rulebook.restype = C.c_void_p
        lib.ref_md_rulebook.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.ref_rb_nlists.restype = C.c_long
        lib.ref_rb_nlists.argtypes = [C.c_void_p]
        lib.ref_rb_list_size.restype = C.c_long
        lib.ref_rb_list_size.argtypes = [C.c_void_p, C.c_long]
        lib.ref_rb_list_copy.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        lib.ref_point_hash.restype = C.c_ulong
        lib.ref_point_hash.argtypes = [C.c_int] * 3
        _rlib = lib
    return _rlib


def _lists(n, size, copy, rb):
    out = []
    for i in range(n(rb)):
        a = np.empty(size(rb, i), dtype=np.int32)
        if a.size:
            copy(rb, i, a.ctypes.data)
        out.append(a)
    return out


def _pairs(lists):
    return [a.reshape(-1, 2) for a in lists]


class OracleMetadata:
    """Our C restatement of Metadata<3> (rulebook side)."""

    kind = "port"

    def __init__(self):
        self.lib = olib()
        self.h = C.c_void_p(self.lib.omd_create())

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.omd_destroy(self.h)
            self.h = None

    def input_layer(self, spatial, coords, batch_size=0, mode=4):
        coords = np.ascontiguousarray(coords, dtype=np.int64)
        self._keep = coords
        return self.lib.omd_input_layer(self.h, _l3(spatial), coords.ctypes.data, coords.shape[0], coords.shape[1], batch_size, mode)

    def input_rules(self):
        ls = _lists(self.lib.orb_nlists, self.lib.orb_list_size, self.lib.orb_list_copy, C.c_void_p(self.lib.omd_input_rules(self.h)))
        return ls

    def nactive(self, spatial):
        return self.lib.omd_nactive(self.h, _l3(spatial))

    def batch_size(self, spatial):
        return self.lib.omd_batch_size(self.h, _l3(spatial))

    def spatial_locations(self, spatial):
        out = np.zeros((self.nactive(spatial), 4), dtype=np.int64)
        if out.size:
            self.lib.omd_spatial_locations(self.h, _l3(spatial), out.ctypes.data)
        return out

    def iteration_order(self, spatial, b=0):
        n = self.lib.omd_iteration_order(self.h, _l3(spatial), b, None)
        out = np.empty(n, dtype=np.int32)
        if n:
            self.lib.omd_iteration_order(self.h, _l3(spatial), b, out.ctypes.data)
        return out

    def submanifold_rules(self, spatial, filt):
        rb = C.c_void_p(self.lib.omd_submanifold_rules(self.h, _l3(spatial), _l3(filt)))
        return _pairs(_lists(self.lib.orb_nlists, self.lib.orb_list_size, self.lib.orb_list_copy, rb))

    def conv_rules(self, in_size, out_size, filt, stride):
        rb = C.c_void_p(self.lib.omd_conv_rules(self.h, _l3(in_size), _l3(out_size), _l3(filt), _l3(stride)))
        return _pairs(_lists(self.lib.orb_nlists, self.lib.orb_list_size, self.lib.orb_list_copy, rb))


    def sparse_to_dense_rules(self, spatial):
        rb = C.c_void_p(self.lib.omd_sparse_to_dense_rules(self.h, _l3(spatial)))
        return _pairs(_lists(self.lib.orb_nlists, self.lib.orb_list_size, self.lib.orb_list_copy, rb))


class RefMetadata:
    """The reference's own Metadata<3> (compiled from /root/reference, oracle/_ref)."""

    kind = "reference"

    def __init__(self):
        self.lib = rlib()
        self.h = C.c_void_p(self.lib.ref_md_create())

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_md_destroy(self.h)
            self.h = None

    def input_layer(self, spatial, coords, batch_size=0, mode=4):
        coords = np.ascontiguousarray(coords, dtype=np.int64)
        return self.lib.ref_md_input_layer(self.h, _l3(spatial), coords.ctypes.data, coords.shape[0], coords.shape[1], batch_size, mode)

    def _rb(self, kind, a=None, b=None, c=None, d=None):
        z = _L3(0, 0, 0)
        args = [x if x is not None else z for x in (a, b, c, d)]
        return C.c_void_p(self.lib.ref_md_rulebook(self.h, kind, *[C.addressof(x) for x in args], 1))

    def input_rules(self):
        return _lists(self.lib.ref_rb_nlists, self.lib.ref_rb_list_size, self.lib.ref_rb_list_copy, self._rb(0))

    def nactive(self, spatial):
        return self.lib.ref_md_nactive(self.h, _l3(spatial))

    def batch_size(self, spatial):
        return self.lib.ref_md_batch_size(self.h, _l3(spatial))

    def spatial_locations(self, spatial):
        out = np.zeros((self.nactive(spatial), 4), dtype=np.int64)
        if out.size:
            self.lib.ref_md_spatial_locations(self.h, _l3(spatial), out.ctypes.data)
        return out

    def iteration_order(self, spatial, b=0):
        n = self.lib.ref_md_iteration_order(self.h, _l3(spatial), b, None)
        out = np.empty(n, dtype=np.int32)
        if n:
            self.lib.ref_md_iteration_order(self.h, _l3(spatial), b, out.ctypes.data)
        return out

    def submanifold_rules(self, spatial, filt):
        rb = self._rb(1, _l3(spatial), _l3(filt))
        return _pairs(_lists(self.lib.ref_rb_nlists, self.lib.ref_rb_list_size, self.lib.ref_rb_list_copy, rb))

    def conv_rules(self, in_size, out_size, filt, stride):
        rb = self._rb(2, _l3(in_size), _l3(out_size), _l3(filt), _l3(stride))
        return _pairs(_lists(self.lib.ref_rb_nlists, self.lib.ref_rb_list_size, self.lib.ref_rb_list_copy, rb))


def _ref_s2d(self, spatial):
    rb = self._rb(3, _l3(spatial))
    return _pairs(_lists(self.lib.ref_rb_nlists, self.lib.ref_rb_list_size, self.lib.ref_rb_list_copy, rb))


RefMetadata.sparse_to_dense_rules = _ref_s2d


def o_sparse_to_dense_forward(feats, rules, spatial, n_planes=None):
    """cpu_SparseToDense_updateOutput (SCN/CPU/SparseToDense.cpp:34-66): [batch, nPlanes, X, Y, Z], zero-filled, one rule list
    (row, offset) per batch item."""
    feats = _f32(feats)
    c = feats.shape[1] if n_planes is None else n_planes
    vol = int(np.prod(spatial))
    out = np.zeros((max(1, len(rules)), c) + tuple(int(v) for v in spatial), dtype=np.float32)
    for b, r in enumerate(rules):
        r = np.ascontiguousarray(r, dtype=np.int32)
        if r.shape[0]:
            olib().o_sparse_to_dense_forward(feats.ctypes.data, out[b].ctypes.data, feats.shape[1], vol, r.ctypes.data, r.shape[0])
    return out


def o_sparse_to_dense_backward(d_out, rules, n_rows):
    """cpu_SparseToDense_updateGradInput (SCN/CPU/SparseToDense.cpp:67-101)."""
    d_out = _f32(d_out)
    c = d_out.shape[1]
    vol = int(np.prod(d_out.shape[2:]))
    d_in = np.zeros((n_rows, c), dtype=np.float32)
    for b, r in enumerate(rules):
        r = np.ascontiguousarray(r, dtype=np.int32)
        if r.shape[0]:
            olib().o_sparse_to_dense_backward(d_in.ctypes.data, d_out[b].ctypes.data, c, vol, r.ctypes.data, r.shape[0])
    return d_in


def o_roi_align_rotated_3d_forward(dense, rois, scale, pooled, sampling):
    """RoIAlignRotated3DForward (maskrcnn_benchmark/csrc/cuda/ROIAlignRotated3D_cuda.cu:89-172) on a dense [B, C, H, W, Z] input;
    rois [n, 8] = (batch, center_w, center_h, center_z, w, h, z, theta in degrees).  -> [n, C, ph, pw, pz]."""
    dense, rois = _f32(dense), _f32(rois)
    _, c, h, w, z = dense.shape
    out = np.zeros((rois.shape[0], c) + tuple(pooled), dtype=np.float32)
    olib().o_roi_align_rotated_3d(dense.ctypes.data, None, c, h, w, z, rois.ctypes.data, rois.shape[0], scale, pooled[0], pooled[1], pooled[2], sampling,
                                  out.ctypes.data, 0)
    return out


def o_roi_align_rotated_3d_backward(grad, rois, scale, pooled, sampling, dense_shape):
    """RoIAlignRotated3DBackwardFeature (:235-346) -> d_input [B, C, H, W, Z]."""
    grad, rois = _f32(grad), _f32(rois)
    b, c, h, w, z = dense_shape
    d = np.zeros(dense_shape, dtype=np.float32)
    olib().o_roi_align_rotated_3d(None, d.ctypes.data, c, h, w, z, rois.ctypes.data, rois.shape[0], scale, pooled[0], pooled[1], pooled[2], sampling,
                                  grad.ctypes.data, 1)
    return d


_roi_ref = None


def roi_align_ref_lib():
    """The reference's own ROIAlignRotated3D CUDA kernels (oracle/_ref/libroialign3d_ref.so, built by `make -C oracle ref`); GPU box only."""
    global _roi_ref
    if _roi_ref is None:
        path = os.path.join(HERE, "_ref", "libroialign3d_ref.so")
        if not os.path.exists(path):
            return None
        import torch  # noqa: F401  (libtorch must be loaded first)
        lib = C.CDLL(path)
        vp, l, f, i = C.c_void_p, C.c_long, C.c_float, C.c_int
        lib.ref_roi_align_rotated_3d_forward.argtypes = [vp, l, l, l, l, l, vp, l, f, i, i, i, i, vp]
        lib.ref_roi_align_rotated_3d_backward.argtypes = [vp, vp, l, f, i, i, i, l, l, l, l, l, i, vp]
        _roi_ref = lib
    return _roi_ref


def point_hash(x, y, z):
    return olib().oracle_point_hash(x, y, z)


# ------------------------------------------------------------------ compute side
def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def o_input_layer_forward(feats, rules_hdr, rules_tab):
    """feats [nIn,C]; returns [nOut,C] (mode 3 = sum, mode 4 = mean)."""
    feats = _f32(feats)
    mode, max_active, _, n_out = [int(x) for x in rules_hdr]
    out = np.zeros((n_out, feats.shape[1]), dtype=np.float32)
    tab = np.ascontiguousarray(rules_tab, dtype=np.int32)
    olib().o_input_layer_forward(feats.ctypes.data, out.ctypes.data, n_out, max_active, feats.shape[1], tab.ctypes.data, int(mode == 4))
    return out


def o_input_layer_backward(d_out, rules_hdr, rules_tab):
    d_out = _f32(d_out)
    mode, max_active, n_in, n_out = [int(x) for x in rules_hdr]
    d_in = np.zeros((n_in, d_out.shape[1]), dtype=np.float32)
    tab = np.ascontiguousarray(rules_tab, dtype=np.int32)
    olib().o_input_layer_backward(d_in.ctypes.data, d_out.ctypes.data, n_in, n_out, max_active, d_out.shape[1], tab.ctypes.data, int(mode == 4))
    return d_in


def o_conv_forward(feats, weight, rules, n_out_rows, deconv=False, bias=None):
    """sum_k scatter(gather(feats, rules[k]) @ weight[k]); weight [K,1,Cin,Cout] or [K,Cin,Cout]."""
    feats = _f32(feats)
    w = _f32(weight).reshape(len(rules), feats.shape[1], -1)
    n_out = w.shape[2]
    out = np.zeros((n_out_rows, n_out), dtype=np.float32)
    if bias is not None and n_out_rows:
        out[:] = _f32(bias)[None, :]
    src, dst = (1, 0) if deconv else (0, 1)
    macs = 0.0
    for k, r in enumerate(rules):
        r = np.ascontiguousarray(r, dtype=np.int32)
        if r.shape[0]:
            macs += r.shape[0] * feats.shape[1] * n_out
            olib().o_conv_list_forward(feats.ctypes.data, out.ctypes.data, w[k].ctypes.data, feats.shape[1], n_out, r.ctypes.data, r.shape[0], src, dst)
    return out, macs


def o_conv_backward(feats, d_out, weight, rules, deconv=False):
    feats, d_out = _f32(feats), _f32(d_out)
    w = _f32(weight).reshape(len(rules), feats.shape[1], -1)
    d_in = np.zeros_like(feats)
    dw = np.zeros_like(w)
    src, dst = (1, 0) if deconv else (0, 1)
    for k, r in enumerate(rules):
        r = np.ascontiguousarray(r, dtype=np.int32)
        if r.shape[0]:
            olib().o_conv_list_backward(feats.ctypes.data, d_in.ctypes.data, d_out.ctypes.data, w[k].ctypes.data, dw[k].ctypes.data,
                                        feats.shape[1], w.shape[2], r.ctypes.data, r.shape[0], src, dst)
    return d_in, dw.reshape(np.shape(weight))


def o_bn_forward(feats, weight, bias, running_mean, running_var, eps=1e-4, momentum=0.9, train=False, leakiness=0.0):
    """Returns (out, saveMean, saveInvStd); running stats updated in place when train."""
    feats = _f32(feats)
    n, c = feats.shape
    out = np.empty_like(feats)
    sm, si = np.zeros(c, np.float32), np.zeros(c, np.float32)
    assert running_mean.dtype == np.float32 and running_var.dtype == np.float32
    wp = _f32(weight).ctypes.data if weight is not None else None
    bp = _f32(bias).ctypes.data if bias is not None else None
    olib().o_bn_forward(feats.ctypes.data, out.ctypes.data, c, n, sm.ctypes.data, si.ctypes.data, running_mean.ctypes.data,
                        running_var.ctypes.data, wp, bp, eps, momentum, int(train), leakiness)
    return out, sm, si


def o_bn_backward(feats, out, d_out, save_mean, save_invstd, weight, leakiness=0.0):
    feats, out = _f32(feats), _f32(out)
    d_out = _f32(d_out).copy()
    n, c = feats.shape
    d_in = np.empty_like(feats)
    dw, db = np.zeros(c, np.float32), np.zeros(c, np.float32)
    olib().o_bn_backward(feats.ctypes.data, d_in.ctypes.data, out.ctypes.data, d_out.ctypes.data, c, n, _f32(save_mean).ctypes.data,
                         _f32(save_invstd).ctypes.data, _f32(weight).ctypes.data, dw.ctypes.data, db.ctypes.data, leakiness)
    return d_in, dw, db


def o_output_layer_forward(feats, rules_hdr, rules_tab):
    """cpu_OutputLayer_updateOutput (SCN/CPU/IOLayers.cpp:97-118): InputLayer_BackwardPass with average = false --
    every input row receives its voxel's feature row; rows dropped by modes 1 / 2 stay zero.  mode 0: copy."""
    feats = _f32(feats)
    mode, max_active, n_in, n_out = [int(x) for x in rules_hdr]
    if mode == 0:
        return feats.copy()
    out = np.zeros((n_in, feats.shape[1]), dtype=np.float32)
    tab = np.ascontiguousarray(rules_tab, dtype=np.int32)
    olib().o_input_layer_backward(out.ctypes.data, feats.ctypes.data, n_in, n_out, max_active, feats.shape[1], tab.ctypes.data, 0)
    return out


def o_output_layer_backward(d_out, rules_hdr, rules_tab):
    """cpu_OutputLayer_updateGradInput (SCN/CPU/IOLayers.cpp:119-140): InputLayer_ForwardPass with average = false."""
    d_out = _f32(d_out)
    mode, max_active, _, n_out = [int(x) for x in rules_hdr]
    if mode == 0:
        return d_out.copy()
    d_in = np.zeros((n_out, d_out.shape[1]), dtype=np.float32)
    tab = np.ascontiguousarray(rules_tab, dtype=np.int32)
    olib().o_input_layer_forward(d_out.ctypes.data, d_in.ctypes.data, n_out, max_active, d_out.shape[1], tab.ctypes.data, 0)
    return d_in


def o_nin_forward(feats, weight, bias=None):
    """cpu_NetworkInNetwork_updateOutput (SCN/CPU/NetworkInNetwork.cpp:7-24): out = bias + feats @ weight; returns (out, macs)."""
    feats, weight = _f32(feats), _f32(weight)
    out = feats @ weight
    if bias is not None:
        out = out + _f32(bias)[None, :]
    return out.astype(np.float32), float(feats.shape[0]) * weight.shape[0] * weight.shape[1]


def o_nin_backward(feats, d_out, weight):
    """cpu_NetworkInNetwork_updateGradInput / _accGradParameters (SCN/CPU/NetworkInNetwork.cpp:25-46):
    d_in = d_out @ W^T, dW = feats^T @ d_out, d_bias = column sums of d_out."""
    feats, d_out, weight = _f32(feats), _f32(d_out), _f32(weight)
    return d_out @ weight.T, feats.T @ d_out, d_out.sum(0)
