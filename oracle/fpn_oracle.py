"""ORACLE / TEST INFRASTRUCTURE ONLY.

CPU forward of the Detection_3D sparse FPN backbone, layer by layer, for two interchangeable
checkers:

 * ``run_fpn_port``  -- our plain-C restatement (oracle/scn_oracle.c) driven from numpy;
 * ``run_fpn_ref``   -- the reference's own compiled CPU extension (oracle/_ref/SCN.so: the
   unmodified SCN/pybind.cpp + sparseconvnet_cpu.cpp) called through its pybind functions.  The
   reference's *Python* layer files cannot travel to the GPU box, so the layer sequence of
   sparseconvnet/fpn_net.py:140-265 is restated here; tests/test_oracle_cpu.py checks in the
   build container that this driver reproduces the real `scn.FPN_Net.forward` bit for bit.

Both take the reference's state_dict (same key names) as numpy / torch tensors.
"""
import numpy as np

from . import scn_oracle as so


def _sizes(full_scale, n):
    return [[int(s) // (2 ** l) for s in full_scale] for l in range(n)]


class _PortBackend:
    """Layer primitives on numpy arrays, OracleMetadata underneath."""

    kind = "port"

    def __init__(self, leakiness=0.0, eps=1e-4):
        self.md = so.OracleMetadata()
        self.macs = 0.0
        self.leak, self.eps = leakiness, eps

    def input(self, full, coords, feats, mode=4):
        self.md.input_layer(full, coords, 0, mode)
        hdr, tab = self.md.input_rules()
        return so.o_input_layer_forward(feats, hdr, tab)

    def subm(self, x, w, sz, f):
        rules = self.md.submanifold_rules(sz, [f, f, f])
        out, m = so.o_conv_forward(x, w, rules, self.md.nactive(sz))
        self.macs += m
        return out

    def conv(self, x, w, in_sz, out_sz, f, s):
        rules = self.md.conv_rules(in_sz, out_sz, f, s)
        out, m = so.o_conv_forward(x, w, rules, self.md.nactive(out_sz))
        self.macs += m
        return out

    def deconv(self, x, w, in_sz, out_sz, f, s):
        rules = self.md.conv_rules(out_sz, in_sz, f, s)  # CPU/Deconvolution.cpp:15-16
        out, m = so.o_conv_forward(x, w, rules, self.md.nactive(out_sz), deconv=True)
        self.macs += m
        return out

    def bn_eval_instance(self, x, gamma, beta):
        # batchNormalization.py:51-56: eval + track_running_stats=False -> mean(0), unbiased var(0)
        x = np.asarray(x, np.float32)
        mean = x.mean(0, dtype=np.float64).astype(np.float32)
        var = x.var(0, ddof=1, dtype=np.float64).astype(np.float32) if x.shape[0] > 1 else np.full(x.shape[1], np.nan, np.float32)
        out, _, _ = so.o_bn_forward(x, gamma, beta, mean.copy(), var.copy(), self.eps, 0.9, False, self.leak)
        return out

    def locations(self, sz):
        return self.md.spatial_locations(sz)


class _RefBackend:
    """Same primitives on torch CPU tensors through the reference's compiled extension."""

    kind = "reference"

    def __init__(self, leakiness=0.0, eps=1e-4):
        import torch
        from . import ref_python
        self.t = torch
        self.SCN = ref_python.load_scn_native()
        self.md = self.SCN.Metadata_3()
        self.macs = 0.0
        self.leak, self.eps = leakiness, eps

    def _L(self, v):
        return self.t.LongTensor([int(a) for a in v])

    def _T(self, a):
        return a if isinstance(a, self.t.Tensor) else self.t.from_numpy(np.ascontiguousarray(a))

    def input(self, full, coords, feats, mode=4):
        out = self.t.empty(0)
        self.SCN.InputLayer_updateOutput(self.md, self._L(full), self._T(coords).long(), self._T(feats).float().contiguous(), out, 0, mode)
        return out

    def subm(self, x, w, sz, f):
        out = self.t.empty(0)
        self.macs += self.SCN.SubmanifoldConvolution_updateOutput(self._L(sz), self._L([f, f, f]), self.md, x, out, self._T(w), self.t.Tensor())
        return out

    def conv(self, x, w, in_sz, out_sz, f, s):
        out = self.t.empty(0)
        self.macs += self.SCN.Convolution_updateOutput(self._L(in_sz), self._L(out_sz), self._L(f), self._L(s), self.md, x, out, self._T(w), self.t.Tensor())
        return out

    def deconv(self, x, w, in_sz, out_sz, f, s):
        out = self.t.empty(0)
        self.macs += self.SCN.Deconvolution_updateOutput(self._L(in_sz), self._L(out_sz), self._L(f), self._L(s), self.md, x, out, self._T(w), self.t.Tensor())
        return out

    def bn_eval_instance(self, x, gamma, beta):
        t = self.t
        mean, var = x.mean(0), x.var(0)  # batchNormalization.py:55-56
        out, sm, si = t.empty(0), t.empty(x.size(1)), t.empty(x.size(1))
        self.SCN.BatchNormalization_updateOutput(x, out, sm, si, mean, var, self._T(gamma), self._T(beta), self.eps, 0.9, False, self.leak)
        return out

    def locations(self, sz):
        return self.md.getSpatialLocations(self._L(sz)).numpy()


def _run(be, cfg, state, coords, feats, taps=None):
    """Layer sequence of FPN_Net.forward / forward_fpn (fpn_net.py:140-203) with reps == 1 and
    residual blocks.  `taps`, if a dict, receives every intermediate feature matrix by name."""
    assert cfg['reps'] == 1 and cfg['residual_blocks']
    P, n = cfg['nPlanesF'], len(cfg['nPlanesF'])
    sizes = _sizes(cfg['full_scale'], n)
    K, S = cfg['downsample']
    g = lambda k: state[k]
    tap = (lambda name, v: taps.__setitem__(name, np.array(v))) if taps is not None else (lambda name, v: None)

    x = be.input(cfg['full_scale'], coords, feats)
    tap("input", x)
    x = be.subm(x, g('layers_in.1.weight'), sizes[0], 3)
    tap("layers_in", x)
    downs = []
    for k in range(n):
        pre = f"m_downs.{k}."
        j = 0
        if k > 0:
            x = be.bn_eval_instance(x, g(pre + "0.0.weight"), g(pre + "0.0.bias"))
            x = be.conv(x, g(pre + "0.1.weight"), sizes[k - 1], sizes[k], K[k - 1], S[k - 1])
            tap(f"down{k}_conv", x)
            j = 1
        b = pre + f"{j}.1."
        y = be.bn_eval_instance(x, g(b + "0.weight"), g(b + "0.bias"))
        y = be.subm(y, g(b + "1.weight"), sizes[k], 3)
        y = be.bn_eval_instance(y, g(b + "2.weight"), g(b + "2.bias"))
        y = be.subm(y, g(b + "3.weight"), sizes[k], 3)
        x = x + y
        tap(f"down{k}", x)
        downs.append(x)
    net = be.subm(x, g(f"m_shortcuts.{n - 1}.weight"), sizes[n - 1], 1)
    ups = [(net, n - 1)]
    for k in range(n - 1):
        j = n - 2 - k
        net = be.bn_eval_instance(net, g(f"m_ups.{k}.0.weight"), g(f"m_ups.{k}.0.bias"))
        net = be.deconv(net, g(f"m_ups.{k}.1.weight"), sizes[j + 1], sizes[j], K[j], S[j])
        net = net + be.subm(downs[j], g(f"m_shortcuts.{j}.weight"), sizes[j], 1)
        merged = be.subm(net, g(f"m_mergeds.{k}.weight"), sizes[j], 3)  # fpn_net.py:196: `net` stays un-merged
        tap(f"up{k + 1}", merged)
        ups.append((merged, j))
    rpn3d = [ups[i] for i in cfg['fpn_scales_from_top']]
    rpn2d = []
    for i, (f, lvl) in enumerate(rpn3d):
        sz = sizes[lvl]
        osz = [sz[0], sz[1], 1]
        rpn2d.append((be.conv(f, g(f"convs_pro2d.{i}.weight"), sz, osz, [1, 1, sz[2]], [1, 1, 1]), osz))
    maps = [(f, sizes[lvl]) for f, lvl in rpn3d] + rpn2d
    rpn = [maps[i] for i in cfg['rpn_3d_2d_selector']]
    roi = [(ups[i][0], sizes[ups[i][1]]) for i in cfg['roi_scales_from_top']]
    pack = lambda lst: [dict(features=np.array(f), spatial_size=list(sz), locations=np.array(be.locations(sz))) for f, sz in lst]
    return pack(rpn), pack(roi), be.macs


def run_fpn_port(cfg, state, coords, feats, taps=None):
    st = {k: np.asarray(v, dtype=np.float32) for k, v in state.items() if 'num_batches' not in k}
    return _run(_PortBackend(), cfg, st, np.asarray(coords, np.int64), np.asarray(feats, np.float32), taps)


def run_fpn_ref(cfg, state, coords, feats, taps=None):
    import torch
    st = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))).float() for k, v in state.items()}
    with torch.no_grad():
        return _run(_RefBackend(), cfg, st, np.asarray(coords, np.int64), np.asarray(feats, np.float32), taps)
