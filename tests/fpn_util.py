"""Shared helpers for the FPN parity tests: deterministic weights that do not depend on any module's
RNG consumption order, the two test configurations, and rulebook digests."""
import hashlib

import numpy as np
import torch


def mini4_config():
    """A 4-level miniature of sw4c_fpn432 (same block structure, small planes) for fast parity runs."""
    return dict(full_scale=[64, 64, 32], dimension=3, raw_elements=['xyz', 'color', 'normal'], reps=1, nPlanesF=[16, 32, 32, 64],
                nPlaneM=32, residual_blocks=True, fpn_scales_from_top=[2, 1], roi_scales_from_top=[1, 0],
                downsample=[[[2, 2, 2]] * 3, [[2, 2, 2]] * 3], rpn_map_sizes=[[32, 32, 16], [16, 16, 8]],
                rpn_3d_2d_selector=[0, 1, 2, 3], bn_momentum=0.9, track_running_stats=False)


def ref_ctor_args(cfg):
    """Positional/keyword arguments of the reference FPN_Net constructor (fpn_net.py:15-17)."""
    return ((cfg['full_scale'], cfg['dimension'], cfg['raw_elements'], cfg['reps'], cfg['nPlanesF'], cfg['nPlaneM'], cfg['residual_blocks'],
             cfg['fpn_scales_from_top'], cfg['roi_scales_from_top'], cfg['downsample'], cfg['rpn_map_sizes'], cfg['rpn_3d_2d_selector']),
            dict(bn_momentum=cfg['bn_momentum'], track_running_stats=cfg['track_running_stats']))


def deterministic_state(module, seed=0):
    """state_dict filled tensor by tensor from a numpy RandomState keyed by (seed, key name)."""
    out = {}
    for name, t in module.state_dict().items():
        h = int(hashlib.sha1(f"{seed}:{name}".encode()).hexdigest()[:8], 16)
        rs = np.random.RandomState(h)
        shape = tuple(t.shape)
        if name.endswith("running_mean"):
            v = 0.1 * rs.randn(*shape)
        elif name.endswith("running_var"):
            v = 1.0 + 0.2 * rs.rand(*shape)
        elif t.dim() == 4:  # conv weight (K, 1, Cin, Cout): reference init scale
            v = rs.randn(*shape) * (2.0 / (shape[0] * shape[2])) ** 0.5
        elif t.dim() == 1 and name.endswith("weight"):  # BN gamma
            v = 1.0 + 0.1 * rs.randn(*shape)
        elif t.dim() == 1:  # BN beta / biases
            v = 0.1 * rs.randn(*shape)
        else:
            v = 0.05 * rs.randn(*shape)
        out[name] = torch.from_numpy(np.asarray(v, dtype=np.float32)).reshape(shape)
    return out


def features_for(coords, channels=9, seed=0):
    rs = np.random.RandomState(seed + 12345)
    return rs.randn(coords.shape[0], channels).astype(np.float32)


def digest(arr):
    a = np.ascontiguousarray(arr)
    return hashlib.sha1(a.tobytes()).hexdigest()


def rulebook_digest(lists):
    """checksum of checksums over the lists of one rulebook (int32 pairs, reference layout)."""
    h = hashlib.sha1()
    for a in lists:
        h.update(digest(np.asarray(a, dtype=np.int32)).encode())
        h.update(str(int(np.asarray(a).size)).encode())
    return h.hexdigest()


def pyramid_sizes(full_scale, n_levels):
    return [[int(s) // (2 ** l) for s in full_scale] for l in range(n_levels)]


def metadata_digests(md, full_scale, n_levels, pro2d_levels=()):
    """Digest every structure the rulebook parity contract covers, for an object with the
    OracleMetadata / RefMetadata / GPU-adapter interface.  The input layer must have been run."""
    out = {}
    sizes = pyramid_sizes(full_scale, n_levels)
    for l, sz in enumerate(sizes):
        out[f"n{l}"] = int(md.nactive(sz))
        out[f"loc{l}"] = digest(np.asarray(md.spatial_locations(sz), dtype=np.int64))
        it = [np.asarray(md.iteration_order(sz, b), dtype=np.int32) for b in range(md.batch_size(sz))]
        out[f"iter{l}"] = digest(np.concatenate(it) if it else np.zeros(0, np.int32))
        out[f"subm{l}"] = rulebook_digest(md.submanifold_rules(sz, [3, 3, 3]))
        if l + 1 < n_levels:
            out[f"conv{l}"] = rulebook_digest(md.conv_rules(sz, sizes[l + 1], [2, 2, 2], [2, 2, 2]))
    for l in pro2d_levels:
        sz = sizes[l]
        o = [sz[0], sz[1], 1]
        out[f"pro2d{l}"] = rulebook_digest(md.conv_rules(sz, o, [1, 1, sz[2]], [1, 1, 1]))
        out[f"pro2d{l}_loc"] = digest(np.asarray(md.spatial_locations(o), dtype=np.int64))
    return out
