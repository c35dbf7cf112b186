"""RPN head and anchor generation on the backbone's sparse feature maps -- the first consumers of the hot path
(SURVEY.md section 8f rank 1), behind the reference's own class names, constructor arguments, parameter names and output shapes:

  * RPNHead           maskrcnn_benchmark/modeling/rpn/rpn_sparse3d.py:81-131   (conv / cls_logits / bbox_pred are nn.Conv2d
                      modules, so a Detection_3D checkpoint loads unchanged; forward = ONE native kernel per map)
  * AnchorGenerator   maskrcnn_benchmark/modeling/rpn/anchor_generator_sparse3d.py:39-168 (grid_anchors on the device from
                      device-resident get_spatial_locations; the reference copies the whole hash map to the host first,
                      SCN/Metadata/Metadata.cpp:147-168)
  * generate_anchors_3d / _yaws / _ratio, examples_bidx_2_sizes, cat_scales_obj_reg  (same files)

No CPU fallback: the forward raises without the CUDA library.
"""
import ctypes as C

import numpy as np
import torch
from torch import nn

from ._lib import check, lib


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class RPNHead(nn.Module):
    """`cfg`-free constructor: seperate_rpn = len(SEPARATE_CLASSES) * SEPARATE_RPN + 1 (rpn_sparse3d.py:95)."""

    def __init__(self, in_channels, num_anchors_per_location, seperate_rpn=1):
        super().__init__()
        self.num_anchors_per_location = num_anchors_per_location
        self.seperate_rpn = seperate_rpn
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.cls_logits = nn.Conv2d(in_channels, num_anchors_per_location * seperate_rpn, kernel_size=1, stride=1)
        self.bbox_pred = nn.Conv2d(in_channels, num_anchors_per_location * 7 * seperate_rpn, kernel_size=1, stride=1)
        for l in (self.conv, self.cls_logits, self.bbox_pred):
            torch.nn.init.normal_(l.weight, std=0.01)
            torch.nn.init.constant_(l.bias, 0)

    def forward(self, x):
        """x: list of feature maps, each [1, C, n, 1] (the reference's layout, rpn_sparse3d.py:172-176) or [n, C] rows.
        -> (logits: list of [1, n, A, sep], bbox_reg: list of [1, n, A, 7 sep])."""
        logits, bbox_reg = [], []
        A, sep = self.num_anchors_per_location, self.seperate_rpn
        for f in x:
            rows = f if f.dim() == 2 else f[0, :, :, 0].t()
            rows = rows.contiguous().float()
            if not rows.is_cuda:
                raise RuntimeError("RPNHead: expected CUDA features (this implementation has no CPU path)")
            n, c = rows.shape
            lg = torch.empty((n, A * sep), device=rows.device)
            rg = torch.empty((n, A * 7 * sep), device=rows.device)
            wc, wl, wr = (m.weight.detach().reshape(m.weight.size(0), -1).contiguous() for m in (self.conv, self.cls_logits, self.bbox_pred))
            check(lib().scn_rpn_head_forward(_p(rows), n, c, _p(wc), _p(self.conv.bias), _p(wl), _p(self.cls_logits.bias), A * sep, _p(wr),
                                             _p(self.bbox_pred.bias), A * 7 * sep, _p(lg), _p(rg), _stream()))
            logits.append(lg.view(1, n, A, sep))
            bbox_reg.append(rg.view(1, n, A, 7 * sep))
        return logits, bbox_reg


def generate_anchors_3d_yaws(size, yaws, centroids=np.array([[0, 0, 0]])):
    """yx_zb boxes [xc, yc, z_bot, y_size, x_size, z_size, yaw], one per yaw (anchor_generator_sparse3d.py:238-250)."""
    out = [np.concatenate([centroids[k], size, yaws[j]]).reshape(1, -1) for j in range(yaws.shape[0]) for k in range(centroids.shape[0])]
    return torch.from_numpy(np.concatenate(out, 0))


def generate_anchors_3d_ratio(size, ratios, centroids=np.array([[0, 0, 0]])):
    """yaw 0, one box per size ratio (anchor_generator_sparse3d.py:220-236)."""
    yaw = np.array([0], dtype=np.float32)
    out = [np.concatenate([centroids[k], size * ratios[j], yaw]).reshape(1, -1) for j in range(ratios.shape[0]) for k in range(centroids.shape[0])]
    return torch.from_numpy(np.concatenate(out, 0))


def generate_anchors_3d(size, yaws, ratios, use_yaw, centroids=np.array([[0, 0, 0]])):
    return generate_anchors_3d_yaws(size, yaws, centroids) if use_yaw else generate_anchors_3d_ratio(size, ratios, centroids)


def examples_bidx_2_sizes(examples_bidx):
    """[batch, 2] (start, end) row ranges of the batch items (anchor_generator_sparse3d.py:171-181)."""
    batch_size = int(examples_bidx[-1]) + 1
    counts = torch.bincount(examples_bidx.cpu().long(), minlength=batch_size)
    end = torch.cumsum(counts, 0)
    return torch.stack([end - counts, end], 1)


class AnchorGenerator(nn.Module):
    def __init__(self, voxel_scale=20, sizes_3d=[[0.2, 1, 3], [0.5, 2, 3], [1, 3, 3]], yaws=(0, -1.57), ratios=[(1, 1, 1), (1, 2, 1)],
                 use_yaws=[1, 1, 1], anchor_strides=[[8, 8, 729], [16, 16, 729], [32, 32, 729]], scene_size=[8, 8, 5], straddle_thresh=0):
        super().__init__()
        sizes_3d = np.array(sizes_3d, dtype=np.float32)
        anchor_strides = np.array(anchor_strides, dtype=np.float32)
        assert sizes_3d.shape[1] == 3 and anchor_strides.shape[1] == 3 and sizes_3d.shape[0] == anchor_strides.shape[0]
        yaws = np.array(yaws, dtype=np.float32).reshape([-1, 1])
        ratios = np.array(ratios, dtype=np.float32)
        self.cell_anchors = [generate_anchors_3d(size, yaws, ratios, uy).float() for size, uy in zip(sizes_3d, use_yaws)]
        self.anchor_num_per_loc = len(yaws)
        self.voxel_scale = voxel_scale
        self.strides = torch.from_numpy(anchor_strides)
        self.straddle_thresh = straddle_thresh
        self.anchor_mode = 'yx_zb'
        self.scene_size = torch.tensor(scene_size, dtype=torch.float)
        self._dev_cells = {}

    def num_anchors_per_location(self):
        return self.anchor_num_per_loc

    def grid_anchors(self, locations, device=None):
        """locations: one int64 [n, 4] tensor per scale (CPU, as the reference's get_spatial_locations returns them, or CUDA).
        -> list of [n * A, 7] CUDA tensors, flatten order [location, yaw] (anchor_generator_sparse3d.py:88-104)."""
        assert len(self.cell_anchors) == len(locations), "scales num not right"
        if device is None:
            device = next((l.device for l in locations if l.is_cuda), torch.device("cuda", torch.cuda.current_device()))
        out = []
        for i, (base, loc, stride) in enumerate(zip(self.cell_anchors, locations, self.strides)):
            key = (i, str(device))
            if key not in self._dev_cells:
                self._dev_cells[key] = base.contiguous().to(device)
            loc = loc.to(device).contiguous()
            n, A = loc.size(0), base.size(0)
            anchors = torch.empty((n * A, 7), device=device)
            st = (C.c_float * 3)(*[float(v) for v in stride])
            check(lib().scn_rpn_grid_anchors(_p(loc), n, float(self.voxel_scale), st, _p(self._dev_cells[key]), A, _p(anchors), _stream()))
            out.append(anchors)
        return out

    def forward(self, points_sparse, feature_maps_sparse, targets=None):
        """-> (anchors per scale, examples_idxscope per scale); the reference wraps the pair in BoxList3D (:137-147)."""
        locations = [f.metadata.getSpatialLocations(f.spatial_size, device="cuda") if hasattr(f.metadata, "getSpatialLocations") else f.get_spatial_locations()
                     for f in feature_maps_sparse]
        anchors = self.grid_anchors(locations)
        scopes = []
        for f, l in zip(feature_maps_sparse, locations):
            # one batch item (every shipped config): the scope is the whole map -- known on the host, where the reference reads the
            # batch column back (two device round trips per map)
            one = hasattr(f.metadata, "getBatchSize") and f.metadata.getBatchSize(f.spatial_size) == 1
            scopes.append((torch.tensor([[0, l.shape[0]]]) if one else examples_bidx_2_sizes(l[:, -1])) * self.anchor_num_per_loc)
        return anchors, scopes


def cat_scales_obj_reg(objectness, rpn_box_regression, examples_idxscopes):
    """Flatten order [batch, scale, location, yaw] (rpn_sparse3d.py:20-76): objectness [sum n A, sep], regression [sum n A, 7 sep]."""
    batch = examples_idxscopes[0].shape[0]
    obj_new, reg_new = [[] for _ in range(batch)], [[] for _ in range(batch)]
    for s in range(len(objectness)):
        sep = objectness[s].shape[-1]
        obj_s = objectness[s].reshape(-1, sep)
        reg_s = rpn_box_regression[s].reshape(-1, 7 * sep)
        for b in range(batch):
            begin, end = [int(v) for v in examples_idxscopes[s][b]]
            obj_new[b].append(obj_s[begin:end])
            reg_new[b].append(reg_s[begin:end])
    return torch.cat([torch.cat(o, 0) for o in obj_new], 0), torch.cat([torch.cat(r, 0) for r in reg_new], 0)


def sw4c_anchor_generator():
    """configs/sw4c/sw4c_fpn432_bs1_lr5.yaml:12-15,31 with ANCHOR_STRIDE as tools/train_net_sparse3d.py:254-270 resolves it
    (RPN_SCALES_FROM_TOP [4,3,2], RPN_3D_2D_SELECTOR [1,3,4,5] -> strides 32, 16, 32, 64 for the maps FPN_Net returns)."""
    return AnchorGenerator(voxel_scale=50, sizes_3d=[[0.4, 1.5, 1.5], [0.2, 0.5, 3], [0.4, 1.5, 3], [0.6, 2.5, 3]], yaws=(0, -1.57, -0.785, 0.785),
                           ratios=[[1, 1, 1], [1, 2, 1], [2, 1, 1], [1.7, 1.7, 1]], use_yaws=[1, 1, 1, 1],
                           anchor_strides=[[32, 32, 32], [16, 16, 16], [32, 32, 32], [64, 64, 64]], scene_size=[40.96, 40.96, 10.24])
