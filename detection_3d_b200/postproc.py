"""The steps either side of the backbone path, behind the reference's names and argument meaning:

  * BoxCoder3D.decode          maskrcnn_benchmark/modeling/box_coder_3d.py:9-65
  * boxes_iou_3d               utils3d/rotate_nms_3d_torch.py:22-84 (+ rotate_iou_gpu_eval, second/core/non_max_suppression/nms_gpu.py:611-654)
  * rotate_nms_3d              second/pytorch/core/box_torch_ops.py:489-514 (+ rotate_nms_3d_cc, second/core/non_max_suppression/nms_cpu.py:32-44)
  * boxlist_nms_3d             maskrcnn_benchmark/structures/boxlist_ops_3d.py:14-61
  * RPNPostProcessor           maskrcnn_benchmark/modeling/rpn/inference_3d.py:17-185 (forward_for_single_feature_map, eval path)
  * voxelize                   data3d/suncg_utils/suncg_dataset.py:115-177 (the deterministic part: affine map, offset, bounds mask, .long())

The reference mixes torch GPU ops, a numba CUDA kernel with host round trips and a C++ host loop (spconv); here every step is a CUDA
kernel of libscn_b200.so (csrc/postproc.cu) on the caller's stream.  No CPU fallback: CPU tensors are moved to the current device,
and without the library the calls raise.
"""
import ctypes as C
import math

import numpy as np
import torch

from ._lib import check, lib


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t, dtype=torch.float32):
    if not torch.cuda.is_available():
        raise RuntimeError("detection_3d_b200.postproc: no CUDA device (this implementation has no CPU path)")
    t = torch.as_tensor(t)
    return t.to(device="cuda" if not t.is_cuda else t.device, dtype=dtype).contiguous()


def top_k(values, k, sigmoid=False):
    """torch.topk(values.sigmoid() if sigmoid else values, k, sorted=True) on the device -> (values, indices); ties: lower index first."""
    v = _dev(values).reshape(-1)
    k = int(k)
    out_v = torch.empty(k, dtype=torch.float32, device=v.device)
    out_i = torch.empty(k, dtype=torch.int64, device=v.device)
    check(lib().scn_top_k_descending(_p(v), v.numel(), int(sigmoid), k, _p(out_v), _p(out_i), _stream()))
    return out_v, out_i


class BoxCoder3D(object):
    """box_coder_3d.py:9-28: weights, smooth_dim = True, bbox_xform_clip = 10000 (smooth) or log(1000) (exp)."""

    def __init__(self, weights=(1.0,) * 7, smooth_dim=True):
        self.smooth_dim = smooth_dim
        self.weights = torch.tensor(weights, dtype=torch.float32).view(1, 7)
        self.bbox_xform_clip = 10000. / 1 if smooth_dim else math.log(1000. / 1)

    def decode(self, box_encodings, anchors, indices=None):
        """box_coder_3d.py:38-65.  box_encodings [N, 7 C], anchors [N, 7] -> [N, 7 C] (C classes share an anchor, :50-54).
        `indices` (int64, optional; one class only) decodes rows `indices` of both tensors in the same launch (the RPN's top-k gather)."""
        enc, anc = _dev(box_encodings), _dev(anchors)
        assert enc.shape[0] == anc.shape[0]
        assert anc.shape[1] == 7
        num_classes = int(enc.shape[1] / 7)
        idx = None
        if num_classes != 1:
            assert indices is None
            num_loc = enc.shape[0]
            enc = enc.view(-1, 7)
            anc = anc.view(num_loc, 1, 7).repeat(1, num_classes, 1).view(-1, 7).contiguous()
        elif indices is not None:
            idx = _dev(indices, torch.int64)
        n = idx.numel() if idx is not None else enc.shape[0]
        out = torch.empty((n, 7), dtype=torch.float32, device=enc.device)
        w = (C.c_float * 7)(*[float(v) for v in self.weights.view(-1).tolist()])
        check(lib().scn_box_decode_3d(_p(enc), _p(anc), _p(idx), n, w, float(self.bbox_xform_clip), int(self.smooth_dim), _p(out), _stream()))
        return out.view(-1, num_classes * 7) if num_classes != 1 else out


_FLAG_RULES = {  # utils3d/rotate_nms_3d_torch.py:32-46
    'rpn_label_generation': lambda a: a['anchor_Y'] == 0 and a['target_Y'] >= 0.3,
    'roi_label_generation': lambda a: a['anchor_Y'] >= 0.3 and a['target_Y'] >= 0.3,
    'eval': lambda a: a['anchor_Y'] <= 0.3 and a['target_Y'] <= 0.3,
}


def boxes_iou_3d(targets_bbox3d, anchors_bbox3d, aug_thickness=None, criterion=-1, only_xy=False, flag=''):
    """utils3d/rotate_nms_3d_torch.py:22-84 -> [n_targets, n_anchors] float32 on the device."""
    if flag in _FLAG_RULES:
        assert _FLAG_RULES[flag](aug_thickness)
    elif flag in ('rpn_post', 'roi_post'):
        assert aug_thickness is None
    else:
        raise NotImplementedError(flag)
    if aug_thickness is None:
        aug_thickness = {'target_Y': 0.0, 'target_Z': 0.0, 'anchor_Y': 0.0, 'anchor_Z': 0.0}
    t, a = _dev(targets_bbox3d), _dev(anchors_bbox3d)
    out = torch.zeros((t.shape[0], a.shape[0]), dtype=torch.float32, device=t.device)
    aug = (C.c_float * 4)(float(aug_thickness['target_Y']), float(aug_thickness['target_Z']), float(aug_thickness['anchor_Y']), float(aug_thickness['anchor_Z']))
    for s in range(0, t.shape[0], 65535):
        part = t[s:s + 65535]
        check(lib().scn_boxes_iou_3d(_p(part), part.shape[0], _p(a), a.shape[0], aug, int(criterion), int(bool(only_xy)), _p(out[s:]), _stream()))
    return out


def rotate_nms_3d(rbboxes, scores, pre_max_size=None, post_max_size=None, iou_threshold=0.5, flag='', lazy=False):
    """second/pytorch/core/box_torch_ops.py:489-514 -> int64 indices into rbboxes, in descending score order.
    lazy=True: no device round trip -- returns (keep [post] padded with index 0, n_keep 1-element device tensor); the caller truncates."""
    b, s = _dev(rbboxes), _dev(scores).reshape(-1)
    n = b.shape[0]
    if n == 0:
        empty = torch.zeros([0]).long().to(b.device)
        return (empty, torch.zeros(1, dtype=torch.int64, device=b.device)) if lazy else empty
    pre = min(n, pre_max_size) if pre_max_size is not None else n
    post = pre if post_max_size is None else min(pre, int(post_max_size))
    keep = torch.zeros(max(post, 1), dtype=torch.int64, device=b.device)
    n_keep = torch.zeros(1, dtype=torch.int64, device=b.device)
    check(lib().scn_rotate_nms_3d(_p(b), _p(s), n, pre, post, float(iou_threshold), _p(keep), _p(n_keep), _stream()))
    if lazy:
        return keep, n_keep
    return keep[:int(n_keep.item())]


class Boxes3D(object):
    """The few members of the reference's BoxList3D (maskrcnn_benchmark/structures/bounding_box_3d.py) this path touches:
    bbox3d [N, 7] yx_zb rows, per-example index scopes, extra fields, indexing by a LongTensor."""

    def __init__(self, bbox3d, size3d=None, mode="yx_zb", examples_idxscope=None, constants=None):
        self.bbox3d = bbox3d
        self.size3d = size3d
        self.mode = mode
        self.examples_idxscope = examples_idxscope if examples_idxscope is not None else torch.tensor([[0, bbox3d.shape[0]]])
        self.constants = constants or {}
        self.extra_fields = {}

    def add_field(self, name, value):
        self.extra_fields[name] = value

    def get_field(self, name):
        return self.extra_fields[name]

    def batch_size(self):
        return int(self.examples_idxscope.shape[0])

    def __len__(self):
        return int(self.bbox3d.shape[0])

    def __getitem__(self, item):
        out = Boxes3D(self.bbox3d[item], self.size3d, self.mode, torch.tensor([[0, int(self.bbox3d[item].shape[0])]]), self.constants)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out


def boxlist_nms_3d(boxlist, nms_thresh, nms_aug_thickness=None, max_proposals=-1, score_field="score", flag='', lazy=False):
    """maskrcnn_benchmark/structures/boxlist_ops_3d.py:14-61.  lazy=True -> (boxlist of max_proposals rows, n_keep device tensor): the rows
    beyond n_keep are padding, `truncate_lazy` cuts them off after ONE device round trip for all pending lists."""
    if nms_aug_thickness is None:
        nms_aug_thickness = [0, 0]
    if flag == 'rpn_post':
        assert max_proposals > 100, max_proposals
    elif flag == 'roi_post':
        assert max_proposals == -1
    else:
        raise NotImplementedError
    if max_proposals < 0:
        max_proposals = 500
    objectness = boxlist.get_field(score_field)
    bbox3d = boxlist.bbox3d.clone().detach()
    bbox3d[:, 3:5] = torch.clamp(bbox3d[:, 3:5], min=nms_aug_thickness[0])
    bbox3d[:, 5] = torch.clamp(bbox3d[:, 5], min=nms_aug_thickness[1])
    keep = rotate_nms_3d(bbox3d, objectness, pre_max_size=2000, post_max_size=max_proposals, iou_threshold=nms_thresh, flag=flag, lazy=lazy)
    if lazy:
        return boxlist[keep[0]], keep[1]
    return boxlist[keep]


def truncate_lazy(pending):
    """pending: list of (Boxes3D padded, n_keep device tensor) from boxlist_nms_3d(lazy=True) -> list of Boxes3D, one host read for all."""
    if not pending:
        return []
    counts = torch.cat([n for _, n in pending]).tolist()
    return [b[torch.arange(c, device=b.bbox3d.device)] if c < len(b) else b for (b, _), c in zip(pending, counts)]


class RPNPostProcessor(torch.nn.Module):
    """maskrcnn_benchmark/modeling/rpn/inference_3d.py:17-185 (inference: no ground-truth proposals are appended)."""

    def __init__(self, batch_size, fpn_pre_nms_top_n, fpn_post_nms_top_n, nms_thresh, nms_aug_thickness, min_size, box_coder=None):
        super().__init__()
        self.batch_size = batch_size
        self.fpn_pre_nms_top_n = fpn_pre_nms_top_n
        self.fpn_post_nms_top_n = fpn_post_nms_top_n
        self.nms_thresh = nms_thresh
        self.nms_aug_thickness = nms_aug_thickness
        self.min_size = min_size
        self.box_coder = box_coder if box_coder is not None else BoxCoder3D()

    def forward_for_single_feature_map(self, anchors, objectness, box_regression, targets=None, lazy=False):
        """anchors: Boxes3D (all examples of the batch concatenated), objectness [N], box_regression [N, 7] (:82-161).
        -> list of Boxes3D, one per example, with the field "objectness"."""
        assert objectness.shape[0] == box_regression.shape[0] == len(anchors)
        objectness, box_regression = _dev(objectness).reshape(-1), _dev(box_regression)
        anchor_boxes = _dev(anchors.bbox3d)
        result = []
        for bi in range(anchors.batch_size()):
            s, e = [int(v) for v in anchors.examples_idxscope[bi]]
            n_top = min(self.fpn_pre_nms_top_n, e - s)
            objectness_i, topk_idx = top_k(objectness[s:e], n_top, sigmoid=True)                       # :97-104
            proposals_i = self.box_coder.decode(box_regression[s:e], anchor_boxes[s:e], indices=topk_idx)  # :107-120
            size3d = None if anchors.size3d is None else anchors.size3d[bi:bi + 1]
            boxlist = Boxes3D(proposals_i, size3d, mode="yx_zb", examples_idxscope=torch.tensor([[0, proposals_i.shape[0]]]), constants={'prediction': True})
            boxlist.add_field("objectness", objectness_i)
            result.append(boxlist_nms_3d(boxlist, self.nms_thresh, nms_aug_thickness=self.nms_aug_thickness, max_proposals=self.fpn_post_nms_top_n,
                                         score_field="objectness", flag='rpn_post', lazy=lazy))      # :140-147
        return result

    def forward(self, anchors, objectness, box_regression, targets=None, add_gt_proposals=False, lazy=False):
        if self.training and add_gt_proposals:
            raise NotImplementedError("appending ground-truth boxes (training) is outside this path")
        return self.forward_for_single_feature_map(anchors, objectness, box_regression, targets, lazy=lazy)


def voxelize(xyz, feats, scale, full_scale, matrix=None, offset=None, xyz_in_feats=True, batch_index=None):
    """data3d/suncg_utils/suncg_dataset.py:115-177 for one scene: a = xyz @ matrix (default eye * scale, :117-119 without augmentation),
    a += offset (default -a.min(0), :127-134), feats[:, 0:3] = a / scale (:149-151), rows outside [0, full_scale) dropped (:163-174),
    locs = a.long() (:175).  xyz float32 [N, 3], feats float32 [N, F].  -> (locs int64 [n, 3] or [n, 4] with batch_index, feats [n, F],
    size3d float32 [1, 6] = min and max of the shifted points / scale, :137-140)."""
    xyz, feats = _dev(xyz), _dev(feats)
    n = xyz.shape[0]
    m = np.eye(3) * scale if matrix is None else np.asarray(matrix, np.float64)
    mat = (C.c_double * 9)(*m.reshape(-1).tolist())
    lo, hi = (C.c_double * 3)(), (C.c_double * 3)()
    check(lib().scn_voxelize_extent(_p(xyz), n, mat, lo, hi, _stream()))
    lo, hi = np.array(list(lo)), np.array(list(hi))
    off = -lo if offset is None else np.asarray(offset, np.float64)
    size3d = torch.from_numpy(np.expand_dims(np.concatenate([(lo + off) / scale, (hi + off) / scale], 0), 0).astype(np.float32))
    cols = 3 if batch_index is None else 4
    locs = torch.empty((n, cols), dtype=torch.int64, device=xyz.device)
    fout = torch.empty_like(feats)
    kept = C.c_long()
    check(lib().scn_voxelize(_p(xyz), _p(feats), n, feats.shape[1], mat, (C.c_double * 3)(*off.tolist()), float(scale),
                             (C.c_double * 3)(*[float(v) for v in full_scale]), int(bool(xyz_in_feats)), int(batch_index or 0), cols, _p(locs), _p(fout),
                             C.byref(kept), _stream()))
    return locs[:kept.value], fout[:kept.value], size3d
