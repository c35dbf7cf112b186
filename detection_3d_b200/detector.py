"""BASELINE.json config 3: backbone + FPN + RPN inference per building, every step on the device.

  * RPNModule     maskrcnn_benchmark/modeling/rpn/rpn_sparse3d.py:133-201,286-305 (inference path: head -> anchors -> cat_scales_obj_reg /
                  cat_scales_anchor -> box_selector_test, once per class group when SEPARATE_CLASSES is set,
                  maskrcnn_benchmark/modeling/seperate_classifier.py:58-81)
  * SparseRPNDetector   modeling/detector/sparse_rcnn.py with RPN_ONLY: backbone(points) -> rpn(points, features)

Built from detection_3d_b200.sparseconvnet.FPN_Net (the hot path), .rpn (head + anchors) and .postproc (decode + rotated NMS).
The reference's BoxList3D (needs open3d) is replaced by postproc.Boxes3D, which carries the members this path uses.
"""
import torch
from torch import nn

from . import postproc, rpn


def cat_scales_anchor(anchors, examples_idxscopes):
    """Flatten order [batch, scale, location, yaw] (bounding_box_3d.py cat_scales_anchor): per example the anchors of all scales."""
    batch = examples_idxscopes[0].shape[0]
    per_ex = [[] for _ in range(batch)]
    for s in range(len(anchors)):
        for b in range(batch):
            begin, end = [int(v) for v in examples_idxscopes[s][b]]
            per_ex[b].append(anchors[s][begin:end])
    flat = [torch.cat(a, 0) for a in per_ex]
    ends = torch.cumsum(torch.tensor([f.shape[0] for f in flat]), 0)
    scope = torch.stack([ends - torch.tensor([f.shape[0] for f in flat]), ends], 1)
    return postproc.Boxes3D(torch.cat(flat, 0), None, "yx_zb", scope, {})


class RPNModule(nn.Module):
    """cfg-free constructor: the values tools/train_net_sparse3d.py:233-255 and config/defaults.py:150-168 resolve for a config."""

    def __init__(self, in_channels=128, anchor_generator=None, seperate_rpn=2, pre_nms_top_n=1500, post_nms_top_n=750, nms_thresh=0.5,
                 nms_aug_thickness=(0.3, 0.3), min_size=0, batch_size=1):
        super().__init__()
        self.anchor_generator = anchor_generator if anchor_generator is not None else rpn.sw4c_anchor_generator()
        self.head = rpn.RPNHead(in_channels, self.anchor_generator.num_anchors_per_location(), seperate_rpn)
        self.box_coder = postproc.BoxCoder3D()
        self.box_selector_test = postproc.RPNPostProcessor(batch_size, pre_nms_top_n, post_nms_top_n, nms_thresh, list(nms_aug_thickness), min_size, self.box_coder)
        self.group_num = seperate_rpn

    def forward(self, inputs_sparse, features_sparse, targets=None):
        """-> (boxes, {}): boxes = one list (per example) of Boxes3D per class group (a single list when there is one group)."""
        if self.training:
            raise NotImplementedError("RPN training (loss evaluator, ground-truth proposals) is outside this path")
        objectness, rpn_box_regression = self.head([fs.features for fs in features_sparse])
        anchors, scopes = self.anchor_generator(inputs_sparse, features_sparse, targets)
        objectness, rpn_box_regression = rpn.cat_scales_obj_reg(objectness, rpn_box_regression, scopes)
        anchors = cat_scales_anchor(anchors, scopes)
        anchors.constants['scale_num'] = len(scopes)
        anchors.constants['num_anchors_per_location'] = self.head.num_anchors_per_location
        self.box_selector_test.eval()
        if self.group_num == 1:
            return self.box_selector_test(anchors, objectness.squeeze(1), rpn_box_regression, targets), {}
        # seperate_classifier.py:68-71.  The class groups are independent: each is queued on a stream of its own (the greedy NMS sweep
        # is one CTA) and nothing is read back until all of them are queued -- one device round trip for the proposal counts instead
        # of one per group in the middle of the work.
        cur = torch.cuda.current_stream()
        if not hasattr(self, "_streams") or len(self._streams) != self.group_num:
            self._streams = [torch.cuda.Stream() for _ in range(self.group_num)]
        pending = []
        for gi, st in enumerate(self._streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                obj_g, reg_g = objectness[:, gi].contiguous(), rpn_box_regression[:, gi * 7:gi * 7 + 7].contiguous()
                for t in (objectness, rpn_box_regression, anchors.bbox3d):
                    t.record_stream(st)
                pending.append(self.box_selector_test(anchors, obj_g, reg_g, targets, lazy=True))
        for st in self._streams:
            cur.wait_stream(st)
        for group in pending:  # allocated on the side streams, consumed on the caller's from here on
            for b, n_keep in group:
                for t in [b.bbox3d, n_keep] + list(b.extra_fields.values()):
                    t.record_stream(cur)
        flat = postproc.truncate_lazy([pb for group in pending for pb in group])
        boxes_g, k = [], 0
        for group in pending:
            boxes_g.append(flat[k:k + len(group)])
            k += len(group)
        return boxes_g, {}


class SparseRPNDetector(nn.Module):
    """Backbone (scn.FPN_Net) + RPNModule: points in, proposals out."""

    def __init__(self, backbone, rpn_module=None):
        super().__init__()
        self.backbone = backbone
        self.rpn = rpn_module if rpn_module is not None else RPNModule()

    def forward(self, points):
        rpn_maps, _roi_maps = self.backbone(points)
        return self.rpn(points, rpn_maps)[0]
