"""ORACLE / TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the RPN post-processing chain and the input voxeliser.
Never imported by the product (detection_3d_b200/); only tests/ use it.

  * box_decode            maskrcnn_benchmark/modeling/box_coder_3d.py:38-65 + second/pytorch/core/box_torch_ops.py:51-88 + utils3d/geometric_torch.py:4-10
  * rotated_iou_2d        second/core/non_max_suppression/nms_gpu.py:166-420 (inter: corners, inside points + edge intersections, vertex sort,
                          triangle fan), :548-570 (devRotateIoUEval criteria), :657-667 (check_same_boxes)
  * boxes_iou_3d          utils3d/rotate_nms_3d_torch.py:7-84
  * rotate_nms_3d         second/pytorch/core/box_torch_ops.py:489-514, second/core/non_max_suppression/nms_cpu.py:32-44
  * rpn_post_process      maskrcnn_benchmark/modeling/rpn/inference_3d.py:82-161 + maskrcnn_benchmark/structures/boxlist_ops_3d.py:14-61
  * voxelize              data3d/suncg_utils/suncg_dataset.py:115-177

Pinning: tests/golden/postproc.npz is produced by tests/golden/make_golden_postproc.py, which EXECUTES the reference's own functions
(taken from /root/reference with `ast`; the numba CUDA kernels of nms_gpu.py run under numba's CUDA simulator) -- this restatement is
checked against those vectors in tests/test_oracle_cpu.py.

Third-party dependency absent from /root/reference: `spconv.utils.rotate_non_max_suppression_cpu` (traveller59/spconv v1.x,
include/spconv/nms.h; version not pinned by the reference).  Its published algorithm, restated here by recollection: visit the boxes
in the given order; a box not yet suppressed is kept and suppresses every later box j with standup_iou[i, j] > 0 whose polygon
intersection-over-union with it (boost::geometry intersection and union_ of the two corner quadrilaterals) is >= thresh.  The reference
passes its 3-D IoU matrix as `standup_iou` (nms_cpu.py:35-43).  For this one function parity is therefore pinned to the restatement,
not to an execution of spconv.
"""
import math

import numpy as np


# ------------------------------------------------------------------ box decode
def box_decode(enc, anchors, weights=(1.0,) * 7, smooth_dim=True):
    enc = np.asarray(enc, np.float32).copy()
    anchors = np.asarray(anchors, np.float32)
    clip = np.float32(10000.0 if smooth_dim else math.log(1000.0))
    enc = enc / np.asarray(weights, np.float32).reshape(1, 7)
    enc[:, 3:6] = np.minimum(enc[:, 3:6], clip)
    xa, ya, za, wa, la, ha, ra = [anchors[:, i] for i in range(7)]
    xt, yt, zt, wt, lt, ht, rt = [enc[:, i] for i in range(7)]
    diagonal = np.sqrt(la ** 2 + wa ** 2)
    xg, yg, zg = xt * diagonal + xa, yt * diagonal + ya, zt * ha + za
    if smooth_dim:
        lg, wg, hg = (lt + 1) * la, (wt + 1) * wa, (ht + 1) * ha
    else:
        lg, wg, hg = np.exp(lt) * la, np.exp(wt) * wa, np.exp(ht) * ha
    rg = rt + ra
    out = np.stack([xg, yg, zg, wg, lg, hg, rg], 1).astype(np.float32)
    pi = np.float32(math.pi)
    out[:, 6] = out[:, 6] - np.floor(out[:, 6] / pi + np.float32(0.5)) * pi
    return out


# ------------------------------------------------------------------ rotated IoU of two rectangles (float32 like the kernel's arrays)
def _corners(b):
    cx, cy, dx, dy, ang = [np.float32(v) for v in b]
    c, s = np.float32(math.cos(ang)), np.float32(math.sin(ang))
    lx = np.array([-dx / 2, -dx / 2, dx / 2, dx / 2], np.float32)
    ly = np.array([-dy / 2, dy / 2, dy / 2, -dy / 2], np.float32)
    return np.stack([c * lx + s * ly + cx, -s * lx + c * ly + cy], 1).astype(np.float32)


def _inside(p, q):  # point_in_quadrilateral
    ab, ad, ap = q[1] - q[0], q[3] - q[0], p - q[0]
    abab, abap, adad, adap = ab @ ab, ab @ ap, ad @ ad, ad @ ap
    return abab >= abap and abap >= 0 and adad >= adap and adap >= 0


def _segment(A, B, Cc, D):  # line_segment_intersection
    BA, DA, CA = B - A, D - A, Cc - A
    acd = DA[1] * CA[0] > CA[1] * DA[0]
    bcd = (D[1] - B[1]) * (Cc[0] - B[0]) > (Cc[1] - B[1]) * (D[0] - B[0])
    if acd != bcd:
        abc = CA[1] * BA[0] > BA[1] * CA[0]
        abd = DA[1] * BA[0] > BA[1] * DA[0]
        if abc != abd:
            DC = D - Cc
            ABBA = A[0] * B[1] - B[0] * A[1]
            CDDC = Cc[0] * D[1] - D[0] * Cc[1]
            DH = BA[1] * DC[0] - BA[0] * DC[1]
            return np.array([(ABBA * DC[0] - BA[0] * CDDC) / DH, (ABBA * DC[1] - BA[1] * CDDC) / DH], np.float32)
    return None


def intersection_area(b1, b2):
    p1, p2 = _corners(b1), _corners(b2)
    pts = []
    for i in range(4):
        if _inside(p1[i], p2):
            pts.append(p1[i])
        if _inside(p2[i], p1):
            pts.append(p2[i])
    for i in range(4):
        for j in range(4):
            t = _segment(p1[i], p1[(i + 1) % 4], p2[j], p2[(j + 1) % 4])
            if t is not None:
                pts.append(t)
    if len(pts) < 3:
        return np.float32(0)
    pts = np.array(pts, np.float32)
    v = pts - pts.mean(0)
    d = np.sqrt((v * v).sum(1))
    v = v / d[:, None]
    key = np.where(v[:, 1] < 0, -2 - v[:, 0], v[:, 0])  # sort_vertex_in_convex_polygon
    pts = pts[np.argsort(key, kind="stable")]
    area = np.float32(0)
    for i in range(len(pts) - 2):
        a, b, c = pts[0], pts[i + 1], pts[i + 2]
        area += abs(((a[0] - c[0]) * (b[1] - c[1]) - (a[1] - c[1]) * (b[0] - c[0])) / 2.0)
    return np.float32(area)


def rotated_iou_2d(boxes, query, criterion=-1):
    """rotate_iou_gpu_eval(boxes [N, 5], query_boxes [K, 5]) -> [N, K]; rbox1 = query, rbox2 = box (nms_gpu.py:601-608)."""
    boxes, query = np.asarray(boxes, np.float32), np.asarray(query, np.float32)
    out = np.zeros((boxes.shape[0], query.shape[0]), np.float32)
    for n in range(boxes.shape[0]):
        for k in range(query.shape[0]):
            r1, r2 = query[k], boxes[n]
            a1, a2 = r1[2] * r1[3], r2[2] * r2[3]
            inter = intersection_area(r1, r2)
            with np.errstate(divide="ignore", invalid="ignore"):
                if criterion == -1:
                    v = inter / (a1 + a2 - inter)
                elif criterion == 0:
                    v = inter / a1
                elif criterion == 1:
                    v = inter / a2
                elif criterion == 2:
                    thin = min(r2[2], r2[3]) / max(r2[2], r2[3]) < 0.25
                    v = inter / (a2 + max(0, a1 * 0.5 - inter)) if thin else inter / (a1 + a2 - inter)
                else:
                    v = inter
            if np.all(np.abs(r1 - r2) < 1e-6):
                v = 1
            out[n, k] = v
    return out


def boxes_iou_3d(targets, anchors, aug=None, criterion=-1, only_xy=False):
    t, a = np.asarray(targets, np.float32).copy(), np.asarray(anchors, np.float32).copy()
    aug = aug or {'target_Y': 0.0, 'target_Z': 0.0, 'anchor_Y': 0.0, 'anchor_Z': 0.0}
    t[:, 3] = np.maximum(t[:, 3], aug['target_Y']); a[:, 3] = np.maximum(a[:, 3], aug['anchor_Y'])
    t[:, 5] = np.maximum(t[:, 5], aug['target_Z']); a[:, 5] = np.maximum(a[:, 5], aug['anchor_Z'])
    tz = np.stack([t[:, 2], t[:, 2] + t[:, 5]], 1)[:, None, :]
    az = np.stack([a[:, 2], a[:, 2] + a[:, 5]], 1)[None, :, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        iouz = (np.minimum(az[..., 1], tz[..., 1]) - np.maximum(az[..., 0], tz[..., 0])) / (np.maximum(az[..., 1], tz[..., 1]) - np.minimum(az[..., 0], tz[..., 0]))
    iou2d = rotated_iou_2d(t[:, [0, 1, 3, 4, 6]], a[:, [0, 1, 3, 4, 6]], criterion)
    return iou2d if only_xy else (iou2d * iouz).astype(np.float32)


# ------------------------------------------------------------------ NMS
def rotate_non_max_suppression_cpu(corners, order, standup_iou, thresh):
    """spconv v1.x restated (see the header): polygon IoU = intersection / union of the two corner quadrilaterals."""
    n = corners.shape[0]
    suppressed = np.zeros(n, bool)
    areas = [abs(_poly_area(corners[i])) for i in range(n)]
    keep = []
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(int(i))
        for _j in range(_i + 1, n):
            j = order[_j]
            if suppressed[j] or not (standup_iou[i, j] > 0):
                continue
            inter = _clip_area(corners[i], corners[j])
            union = areas[i] + areas[j] - inter
            if union > 0 and inter > 0 and inter / union >= thresh:
                suppressed[j] = True
    return keep


def _poly_area(p):
    x, y = p[:, 0].astype(np.float64), p[:, 1].astype(np.float64)
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def _clip_area(p, q):
    """area of the intersection of two convex polygons (float64 Sutherland-Hodgman)."""
    poly = [tuple(map(float, v)) for v in p]
    sgn = 1.0 if _poly_area(q) >= 0 else -1.0
    for e in range(len(q)):
        a, b = q[e].astype(np.float64), q[(e + 1) % len(q)].astype(np.float64)
        ex, ey = b - a
        new = []
        if not poly:
            break
        s = poly[-1]
        sd = sgn * (ex * (s[1] - a[1]) - ey * (s[0] - a[0]))
        for c in poly:
            cd = sgn * (ex * (c[1] - a[1]) - ey * (c[0] - a[0]))
            if (cd >= 0) != (sd >= 0):
                t = sd / (sd - cd)
                new.append((s[0] + t * (c[0] - s[0]), s[1] + t * (c[1] - s[1])))
            if cd >= 0:
                new.append(c)
            s, sd = c, cd
        poly = new
    if len(poly) < 3:
        return 0.0
    return abs(_poly_area(np.array(poly)))


def center_to_corner_box2d(centers, dims, angles):
    """second/core/box_np_ops.py:374-394 with corners_nd (:176-207) and rotation_2d (:313-326)."""
    norm = np.array([[0, 0], [0, 1], [1, 1], [1, 0]], dims.dtype) - np.array(0.5, dims.dtype)
    corners = dims.reshape(-1, 1, 2) * norm.reshape(1, 4, 2)
    s, c = np.sin(angles), np.cos(angles)
    rot_t = np.stack([[c, -s], [s, c]])
    return np.einsum('aij,jka->aik', corners, rot_t) + centers.reshape(-1, 1, 2)


def rotate_nms_3d(boxes, scores, pre_max_size=None, post_max_size=None, iou_threshold=0.5):
    """-> indices into boxes (descending score).  Ties in the scores: lower index first (the CUDA path's documented order)."""
    boxes, scores = np.asarray(boxes, np.float32), np.asarray(scores, np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros(0, np.int64)
    indices = np.argsort(-scores, kind="stable")[:min(n, pre_max_size) if pre_max_size is not None else n]
    b = boxes[indices]
    ious = boxes_iou_3d(b, b)
    corners = center_to_corner_box2d(b[:, :2], b[:, 3:5], b[:, 6])
    keep = rotate_non_max_suppression_cpu(corners, np.arange(b.shape[0]), ious, iou_threshold)
    keep = np.array(keep[:post_max_size], np.int64)
    return indices[keep] if keep.size else np.zeros(0, np.int64)


def rpn_post_process(anchors, objectness, regression, pre_nms_top_n, post_nms_top_n, nms_thresh, nms_aug_thickness=(0, 0), weights=(1.0,) * 7):
    """inference_3d.py:82-161 for one example -> (boxes [n, 7], objectness [n])."""
    obj = (1.0 / (1.0 + np.exp(-np.asarray(objectness, np.float32)))).astype(np.float32)
    k = min(pre_nms_top_n, obj.shape[0])
    idx = np.argsort(-obj, kind="stable")[:k]
    props = box_decode(np.asarray(regression, np.float32)[idx], np.asarray(anchors, np.float32)[idx], weights)
    b = props.copy()
    b[:, 3:5] = np.maximum(b[:, 3:5], nms_aug_thickness[0])
    b[:, 5] = np.maximum(b[:, 5], nms_aug_thickness[1])
    keep = rotate_nms_3d(b, obj[idx], pre_max_size=2000, post_max_size=post_nms_top_n, iou_threshold=nms_thresh)
    return props[keep], obj[idx][keep]


# ------------------------------------------------------------------ voxeliser
def voxelize(xyz, feats, scale, full_scale, matrix=None, offset=None, xyz_in_feats=True):
    a = np.asarray(xyz, np.float32)
    m = np.eye(3) * scale if matrix is None else np.asarray(matrix, np.float64)
    a = np.matmul(a, m)                                  # :121
    lo = a.min(0)
    off = -lo if offset is None else np.asarray(offset, np.float64)
    a = a + off                                          # :134
    size3d = np.expand_dims(np.concatenate([a.min(0) / scale, a.max(0) / scale], 0), 0).astype(np.float32)  # :136-139
    b = np.asarray(feats, np.float32).copy()
    if xyz_in_feats:
        b[:, 0:3] = a / scale                            # :151
    full = np.asarray(full_scale, np.float64)
    idxs = (a.min(1) >= 0) * np.all(a < full[np.newaxis, :], 1)  # :163,173
    return a[idxs].astype(np.int64), b[idxs], size3d     # :174-177 (torch .long() truncates; values are >= 0)
