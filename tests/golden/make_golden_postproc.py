"""Generates tests/golden/postproc.npz from the REFERENCE'S OWN post-processing / voxeliser code (build container only).

The reference's packages do not import here (open3d, yacs, spconv, Python-3.12-incompatible imports), so the functions of the path are
taken out of the reference files with `ast` AT GENERATION TIME and executed unmodified:
  * second_box_decode, rotate_nms_3d      /root/reference/second/pytorch/core/box_torch_ops.py:51-88,489-514
  * limit_period                          /root/reference/utils3d/geometric_torch.py:4-10
  * BoxCoder3D                            /root/reference/maskrcnn_benchmark/modeling/box_coder_3d.py:9-65
  * rotate_iou_gpu_eval (+ device fns)    /root/reference/second/core/non_max_suppression/nms_gpu.py -- the module itself is imported
                                          (with an empty `spconv` stand-in) and its numba CUDA kernels run under NUMBA_ENABLE_CUDASIM=1
  * iou_one_dim, boxes_iou_3d             /root/reference/utils3d/rotate_nms_3d_torch.py:7-84
  * rotate_nms_3d_cc                      /root/reference/second/core/non_max_suppression/nms_cpu.py:32-44
  * center_to_corner_box2d, corners_nd, rotation_2d   /root/reference/second/core/box_np_ops.py
  * boxlist_nms_3d                        /root/reference/maskrcnn_benchmark/structures/boxlist_ops_3d.py:14-61
  * RPNPostProcessor                      /root/reference/maskrcnn_benchmark/modeling/rpn/inference_3d.py:17-185
  * the voxeliser statements              /root/reference/data3d/suncg_utils/suncg_dataset.py:115-177 (body of __getitem__)
Absent third-party piece: spconv.utils.rotate_non_max_suppression_cpu -> oracle/postproc_oracle.py's restatement (see its header).
BoxList3D (bounding_box_3d.py needs open3d) is replaced by a stand-in with the handful of members this path touches.
Inputs: anchors / logits / regression of tests/golden/rpn_sw4c_mid.npz (the reference's own RPN head on its own backbone outputs).

  NUMBA_ENABLE_CUDASIM=1 python tests/golden/make_golden_postproc.py      (a few minutes: the simulator interprets the kernels in Python)
"""
import ast
import importlib
import os
import sys
import types

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import postproc_oracle as po  # noqa: E402

REF = "/root/reference"
# the golden RPN head has random weights: its raw outputs are scaled to the magnitudes a trained head produces (regression deltas of
# ~0.1-0.3, logits of a few units -- sigmoid stays away from its float32 saturation, so the ranking has no ties)
OBJ_SCALE, REG_SCALE = 0.25, 0.08


def extract(path, names, ns):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in names:
            node.decorator_list = []
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


class BoxList3D(object):
    """stand-in for maskrcnn_benchmark/structures/bounding_box_3d.py (imports open3d): only what inference_3d.py / boxlist_ops_3d.py use"""

    def __init__(self, bbox3d, size3d, mode, examples_idxscope, constants):
        self.bbox3d, self.size3d, self.mode, self.examples_idxscope, self.constants = bbox3d, size3d, mode, examples_idxscope, constants
        self.extra_fields = {}

    def add_field(self, k, v):
        self.extra_fields[k] = v

    def get_field(self, k):
        return self.extra_fields[k]

    def set_as_prediction(self):
        self.constants['prediction'] = True

    def batch_size(self):
        return self.examples_idxscope.shape[0]

    def __len__(self):
        return self.bbox3d.shape[0]

    def __getitem__(self, item):
        out = BoxList3D(self.bbox3d[item], self.size3d, self.mode, torch.tensor([[0, self.bbox3d[item].shape[0]]]), self.constants)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out


def reference_namespace():
    class Stub(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return None
    sys.modules["spconv"] = Stub("spconv")
    sys.modules["spconv.utils"] = Stub("spconv.utils")
    sys.path.insert(0, REF)
    importlib.import_module("second.core.non_max_suppression.nms_gpu")
    nms_gpu = sys.modules["second.core.non_max_suppression.nms_gpu"]
    # rotate_iou_gpu_eval (nms_gpu.py:611-654) = pad to 64-thread tiles, devRotateIoUEval(query box, box, criterion) per pair, then
    # check_same_boxes.  numba's simulator runs the 64 CUDA threads of a block as Python threads and patches module globals per
    # device-function call, which races with itself ("dictionary changed size during iteration"); so the reference's device function
    # is called pair by pair from a ONE-thread simulator kernel instead, followed by the reference's check_same_boxes.
    from numba import cuda
    dev_iou = nms_gpu.devRotateIoUEval

    @cuda.jit
    def one_pair(boxes, query, iou, n, k, K, criterion):
        iou[n * K + k] = dev_iou(query[k * 5:k * 5 + 5], boxes[n * 5:n * 5 + 5], criterion)

    def eval_fn(boxes, query_boxes, criterion=-1, device_id=0):
        boxes, query_boxes = boxes.astype(np.float32), query_boxes.astype(np.float32)
        N, K = boxes.shape[0], query_boxes.shape[0]
        iou = np.zeros((N, K), np.float32)
        if N == 0 or K == 0:
            return iou
        flat = iou.reshape(-1)
        for n in range(N):
            for k in range(K):
                try:
                    one_pair[1, 1](boxes.reshape(-1), query_boxes.reshape(-1), flat, n, k, K, criterion)
                except IndexError:
                    # the reference kernel found more than 8 polygon vertices and wrote past its 16-float local array (nms_gpu.py:335-350;
                    # coinciding edges).  Only legitimate for identical rectangles, which check_same_boxes overwrites with 1 below.
                    flat[n * K + k] = np.nan
        iou = flat.reshape(N, K)
        nms_gpu.check_same_boxes(iou, boxes, query_boxes)
        bad = np.argwhere(np.isnan(iou))
        for n, k in bad[:5]:
            print("overflow pair", boxes[n], query_boxes[k], file=sys.stderr)
        assert not len(bad), "vertex overflow of the reference kernel on a pair of different boxes"
        return iou
    ns = {"torch": torch, "np": np, "math": __import__("math"), "DEBUG": False, "rotate_iou_gpu_eval": eval_fn}
    extract(f"{REF}/utils3d/geometric_torch.py", ["limit_period"], ns)
    extract(f"{REF}/second/pytorch/core/box_torch_ops.py", ["second_box_decode", "second_box_encode", "rotate_nms_3d"], ns)
    extract(f"{REF}/maskrcnn_benchmark/modeling/box_coder_3d.py", ["BoxCoder3D"], ns)
    extract(f"{REF}/utils3d/rotate_nms_3d_torch.py", ["iou_one_dim", "boxes_iou_3d"], ns)
    bnp = {"np": np}
    extract(f"{REF}/second/core/box_np_ops.py", ["corners_nd", "rotation_2d", "center_to_corner_box2d"], bnp)
    ns["box_np_ops"] = types.SimpleNamespace(**{k: bnp[k] for k in ("corners_nd", "rotation_2d", "center_to_corner_box2d")})
    ns["rotate_non_max_suppression_cpu"] = po.rotate_non_max_suppression_cpu  # spconv is absent: the oracle's restatement
    extract(f"{REF}/second/core/non_max_suppression/nms_cpu.py", ["rotate_nms_3d_cc"], ns)
    ns["rotate_nms"] = None
    extract(f"{REF}/maskrcnn_benchmark/structures/boxlist_ops_3d.py", ["boxlist_nms_3d"], ns)
    ns.update(BoxList3D=BoxList3D, cat_boxlist_3d=lambda lst, per_example=True: lst, SHOW_RPN_OUT_BEFORE_NMS=False, SHOW_NMS_OUT=False, SHOW_PRO_NUMS=False)
    extract(f"{REF}/maskrcnn_benchmark/modeling/rpn/inference_3d.py", ["RPNPostProcessor"], ns)
    return ns


def voxeliser_reference(a, b, scale, full_scale, elements_ids, elements):
    """executes the statements of SUNCGDataset.__getitem__ between :115 and :177 (augmentation switches off, as the file sets them)"""
    path = f"{REF}/data3d/suncg_utils/suncg_dataset.py"
    tree = ast.parse(open(path).read())
    body = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "__getitem__":
            body = [s for s in node.body if s.lineno >= 115 and s.end_lineno <= 177]
    assert body
    ns = {"np": np, "torch": torch, "math": __import__("math"), "a": a.copy(), "b": b.copy(), "scale": scale, "full_scale": full_scale, "zoom_rate": 0.0,
          "flip_x": False, "random_rotate": False, "distortion": False, "origin_offset": False, "norm_noise": 0.0, "bboxes_dic_i": {}, "ENABLE_POINTS_MISSED": True,
          "self": types.SimpleNamespace(elements_ids=elements_ids, elements=elements, files=["synthetic"]), "index": 0, "print": lambda *x, **k: None}
    exec(compile(ast.Module(body, []), path, "exec"), ns)
    return ns["locs"].numpy(), ns["feats"].numpy(), ns["size3d"].numpy(), ns["offset"]


def main():
    ns = reference_namespace()
    rs = np.random.RandomState(11)
    out = {"obj_scale": np.float32(OBJ_SCALE), "reg_scale": np.float32(REG_SCALE)}
    # ---- 1. rotated BEV IoU, every criterion (numba kernel under the simulator), incl. identical, touching, nested and zero-size boxes
    n = 26
    b5 = np.stack([rs.uniform(0, 6, n), rs.uniform(0, 6, n), rs.uniform(0.2, 4, n), rs.uniform(0.2, 4, n), rs.uniform(-3.2, 3.2, n)], 1).astype(np.float32)
    b5[1] = b5[0]                                   # identical (forced to 1 by check_same_boxes)
    b5[2] = [1, 1, 2, 2, 0]; b5[3] = [3, 1, 2, 2, 0]  # sharing an edge
    b5[4] = [1.5, 1, 2, 2, 0]                       # overlapping, same orientation
    b5[5] = [1, 1, 0.5, 0.5, 0.3]                   # nested in box 2
    b5[6] = [1, 1, 2, 2, 1.5707964]                 # the square of box 2 rotated by 90 degrees
    out["iou2d_boxes"] = b5
    # (queries: no axis-aligned box against itself -- with exactly coinciding edges the reference kernel finds more than 8 polygon
    #  vertices and writes past its 16-float local array, nms_gpu.py:335-350; the simulator raises where the GPU would scribble)
    q_sel = np.array([0, 1, 3, 5, 7, 8, 9, 10, 11])
    out["iou2d_query_sel"] = q_sel
    for crit in (-1, 0, 1, 2, 3):
        rows = np.array([i for i in range(n) if i not in (2, 6)])
        out[f"iou2d_c{crit}"] = ns["rotate_iou_gpu_eval"](b5[rows], b5[q_sel], criterion=crit, device_id=0)
    out["iou2d_rows"] = rows
    # ---- 2. boxes_iou_3d with and without thickness augmentation
    t7 = np.concatenate([b5[:12, :2], rs.uniform(0, 2, (12, 1)), b5[:12, 2:4], rs.uniform(0.05, 2.5, (12, 1)), b5[:12, 4:]], 1).astype(np.float32)
    a7 = np.concatenate([b5[12:, :2], rs.uniform(0, 2, (n - 12, 1)), b5[12:, 2:4], rs.uniform(0.05, 2.5, (n - 12, 1)), b5[12:, 4:]], 1).astype(np.float32)
    a7[0, [0, 1, 3, 4, 6]] = t7[0, [0, 1, 3, 4, 6]]  # one identical BEV rectangle with a different z extent
    out["iou3d_targets"], out["iou3d_anchors"] = t7, a7
    out["iou3d_plain"] = ns["boxes_iou_3d"](torch.from_numpy(t7), torch.from_numpy(a7), aug_thickness=None, criterion=-1, flag='rpn_post').numpy()
    aug = {'target_Y': 0.3, 'target_Z': 0.4, 'anchor_Y': 0.0, 'anchor_Z': 0.0}
    out["iou3d_aug"] = ns["boxes_iou_3d"](torch.from_numpy(t7), torch.from_numpy(a7), aug_thickness=aug, criterion=1, flag='rpn_label_generation').numpy()
    out["iou3d_xy"] = ns["boxes_iou_3d"](torch.from_numpy(t7), torch.from_numpy(a7), aug_thickness=None, criterion=-1, only_xy=True, flag='roi_post').numpy()
    # ---- 3. BoxCoder3D.decode: one class and two classes per anchor
    g = np.load(os.path.join(HERE, "rpn_sw4c_mid.npz"))
    anchors = np.concatenate([g[f"anchors{i}"] for i in range(int(g["n_maps"]))], 0)
    logits = np.concatenate([g[f"logits{i}"].reshape(-1, g[f"logits{i}"].shape[-1]) for i in range(int(g["n_maps"]))], 0)
    regs = np.concatenate([g[f"reg{i}"].reshape(-1, g[f"reg{i}"].shape[-1]) for i in range(int(g["n_maps"]))], 0)
    coder = ns["BoxCoder3D"]()
    out["decode_1"] = coder.decode(torch.from_numpy(regs[:, :7] * REG_SCALE), torch.from_numpy(anchors)).numpy()
    out["decode_2"] = coder.decode(torch.from_numpy(regs[:500] * REG_SCALE), torch.from_numpy(anchors[:500])).numpy()
    coder_w = ns["BoxCoder3D"](weights=(10., 10., 5., 5., 5., 5., 2.))
    out["decode_w"] = coder_w.decode(torch.from_numpy(regs[:500, :7] * REG_SCALE), torch.from_numpy(anchors[:500])).numpy()
    # ---- 4. the whole RPN post-processing of one example: sigmoid, top-k, decode, clamp, rotate_nms_3d, post top-n
    sel = np.sort(rs.permutation(anchors.shape[0])[:900])
    anc, obj, reg = anchors[sel], logits[sel, 0] * OBJ_SCALE, regs[sel, :7] * REG_SCALE
    out["rpn_sel"] = sel
    post = ns["RPNPostProcessor"](batch_size=1, fpn_pre_nms_top_n=130, fpn_post_nms_top_n=105, nms_thresh=0.1, nms_aug_thickness=[0.3, 0.3], min_size=0, box_coder=coder)
    post.eval()
    boxl = BoxList3D(torch.from_numpy(anc), None, "yx_zb", torch.tensor([[0, anc.shape[0]]]), {})
    res = post.forward_for_single_feature_map(boxl, torch.from_numpy(obj), torch.from_numpy(reg))
    out["rpn_boxes"], out["rpn_objectness"] = res[0].bbox3d.numpy(), res[0].get_field("objectness").numpy()
    # and rotate_nms_3d on its own with a threshold that suppresses more
    dec = coder.decode(torch.from_numpy(reg), torch.from_numpy(anc))
    out["nms_keep_03"] = ns["rotate_nms_3d"](dec[:150], torch.from_numpy(obj[:150]), pre_max_size=120, post_max_size=60, iou_threshold=0.3, flag='rpn_post').numpy()
    # ---- 5. voxeliser
    pts = rs.uniform(-3, 37.5, (4000, 3)).astype(np.float32)
    pts[:, 2] = rs.uniform(0, 9, 4000)
    pcl = np.concatenate([pts, rs.randn(4000, 6).astype(np.float32)], 1)
    full = np.array([2048, 2048, 512])
    locs, feats, size3d, offset = voxeliser_reference(pts, pcl, 50, full, np.arange(9), ['xyz', 'color', 'normal'])
    out.update(vox_pcl=pcl, vox_locs=locs, vox_feats=feats, vox_size3d=size3d, vox_offset=offset)
    full_small = np.array([1900, 2048, 512])  # some points fall outside (ENABLE_POINTS_MISSED): the bounds mask drops them
    locs, feats, size3d, offset = voxeliser_reference(pts, pcl, 50, full_small, np.array([0, 1, 2, 6, 7, 8]), ['xyz', 'normal'])
    out.update(vox2_locs=locs, vox2_feats=feats, vox2_size3d=size3d)
    np.savez_compressed(os.path.join(HERE, "postproc.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
