"""Generates tests/golden/rpn_sw4c_mid.npz from the REFERENCE'S OWN RPN code (build container only).

The reference's maskrcnn_benchmark package does not import here (open3d, yacs and Python-3.12-incompatible imports), so the
classes / functions of the path are taken out of the reference files with `ast` AT GENERATION TIME and executed unmodified
(decorators stripped, `cfg` / debug dependencies stubbed):
  * RPNHead                                  /root/reference/maskrcnn_benchmark/modeling/rpn/rpn_sparse3d.py:81-131
  * AnchorGenerator, generate_anchors_3d*    /root/reference/maskrcnn_benchmark/modeling/rpn/anchor_generator_sparse3d.py:39-250
Inputs: the four rpn maps of tests/golden/fpn_sw4c_mid.npz (features + locations from the reference's own backbone),
deterministic head weights (fpn_util.deterministic_state), the sw4c anchor configuration.  Nothing of the reference is copied
into the repository.

  python tests/golden/make_golden_rpn.py
"""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import fpn_util  # noqa: E402

REF = "/root/reference/maskrcnn_benchmark/modeling/rpn"


def extract(path, names, ns):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in names:
            node.decorator_list = []
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def main():
    import torch.nn.functional as F
    from torch import nn
    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "math": __import__("math"),
          "OBJ_DEF": types.SimpleNamespace(check_bboxes=lambda *a, **k: None), "BoxList3D": lambda *a: a, "DEBUG": False,
          "SHOW_ANCHOR_EACH_SCALE": False, "CHECK_ANCHOR_STRIDES": False}
    extract(os.path.join(REF, "rpn_sparse3d.py"), ["RPNHead"], ns)
    extract(os.path.join(REF, "anchor_generator_sparse3d.py"),
            ["AnchorGenerator", "generate_anchors_3d", "generate_anchors_3d_ratio", "generate_anchors_3d_yaws", "examples_bidx_2_sizes"], ns)
    g = np.load(os.path.join(HERE, "fpn_sw4c_mid.npz"))
    n_rpn = int(g["n_rpn"])
    cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(SEPARATE_CLASSES=[["wall"]], SEPARATE_RPN=True))
    head = ns["RPNHead"](cfg, 128, 4)
    state = fpn_util.deterministic_state(head, seed=7)
    for k in state:  # Conv2d weights [out, in, 1, 1] come out of the generic branch (0.05 sigma); biases 0.1 sigma
        pass
    head.load_state_dict(state)
    head.eval()
    gen = ns["AnchorGenerator"](voxel_scale=50, sizes_3d=[[0.4, 1.5, 1.5], [0.2, 0.5, 3], [0.4, 1.5, 3], [0.6, 2.5, 3]], yaws=(0, -1.57, -0.785, 0.785),
                                ratios=[[1, 1, 1], [1, 2, 1], [2, 1, 1], [1.7, 1.7, 1]], use_yaws=[1, 1, 1, 1],
                                anchor_strides=[[32, 32, 32], [16, 16, 16], [32, 32, 32], [64, 64, 64]], scene_size=[40.96, 40.96, 10.24])
    feats = [torch.from_numpy(g[f"rpn{i}_features"]) for i in range(n_rpn)]
    locs = [torch.from_numpy(g[f"rpn{i}_locations"].astype(np.int64)) for i in range(n_rpn)]
    with torch.no_grad():
        logits, regs = head([f.t().unsqueeze(0).unsqueeze(3) for f in feats])  # RPNModule.forward's reshape, rpn_sparse3d.py:172-176
        anchors = gen.grid_anchors(locs)
    out = {"n_maps": n_rpn}
    for k, v in state.items():
        out["w:" + k] = v.numpy()
    for i in range(n_rpn):
        out[f"logits{i}"] = logits[i].numpy()
        out[f"reg{i}"] = regs[i].numpy()
        out[f"anchors{i}"] = anchors[i].numpy()
        out[f"cell{i}"] = gen.cell_anchors[i].numpy()
    # the ratio branch of generate_anchors_3d (use_yaw = 0), which sw4c does not take
    out["cell_ratio"] = ns["generate_anchors_3d"](np.array([0.4, 1.5, 3], np.float32), np.array([[0], [-1.57]], np.float32),
                                                  np.array([[1, 1, 1], [1, 2, 1]], np.float32), 0).float().numpy()
    np.savez_compressed(os.path.join(HERE, "rpn_sw4c_mid.npz"), **out)
    print({k: v.shape for k, v in out.items() if hasattr(v, "shape") and not k.startswith("w:")})


if __name__ == "__main__":
    main()
