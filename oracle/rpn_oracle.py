"""ORACLE / TEST INFRASTRUCTURE ONLY -- CPU restatement of the RPN head and anchor generation on sparse maps.

  * rpn_head_forward     maskrcnn_benchmark/modeling/rpn/rpn_sparse3d.py:104-131 (RPNHead.forward): the same torch ops on the CPU
                         (F.conv2d with the 1x1 kernels on [1, C, n, 1], relu, permute, reshape)
  * grid_anchors         maskrcnn_benchmark/modeling/rpn/anchor_generator_sparse3d.py:88-104
  * generate_anchors_3d* anchor_generator_sparse3d.py:213-250
Pinned against the reference's own code (executed from /root/reference by tests/golden/make_golden_rpn.py) in
tests/test_oracle_cpu.py.
"""
import numpy as np
import torch
import torch.nn.functional as F


def rpn_head_forward(rows, w_conv, b_conv, w_cls, b_cls, w_reg, b_reg, num_anchors, seperate_rpn):
    """rows [n, C] float32; weights in Conv2d layout [out, in, 1, 1].  -> (logit [1, n, A, sep], reg [1, n, A, 7 sep]) as numpy."""
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    feature = t(rows).t().unsqueeze(0).unsqueeze(3)                                  # rpn_sparse3d.py:172-176
    h = F.relu(F.conv2d(feature, t(w_conv), t(b_conv)))                             # :110
    logit = F.conv2d(h, t(w_cls), t(b_cls)).permute(0, 2, 1, 3)                     # :112-113
    logit = logit.reshape(1, logit.shape[1], num_anchors, seperate_rpn)             # :114
    reg = F.conv2d(h, t(w_reg), t(b_reg)).permute(0, 2, 1, 3)                       # :116-117
    reg = reg.reshape(1, reg.shape[1], num_anchors, 7 * seperate_rpn)               # :119 ('box_toghter')
    return logit.numpy(), reg.numpy()


def generate_anchors_3d(size, yaws, ratios, use_yaw):
    """anchor_generator_sparse3d.py:213-250 with centroids = [[0, 0, 0]]; yaws [A, 1], ratios [A, 3] float32."""
    size = np.asarray(size, np.float32)
    zero = np.zeros(3, np.float64)  # (the reference's default `centroids` is an int64 array: the concatenation promotes to float64)
    if use_yaw:
        rows = [np.concatenate([zero, size, np.asarray(y).reshape(-1)]) for y in yaws]
    else:
        rows = [np.concatenate([zero, size * np.asarray(r, np.float32), np.zeros(1, np.float32)]) for r in ratios]
    return np.stack(rows, 0).astype(np.float32)


def grid_anchors(locations, base_anchors, voxel_scale, stride):
    """locations int64 [n, 4], base_anchors float32 [A, 7], stride float32 [3] -> [n * A, 7] (flatten order location, yaw)."""
    loc = torch.from_numpy(np.ascontiguousarray(locations, dtype=np.int64))
    centroids = (loc[:, 0:3].float() + 0) / voxel_scale * torch.from_numpy(np.asarray(stride, np.float32)).view(1, 3)   # :94
    centroids = torch.cat([centroids, torch.zeros(centroids.shape[0], 4)], 1).view(-1, 1, 7)                          # :95-96
    out = centroids + torch.from_numpy(np.asarray(base_anchors, np.float32)).view(1, -1, 7)                           # :101
    return out.reshape(-1, 7).numpy()
