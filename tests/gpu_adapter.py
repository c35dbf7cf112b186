"""Gives the product's Metadata the interface of the CPU checkers (OracleMetadata / RefMetadata)
so one set of comparison helpers serves all three.  Everything goes through the C-ABI."""
import numpy as np
import torch

import detection_3d_b200.sparseconvnet as scn


class GpuMetadata:
    kind = "cuda"

    def __init__(self):
        self.m = scn.Metadata(3)

    def input_layer(self, spatial, coords, batch_size=0, mode=4, feats=None, coords_on_device=False):
        c = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.int64))
        if coords_on_device:
            c = c.cuda()
        f = torch.zeros(c.size(0), 1, device="cuda") if feats is None else torch.from_numpy(np.ascontiguousarray(feats, np.float32)).cuda()
        out = torch.empty(0, device="cuda")
        scn.SCN.InputLayer_updateOutput(self.m, torch.LongTensor(list(spatial)), c, f, out, batch_size, mode)
        self.input_features = out
        return out.size(0)

    def input_rules(self):
        return [t.numpy() for t in self.m.inputLayerRuleBook()]

    def nactive(self, sz):
        return self.m.getNActive(torch.LongTensor(list(sz)))

    def batch_size(self, sz):
        return 1  # iteration_order() already concatenates the batch items

    def spatial_locations(self, sz):
        return self.m.getSpatialLocations(torch.LongTensor(list(sz))).numpy()

    def iteration_order(self, sz, b=0):
        return self.m.iterationOrder(torch.LongTensor(list(sz))).numpy()

    def submanifold_rules(self, sz, f):
        return [t.numpy() for t in self.m.submanifoldRuleBook(list(sz), list(f))]

    def conv_rules(self, in_sz, out_sz, f, s):
        return [t.numpy() for t in self.m.ruleBook(list(in_sz), list(out_sz), list(f), list(s))]

    def sparse_to_dense_rules(self, sz):
        return [t.numpy() for t in self.m.sparseToDenseRuleBook(list(sz))]
