"""CPU suite (no GPU): the oracle against the golden vectors generated from the reference itself,
the host-side logic, and the C-ABI surface."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import fpn_util
from detection_3d_b200 import synthetic
from oracle import fpn_oracle, ref_python
from oracle import scn_oracle as so

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
ROOT = os.path.dirname(HERE)


def test_hash_known_answers():
    g = np.load(os.path.join(GOLD, "rulebook_small.npz"))
    for (x, y, z), want in zip(g["hash_points"], g["hash_values"]):
        assert so.point_hash(int(x), int(y), int(z)) & 0xffffffff == int(want)
    # SURVEY.md A.4 (values recomputed from SCN/Metadata/32bits.h:57-66)
    assert so.point_hash(0, 0, 0) & 0xffffffff == 0x482c8e87
    assert so.point_hash(541, 541, 67) & 0xffffffff == 0xc46ab370
    assert so.point_hash(-1, 0, 0) & 0xffffffff == 0x498037e0


def test_small_rulebook_matches_reference_dump():
    g = np.load(os.path.join(GOLD, "rulebook_small.npz"))
    md = so.OracleMetadata()
    assert md.input_layer([8, 8, 8], g["coords"], 0, 4) == 6
    hdr, tab = md.input_rules()
    assert hdr.tolist() == g["input0"].tolist() == [4, 2, 7, 6]
    assert np.array_equal(tab, g["input1"])
    assert np.array_equal(md.iteration_order([8, 8, 8], 0), g["iter"])
    for k, r in enumerate(md.submanifold_rules([8, 8, 8], [3, 3, 3])):
        assert np.array_equal(r, g[f"subm{k}"]), k
    for k, r in enumerate(md.conv_rules([8, 8, 8], [4, 4, 4], [2, 2, 2], [2, 2, 2])):
        assert np.array_equal(r, g[f"conv{k}"]), k
    assert np.array_equal(md.spatial_locations([4, 4, 4]), g["loc4"])
    for k, r in enumerate(md.conv_rules([4, 4, 4], [4, 4, 1], [1, 1, 4], [1, 1, 1])):
        assert np.array_equal(r, g[f"pro{k}"]), k


@pytest.mark.parametrize("name,bld,full,nl,pro", [
    ("mini4", dict(nx=60, ny=56, nz=24, n_walls=3, seed=3), [64, 64, 32], 4, (1, 2)),
    ("sw4c_mid", dict(nx=300, ny=280, nz=40, n_walls=5, seed=5), [2048, 2048, 512], 9, (4, 5, 6)),
    ("b470", dict(), [2048, 2048, 512], 9, (4, 5, 6)),
])
def test_rulebook_digests_match_reference(name, bld, full, nl, pro):
    gold = json.load(open(os.path.join(GOLD, "rulebooks.json")))[name]
    coords = synthetic.building_coords(**bld)
    assert coords.shape[0] == gold["n_input_rows"]
    md = so.OracleMetadata()
    md.input_layer(full, coords, 0, 4)
    got = fpn_util.metadata_digests(md, full, nl, pro)
    got["input_rules"] = fpn_util.rulebook_digest(md.input_rules())
    got["n_input_rows"] = int(coords.shape[0])
    assert got == gold


def test_b470_level_counts():
    # BASELINE.md / SURVEY.md section 8: per-level active sites of the benchmark building
    gold = json.load(open(os.path.join(GOLD, "rulebooks.json")))["b470"]
    assert [gold[f"n{l}"] for l in range(9)] == [1155656, 283586, 68672, 16416, 3752, 786, 162, 25, 9]
    assert gold["n_input_rows"] == 1177224


@pytest.mark.parametrize("name,cfgf,bld", [
    ("mini4", fpn_util.mini4_config, dict(nx=60, ny=56, nz=24, n_walls=3, seed=3)),
])
def test_port_fpn_forward_matches_reference_golden(name, cfgf, bld):
    import detection_3d_b200.sparseconvnet.fpn as fpn
    cfg = cfgf()
    g = np.load(os.path.join(GOLD, f"fpn_{name}.npz"))
    net = fpn.FPN_Net(**cfg)
    state = fpn_util.deterministic_state(net, seed=1)
    coords = synthetic.building_coords(**bld)
    rpn, roi, macs = fpn_oracle.run_fpn_port(cfg, state, coords, fpn_util.features_for(coords))
    assert macs == float(g["macs"])
    for tag, maps in (("rpn", rpn), ("roi", roi)):
        assert len(maps) == int(g[f"n_{tag}"])
        for i, m in enumerate(maps):
            assert np.array_equal(m["locations"], g[f"{tag}{i}_locations"])
            ref = g[f"{tag}{i}_features"]
            # fp32 tolerance: naive C dot products vs the reference's sgemm
            np.testing.assert_allclose(m["features"], ref, rtol=2e-4, atol=2e-5 * np.abs(ref).max())


def test_state_dict_layout_matches_reference_checkpoint_format():
    import detection_3d_b200.sparseconvnet as scn
    net = scn.FPN_Net(**scn.sw4c_fpn432_config())
    sd = net.state_dict()
    assert len(sd) == 197 and sum(p.numel() for p in net.parameters()) == 21147124
    # a few keys / shapes read off the reference module tree (fpn_net.py:40-135)
    assert tuple(sd["layers_in.1.weight"].shape) == (27, 1, 9, 32)
    assert tuple(sd["convs_pro2d.0.weight"].shape) == (32, 1, 128, 128)
    assert tuple(sd["m_downs.1.0.1.weight"].shape) == (8, 1, 32, 64)
    assert tuple(sd["m_downs.0.0.1.3.weight"].shape) == (27, 1, 32, 32)
    assert tuple(sd["m_ups.7.1.weight"].shape) == (8, 1, 128, 128)
    assert tuple(sd["m_shortcuts.8.weight"].shape) == (1, 1, 256, 128)
    if ref_python.available():
        ref = ref_python.load_reference_package()
        args, kw = fpn_util.ref_ctor_args(scn.sw4c_fpn432_config())
        rsd = ref.FPN_Net(*args, **kw).state_dict()
        assert list(rsd.keys()) == list(sd.keys())
        assert all(tuple(rsd[k].shape) == tuple(sd[k].shape) for k in sd)


def test_c_abi_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "scn_b200.h")).read()
    declared = set(re.findall(r"\b(scn_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 25
    so_path = os.path.join(ROOT, "detection_3d_b200", "libscn_b200.so")
    assert os.path.exists(so_path), "run __graft_entry__.build() first"
    handle = ctypes.CDLL(so_path)
    missing = [s for s in declared if not hasattr(handle, s)]
    assert not missing, missing
    from detection_3d_b200 import _lib
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert handle.scn_n_rulebook_bits() == 32


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import detection_3d_b200.sparseconvnet as scn
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        scn.Metadata(3)
    net = scn.FPN_Net(**fpn_util.mini4_config())
    c = torch.zeros(4, 4, dtype=torch.int64)
    with pytest.raises(RuntimeError):
        net([c, torch.zeros(4, 9)])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "detection_3d_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"(import|from)\s+oracle|oracle[./](scn_oracle|_ref|_build)|liboracle", txt), os.path.join(dp, f)


@pytest.mark.skipif(not (so.have_ref() and ref_python.available()), reason="reference build only exists in the build container")
class TestAgainstCompiledReference:
    """Pins the C restatement (and the restated layer sequence) against the reference's own code."""

    @pytest.mark.parametrize("seed", range(6))
    def test_random_rulebooks(self, seed):
        rs = np.random.RandomState(seed)
        ext = [rs.randint(6, 40) for _ in range(3)]
        n = rs.randint(1, 4000)
        c = np.stack([rs.randint(0, e, n) for e in ext], 1).astype(np.int64)
        bs = rs.randint(1, 4)
        if seed % 2:
            c = np.concatenate([c, np.sort(rs.randint(0, bs, (n, 1)), 0)], 1)
        full = [64, 64, 64]
        mode = [4, 3, 4, 1, 2, 4][seed]
        O, R = so.OracleMetadata(), so.RefMetadata()
        hint = bs if seed % 2 else 0
        assert O.input_layer(full, c, hint, mode) == R.input_layer(full, c, hint, mode)
        for a, b in zip(O.input_rules(), R.input_rules()):
            assert np.array_equal(a, b)
        d1 = fpn_util.metadata_digests(O, full, 4, (1, 2))
        d2 = fpn_util.metadata_digests(R, full, 4, (1, 2))
        assert d1 == d2
        for f, s in (([3, 3, 3], [2, 2, 2]), ([4, 4, 4], [2, 2, 2]), ([3, 1, 2], [2, 1, 2])):
            O2, R2 = so.OracleMetadata(), so.RefMetadata()
            O2.input_layer(full, c, hint, 4), R2.input_layer(full, c, hint, 4)
            out = [(full[d] - f[d]) // s[d] + 1 for d in range(3)]
            for a, b in zip(O2.conv_rules(full, out, f, s), R2.conv_rules(full, out, f, s)):
                assert np.array_equal(a, b)
            assert np.array_equal(O2.spatial_locations(out), R2.spatial_locations(out))

    def test_ref_driver_reproduces_reference_fpn_bitwise(self):
        import detection_3d_b200.sparseconvnet.fpn as fpn
        cfg = fpn_util.mini4_config()
        g = np.load(os.path.join(GOLD, "fpn_mini4.npz"))
        state = fpn_util.deterministic_state(fpn.FPN_Net(**cfg), seed=1)
        coords = synthetic.building_coords(nx=60, ny=56, nz=24, n_walls=3, seed=3)
        rpn, roi, macs = fpn_oracle.run_fpn_ref(cfg, state, coords, fpn_util.features_for(coords))
        assert macs == float(g["macs"])
        for tag, maps in (("rpn", rpn), ("roi", roi)):
            for i, m in enumerate(maps):
                assert np.array_equal(m["features"], g[f"{tag}{i}_features"])
                assert np.array_equal(m["locations"], g[f"{tag}{i}_locations"])

    def test_compute_kernels_of_port_vs_reference_extension(self):
        import torch
        SCN = ref_python.load_scn_native()
        rs = np.random.RandomState(0)
        x = rs.randn(500, 24).astype(np.float32)
        gam, bet = (1 + 0.1 * rs.randn(24)).astype(np.float32), (0.1 * rs.randn(24)).astype(np.float32)
        rm, rv = np.zeros(24, np.float32), np.ones(24, np.float32)
        out, sm, si = so.o_bn_forward(x, gam, bet, rm, rv, 1e-4, 0.9, True, 0.0)
        t = torch.from_numpy
        o2, sm2, si2, rm2, rv2 = torch.empty(0), torch.empty(24), torch.empty(24), torch.zeros(24), torch.ones(24)
        SCN.BatchNormalization_updateOutput(t(x), o2, sm2, si2, rm2, rv2, t(gam), t(bet), 1e-4, 0.9, True, 0.0)
        np.testing.assert_allclose(out, o2.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(rv, rv2.numpy(), rtol=1e-5)
        dy = rs.randn(500, 24).astype(np.float32)
        din, dw, db = so.o_bn_backward(x, out, dy, sm, si, gam, 0.0)
        din2, dw2, db2 = torch.empty(0), torch.zeros(24), torch.zeros(24)
        SCN.BatchNormalization_backward(t(x), din2, o2, t(dy.copy()), sm2, si2, rm2, rv2, t(gam), t(bet), dw2, db2, 0.0)
        np.testing.assert_allclose(din, din2.numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(dw, dw2.numpy(), rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(db, db2.numpy(), rtol=1e-4, atol=1e-4)


def test_program_recorder_hoists_lateral_convolutions():
    """Host logic of the recorded programs (sparseconvnet/program.py): a 1x1x1 SubmanifoldConvolution is moved right behind
    the op that produces its input, everything else keeps its order (no GPU needed: pure list manipulation)."""
    from detection_3d_b200.sparseconvnet import program
    sz, f3, f1 = [64, 64, 32], [3, 3, 3], [1, 1, 1]
    ops = [
        (0, [0] + sz + [4, 0, 9], []),                       # input -> r0
        (1, [0, 1] + sz + f3 + [0, -1, 9, 32], []),          # subm 3^3 r0 -> r1
        (4, [1, 2, 32, 1, 2, -1, -1, 2], [1e-4, 0.9, 0.0]),  # bn r1 -> r2
        (1, [2, 3] + sz + f3 + [3, -1, 32, 32], []),         # subm 3^3 r2 -> r3
        (2, [3, 4] + sz + [32, 32, 16] + [2, 2, 2] + [2, 2, 2] + [4, -1, 32, 64], []),  # conv r3 -> r4
        (1, [1, 5] + sz + f1 + [5, -1, 32, 128], []),        # lateral 1x1x1 on r1 -> r5 (recorded late)
    ]
    out = program._hoist_pointwise(ops)
    assert len(out) == len(ops)
    kinds = [(k, i[0], i[1]) for k, i, _ in out]
    assert kinds[1] == (1, 0, 1) and kinds[2] == (1, 1, 5), kinds  # the lateral now follows its producer
    rest = [o for o in out if o is not ops[5]]
    assert rest == ops[:5]                                          # relative order of the others unchanged


# ------------------------------------------------------------------ next rows: RPN head / anchors, SparseToDense
def test_rpn_oracle_vs_reference_golden():
    """oracle/rpn_oracle.py against the outputs of the reference's own RPNHead / AnchorGenerator code
    (tests/golden/rpn_sw4c_mid.npz, generated by tests/golden/make_golden_rpn.py from /root/reference)."""
    from oracle import rpn_oracle as ro
    g = np.load(os.path.join(GOLD, "rpn_sw4c_mid.npz"))
    f = np.load(os.path.join(GOLD, "fpn_sw4c_mid.npz"))
    sizes = [[0.4, 1.5, 1.5], [0.2, 0.5, 3], [0.4, 1.5, 3], [0.6, 2.5, 3]]
    yaws = np.array((0, -1.57, -0.785, 0.785), np.float32).reshape(-1, 1)
    ratios = np.array([[1, 1, 1], [1, 2, 1], [2, 1, 1], [1.7, 1.7, 1]], np.float32)
    strides = [32, 16, 32, 64]
    for i in range(int(g["n_maps"])):
        lg, rg = ro.rpn_head_forward(f[f"rpn{i}_features"], g["w:conv.weight"], g["w:conv.bias"], g["w:cls_logits.weight"], g["w:cls_logits.bias"],
                                     g["w:bbox_pred.weight"], g["w:bbox_pred.bias"], 4, 2)
        assert lg.shape == g[f"logits{i}"].shape and rg.shape == g[f"reg{i}"].shape
        np.testing.assert_allclose(lg, g[f"logits{i}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(rg, g[f"reg{i}"], rtol=1e-5, atol=1e-6)
        cell = ro.generate_anchors_3d(sizes[i], yaws, ratios, 1)
        assert np.array_equal(cell, g[f"cell{i}"])
        a = ro.grid_anchors(f[f"rpn{i}_locations"], cell, 50, [strides[i]] * 3)
        assert np.array_equal(a, g[f"anchors{i}"])  # three separately rounded fp32 operations: bit exact
    assert np.array_equal(ro.generate_anchors_3d([0.4, 1.5, 3], np.array([[0], [-1.57]], np.float32), np.array([[1, 1, 1], [1, 2, 1]], np.float32), 0), g["cell_ratio"])


def test_rpn_module_mirrors_reference_parameters():
    """detection_3d_b200.rpn.RPNHead keeps the reference's parameter names / shapes (checkpoint compatibility) and the anchor
    generator its base anchors (host logic only: no GPU needed)."""
    from detection_3d_b200 import rpn
    g = np.load(os.path.join(GOLD, "rpn_sw4c_mid.npz"))
    head = rpn.RPNHead(128, 4, 2)
    want = {k[2:]: g[k].shape for k in g.files if k.startswith("w:")}
    assert {k: tuple(v.shape) for k, v in head.state_dict().items()} == want
    gen = rpn.sw4c_anchor_generator()
    assert gen.num_anchors_per_location() == 4
    for i in range(4):
        assert np.array_equal(gen.cell_anchors[i].numpy(), g[f"cell{i}"])
    import torch
    assert rpn.examples_bidx_2_sizes(torch.tensor([0, 0, 0, 1, 1, 3])).tolist() == [[0, 3], [3, 5], [5, 5], [5, 6]]


def test_sparse_to_dense_oracle_vs_reference_golden():
    """C port of the SparseToDense rules + passes against the reference package's scn.SparseToDense (tests/golden/sparse_to_dense.npz)."""
    g = np.load(os.path.join(GOLD, "sparse_to_dense.npz"))
    O = so.OracleMetadata()
    sz = [16, 16, 8]
    n = O.input_layer(sz, g["coords"], 0, 4)
    assert np.array_equal(O.spatial_locations(sz), g["locations"]) and n == g["rows"].shape[0]
    hdr, tab = O.input_rules()
    rows = so.o_input_layer_forward(g["feats"], hdr, tab)
    np.testing.assert_allclose(rows, g["rows"], rtol=1e-6, atol=1e-7)
    rules = O.sparse_to_dense_rules(sz)
    assert len(rules) == 2 and sum(r.shape[0] for r in rules) == n
    dense = so.o_sparse_to_dense_forward(g["rows"], rules, sz)
    assert np.array_equal(dense, g["dense"])
    d_rows = so.o_sparse_to_dense_backward(g["w"], rules, n)
    np.testing.assert_allclose(so.o_input_layer_backward(d_rows, hdr, tab), g["grad_feats"], rtol=1e-6, atol=1e-7)
    x_size, y_size, z_size = (g["locations"][:, :3].max(0) + 1).tolist()
    assert np.array_equal(dense[:, :, :x_size, :y_size, :z_size], g["crop"])
    if so.have_ref():
        R = so.RefMetadata()
        R.input_layer(sz, g["coords"], 0, 4)
        for a, b in zip(rules, R.sparse_to_dense_rules(sz)):
            assert np.array_equal(a, b)


# ------------------------------------------------------------------ RPN post-processing / voxeliser
def test_postproc_oracle_vs_reference_golden():
    """oracle/postproc_oracle.py against tests/golden/postproc.npz = outputs of the reference's own decode / rotated IoU / 3-D IoU /
    rotate_nms_3d / RPNPostProcessor / voxeliser code (tests/golden/make_golden_postproc.py).  Float outputs: the reference's numba
    kernel mixes float32 arrays with double scalars, the restatement is float32 throughout -> 2e-5 absolute on IoUs in [0, 1]."""
    from oracle import postproc_oracle as po
    g = np.load(os.path.join(GOLD, "postproc.npz"))
    OS, RS = np.float32(g["obj_scale"]), np.float32(g["reg_scale"])
    b5, rows, qs = g["iou2d_boxes"], g["iou2d_rows"], g["iou2d_query_sel"]
    for crit in (-1, 0, 1, 2, 3):
        np.testing.assert_allclose(po.rotated_iou_2d(b5[rows], b5[qs], crit), g[f"iou2d_c{crit}"], rtol=1e-4, atol=2e-5)
    t7, a7 = g["iou3d_targets"], g["iou3d_anchors"]
    np.testing.assert_allclose(po.boxes_iou_3d(t7, a7), g["iou3d_plain"], rtol=1e-4, atol=2e-5)
    aug = {'target_Y': 0.3, 'target_Z': 0.4, 'anchor_Y': 0.0, 'anchor_Z': 0.0}
    np.testing.assert_allclose(po.boxes_iou_3d(t7, a7, aug, criterion=1), g["iou3d_aug"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(po.boxes_iou_3d(t7, a7, only_xy=True), g["iou3d_xy"], rtol=1e-4, atol=2e-5)
    r = np.load(os.path.join(GOLD, "rpn_sw4c_mid.npz"))
    nm = int(r["n_maps"])
    anchors = np.concatenate([r[f"anchors{i}"] for i in range(nm)], 0)
    logits = np.concatenate([r[f"logits{i}"].reshape(-1, r[f"logits{i}"].shape[-1]) for i in range(nm)], 0)
    regs = np.concatenate([r[f"reg{i}"].reshape(-1, r[f"reg{i}"].shape[-1]) for i in range(nm)], 0)
    np.testing.assert_allclose(po.box_decode(regs[:, :7] * RS, anchors), g["decode_1"], rtol=1e-6, atol=1e-6)
    two = np.concatenate([po.box_decode(regs[:500, 7 * c:7 * c + 7] * RS, anchors[:500]) for c in range(2)], 1)
    np.testing.assert_allclose(two, g["decode_2"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(po.box_decode(regs[:500, :7] * RS, anchors[:500], (10., 10., 5., 5., 5., 5., 2.)), g["decode_w"], rtol=1e-6, atol=1e-6)
    sel = g["rpn_sel"]
    boxes, obj = po.rpn_post_process(anchors[sel], logits[sel, 0] * OS, regs[sel, :7] * RS, 130, 105, 0.1, (0.3, 0.3))
    assert boxes.shape == g["rpn_boxes"].shape  # the same boxes survive, in the same order
    np.testing.assert_allclose(boxes, g["rpn_boxes"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(obj, g["rpn_objectness"], rtol=1e-6, atol=1e-7)
    dec = po.box_decode(regs[sel, :7] * RS, anchors[sel])
    assert np.array_equal(po.rotate_nms_3d(dec[:150], (logits[sel, 0] * OS)[:150], 120, 60, 0.3), g["nms_keep_03"])
    pcl = g["vox_pcl"]
    locs, feats, size3d = po.voxelize(pcl[:, :3], pcl, 50, [2048, 2048, 512])
    assert np.array_equal(locs, g["vox_locs"]) and np.array_equal(feats, g["vox_feats"])
    np.testing.assert_allclose(size3d, g["vox_size3d"], rtol=1e-6)
    locs, feats, size3d = po.voxelize(pcl[:, :3], pcl[:, [0, 1, 2, 6, 7, 8]], 50, [1900, 2048, 512])
    assert np.array_equal(locs, g["vox2_locs"]) and np.array_equal(feats, g["vox2_feats"])
    assert locs.shape[0] < pcl.shape[0]  # points outside the full scale were dropped
