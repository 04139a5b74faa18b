"""CPU: the fused plan (BN folding, border-class bias tables, residual / upsample fusion, FC-as-conv)
reproduces the node-by-node fp32 oracle, and the ONNX reader/writer round-trips."""
import numpy as np
import pytest
import torch

import plan_sim
from oracle import restate
from oracle.torch_exec import TorchGraph
from scrfd_arcface_facerecognition_b200 import archs, graph, onnx_wire
from tests.golden import inputs


def _check(key, hw, n=1, tol=2e-5):
    g = archs.build_arch(key)
    plan = graph.compile_graph(g, hw)
    det = key.startswith("scrfd")
    x = restate.blob_from_bgr(np.stack([inputs.frame(50 + i, hw[0], hw[1]) for i in range(n)]),
                              1 / 128 if det else 1 / 127.5, 127.5)
    tg = TorchGraph(g)
    ref = tg.run(x)
    out = plan_sim.run_plan(plan, x)
    for name in tg.output_names:
        r = ref[name]
        o = out[name]
        np.testing.assert_allclose(o.reshape(r.shape), r, rtol=0, atol=tol * max(1.0, np.abs(r).max()))
    return plan


def test_scrfd_500m_plan_matches_oracle():
    plan = _check("scrfd_500m", (160, 160))
    kinds = {o.kind for o in plan.ops}
    assert kinds == {"im2col", "conv", "dwconv"}
    assert sum(o.attrs.get("sig_hi", 0) == 2 for o in plan.ops) == 3            # one merged head conv per level


def test_scrfd_2p5g_plan_matches_oracle():
    plan = _check("scrfd_2.5g", (160, 192))
    assert sum(o.res_mode == 2 for o in plan.ops) == 2            # both top-down upsample-adds are fused
    # the three avg_down shortcuts (AveragePool 2x2 / s2 -> Conv1x1) are 2x2 / stride 2 convolutions, no pooling pass
    assert sum(o.kind == "pool" for o in plan.ops) == 0
    assert sum(o.kind == "conv" and (o.attrs["kh"], o.attrs["stride"]) == (2, 2) for o in plan.ops) == 3
    assert sum(o.attrs.get("pool", 0) for o in plan.ops) == 1      # the stem max-pool rides in its convolution's epilogue


def test_avgpool_fold_is_optional(monkeypatch):
    """B2F_FOLD_AVGPOOL=0 keeps AveragePool + Conv1x1 as two ops; both plans reproduce the oracle.  (Detector inputs are
    multiples of 32, so the pooled maps are always even and the fold's odd-size guard never triggers there.)"""
    monkeypatch.setenv("B2F_FOLD_AVGPOOL", "0")
    plan = _check("scrfd_2.5g", (160, 192))
    assert sum(o.kind == "pool" for o in plan.ops) == 3


def test_arcface_mbf_plan_matches_oracle():
    _check("arcface_mbf", (112, 112), n=2)


@pytest.mark.slow
def test_arcface_r50_plan_matches_oracle_and_uses_border_tables():
    plan = _check("arcface_r50", (112, 112), n=1)
    assert sum(o.attrs.get("bias_classes") == 9 for o in plan.ops) == 24        # one pre-BN per IBasicBlock
    # 55 fused launches, minus the four projection shortcuts that ride as extra K of their block's stride-2 conv
    assert all(o.kind in ("conv", "im2col") for o in plan.ops) and len(plan.ops) == 51
    assert sum(1 for o in plan.ops if o.sc_src) == 4
    assert abs(plan.conv_flops() / 1e9 - 12.62) < 0.01                            # SURVEY 8d per-face figure
    g = archs.build_arch("arcface_r50")
    assert len(graph.compile_graph(g, (112, 112), fuse_shortcuts=False).ops) == 55


def test_unfused_variants_still_match_oracle():
    g = archs.build_arch("scrfd_500m")
    x = restate.blob_from_bgr(inputs.frame(50, 96, 128)[None], 1 / 128, 127.5)
    ref = TorchGraph(g).run(x)
    plan = graph.compile_graph(g, (96, 128), stem_im2col=False, merge_heads=False)
    assert {o.kind for o in plan.ops} == {"stem", "conv", "dwconv"}
    out = plan_sim.run_plan(plan, x)
    for name, r in ref.items():
        np.testing.assert_allclose(out[name].reshape(r.shape), r, rtol=0, atol=2e-5 * max(1.0, np.abs(r).max()))


def test_scrfd_10g_flops():
    plan = graph.compile_graph(archs.build_arch("scrfd_10g"), (640, 640))
    assert [t.cp for t in plan.tensors.values() if t.c in (28, 56, 80, 88, 224)] and \
        {t.c: t.cp for t in plan.tensors.values()}.get(80) == 96
    assert abs(plan.conv_flops() / 1e9 - 26.68) < 0.01                            # SURVEY 8d per-frame figure
    assert [o[2] for o in plan.outputs] == [2, 2, 2, 8, 8, 8, 20, 20, 20]


def test_fp16_activations_keep_embedding_cosine():
    g = archs.build_arch("arcface_mbf")
    plan = graph.compile_graph(g, (112, 112))
    x = restate.blob_from_bgr(np.stack([inputs.smooth_frame(60 + i, 112, 112) for i in range(2)]), 1 / 127.5, 127.5)
    ref = TorchGraph(g).run(x)
    ref = next(iter(ref.values()))
    for q in (torch.float16, torch.bfloat16):
        out = next(iter(plan_sim.run_plan(plan, x, quantize=q).values())).reshape(2, -1)
        cos = (out * ref).sum(1) / np.linalg.norm(out, axis=1) / np.linalg.norm(ref, axis=1)
        assert cos.min() >= 0.999


def test_onnx_roundtrip_and_hand_built_model(tmp_path):
    g = archs.build_arch("scrfd_500m")
    path = tmp_path / "m.onnx"
    onnx_wire.save_model(g, str(path))
    g2 = onnx_wire.load_model(str(path))
    assert [n.op_type for n in g2.nodes] == [n.op_type for n in g.nodes]
    assert [n.inputs for n in g2.nodes] == [n.inputs for n in g.nodes]
    assert all(np.array_equal(g.initializers[k], g2.initializers[k]) for k in g.initializers)
    assert [(v.name, v.shape) for v in g2.outputs] == [(v.name, v.shape) for v in g.outputs]
    for a, b in zip(g.nodes, g2.nodes):
        for k, v in a.attrs.items():
            if isinstance(v, float):
                assert abs(v - b.attrs[k]) < 1e-6
            else:
                assert v == b.attrs[k]
    # a file that is not a model fails loudly
    bad = tmp_path / "bad.onnx"
    bad.write_bytes(b"\x08\x07")
    with pytest.raises(ValueError):
        onnx_wire.load_model(str(bad))


def test_known_weight_names_resolve_to_architectures():
    assert archs.arch_for_path("./weights/det_10g.onnx") == "scrfd_10g"
    assert archs.arch_for_path("/x/w600k_r50.onnx") == "arcface_r50"
    assert archs.arch_for_path("other.onnx") is None
    assert abs(archs.count_macs(archs.build_arch("scrfd_10g"), (1, 3, 480, 640)) / 1e9 - 10.006) < 0.01


def test_stem_form_weights_equal_patch_gemm():
    """engine.stem8_weights: the first convolution's 1x1-over-patches weights regrouped as [10 taps][cout][8 channels]
    compute the same 3x3 convolution (CPU, fp32): slot t = tap (ky*3+kx), 3 of 8 channels used, slot 9 zero"""
    from scrfd_arcface_facerecognition_b200.engine import stem8_weights
    plan = graph.compile_graph(archs.build_arch("arcface_mbf"), (112, 112))
    assert plan.ops[0].kind == "im2col" and plan.ops[1].kind == "conv" and plan.ops[1].attrs["kh"] == 1
    w = torch.from_numpy(np.ascontiguousarray(plan.ops[1].arrays["weight"])).float()      # [1][cout_p][32]
    w8 = stem8_weights(w)
    cout_p = w.shape[1]
    assert w8.shape == (10, cout_p, 8) and (w8[9] == 0).all() and (w8[:, :, 3:] == 0).all()
    g = torch.Generator().manual_seed(0)
    x = torch.randn((2, 3, 20, 24), generator=g)
    stride = plan.ops[0].attrs["stride"]
    # patches: k = tap * 3 + channel, then the 1x1 GEMM
    cols = torch.nn.functional.unfold(x, 3, padding=1, stride=stride)                     # [n, c*9, L] with k = c*9 + tap
    cols = cols.reshape(2, 3, 9, -1).permute(0, 2, 1, 3).reshape(2, 27, -1)               # k = tap*3 + c
    want = torch.einsum("ok,nkl->nol", w[0, :, :27], cols)
    # stem form: a 3x3 convolution over the 8-channel image with the regrouped weights
    w33 = w8[:9].reshape(3, 3, cout_p, 8).permute(2, 3, 0, 1)                             # [cout][8][ky][kx]
    x8 = torch.zeros((2, 8, 20, 24))
    x8[:, :3] = x
    got = torch.nn.functional.conv2d(x8, w33, None, stride, 1).reshape(2, cout_p, -1)
    assert torch.allclose(got, want, atol=1e-5)


def test_missing_model_file_raises_unless_synthetic_weights_are_opted_in(monkeypatch):
    """reference models/scrfd.py:59-68: a wrong weights path fails at construction, it does not run on random weights"""
    import pytest
    from scrfd_arcface_facerecognition_b200.scrfd import load_graph
    monkeypatch.delenv("B2F_SYNTHETIC_WEIGHTS", raising=False)
    with pytest.raises(FileNotFoundError):
        load_graph("weights/det_10g.onnx")
    with pytest.raises(FileNotFoundError):
        load_graph("weights/not_a_model.onnx")
    monkeypatch.setenv("B2F_SYNTHETIC_WEIGHTS", "1")
    with pytest.warns(RuntimeWarning, match="RANDOM weights"):
        g = load_graph("weights/det_500m.onnx")
    assert len(g.outputs) == 9
    with pytest.raises(FileNotFoundError):
        load_graph("weights/not_a_model.onnx")


def _with_export_glue(g):
    """The same network written the way exporters write it for dynamic input sizes: every Resize computes its `sizes`
    from Shape / Slice / Gather / Cast / Mul / Floor / Unsqueeze / Concat arithmetic, every head Reshape gets its shape
    from a Concat of Constant nodes, and small initializers become Constant nodes."""
    from scrfd_arcface_facerecognition_b200.onnx_wire import Graph, Node
    nodes, init = [], dict(g.initializers)
    k = [0]

    def name(p):
        k[0] += 1
        return f"glue_{p}_{k[0]}"

    def const(arr):
        out = name("const")
        nodes.append(Node("Constant", [], [out], {"value": np.asarray(arr)}))
        return out
    for ri, n in enumerate(g.nodes):
        if n.op_type == "Resize":
            x = n.inputs[0]
            shp = name("shape")
            nodes.append(Node("Shape", [x], [shp]))
            if ri % 2 == 0:                                  # Slice + integer Mul form
                hw, nc, hw2, sizes = name("hw"), name("nc"), name("hw2"), name("sizes")
                nodes.append(Node("Slice", [shp, const(np.asarray([2], np.int64)), const(np.asarray([4], np.int64)),
                                            const(np.asarray([0], np.int64))], [hw]))
                nodes.append(Node("Slice", [shp, const(np.asarray([0], np.int64)), const(np.asarray([2], np.int64)),
                                            const(np.asarray([0], np.int64))], [nc]))
                nodes.append(Node("Mul", [hw, const(np.asarray([2, 2], np.int64))], [hw2]))
                nodes.append(Node("Concat", [nc, hw2], [sizes], {"axis": 0}))
            else:                                            # Gather + Cast + float Mul + Floor + Unsqueeze form
                parts = []
                for axis, mul in ((0, 1.0), (1, 1.0), (2, 2.0), (3, 2.0)):
                    d, f, m, fl, i64, u = (name(s) for s in ("dim", "f", "mul", "floor", "i64", "unsq"))
                    nodes.append(Node("Gather", [shp, const(np.asarray(axis, np.int64))], [d], {"axis": 0}))
                    nodes.append(Node("Cast", [d], [f], {"to": 1}))
                    nodes.append(Node("Mul", [f, const(np.asarray(mul, np.float32))], [m]))
                    nodes.append(Node("Floor", [m], [fl]))
                    nodes.append(Node("Cast", [fl], [i64], {"to": 7}))
                    nodes.append(Node("Unsqueeze", [i64], [u], {"axes": [0]}))
                    parts.append(u)
                sizes = name("sizes")
                nodes.append(Node("Concat", parts, [sizes], {"axis": 0}))
            nodes.append(Node("Resize", [x, "", "", sizes], n.outputs, dict(n.attrs)))
            for i in n.inputs[1:]:
                init.pop(i, None)
        elif n.op_type == "Reshape":
            tgt = init.pop(n.inputs[1])
            shape = name("shape")
            nodes.append(Node("Concat", [const(tgt[:1]), const(tgt[1:])], [shape], {"axis": 0}))
            nodes.append(Node("Reshape", [n.inputs[0], shape], n.outputs, dict(n.attrs)))
        else:
            nodes.append(n)
    return Graph(nodes, init, g.inputs, g.outputs, g.name)


def test_exporter_shape_glue_is_constant_folded(tmp_path):
    """reference models/scrfd.py:52-68 loads whatever the exporter wrote; onnxruntime folds the shape arithmetic at
    session creation.  Here: a detector with Shape/Slice/Gather/Cast/Mul/Floor/Unsqueeze/Concat glue in front of every
    Resize, Concat-of-Constant shapes on the head Reshapes and unfused BatchNormalization, through the protobuf
    writer / reader, the folding pass and the compiler -- against the node-by-node oracle on the ORIGINAL graph."""
    from scrfd_arcface_facerecognition_b200.onnx_fold import fold_shape_glue
    hw = (160, 192)
    g = archs.build_arch("scrfd_2.5g")
    glued = _with_export_glue(g)
    assert sum(n.op_type == "Shape" for n in glued.nodes) == 2 and sum(n.op_type == "Constant" for n in glued.nodes) > 20
    path = tmp_path / "det_glue.onnx"
    onnx_wire.save_model(glued, str(path))
    loaded = onnx_wire.load_model(str(path))
    assert [n.op_type for n in loaded.nodes] == [n.op_type for n in glued.nodes]
    folded = fold_shape_glue(loaded, hw)
    left = {n.op_type for n in folded.nodes}
    assert not left & {"Shape", "Gather", "Slice", "Concat", "Cast", "Floor", "Unsqueeze", "Constant"}, left
    for n in folded.nodes:
        if n.op_type == "Resize":
            sizes = folded.initializers[n.inputs[3]]
            assert sizes.dtype == np.int64 and sizes.shape == (4,) and sizes[0] == 1
    x = restate.blob_from_bgr(inputs.frame(50, hw[0], hw[1])[None], 1 / 128, 127.5)
    ref = TorchGraph(g).run(x)
    plan = graph.compile_graph(loaded, hw)
    assert sum(o.res_mode == 2 for o in plan.ops) == 2                 # both upsample-adds still fuse into their convolutions
    out = plan_sim.run_plan(plan, x)
    glue_ref = TorchGraph(loaded).run(x)                               # the oracle executes the glue node by node
    for name, r in ref.items():
        np.testing.assert_allclose(out[name].reshape(r.shape), r, rtol=0, atol=2e-5 * max(1.0, np.abs(r).max()))
        np.testing.assert_array_equal(glue_ref[name], r)


def test_data_dependent_shape_ops_are_rejected():
    """a Gather over activations is not glue: it must not be folded away silently"""
    from scrfd_arcface_facerecognition_b200.onnx_wire import Node
    g = archs.build_arch("scrfd_500m")
    first = g.nodes[0].outputs[0]
    g.nodes.insert(1, Node("Gather", [first, "idx"], ["gathered"], {"axis": 1}))
    g.initializers["idx"] = np.asarray([0], np.int64)
    for n in g.nodes[2:]:
        n.inputs[:] = ["gathered" if i == first else i for i in n.inputs]
    with pytest.raises((NotImplementedError, KeyError, RuntimeError)):
        graph.compile_graph(g, (160, 160))
