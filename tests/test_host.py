"""CPU-side checks: the C-ABI library loads and exports every symbol include/vct.h declares, the
drop-in package keeps the reference's API / state_dict surface, host logic (arena, buckets,
tokeniser, synthetic inputs) behaves, and the product refuses to run without CUDA."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import importlib.util
    spec = importlib.util.spec_from_file_location("vct_build", os.path.join(ROOT, "video-captioning-transformer_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def test_library_exports_every_declared_symbol(built):
    from vct import lib as L
    header = open(os.path.join(ROOT, "include", "vct.h")).read()
    declared = set(re.findall(r"\b(vct_[a-z0-9_]+)\s*\(", header))
    declared -= {n for n in declared if n.endswith("_args") or n.endswith("_t")}
    assert declared, "no declarations parsed"
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    lib = ctypes.CDLL(built)
    for name in declared:
        assert hasattr(lib, name), name
    bound = L.load()
    assert bound.vct_version() >= 100
    assert bound.vct_ln_bwd_workspace_floats(1280, 768) > 0      # host-only helper, no GPU needed
    assert bound.vct_colsum_workspace_floats(1280, 30522) == 20 * 30522


def test_ctypes_structs_match_header_field_order():
    from vct import lib as L
    header = open(os.path.join(ROOT, "include", "vct.h")).read()

    def fields(struct_name):
        body = [c for c in header.split("typedef struct {")[1:] if c.split("}")[1].strip().startswith(struct_name + ";")][0]
        body = body.split("}")[0]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            parts = [p.strip() for p in decl.split(",")]
            for i, p in enumerate(parts):
                names.append(re.sub(r"[\*\s]", " ", p).split()[-1])
        return names
    assert fields("vct_gemm_args") == [f[0] for f in L.GemmArgs._fields_]
    assert fields("vct_attn_args") == [f[0] for f in L.AttnArgs._fields_]
    assert fields("vct_mha_args") == [f[0] for f in L.MhaArgs._fields_]


def test_ctypes_struct_sizes_match_the_c_compiler(tmp_path):
    """Field ORDER is checked above; sizes / padding are checked by compiling include/vct.h with gcc (plain C: the header
    must stay free of C++ and torch types) and comparing sizeof with the ctypes mirrors."""
    import ctypes as C
    import shutil
    import subprocess
    from vct import lib as L
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "vct.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(vct_gemm_args), '
                   'sizeof(vct_attn_args), sizeof(vct_mha_args));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(L.GemmArgs), C.sizeof(L.AttnArgs), C.sizeof(L.MhaArgs)]


def test_dropin_api_surface_and_state_dict(tokenizer_dir):
    from model.MMT4Caption import MMT4Caption
    from vct.synthetic import shipped_model_config
    torch.manual_seed(666)
    m = MMT4Caption(shipped_model_config(tokenizer_dir), device=torch.device("cpu"))
    m.mode("caption")
    keys = set(m.state_dict())
    for k in ("video_encoder.unify.0.weight", "video_encoder.temp_emb.pe",
              "video_encoder.transformer_encoder.layers.0.self_attn.in_proj_weight",
              "video_encoder.transformer_encoder.norm.bias", "cap_decoder.decoder.layers.2.multihead_attn.out_proj.weight",
              "cap_decoder.decoder.layers.0.norm3.weight", "cap_decoder.decoder.norm.weight", "cap_decoder.generator.bias",
              "cap_decoder.tgt_to_emb.weight", "cap_decoder.positional_encoding.pos_embedding", "matching.v_proj.weight"):
        assert k in keys, k
    assert len(keys) == 79
    assert sum(p.numel() for p in m.parameters()) == 76850746
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 76457018
    assert not any(p.requires_grad for p in m.matching.parameters())
    pp = m.cap_preprocessor
    assert (pp.pad_id, pp.start_id, pp.end_id) == (0, 101, 102)
    layer = m.cap_decoder.decoder.layers[0]
    for attr in ("self_attn", "multihead_attn", "linear1", "linear2", "activation", "dropout", "dropout1", "dropout2",
                 "dropout3", "norm1", "norm2", "norm3"):
        assert hasattr(layer, attr), attr                                        # predict_video.py:60-69
    with pytest.raises(ValueError):
        m.mode("nonsense")
    m.f_type = None
    with pytest.raises(ValueError):
        m([torch.zeros(1, 12, 512)], [torch.zeros(1, 12, dtype=torch.bool)], ["w1000"])
    m.mode("caption")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m([torch.zeros(2, 12, 512)], [torch.zeros(2, 12, dtype=torch.bool)], ["w1000 w1001", "w1002"])


def test_out_of_scope_variants_raise():
    from model.MMEncoder import MultiModalEncoder, SimpleSepEncoder, HMMEncoder
    from model.CapDecoder import CapDecoder
    cpu = torch.device("cpu")
    with pytest.raises(NotImplementedError):
        MultiModalEncoder([512, 128], 64, 2, 64, 1, 0.1, "gelu", "avg", True, "encoding", False, cpu)
    with pytest.raises(NotImplementedError):
        MultiModalEncoder([512], 64, 2, 64, 1, 0.1, "gelu", "avg", True, "embedding", False, cpu)
    with pytest.raises(NotImplementedError):
        MultiModalEncoder([512], 64, 2, 64, 1, 0.1, "gelu", "GRU", True, "encoding", False, cpu)
    with pytest.raises(NotImplementedError):
        SimpleSepEncoder()
    with pytest.raises(NotImplementedError):
        HMMEncoder()
    with pytest.raises(NotImplementedError):
        CapDecoder(1, 64, 2, 64, 0.1, 100, 0, 0.5, "vis", "gelu", cpu)


def test_tokeniser_batches_like_the_reference(tokenizer_dir):
    from model.CapPreprocessor import CapPreprocessor
    pp = CapPreprocessor(tokenizer_dir, device=torch.device("cpu"))
    ids, mask = pp(["w1000 w1001 w1002", "w2000"])
    assert ids.tolist() == [[101, 1000, 1001, 1002, 102], [101, 2000, 102, 0, 0]]
    assert mask.tolist() == [[False] * 5, [False, False, False, True, True]]


def test_param_arena_aliases_parameters_and_tracks_versions():
    from vct.arena import ParamArena, ALIGN
    lin, emb = torch.nn.Linear(10, 7), torch.nn.Embedding(5, 3)
    before = {n: p.detach().clone() for n, p in list(lin.named_parameters()) + list(emb.named_parameters())}
    named = [("lin." + n, p) for n, p in lin.named_parameters()] + [("emb." + n, p) for n, p in emb.named_parameters()]
    a = ParamArena(named, torch.device("cpu"))
    assert a.numel % ALIGN == 0 and all(o % ALIGN == 0 for o in a.offset.values())
    assert a.is_current()
    torch.testing.assert_close(lin.weight.detach(), before["weight"].view(7, 10) if before["weight"].numel() == 70 else lin.weight.detach())
    lin.weight.data.fill_(3.0)
    assert float(a.view(a.p32, "lin.weight").sum()) == 3.0 * 70       # parameter storage IS the arena
    v0 = a.version()
    with torch.no_grad():
        lin.bias.add_(1.0)
    assert a.version() > v0
    lin.to(torch.float64)
    assert not a.is_current()


def test_gradient_buckets_cover_the_arena():
    from vct.arena import ParamArena
    from vct.trainer import gradient_buckets
    enc, dec = torch.nn.Linear(100, 300), torch.nn.Linear(300, 50)
    named = [("video_encoder." + n, p) for n, p in enc.named_parameters()] + \
            [("cap_decoder." + n, p) for n, p in dec.named_parameters()]
    a = ParamArena(named, torch.device("cpu"))
    b = gradient_buckets(a, ["video_encoder.", "cap_decoder."], max_bytes=4 * 4096)
    assert sorted(b)[0][0] == 0 and sorted(b)[-1][1] == a.numel
    assert sum(hi - lo for lo, hi in b) == a.numel and all(hi - lo <= 4096 for lo, hi in b)
    with pytest.raises(ValueError):
        gradient_buckets(a, ["cap_decoder."])


def test_synthetic_batch_shapes_and_tokens():
    from vct.synthetic import synth_batch
    x, vm, tok = synth_batch(8, padded=True)
    assert x.shape == (8, 12, 512) and vm.shape == (8, 12) and tok.shape == (8, 21)
    assert (tok[:, 0] == 101).all() and not vm.any()
    for row in tok.tolist():
        n = row.index(102)
        assert all(t == 0 for t in row[n + 1:]) and all(t >= 1000 for t in row[1:n])


def test_backward_plan_is_split_at_optimizer_lane_runs():
    """vct.trainer.split_segments: the N > 1 trainer captures the launches between two optimizer slices as one CUDA graph
    and issues the slice's (all-reduce, Adam) eagerly; every call must land in exactly one segment, in order."""
    from vct.trainer import split_segments
    c = lambda name, lane: (name, None, (), lane)          # noqa: E731
    calls = [c("a", 0), c("w1", 1), c("b", 0), c("py:all_reduce:x", 2), c("vct_adam:x", 2), c("c", 0), c("join", 0), c("w2", 1),
             c("py:all_reduce:y", 2), c("vct_adam:y", 2), c("vct_adam:z", 2), c("d", 0)]
    segs = split_segments(calls)
    assert [[x[0] for x in g] for g, _ in segs] == [["a", "w1", "b"], ["c", "join", "w2"], ["d"]]
    assert [[x[0] for x in o] for _, o in segs] == [["py:all_reduce:x", "vct_adam:x"], ["py:all_reduce:y", "vct_adam:y", "vct_adam:z"], []]
    flat = [x for g, o in segs for x in (g + o)]
    assert flat == calls
    # optimizer calls first (nothing to capture before them) and no trailing main-lane work
    segs = split_segments([c("vct_adam:p", 2), c("a", 0), c("vct_adam:q", 2)])
    assert [[x[0] for x in g] for g, _ in segs] == [[], ["a"]] and [[x[0] for x in o] for _, o in segs] == [["vct_adam:p"], ["vct_adam:q"]]
    assert split_segments([]) == []
