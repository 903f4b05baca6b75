"""Phase timestamps of CTA (0,0) of the fused attention kernels + warm back-to-back timing (graph of 10 launches)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-captioning-transformer_b200"))
import torch
from vct import lib as L
lib = L.load()
dev = "cuda"
trace = torch.zeros(128, dtype=torch.int64, device=dev)
rng = torch.tensor([77, 5], dtype=torch.int64, device=dev)
NAMES = ["entry", "pdl_wait", "proj_committed", "proj_visible", "tiles_written", "scores_visible", "P_written", "O_visible", "O_stored", "exit"]
def run(tag, B, Lq, Lk, d, H, cross, causal, p):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B * Lq, d, generator=g).to(dev, torch.bfloat16)
    mem = torch.randn(B * Lk, d, generator=g).to(dev, torch.bfloat16)
    w = (torch.randn(3 * d, d, generator=g) * 0.04).to(dev, torch.bfloat16)
    b = (torch.randn(3 * d, generator=g) * 0.1).to(dev)
    kv = torch.randn(B * Lk, 2 * d, generator=g).to(dev, torch.bfloat16)
    pad = torch.zeros(B, Lq, dtype=torch.uint8, device=dev)
    m = L.MhaArgs()
    q = torch.empty((B * Lq, 3 * d), device=dev, dtype=torch.bfloat16)
    o = torch.empty((B * Lq, d), device=dev, dtype=torch.bfloat16)
    m.B, m.L, m.Lk, m.d, m.H, m.dtype = B, Lq, Lk, d, H, L.BF16
    m.x, m.mem, m.w_in, m.b_in = x.data_ptr(), mem.data_ptr(), w.data_ptr(), b.data_ptr()
    m.qkv, m.kv, m.kv_ready, m.o = q.data_ptr(), kv.data_ptr(), 1, o.data_ptr()
    m.key_pad = None if cross else pad.data_ptr()
    m.drop_p, m.rng_state, m.site = p, rng.data_ptr(), 3
    m.gemm_impl = L.GEMM_TCGEN05
    fn = lib.vct_attn_dec_cross_fwd if cross else (lib.vct_attn_dec_self_fwd if causal else lib.vct_attn_enc_self_fwd)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        L.check(lib.vct_gemm_trace(trace.data_ptr()))
        for _ in range(4):
            L.check(fn(C.byref(m), st.cuda_stream))
        torch.cuda.synchronize()
        t = trace.cpu().tolist()
        L.check(lib.vct_gemm_trace(None))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for _ in range(10):
                L.check(fn(C.byref(m), st.cuda_stream))
        graph.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); graph.replay(); e1.record(st); torch.cuda.synchronize()
    z = t[100]
    print(f"{tag}: {e0.elapsed_time(e1) * 100:.2f} us/launch | " + "  ".join(f"{n} {t[100 + k] - z}" for k, n in enumerate(NAMES)))
run("dec self  B64 L20", 64, 20, 20, 768, 8, False, 1, 0.3)
run("enc self  B64 L13", 64, 13, 13, 768, 8, False, 0, 0.3)
run("dec cross B64 20x13", 64, 20, 13, 768, 8, True, 0, 0.3)
run("dec cross B128 20x33", 128, 20, 33, 768, 8, True, 0, 0.3)
run("dec self  B64 L20 p0", 64, 20, 20, 768, 8, False, 1, 0.0)
