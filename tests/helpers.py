"""Shared test helpers (loading golden fixtures, synthetic inputs identical to
oracle/make_golden.py)."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def load_tiny(name):
    """-> (cfg dict, state_dict with FULL constant tables rebuilt, inputs, outputs, grads)."""
    from oracle import vct_oracle as O
    z = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = json.loads(str(z["cfg"]))
    sd, grads, ins, outs = {}, {}, {}, {}
    for k in z.files:
        if k.startswith("sd/"):
            sd[k[3:]] = torch.from_numpy(z[k])
        elif k.startswith("grad/"):
            grads[k[5:]] = torch.from_numpy(z[k])
        elif k.startswith("in/"):
            ins[k[3:]] = torch.from_numpy(z[k])
        elif k.startswith("out/"):
            outs[k[4:]] = z[k]
    d = cfg["d"]
    pos_rows = sd["cap_decoder.positional_encoding.pos_embedding"]
    pe_rows = sd["video_encoder.temp_emb.pe"]
    full_pos = O.sinusoid_table(5000, d)
    full_pe = O.temporal_sinusoid_table(512, d).unsqueeze(0)
    # the committed rows pin the table formula bit-exactly
    assert torch.equal(full_pos[:64], pos_rows), "pos_embedding table formula drifted from the reference"
    assert torch.equal(full_pe[:, :64], pe_rows), "temporal pe table formula drifted from the reference"
    sd["cap_decoder.positional_encoding.pos_embedding"] = full_pos
    sd["video_encoder.temp_emb.pe"] = full_pe
    return cfg, sd, ins, outs, grads


def load_anchors():
    with open(os.path.join(GOLD, "fullsize_anchors.json")) as f:
        return json.load(f)


def synth_inputs(B, T, Din, S1, V, seed=1234, padded=True, vid_padded=False):
    """Same generator as oracle/make_golden.py:synth_inputs."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, Din, generator=g)
    lo = 1000 if V > 2000 else 104
    tok = torch.randint(lo, V, (B, S1), generator=g)
    tok[:, 0] = 101
    if padded:
        lens = torch.randint(min(6, S1 - 1), S1 + 1, (B,), generator=g)
        lens[0] = S1
        for b in range(B):
            tok[b, lens[b] - 1] = 102
            tok[b, lens[b]:] = 0
    else:
        tok[:, -1] = 102
    vm = torch.zeros(B, T, dtype=torch.bool)
    if vid_padded:
        vlen = torch.randint(2, T + 1, (B,), generator=g)
        vlen[0] = T
        for b in range(B):
            vm[b, vlen[b]:] = True
            x[b, vlen[b]:] = 0.0
    return x, vm, tok


def load_extra():
    """Round-2 anchors made by ``oracle/make_golden.py --extra`` -> (anchors dict, samples npz)."""
    with open(os.path.join(GOLD, "extra_anchors.json")) as f:
        anchors = json.load(f)
    return anchors, np.load(os.path.join(GOLD, "extra_samples.npz"))


def sampled(t, stride):
    return t[::int(stride[0]), ::int(stride[1])]


def dropin_state_dict(tokenizer_dir, enc_layers=1, dec_layers=3, embed_dim=768):
    """state_dict of OUR MMT4Caption constructed on CPU with the reference's seed: bit-identical to the reference
    constructor's (tests/test_host.py checks the checksums), so the oracle can be run at full size on the GPU box where
    /root/reference does not exist."""
    from model.MMT4Caption import MMT4Caption
    from vct.synthetic import shipped_model_config
    torch.manual_seed(666)
    m = MMT4Caption(shipped_model_config(tokenizer_dir, embed_dim=embed_dim, enc_layers=enc_layers, dec_layers=dec_layers),
                    device=torch.device("cpu"))
    return {k: v.detach().clone() for k, v in m.state_dict().items()}
