"""How peaky is the decoder's cross-attention (se_model.py:196-204, nn.MultiheadAttention over the memory) on the trained
fixture checkpoint?  For every decode position, layer and head: the share of memory rows whose score lies within `margin` of the
row maximum (weight >= exp(-margin) of the largest).  Rows below that cannot change the fp32 result beyond T * exp(-margin).
Analysis script (uses the numpy oracle) - not part of the product path.   python tests/parity/cross_attention_sparsity.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from oracle import recognizer_np as O      # noqa: E402
from helpers import load_fixture_ckpt      # noqa: E402
from workloads import synth                # noqa: E402

sd = load_fixture_ckpt()
imgs, _ = synth.make_lines(6, 200, 1600, seed=5)
H, D = 8, 384
dh = D // H
for margin in (12.0, 16.0, 20.0):
    tot = np.zeros(2); kept = np.zeros(2); kept_union = np.zeros(2); tot_rows = np.zeros(2); kept8 = np.zeros(2)
    for img in imgs:
        chunks = O.preprocess_gray(img)[1]
        enc = O.encoder_forward(sd, O.patch_forward(sd, O.cnn_forward(sd, chunks, "se")))
        mem = O.memory_for_line(sd, enc, "se")
        toks = [int(t) for t in O.greedy_decode(sd, mem)]          # <sos> + decoded ids
        t = len(toks)
        tok = np.asarray(toks, np.int64)
        x = (sd["dec.tok_emb.weight"][tok] + sd["dec.pos_emb"][:t]).astype(np.float32)
        causal = np.where(np.arange(t)[None, :] > np.arange(t)[:, None], -np.inf, 0.0).astype(np.float32)
        for l in range(2):
            pre = f"dec.decoder.layers.{l}."
            a = O.mha(x, x, x, sd[pre + "self_attn.in_proj_weight"], sd[pre + "self_attn.in_proj_bias"],
                      sd[pre + "self_attn.out_proj.weight"], sd[pre + "self_attn.out_proj.bias"], causal)
            x = O.layer_norm(x + a, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
            W, b = sd[pre + "multihead_attn.in_proj_weight"], sd[pre + "multihead_attn.in_proj_bias"]
            q = (x @ W[:D].T + b[:D]).reshape(t, H, dh)
            k = (mem @ W[D:2 * D].T + b[D:2 * D]).reshape(-1, H, dh)
            s = np.einsum("thd,mhd->thm", q, k) / np.sqrt(dh)                  # (t, H, T)
            keep = s >= s.max(-1, keepdims=True) - margin
            tot[l] += keep.size; kept[l] += keep.sum()
            kept_union[l] += keep.any(1).sum(); tot_rows[l] += keep.any(1).size
            # 8-row granularity (one 8-row K/V tile per warp step)
            T = keep.shape[-1]; T8 = (T + 7) // 8 * 8
            kp = np.zeros(keep.shape[:2] + (T8,), bool); kp[..., :T] = keep
            kept8[l] += kp.reshape(t, H, T8 // 8, 8).any(-1).sum() * 8
            a = O.mha(x, mem, mem, W, b, sd[pre + "multihead_attn.out_proj.weight"], sd[pre + "multihead_attn.out_proj.bias"])
            x = O.layer_norm(x + a, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
            ff = O.linear(O.relu(O.linear(x, sd[pre + "linear1.weight"], sd[pre + "linear1.bias"])),
                          sd[pre + "linear2.weight"], sd[pre + "linear2.bias"])
            x = O.layer_norm(x + ff, sd[pre + "norm3.weight"], sd[pre + "norm3.bias"])
    print(f"margin {margin}: share of (row, head) pairs kept  layer0 {kept[0] / tot[0]:.3f}  layer1 {kept[1] / tot[1]:.3f};"
          f"  rows kept by ANY head  {kept_union[0] / tot_rows[0]:.3f} / {kept_union[1] / tot_rows[1]:.3f};"
          f"  (8-row tile, head) kept {kept8[0] / tot[0]:.3f} / {kept8[1] / tot[1]:.3f}")
