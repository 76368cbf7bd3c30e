"""Pack a reference state_dict into the weight blob `kocr_create` consumes (host logic, numpy only).

What the packing does:
  * eval-mode BatchNorm folded into the preceding conv:  w' = w * g/sqrt(var+eps),
    b' = (b - mean) * g/sqrt(var+eps) + beta            (se_model.py:39-58, BatchNorm2d eps 1e-5)
  * conv weights (Cout, Cin, 3, 3) -> K-major bf16 [Cout][tap = r*3+s][Cin] for the implicit GEMM
  * patch projection Conv2d(512->D, (2,1)) -> bf16 [D][kh*512 + c]           (se_model.py:92-97)
  * SE excitation Conv1d(C, C/16, 1) / Conv1d(C/16, C, 1) -> bf16 GEMM operands [128][C] and [C][128]
    (reduced width zero-padded to the 128-wide N tile)                        (se_model.py:12-17)
  * nn.Linear / in_proj weights are already [N][K] K-major: bf16 cast only
  * LSTM: W_ih of both directions stacked [1536][384]; b = b_ih + b_hh; W_hh re-laid out for the
    2-CTA persistent kernel as bf16x2 [dir][rank][k-pair][row]               (se_model.py:228-234)
  * decoder projections / FFN / out_proj: fp32 pre-rounded to TF32 (the decoder GEMMs run in TF32)
  * decoder cross-attention K/V projections of both layers stacked [1536][384] (bf16)
  * out_proj padded from 124 to 128 rows (pad logits are never read by the argmax)
bf16 conversion is round-to-nearest-even, identical to __float2bfloat16_rn.
"""
from __future__ import annotations

import struct
import numpy as np

from .checkpoint import detect_variant, validate_state_dict, RESNET_BLOCKS

BN_EPS = 1e-5
MAGIC = b"KOCRW001"
DT_F32, DT_A16, DT_I32 = 0, 1, 2
DT_BF16 = DT_A16          # historical name of the 16-bit entry type
A16_FORMAT = 1            # 16-bit storage format of libkocr_b200.so: 1 = IEEE fp16 (default build), 0 = bf16 (-DKOCR_A16_BF16)
LSTM_H = 192


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16, returned as uint16 bit patterns."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounding = ((u >> 16) & 1) + np.uint32(0x7FFF)
    out = ((u + rounding) >> 16).astype(np.uint16)
    nan = np.isnan(x)
    if nan.any():
        out[nan] = 0x7FC0
    return out


def round_to_tf32(x: np.ndarray) -> np.ndarray:
    """fp32 -> nearest TF32 (10-bit mantissa, ties away from zero like cvt.rna.tf32.f32), kept as fp32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u.astype(np.uint64) + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return r.view(np.float32).reshape(np.shape(x))


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


def f32_to_a16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> the library's 16-bit operand format (csrc/common.cuh: fp16, round-to-nearest-even, saturating at
    +-65504 like cvt.rn.satfinite.f16x2.f32; bf16 for the -DKOCR_A16_BF16 build), as uint16 bit patterns."""
    if A16_FORMAT == 0:
        return f32_to_bf16_bits(x)
    x = np.ascontiguousarray(x, dtype=np.float32)
    return np.clip(x, -65504.0, 65504.0).astype(np.float16).view(np.uint16)


def a16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    if A16_FORMAT == 0:
        return bf16_bits_to_f32(b)
    return np.ascontiguousarray(b, dtype=np.uint16).view(np.float16).astype(np.float32)


def fold_bn(w, b, gamma, beta, mean, var):
    s = (gamma / np.sqrt(var.astype(np.float32) + np.float32(BN_EPS))).astype(np.float32)
    return (w * s.reshape(-1, 1, 1, 1)).astype(np.float32), ((b - mean) * s + beta).astype(np.float32)


def conv_to_kmajor(w):
    """(Cout, Cin, 3, 3) -> (Cout, 9*Cin) with k = (r*3+s)*Cin + c."""
    co, ci = w.shape[:2]
    return np.ascontiguousarray(w.transpose(0, 2, 3, 1).reshape(co, 9 * ci))


def pack_whh(w_f, w_b):
    """[2 dirs][768][192] -> bf16 [dir][rank][kp=96][row=384][2]; row = gate*96 + jj <-> W row
    gate*192 + rank*96 + jj; element pair = k (2kp, 2kp+1)."""
    out = np.empty((2, 2, LSTM_H // 2, 384, 2), np.float32)
    for d, w in enumerate((w_f, w_b)):
        w4 = w.reshape(4, 2, 96, LSTM_H // 2, 2)            # gate, rank, jj, kp, pair
        out[d] = w4.transpose(1, 3, 0, 2, 4).reshape(2, LSTM_H // 2, 384, 2)
    return out


def pack_whh_mma(w_f, w_b):
    """Recurrent weights as mma.sync m16n8k16 A-fragments for bilstm_mma_kernel:
    [dir][rank][warp][mtile][kstep][lane][reg 0..3][2].  Warp w of CTA `rank` owns hidden units
    rank*96 + w*8 + g (g = lane//4); m-tile 0 = gates (i | f), m-tile 1 = gates (g | o);
    reg0 = (row g,   k0 + 2*tig + {0,1}), reg1 = (row g+8, same k),
    reg2 = (row g,   k0 + 2*tig + 8 + {0,1}), reg3 = (row g+8, same), tig = lane % 4, k0 = 16*kstep."""
    out = np.empty((2, 2, 12, 2, 12, 32, 4, 2), np.float32)
    lane = np.arange(32)
    g, tig = lane // 4, lane % 4
    for d, w in enumerate((w_f, w_b)):
        for rank in range(2):
            for warp in range(12):
                unit = rank * 96 + warp * 8 + g                       # (32,)
                for mt in range(2):
                    row_a = (2 * mt) * LSTM_H + unit                  # gate i / g
                    row_b = (2 * mt + 1) * LSTM_H + unit              # gate f / o
                    for ks in range(12):
                        k = ks * 16 + 2 * tig
                        for e in range(2):
                            out[d, rank, warp, mt, ks, :, 0, e] = w[row_a, k + e]
                            out[d, rank, warp, mt, ks, :, 1, e] = w[row_b, k + e]
                            out[d, rank, warp, mt, ks, :, 2, e] = w[row_a, k + 8 + e]
                            out[d, rank, warp, mt, ks, :, 3, e] = w[row_b, k + 8 + e]
    return out


def pack_tensors(sd: dict) -> dict:
    """Reference state_dict (numpy fp32) -> {blob entry name: (dtype, ndarray)}."""
    variant, D, max_len, dec_max, = validate_state_dict(sd)
    vocab = sd["dec.tok_emb.weight"].shape[0]
    se = variant == "se"
    t: dict[str, tuple[int, np.ndarray]] = {}

    def f32(name, a):
        t[name] = (DT_F32, np.ascontiguousarray(a, dtype=np.float32))

    def bf16(name, a):          # 16-bit tensor-core operand (fp16 by default, see f32_to_a16_bits)
        t[name] = (DT_A16, f32_to_a16_bits(np.ascontiguousarray(a, dtype=np.float32)))

    def tf32(name, a):          # decoder GEMM operands stay fp32 (consumed as TF32 by the tensor cores)
        t[name] = (DT_F32, round_to_tf32(np.ascontiguousarray(a, dtype=np.float32)))

    resnet = variant == "resnet"
    t["meta"] = (DT_I32, np.asarray([0 if se else (2 if resnet else 1), D, max_len, dec_max, vocab, A16_FORMAT, 0, 0], np.int32))
    if resnet:
        # BasicBlock convs have no bias: the folded BN supplies it.  1x1 shortcut convs are [Cout][Cin] linear layers.
        def fold(conv, bnp):
            w = sd[conv + ".weight"]
            return fold_bn(w, np.zeros(w.shape[0], np.float32), sd[bnp + ".weight"], sd[bnp + ".bias"],
                           sd[bnp + ".running_mean"], sd[bnp + ".running_var"])
        w, b = fold("cnn.conv1", "cnn.bn1")
        f32("conv1.w", w.reshape(64, 9))
        w16 = np.zeros((64, 16), np.float32); w16[:, :9] = w.reshape(64, 9)
        bf16("conv1.w16", w16)
        f32("conv1.b", b)
        for bi, (name, cin, cout) in enumerate(RESNET_BLOCKS):
            p = f"cnn.{name}"
            w, b = fold(p + ".conv1", p + ".bn1")
            bf16(f"res{bi}.c1.w", conv_to_kmajor(w)); f32(f"res{bi}.c1.b", b)
            w, b = fold(p + ".conv2", p + ".bn2")
            bf16(f"res{bi}.c2.w", conv_to_kmajor(w)); f32(f"res{bi}.c2.b", b)
            if cin != cout:
                w, b = fold(p + ".shortcut.0", p + ".shortcut.1")
                bf16(f"res{bi}.sc.w", w.reshape(cout, cin)); f32(f"res{bi}.sc.b", b)
    for i in (range(1, 7) if not resnet else ()):
        p = f"cnn.conv{i}"
        w, b = fold_bn(sd[p + ".0.weight"], sd[p + ".0.bias"], sd[p + ".1.weight"], sd[p + ".1.bias"],
                       sd[p + ".1.running_mean"], sd[p + ".1.running_var"])
        if i == 1:
            f32("conv1.w", w.reshape(64, 9))
            w16 = np.zeros((64, 16), np.float32); w16[:, :9] = w.reshape(64, 9)
            bf16("conv1.w16", w16)                      # tensor-core conv1: K = 9 taps zero-padded to 16
            f32("conv1.b", b)
        else:
            bf16(f"conv{i}.w", conv_to_kmajor(w))
            f32(f"conv{i}.b", b)
    w7, b7 = (sd["cnn.conv7.weight"], sd["cnn.conv7.bias"]) if not resnet else (None, None)
    if se:
        w7, b7 = fold_bn(w7, b7, sd["cnn.bn7.weight"], sd["cnn.bn7.bias"], sd["cnn.bn7.running_mean"],
                         sd["cnn.bn7.running_var"])
        for k in (3, 4, 5):
            # excitation FCs as GEMM operands, reduced width R = C/16 zero-padded to 128
            w0 = sd[f"cnn.se{k}.fc.0.weight"][:, :, 0]            # (R, C)
            w2 = sd[f"cnn.se{k}.fc.2.weight"][:, :, 0]            # (C, R)
            R, C = w0.shape
            w0p = np.zeros((128, C), np.float32); w0p[:R] = w0
            b0p = np.zeros(128, np.float32); b0p[:R] = sd[f"cnn.se{k}.fc.0.bias"]
            w2p = np.zeros((C, 128), np.float32); w2p[:, :R] = w2
            bf16(f"se{k}.w0p", w0p)
            f32(f"se{k}.b0p", b0p)
            bf16(f"se{k}.w2p", w2p)
            bf16(f"se{k}.w0f", se_fragments(w0))                  # the same weights in mma.sync B-fragment order
            bf16(f"se{k}.w2f", se_fragments(w2))
            f32(f"se{k}.b2", sd[f"cnn.se{k}.fc.2.bias"])
    if not resnet:
        bf16("conv7.w", conv_to_kmajor(w7))
        f32("conv7.b", b7)
    pw = sd["patch.proj.weight"][:, :, :, 0]                       # (D, 512, 2)
    bf16("patch.w", pw.transpose(0, 2, 1).reshape(D, 1024))       # k = kh*512 + c
    f32("patch.b", sd["patch.proj.bias"])
    f32("patch.pos", sd["patch.pos_emb"][:32])
    for l in range(2):
        p = f"enc.layers.{l}."
        bf16(f"enc{l}.in_w", sd[p + "self_attn.in_proj_weight"]); f32(f"enc{l}.in_b", sd[p + "self_attn.in_proj_bias"])
        bf16(f"enc{l}.out_w", sd[p + "self_attn.out_proj.weight"]); f32(f"enc{l}.out_b", sd[p + "self_attn.out_proj.bias"])
        bf16(f"enc{l}.l1_w", sd[p + "linear1.weight"]); f32(f"enc{l}.l1_b", sd[p + "linear1.bias"])
        bf16(f"enc{l}.l2_w", sd[p + "linear2.weight"]); f32(f"enc{l}.l2_b", sd[p + "linear2.bias"])
        for k in (1, 2):
            f32(f"enc{l}.n{k}_g", sd[p + f"norm{k}.weight"]); f32(f"enc{l}.n{k}_b", sd[p + f"norm{k}.bias"])
    f32("global_pos", sd["global_pos"])
    if se:
        p = "context_bilstm."
        wih = np.concatenate([sd[p + "weight_ih_l0"], sd[p + "weight_ih_l0_reverse"]], 0).astype(np.float32)
        bf16("lstm.w_ih", wih)
        wih_hi = a16_bits_to_f32(f32_to_a16_bits(wih)).reshape(wih.shape)      # split-precision copy, see dec.ca_kv_w3
        bf16("lstm.w_ih3", np.concatenate([wih_hi, wih_hi, wih - wih_hi], 1))
        f32("lstm.b", np.concatenate([sd[p + "bias_ih_l0"] + sd[p + "bias_hh_l0"],
                                      sd[p + "bias_ih_l0_reverse"] + sd[p + "bias_hh_l0_reverse"]]))
        bf16("lstm.w_hh", pack_whh(sd[p + "weight_hh_l0"], sd[p + "weight_hh_l0_reverse"]))
        bf16("lstm.w_hh_mma", pack_whh_mma(sd[p + "weight_hh_l0"], sd[p + "weight_hh_l0_reverse"]))
    f32("dec.tok_emb", sd["dec.tok_emb.weight"])
    f32("dec.pos", sd["dec.pos_emb"])
    kv_w, kv_b = [], []
    for l in range(2):
        p = f"dec.decoder.layers.{l}."
        tf32(f"dec{l}.sa_in_w", sd[p + "self_attn.in_proj_weight"]); f32(f"dec{l}.sa_in_b", sd[p + "self_attn.in_proj_bias"])
        tf32(f"dec{l}.sa_out_w", sd[p + "self_attn.out_proj.weight"]); f32(f"dec{l}.sa_out_b", sd[p + "self_attn.out_proj.bias"])
        cw, cb = sd[p + "multihead_attn.in_proj_weight"], sd[p + "multihead_attn.in_proj_bias"]
        tf32(f"dec{l}.ca_q_w", cw[:D]); f32(f"dec{l}.ca_q_b", cb[:D])
        kv_w.append(cw[D:]); kv_b.append(cb[D:])                   # rows: K (D) then V (D)
        tf32(f"dec{l}.ca_out_w", sd[p + "multihead_attn.out_proj.weight"]); f32(f"dec{l}.ca_out_b", sd[p + "multihead_attn.out_proj.bias"])
        tf32(f"dec{l}.l1_w", sd[p + "linear1.weight"]); f32(f"dec{l}.l1_b", sd[p + "linear1.bias"])
        tf32(f"dec{l}.l2_w", sd[p + "linear2.weight"]); f32(f"dec{l}.l2_b", sd[p + "linear2.bias"])
        for k in (1, 2, 3):
            f32(f"dec{l}.n{k}_g", sd[p + f"norm{k}.weight"]); f32(f"dec{l}.n{k}_b", sd[p + f"norm{k}.bias"])
    kvw = np.concatenate(kv_w, 0).astype(np.float32)
    bf16("dec.ca_kv_w", kvw)
    # split-precision copy for the K = 3 x 384 projection: [w_hi | w_hi | w_lo] against operand rows [hi | lo | hi]
    w_hi = a16_bits_to_f32(f32_to_a16_bits(kvw)).reshape(kvw.shape)
    bf16("dec.ca_kv_w3", np.concatenate([w_hi, w_hi, kvw - w_hi], 1))
    f32("dec.ca_kv_b", np.concatenate(kv_b, 0))
    ow = np.zeros((128, D), np.float32); ow[:vocab] = sd["dec.out_proj.weight"]
    ob = np.zeros(128, np.float32); ob[:vocab] = sd["dec.out_proj.bias"]
    tf32("dec.out_w", ow)
    f32("dec.out_b", ob)
    return t


def se_fragments(w: np.ndarray) -> np.ndarray:
    """w [N][K] (out, in) -> the B operand of `mma.sync.m16n8k16` (B[k][n] = w[n][k]) in FRAGMENT order: for every n-tile of 8
    outputs and every k-step of 16, lane (g = lane >> 2, t = lane & 3) holds {w[n0+g][k0+2t], w[n0+g][k0+2t+1]} (b0) and
    {w[n0+g][k0+8+2t], w[n0+g][k0+9+2t]} (b1); laid out [n-tile][k-step PAIR][lane][8 values] (K = 16: [n-tile][lane][4]), so a
    warp reads its fragments of two k-steps with ONE coalesced 16-byte load per lane (se_excite_kernel, csrc/cnn_misc.cu)."""
    N, K = w.shape
    assert N % 8 == 0 and K % 16 == 0
    lane = np.arange(32)
    g, t = lane >> 2, lane & 3
    koff = np.stack([2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9], axis=1)              # [32][4] within a k-step
    ks = np.arange(K // 16)
    n_idx = (np.arange(N // 8)[:, None, None, None] * 8 + g[None, None, :, None])   # [NT][1][32][1]
    k_idx = ks[None, :, None, None] * 16 + koff[None, None, :, :]                   # [1][KS][32][4]
    frag = w[n_idx, k_idx]                                                          # [NT][KS][32][4]
    if K // 16 >= 2:
        assert (K // 16) % 2 == 0
        frag = frag.reshape(N // 8, K // 32, 2, 32, 4).transpose(0, 1, 3, 2, 4)     # [NT][KP][32][2][4]
    return np.ascontiguousarray(frag, np.float32).reshape(-1)


def pack_blob(sd: dict) -> bytes:
    """Serialise: header {magic[8], u32 n, u32 0}, n entries {name[48], u32 dtype, u32 0, u64 offset,
    u64 nbytes}, then 256-byte-aligned payloads."""
    tensors = pack_tensors(sd)
    n = len(tensors)
    head = 16 + n * 72
    off = (head + 255) // 256 * 256
    entries, payload = [], []
    for name, (dt, arr) in tensors.items():
        raw = arr.tobytes()
        nb = name.encode()
        assert len(nb) < 48
        entries.append(struct.pack("<48sIIQQ", nb, dt, 0, off, len(raw)))
        pad = (-len(raw)) % 256
        payload.append(raw + b"\0" * pad)
        off += len(raw) + pad
    blob = MAGIC + struct.pack("<II", n, 0) + b"".join(entries)
    blob += b"\0" * ((-len(blob)) % 256)
    return blob + b"".join(payload)


def variant_of(sd: dict) -> str:
    return detect_variant(sd)
