"""CPU oracle for the text-line recognition forward path (TEST INFRASTRUCTURE ONLY).

This file is a plain-numpy restatement of the reference algorithm
(netra-ai-lab/Khmer-OCR-CNN-Transformer, `netra_ocr/recognition`).  It exists to CHECK the CUDA
path; it is never on the product path.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.

Parity pinning: the reference ships no golden vectors for this path.  The oracle
is therefore pinned against outputs of the reference itself, generated in the build container by
`tests/golden/make_fixtures.py` (which imports `/root/reference` read-only) and committed under
`tests/golden/`; `tests/test_oracle_golden.py` checks every function below against them.

Third-party arithmetic restated here (the reference only *calls* these):
  * Pillow `Image.convert('L')` and `Image.resize(BILINEAR)`  (reference pins pillow>=10.2,<11;
    verified against the installed 12.2.0) - call sites preprocessor.py:39-49.
  * torchvision `ToTensor`                                     - preprocessor.py:11-14,50.
  * torch.nn Conv2d / BatchNorm2d(eval) / MaxPool2d / AdaptiveAvgPool2d / Conv1d /
    TransformerEncoderLayer / TransformerDecoderLayer / LSTM / Embedding / Linear
    (reference pins torch>=2.7,<3; verified against 2.11.0)    - se_model.py, vgg_model.py.

All reference citations are relative to /root/reference/netra_ocr/recognition/.
"""
from __future__ import annotations

import math
import numpy as np

F32 = np.float32

IMG_H = 48          # config.py:7
CHUNK_W = 100       # config.py:8
OVERLAP = 16        # config.py:9
STRIDE = CHUNK_W - OVERLAP
NHEAD = 8           # se_model.py:217,236
LN_EPS = 1e-5
BN_EPS = 1e-5

PAD, UNK, SOS, EOS = 0, 1, 2, 3   # char2idx.json


# ----------------------------------------------------------------------------------------------
# Stage 1: grey conversion, Pillow-exact bilinear resize, chunking, normalisation
# ----------------------------------------------------------------------------------------------
def rgb_to_l(rgb: np.ndarray) -> np.ndarray:
    """Pillow ImagingConvert rgb2l: L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16.
    Called by `image.convert('L')`, preprocessor.py:39,41."""
    r = rgb[..., 0].astype(np.uint32)
    g = rgb[..., 1].astype(np.uint32)
    b = rgb[..., 2].astype(np.uint32)
    return ((19595 * r + 38470 * g + 7471 * b + 0x8000) >> 16).astype(np.uint8)


def textline_boxes(image_size, polygons, expansion_px=5):
    """extract_textline_crops steps 1-2 (netra_ocr/textline_detection.py:17-34): int() of the polygon extremes,
    expansion, clipping, empty boxes skipped."""
    img_w, img_h = image_size
    boxes = []
    for poly in polygons:
        xs = [p[0] for p in poly]
        ys = [p[1] for p in poly]
        x0, y0, x1, y1 = int(min(xs)), int(min(ys)), int(max(xs)), int(max(ys))
        x0, y0 = max(0, x0 - expansion_px), max(0, y0 - expansion_px)
        x1, y1 = min(img_w, x1 + expansion_px), min(img_h, y1 + expansion_px)
        if x1 - x0 <= 0 or y1 - y0 <= 0:
            continue
        boxes.append((x0, y0, x1, y1))
    return boxes


def crop_line_gray(page: np.ndarray, box, padding_px=10) -> np.ndarray:
    """extract_textline_crops steps 3-4 (textline_detection.py:36-47: crop, paste on a white RGB canvas) followed by the
    `convert('L')` of ImagePreprocessor.process (preprocessor.py:41).  page: uint8 (H, W, 3) or (H, W)."""
    x0, y0, x1, y1 = box
    crop = page[y0:y1, x0:x1]
    gray = rgb_to_l(crop) if crop.ndim == 3 else crop
    out = np.full((gray.shape[0] + 2 * padding_px, gray.shape[1] + 2 * padding_px), 255, np.uint8)
    out[padding_px:padding_px + gray.shape[0], padding_px:padding_px + gray.shape[1]] = gray
    return out


def resized_width(w: int, h: int) -> int:
    """preprocessor.py:45-47: Python float division, truncation, floor of chunk_width//2."""
    aspect = w / h
    return max(CHUNK_W // 2, int(IMG_H * aspect))


def n_chunks_for_width(w: int) -> int:
    """preprocessor.py:21-31: `while start < W` with stride 84."""
    return (w + STRIDE - 1) // STRIDE


def _bilinear_coeffs(in_size: int, out_size: int):
    """Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR (triangle,
    support 1.0) filter.  Returns (xmin[out], count[out], kk[out, ksize] int32), PRECISION_BITS=22."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmins = np.zeros(out_size, np.int32)
    counts = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(xmax, np.float64)
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
        ww = w.sum() if False else 0.0
        for x in range(xmax):      # sequential double accumulation, as in the C loop
            ww += w[x]
        if ww != 0.0:
            w = w / ww
        for x in range(xmax):
            v = w[x] * (1 << 22)
            kk[xx, x] = int(v - 0.5) if w[x] < 0 else int(v + 0.5)
        xmins[xx] = xmin
        counts[xx] = xmax
    return xmins, counts, kk


def _resample_axis_u8(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One separable pass (uint8 in, uint8 out): out = clip8((2^21 + sum pix*k) >> 22)."""
    in_size = img.shape[axis]
    xmins, counts, kk = _bilinear_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, -1).astype(np.int64)           # (..., in)
    out = np.empty(src.shape[:-1] + (out_size,), np.uint8)
    for xx in range(out_size):
        c = int(counts[xx])
        x0 = int(xmins[xx])
        acc = (src[..., x0:x0 + c] * kk[xx, :c].astype(np.int64)).sum(-1) + (1 << 21)
        out[..., xx] = np.clip(acc >> 22, 0, 255).astype(np.uint8)
    return np.moveaxis(out, -1, axis)


def pil_resize_bilinear_u8(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """Pillow `Image.resize((out_w,out_h), BILINEAR)` for mode 'L' (preprocessor.py:49):
    horizontal pass first (skipped when the width is unchanged), then vertical (skipped when the
    height is unchanged), with a uint8 intermediate."""
    h, w = img.shape
    out = img
    if out_w != w:
        out = _resample_axis_u8(out, out_w, axis=1)
    if out_h != h:
        out = _resample_axis_u8(out, out_h, axis=0)
    return np.ascontiguousarray(out)


def to_tensor_normalise(u8: np.ndarray) -> np.ndarray:
    """torchvision ToTensor (u8 -> f32, /255) followed by (c - 0.5) / 0.5, preprocessor.py:50,55.
    Two IEEE fp32 ops; a fused u8*(2/255)-1 is NOT bit-identical."""
    t = u8.astype(F32) / F32(255.0)
    return (t - F32(0.5)) / F32(0.5)


def chunk_resized_line(line_u8: np.ndarray) -> np.ndarray:
    """`_chunk_tensor` + normalise, preprocessor.py:16-33,55-58.  line_u8: (48, W) uint8.
    Returns f32 (n, 1, 48, 100); columns >= W are white (1.0 before normalisation)."""
    H, W = line_u8.shape
    assert H == IMG_H
    n = n_chunks_for_width(W)
    out = np.full((n, 1, IMG_H, CHUNK_W), 255, np.uint8)
    for k in range(n):
        s = k * STRIDE
        e = min(s + CHUNK_W, W)
        out[k, 0, :, : e - s] = line_u8[:, s:e]
    return to_tensor_normalise(out)


def preprocess_gray(img_u8: np.ndarray):
    """`ImagePreprocessor.process` (preprocessor.py:35-58) for an already-grey (h, w) uint8 image.
    Returns (resized uint8 (48, W'), chunks f32 (n,1,48,100))."""
    h, w = img_u8.shape
    new_w = resized_width(w, h)
    resized = pil_resize_bilinear_u8(img_u8, new_w, IMG_H)
    return resized, chunk_resized_line(resized)


# ----------------------------------------------------------------------------------------------
# torch.nn primitives restated
# ----------------------------------------------------------------------------------------------
def conv3x3(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """nn.Conv2d(Cin, Cout, 3, 1, 1) on NCHW fp32 via im2col."""
    N, C, H, W = x.shape
    Co = w.shape[0]
    xp = np.zeros((N, C, H + 2, W + 2), F32)
    xp[:, :, 1:-1, 1:-1] = x
    cols = np.empty((N, H, W, C, 3, 3), F32)
    for r in range(3):
        for s in range(3):
            cols[:, :, :, :, r, s] = xp[:, :, r:r + H, s:s + W].transpose(0, 2, 3, 1)
    y = cols.reshape(N * H * W, C * 9) @ w.reshape(Co, C * 9).T.astype(F32)
    y += b.astype(F32)
    return np.ascontiguousarray(y.reshape(N, H, W, Co).transpose(0, 3, 1, 2))


def batchnorm_eval(x, weight, bias, mean, var):
    """nn.BatchNorm2d in eval mode (running stats), eps 1e-5."""
    inv = (1.0 / np.sqrt(var.astype(F32) + F32(BN_EPS))).astype(F32)
    return ((x - mean.reshape(1, -1, 1, 1)) * inv.reshape(1, -1, 1, 1) * weight.reshape(1, -1, 1, 1)
            + bias.reshape(1, -1, 1, 1)).astype(F32)


def relu(x):
    return np.maximum(x, F32(0))


def maxpool(x, kh, kw):
    N, C, H, W = x.shape
    return x[:, :, : H // kh * kh, : W // kw * kw].reshape(N, C, H // kh, kh, W // kw, kw).max(axis=(3, 5))


def sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(F32)))).astype(F32)


def sequence_se(x, sd, prefix):
    """SequenceSE.forward, se_model.py:19-30: mean over H -> Conv1d(C, C/16, 1) -> ReLU ->
    Conv1d(C/16, C, 1) -> Sigmoid -> broadcast multiply over H."""
    w0 = sd[prefix + ".fc.0.weight"][:, :, 0]
    b0 = sd[prefix + ".fc.0.bias"]
    w2 = sd[prefix + ".fc.2.weight"][:, :, 0]
    b2 = sd[prefix + ".fc.2.bias"]
    y = x.mean(axis=2, dtype=F32)                                   # (N, C, W)
    z = relu(np.einsum("rc,ncw->nrw", w0, y).astype(F32) + b0.reshape(1, -1, 1))
    g = sigmoid(np.einsum("cr,nrw->ncw", w2, z).astype(F32) + b2.reshape(1, -1, 1))
    return (x * g[:, :, None, :]).astype(F32)


def adaptive_avg_pool(x, oh, ow):
    """nn.AdaptiveAvgPool2d((oh, ow)): bin k covers [floor(k*I/O), ceil((k+1)*I/O))."""
    N, C, H, W = x.shape
    out = np.empty((N, C, oh, ow), F32)
    for i in range(oh):
        h0, h1 = (i * H) // oh, -((-(i + 1) * H) // oh)
        for j in range(ow):
            w0, w1 = (j * W) // ow, -((-(j + 1) * W) // ow)
            out[:, :, i, j] = x[:, :, h0:h1, w0:w1].mean(axis=(2, 3), dtype=F32)
    return out


def layer_norm(x, w, b):
    mu = x.mean(-1, keepdims=True, dtype=F32)
    var = ((x - mu) ** 2).mean(-1, keepdims=True, dtype=F32)
    return ((x - mu) / np.sqrt(var + F32(LN_EPS)) * w + b).astype(F32)


def linear(x, w, b=None):
    y = x.astype(F32) @ w.T.astype(F32)
    if b is not None:
        y = y + b
    return y.astype(F32)


def softmax_lastdim(s):
    m = s.max(-1, keepdims=True)
    e = np.exp(s - m)
    return (e / e.sum(-1, keepdims=True)).astype(F32)


def mha(q_in, k_in, v_in, in_w, in_b, out_w, out_b, add_mask=None):
    """nn.MultiheadAttention (batch-less helper): q_in (Lq, D), k_in/v_in (Lk, D).
    add_mask: additive float (Lq, Lk) or None.  8 heads, scale 1/sqrt(D/8)."""
    D = q_in.shape[-1]
    hd = D // NHEAD
    q = linear(q_in, in_w[:D], in_b[:D]).reshape(-1, NHEAD, hd).transpose(1, 0, 2)
    k = linear(k_in, in_w[D:2 * D], in_b[D:2 * D]).reshape(-1, NHEAD, hd).transpose(1, 0, 2)
    v = linear(v_in, in_w[2 * D:], in_b[2 * D:]).reshape(-1, NHEAD, hd).transpose(1, 0, 2)
    s = (q @ k.transpose(0, 2, 1)) * F32(1.0 / math.sqrt(hd))
    if add_mask is not None:
        s = s + add_mask[None]
    p = softmax_lastdim(s)
    o = (p @ v).transpose(1, 0, 2).reshape(-1, D)
    return linear(o, out_w, out_b)


# ----------------------------------------------------------------------------------------------
# Stage 2: CNN backbones
# ----------------------------------------------------------------------------------------------
def _conv_bn_relu(x, sd, name):
    """nn.Sequential(Conv2d, BatchNorm2d, ReLU) - se_model.py:39-52 / vgg_model.py:14-42."""
    y = conv3x3(x, sd[name + ".0.weight"], sd[name + ".0.bias"])
    y = batchnorm_eval(y, sd[name + ".1.weight"], sd[name + ".1.bias"],
                       sd[name + ".1.running_mean"], sd[name + ".1.running_var"])
    return relu(y)


def _bn(x, sd, p):
    return batchnorm_eval(x, sd[p + ".weight"], sd[p + ".bias"], sd[p + ".running_mean"], sd[p + ".running_var"])


def basic_block(x, sd, p):
    """BasicBlock.forward (model/resnet_model.py:28-37): relu(bn1(conv1 x)) -> bn2(conv2 .) + shortcut(x) -> relu; the
    shortcut is a 1x1 conv + BN where the channel count changes (:20-26), identity otherwise.  Convs have no bias."""
    zero = lambda w: np.zeros(w.shape[0], F32)
    w1, w2 = sd[p + ".conv1.weight"], sd[p + ".conv2.weight"]
    out = relu(_bn(conv3x3(x, w1, zero(w1)), sd, p + ".bn1"))
    out = _bn(conv3x3(out, w2, zero(w2)), sd, p + ".bn2")
    if p + ".shortcut.0.weight" in sd:
        ws = sd[p + ".shortcut.0.weight"][:, :, 0, 0]
        sc = _bn(np.einsum("oc,nchw->nohw", ws, x).astype(F32), sd, p + ".shortcut.1")
    else:
        sc = x
    return relu(out + sc).astype(F32)


def resnet_forward(sd, chunks, taps=None):
    """ResNetFeatureExtractor.forward (model/resnet_model.py:75-91)."""
    def tap(k, v):
        if taps is not None:
            taps[k] = v
        return v
    w = sd["cnn.conv1.weight"]
    x = tap("pool1", maxpool(relu(_bn(conv3x3(chunks, w, np.zeros(64, F32)), sd, "cnn.bn1")), 2, 2))
    x = tap("layer1", basic_block(x, sd, "cnn.layer1.0"))
    x = tap("pool2", maxpool(x, 2, 2))
    x = tap("layer2", basic_block(basic_block(x, sd, "cnn.layer2.0"), sd, "cnn.layer2.1"))
    x = tap("pool3", maxpool(x, 2, 1))
    x = tap("layer3", basic_block(basic_block(x, sd, "cnn.layer3.0"), sd, "cnn.layer3.1"))
    x = tap("pool4", maxpool(x, 2, 1))
    x = tap("layer4", basic_block(x, sd, "cnn.layer4.0"))
    return tap("final_pool", adaptive_avg_pool(x, 2, 32))


def cnn_forward(sd, chunks, variant="se", taps=None):
    """ImprovedFeatureExtractor.forward (se_model.py:63-79) or the VGG baseline
    (vgg_model.py:50-59; no SE, conv7 without BN/ReLU).  chunks: (N,1,48,100) -> (N,512,2,32).
    `taps`, if a dict, receives the intermediate activations by name."""
    if variant == "resnet":
        return resnet_forward(sd, chunks, taps)

    def tap(k, v):
        if taps is not None:
            taps[k] = v
        return v
    p = "cnn."
    x = tap("pool1", maxpool(_conv_bn_relu(chunks, sd, p + "conv1"), 2, 2))
    x = tap("pool2", maxpool(_conv_bn_relu(x, sd, p + "conv2"), 2, 2))
    x = tap("conv3", _conv_bn_relu(x, sd, p + "conv3"))
    x = tap("conv4", _conv_bn_relu(x, sd, p + "conv4"))
    if variant == "se":
        x = sequence_se(x, sd, p + "se3")
    x = tap("pool3", maxpool(x, 2, 1))
    x = tap("conv5", _conv_bn_relu(x, sd, p + "conv5"))
    x = tap("conv6", _conv_bn_relu(x, sd, p + "conv6"))
    if variant == "se":
        x = sequence_se(x, sd, p + "se4")
    x = tap("pool4", maxpool(x, 2, 1))
    x = conv3x3(x, sd[p + "conv7.weight"], sd[p + "conv7.bias"])
    if variant == "se":
        x = relu(batchnorm_eval(x, sd[p + "bn7.weight"], sd[p + "bn7.bias"],
                                sd[p + "bn7.running_mean"], sd[p + "bn7.running_var"]))
        x = sequence_se(x, sd, p + "se5")
    tap("conv7", x)
    return tap("final_pool", adaptive_avg_pool(x, 2, 32))


# ----------------------------------------------------------------------------------------------
# Stage 3 + 4: patch projection and per-chunk Transformer encoder
# ----------------------------------------------------------------------------------------------
def patch_forward(sd, f):
    """PatchEncoder.forward, se_model.py:104-117: Conv2d(512->D, (2,1), stride (2,1)) over the two
    pooled rows == GEMM with K = 512*2, then + pos_emb[:32].  f: (N,512,2,32) -> (N,32,D)."""
    w = sd["patch.proj.weight"]            # (D, 512, 2, 1)
    D = w.shape[0]
    N = f.shape[0]
    a = f.transpose(0, 3, 1, 2).reshape(N * 32, 512 * 2)          # (N*32, c*2+kh)
    y = linear(a, w.reshape(D, 1024), sd["patch.proj.bias"]).reshape(N, 32, D)
    return (y + sd["patch.pos_emb"][:32][None]).astype(F32)


def encoder_layer(sd, x, pre):
    """nn.TransformerEncoderLayer (post-norm, relu, eval): x (L, D) for one chunk."""
    a = mha(x, x, x, sd[pre + "self_attn.in_proj_weight"], sd[pre + "self_attn.in_proj_bias"],
            sd[pre + "self_attn.out_proj.weight"], sd[pre + "self_attn.out_proj.bias"])
    x = layer_norm(x + a, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
    ff = linear(relu(linear(x, sd[pre + "linear1.weight"], sd[pre + "linear1.bias"])),
                sd[pre + "linear2.weight"], sd[pre + "linear2.bias"])
    return layer_norm(x + ff, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])


def encoder_forward(sd, p):
    """make_encoder(num_layers=2) applied per chunk (attention never crosses chunks because the
    reference feeds (L=32, N, D) seq-first, predictor.py:169-170).  p: (N,32,D) -> (N,32,D)."""
    out = np.empty_like(p)
    for n in range(p.shape[0]):
        x = p[n]
        for l in range(2):
            x = encoder_layer(sd, x, f"enc.layers.{l}.")
        out[n] = x
    return out


# ----------------------------------------------------------------------------------------------
# Stage 5: merge, BiLSTM, decoder, greedy loop
# ----------------------------------------------------------------------------------------------
def merge_line(sd, enc_line):
    """predictor.py:174-183: reshape (n,32,D) -> (T,D), truncate to global_pos rows, add global_pos."""
    n, L, D = enc_line.shape
    merged = enc_line.reshape(n * L, D)
    limit = min(n * L, sd["global_pos"].shape[0])
    return (merged[:limit] + sd["global_pos"][:limit]).astype(F32)


def lstm_direction(x, w_ih, w_hh, b_ih, b_hh, reverse=False):
    """One direction of nn.LSTM: gates (i,f,g,o), zero initial state.  x: (T, D) -> (T, Hh)."""
    T = x.shape[0]
    Hh = w_hh.shape[1]
    gi = linear(x, w_ih, b_ih + b_hh)
    h = np.zeros(Hh, F32)
    c = np.zeros(Hh, F32)
    out = np.empty((T, Hh), F32)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        g = gi[t] + w_hh.astype(F32) @ h
        i = sigmoid(g[:Hh]); f = sigmoid(g[Hh:2 * Hh])
        gg = np.tanh(g[2 * Hh:3 * Hh]).astype(F32); o = sigmoid(g[3 * Hh:])
        c = (f * c + i * gg).astype(F32)
        h = (o * np.tanh(c)).astype(F32)
        out[t] = h
    return out


def bilstm(sd, merged):
    """context_bilstm (se_model.py:228-234; predictor.py:185-186): concat(fwd, bwd)."""
    p = "context_bilstm."
    f = lstm_direction(merged, sd[p + "weight_ih_l0"], sd[p + "weight_hh_l0"],
                       sd[p + "bias_ih_l0"], sd[p + "bias_hh_l0"])
    b = lstm_direction(merged, sd[p + "weight_ih_l0_reverse"], sd[p + "weight_hh_l0_reverse"],
                       sd[p + "bias_ih_l0_reverse"], sd[p + "bias_hh_l0_reverse"], reverse=True)
    return np.concatenate([f, b], axis=1)


def memory_for_line(sd, enc_line, variant="se"):
    m = merge_line(sd, enc_line)
    return bilstm(sd, m) if variant == "se" else m


def decoder_forward(sd, tokens, memory):
    """TransformerDecoderWrapper.forward (se_model.py:182-208) for one line: tokens list[int] (t),
    memory (T, D) with an all-False memory mask (predictor.py:87).  Returns logits (t, V)."""
    t = len(tokens)
    tok = np.asarray(tokens, np.int64)
    emb = sd["dec.tok_emb.weight"][tok]     # padding_idx only freezes row 0 (zero at init), se_model.py:167
    x = (emb + sd["dec.pos_emb"][:t]).astype(F32)
    causal = np.where(np.arange(t)[None, :] > np.arange(t)[:, None], -np.inf, 0.0).astype(F32)
    keypad = np.where(tok == PAD, -np.inf, 0.0).astype(F32)[None, :]      # se_model.py:190
    self_mask = causal + keypad
    for l in range(2):
        pre = f"dec.decoder.layers.{l}."
        a = mha(x, x, x, sd[pre + "self_attn.in_proj_weight"], sd[pre + "self_attn.in_proj_bias"],
                sd[pre + "self_attn.out_proj.weight"], sd[pre + "self_attn.out_proj.bias"], self_mask)
        x = layer_norm(x + a, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
        a = mha(x, memory, memory, sd[pre + "multihead_attn.in_proj_weight"],
                sd[pre + "multihead_attn.in_proj_bias"], sd[pre + "multihead_attn.out_proj.weight"],
                sd[pre + "multihead_attn.out_proj.bias"])
        x = layer_norm(x + a, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
        ff = linear(relu(linear(x, sd[pre + "linear1.weight"], sd[pre + "linear1.bias"])),
                    sd[pre + "linear2.weight"], sd[pre + "linear2.bias"])
        x = layer_norm(x + ff, sd[pre + "norm3.weight"], sd[pre + "norm3.bias"])
    return linear(x, sd["dec.out_proj.weight"], sd["dec.out_proj.bias"])


def forward_teacher_forced(sd, enc_lines, tgt_tokens, variant="se"):
    """KhmerOCR.forward (se_model.py:240-289; vgg_model.py:214-246 without the BiLSTM) AFTER the per-chunk encoder:
    enc_lines = list of (n_i, 32, D) encoder outputs, tgt_tokens = int array (B, L).  Training-time semantics:
    merged sequences zero-padded to Tmax (pad_sequence :262), global_pos added to EVERY row incl. the pads (:265-273),
    BiLSTM over all Tmax positions of every line (:278-279: no packing, so the backward direction reads the pad rows
    first), memory_key_padding_mask = positions >= T_i (:282-285), decoder with causal + <pad>-key masks.
    Returns logits (B, L, V)."""
    B = len(enc_lines)
    D = enc_lines[0].shape[-1]
    merged = [e.reshape(-1, D) for e in enc_lines]
    T = [m.shape[0] for m in merged]
    limit = min(max(T), sd["global_pos"].shape[0])
    out = []
    for b in range(B):
        mem = np.zeros((limit, D), F32)
        tb = min(T[b], limit)
        mem[:tb] = merged[b][:tb]
        mem = (mem + sd["global_pos"][:limit]).astype(F32)
        if variant == "se":
            mem = bilstm(sd, mem)
        tok = np.asarray(tgt_tokens[b], np.int64)
        L = tok.shape[0]
        x = (sd["dec.tok_emb.weight"][tok] + sd["dec.pos_emb"][:L]).astype(F32)
        causal = np.where(np.arange(L)[None, :] > np.arange(L)[:, None], -np.inf, 0.0).astype(F32)
        self_mask = causal + np.where(tok == PAD, -np.inf, 0.0).astype(F32)[None, :]
        cross_mask = np.broadcast_to(np.where(np.arange(limit) >= tb, -np.inf, 0.0).astype(F32)[None, :], (L, limit))
        for l in range(2):
            pre = f"dec.decoder.layers.{l}."
            a = mha(x, x, x, sd[pre + "self_attn.in_proj_weight"], sd[pre + "self_attn.in_proj_bias"],
                    sd[pre + "self_attn.out_proj.weight"], sd[pre + "self_attn.out_proj.bias"], self_mask)
            x = layer_norm(x + a, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
            a = mha(x, mem, mem, sd[pre + "multihead_attn.in_proj_weight"], sd[pre + "multihead_attn.in_proj_bias"],
                    sd[pre + "multihead_attn.out_proj.weight"], sd[pre + "multihead_attn.out_proj.bias"], cross_mask)
            x = layer_norm(x + a, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
            ff = linear(relu(linear(x, sd[pre + "linear1.weight"], sd[pre + "linear1.bias"])),
                        sd[pre + "linear2.weight"], sd[pre + "linear2.bias"])
            x = layer_norm(x + ff, sd[pre + "norm3.weight"], sd[pre + "norm3.bias"])
        out.append(linear(x, sd["dec.out_proj.weight"], sd["dec.out_proj.bias"]))
    return np.stack(out)


def greedy_decode(sd, memory, max_len=256, return_logits=False):
    """OCRPredictor._greedy_decode (predictor.py:85-99): start [sos]; up to `max_len` iterations;
    argmax of the last position (ties -> lowest index); stop BEFORE appending eos.
    The prefix is re-run every step exactly like the reference (no cache)."""
    generated = [SOS]
    last_logits = []
    for _ in range(max_len):
        logits = decoder_forward(sd, generated, memory)
        nxt = int(np.argmax(logits[-1]))
        last_logits.append(logits[-1])
        if nxt == EOS:
            break
        generated.append(nxt)
    if return_logits:
        return generated, np.stack(last_logits)
    return generated


def log_softmax(x):
    """F.log_softmax(x, dim=-1) in fp32 (predictor.py:115)."""
    x = np.asarray(x, F32)
    z = x - x.max(axis=-1, keepdims=True)
    return (z - np.log(np.exp(z).sum(axis=-1, keepdims=True, dtype=F32))).astype(F32)


def beam_search(sd, memory, beam_width=3, max_len=256):
    """OCRPredictor._beam_search (predictor.py:101-136).  Scores are Python floats (sums of `.item()` values);
    candidates = top-`beam_width` tokens of every live hypothesis, sorted by score with Python's STABLE sort
    (ties keep hypothesis order, then top-k rank); a candidate ending in eos is moved to `completed` with its
    score divided by len(seq) INCLUDING sos (:127) wherever it ranks, the first `beam_width` others survive;
    result = best completed hypothesis, else the first live one.  Returns the id list (sos first, eos last if
    completed) that the reference passes to Tokenizer.decode."""
    beams = [(0.0, [SOS])]
    completed = []
    for _ in range(max_len):
        logp = np.stack([log_softmax(decoder_forward(sd, seq, memory)[-1]) for _, seq in beams])
        candidates = []
        for i, (score, seq) in enumerate(beams):
            order = np.argsort(-logp[i], kind="stable")[:beam_width]          # topk: descending values
            for k in order:
                candidates.append((score + float(logp[i, k]), seq + [int(k)]))
        candidates.sort(key=lambda c: c[0], reverse=True)
        nxt = []
        for s, seq in candidates:
            if seq[-1] == EOS:
                completed.append((s / len(seq), seq))
            elif len(nxt) < beam_width:
                nxt.append((s, seq))
        beams = nxt
        if not beams:
            break
    if completed:
        return sorted(completed, key=lambda c: c[0], reverse=True)[0][1]
    return beams[0][1]


def tokens_to_text(ids, idx2char):
    """Tokenizer.decode, tokenizer.py:26-35."""
    out = []
    for i in ids:
        if i == SOS or i == PAD:
            continue
        if i == EOS:
            break
        out.append(idx2char.get(int(i), ""))
    return "".join(out)


# ----------------------------------------------------------------------------------------------
# Whole path (predict_batch, predictor.py:138-199) for grey uint8 line images
# ----------------------------------------------------------------------------------------------
def recognise_lines(sd, images_u8, variant="se", max_len=256, batch_size=8, stages=None):
    """Returns list[list[int]] of generated ids (including the leading sos, excluding eos)."""
    results = []
    for i in range(0, len(images_u8), batch_size):
        group = images_u8[i:i + batch_size]
        chunk_list = [preprocess_gray(im)[1] for im in group]
        counts = [c.shape[0] for c in chunk_list]
        batch = np.concatenate(chunk_list, 0)
        enc = encoder_forward(sd, patch_forward(sd, cnn_forward(sd, batch, variant)))
        cur = 0
        for n in counts:
            mem = memory_for_line(sd, enc[cur:cur + n], variant)
            cur += n
            results.append(greedy_decode(sd, mem, max_len))
            if stages is not None:
                stages.append({"enc": enc[cur - n:cur], "memory": mem})
    return results


def cer(pred: str, ref: str) -> float:
    """Character error rate = Levenshtein(pred, ref) / len(ref) (training notebook cell 19)."""
    if len(ref) == 0:
        return 0.0 if len(pred) == 0 else 1.0
    prev = list(range(len(ref) + 1))
    for i, pc in enumerate(pred, 1):
        cur = [i]
        for j, rc in enumerate(ref, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (pc != rc)))
        prev = cur
    return prev[-1] / len(ref)
