"""Checkpoint layout of the recogniser (the on-disk format the drop-in keeps loadable).

The key names and shapes are the reference's `state_dict` layout (SURVEY.md §8 a-0;
reference: netra_ocr/recognition/model/se_model.py:35-79,81-117,119-126,162-181,210-238 and
model/vgg_model.py:5-59,198-212).  Nothing here touches the GPU.

`seeded_state_dict` builds a deterministic numpy-RNG checkpoint (rescaled init, perturbed
BatchNorm running statistics so that BN folding is actually exercised) for tests and smoke runs;
`load_checkpoint` reads a reference `.pth` (bare state_dict or {'model_state_dict': ...},
predictor.py:38-40) or a compact `.npz` fixture written by tests/golden/make_fixtures.py.
"""
from __future__ import annotations

from pathlib import Path
import numpy as np

VOCAB_SIZE = 124
NHEAD = 8
ENC_FF = 1024
# BasicBlocks of the ResNet baseline in execution order: (state_dict prefix under cnn., in channels, out channels)
RESNET_BLOCKS = [("layer1.0", 64, 128), ("layer2.0", 128, 256), ("layer2.1", 256, 256), ("layer3.0", 256, 512),
                 ("layer3.1", 512, 512), ("layer4.0", 512, 512)]


def state_dict_spec(variant: str = "se", emb_dim: int = 384, max_global_len: int = 4096,
                    vocab_size: int = VOCAB_SIZE, dec_max_len: int = 256) -> dict:
    """Ordered {name: shape} of every floating-point entry of the reference state_dict
    (`num_batches_tracked` int64 scalars are accepted on load and ignored)."""
    assert variant in ("se", "vgg", "resnet")
    D = emb_dim
    spec: dict[str, tuple] = {"global_pos": (max_global_len, D)}
    if variant == "resnet":
        # ResNetFeatureExtractor (model/resnet_model.py:37-91): conv1/bn1, then BasicBlocks layer1.0, layer2.{0,1},
        # layer3.{0,1}, layer4.0 (conv bias=False; a 1x1 conv + BN shortcut where the channel count changes)
        def bn(p, c):
            for n in ("weight", "bias", "running_mean", "running_var"):
                spec[f"{p}.{n}"] = (c,)
        spec["cnn.conv1.weight"] = (64, 1, 3, 3)
        bn("cnn.bn1", 64)
        for name, cin, cout in RESNET_BLOCKS:
            p = f"cnn.{name}"
            spec[p + ".conv1.weight"] = (cout, cin, 3, 3)
            bn(p + ".bn1", cout)
            spec[p + ".conv2.weight"] = (cout, cout, 3, 3)
            bn(p + ".bn2", cout)
            if cin != cout:
                spec[p + ".shortcut.0.weight"] = (cout, cin, 1, 1)
                bn(p + ".shortcut.1", cout)
    chans = [1, 64, 128, 256, 256, 512, 512]
    for i in (range(1, 7) if variant != "resnet" else ()):
        ci, co = chans[i - 1], chans[i]
        p = f"cnn.conv{i}"
        spec[p + ".0.weight"] = (co, ci, 3, 3)
        spec[p + ".0.bias"] = (co,)
        for n in ("weight", "bias", "running_mean", "running_var"):
            spec[f"{p}.1.{n}"] = (co,)
        if variant == "se" and i in (4, 6):
            s = "cnn.se3" if i == 4 else "cnn.se4"
            spec[s + ".fc.0.weight"] = (co // 16, co, 1)
            spec[s + ".fc.0.bias"] = (co // 16,)
            spec[s + ".fc.2.weight"] = (co, co // 16, 1)
            spec[s + ".fc.2.bias"] = (co,)
    if variant != "resnet":
        spec["cnn.conv7.weight"] = (512, 512, 3, 3)
        spec["cnn.conv7.bias"] = (512,)
    if variant == "se":
        for n in ("weight", "bias", "running_mean", "running_var"):
            spec[f"cnn.bn7.{n}"] = (512,)
        spec["cnn.se5.fc.0.weight"] = (32, 512, 1)
        spec["cnn.se5.fc.0.bias"] = (32,)
        spec["cnn.se5.fc.2.weight"] = (512, 32, 1)
        spec["cnn.se5.fc.2.bias"] = (512,)
    spec["patch.pos_emb"] = (256, D)
    spec["patch.proj.weight"] = (D, 512, 2, 1)
    spec["patch.proj.bias"] = (D,)

    def attn(p):
        spec[p + ".in_proj_weight"] = (3 * D, D)
        spec[p + ".in_proj_bias"] = (3 * D,)
        spec[p + ".out_proj.weight"] = (D, D)
        spec[p + ".out_proj.bias"] = (D,)

    def ffn_norms(p, ff, nn):
        spec[p + ".linear1.weight"] = (ff, D)
        spec[p + ".linear1.bias"] = (ff,)
        spec[p + ".linear2.weight"] = (D, ff)
        spec[p + ".linear2.bias"] = (D,)
        for k in range(1, nn + 1):
            spec[f"{p}.norm{k}.weight"] = (D,)
            spec[f"{p}.norm{k}.bias"] = (D,)

    for l in range(2):
        p = f"enc.layers.{l}"
        attn(p + ".self_attn")
        ffn_norms(p, ENC_FF, 2)
    if variant == "se":
        Hh = D // 2
        for suf in ("", "_reverse"):
            spec[f"context_bilstm.weight_ih_l0{suf}"] = (4 * Hh, D)
            spec[f"context_bilstm.weight_hh_l0{suf}"] = (4 * Hh, Hh)
            spec[f"context_bilstm.bias_ih_l0{suf}"] = (4 * Hh,)
            spec[f"context_bilstm.bias_hh_l0{suf}"] = (4 * Hh,)
    spec["dec.pos_emb"] = (dec_max_len, D)
    spec["dec.tok_emb.weight"] = (vocab_size, D)
    for l in range(2):
        p = f"dec.decoder.layers.{l}"
        attn(p + ".self_attn")
        attn(p + ".multihead_attn")
        ffn_norms(p, 4 * D, 3)
    spec["dec.out_proj.weight"] = (vocab_size, D)
    spec["dec.out_proj.bias"] = (vocab_size,)
    return spec


def seeded_state_dict(variant: str = "se", seed: int = 0, emb_dim: int = 384,
                      max_global_len: int = 1024, gain: float = 1.0) -> dict:
    """Deterministic fp32 checkpoint from numpy's PCG64 (identical on every machine).

    Weights ~ U(-a, a) with a = gain*sqrt(3/fan_in) (variance-preserving-ish); biases small;
    BatchNorm gamma in [0.8, 1.2], beta in [-0.1, 0.1], running_mean in [-0.2, 0.2],
    running_var in [0.5, 1.5]; LayerNorm gamma near 1; embeddings normal(0, 0.02) except
    dec.pos_emb normal(0, 0.1) (se_model.py:171-172); tok_emb row 0 zero (padding_idx)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}
    for name, shape in state_dict_spec(variant, emb_dim, max_global_len).items():
        if name.endswith("running_var"):
            v = rng.uniform(0.5, 1.5, shape)
        elif name.endswith("running_mean"):
            v = rng.uniform(-0.2, 0.2, shape)
        elif ((".1.weight" in name and name.startswith("cnn.conv")) or name == "cnn.bn7.weight"
              or (variant == "resnet" and name.startswith("cnn.") and len(shape) == 1 and name.endswith(".weight"))):
            v = rng.uniform(0.8, 1.2, shape)
        elif ((".1.bias" in name and name.startswith("cnn.conv")) or name == "cnn.bn7.bias"
              or (variant == "resnet" and name.startswith("cnn.") and len(shape) == 1 and name.endswith(".bias"))):
            v = rng.uniform(-0.1, 0.1, shape)
        elif "norm" in name and name.endswith("weight"):
            v = rng.uniform(0.9, 1.1, shape)
        elif "norm" in name and name.endswith("bias"):
            v = rng.uniform(-0.05, 0.05, shape)
        elif name in ("global_pos", "patch.pos_emb"):
            v = rng.normal(0.0, 0.02, shape)
        elif name == "dec.pos_emb":
            v = rng.normal(0.0, 0.1, shape)
        elif name == "dec.tok_emb.weight":
            v = rng.normal(0.0, 1.0, shape)
            v[0] = 0.0
        elif name.endswith("bias") or "bias_" in name:
            v = rng.uniform(-0.05, 0.05, shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            a = gain * np.sqrt(3.0 / fan_in) * (np.sqrt(2.0) if name.startswith("cnn.") and len(shape) == 4 else 1.0)
            v = rng.uniform(-a, a, shape)
        sd[name] = np.ascontiguousarray(v, dtype=np.float32)
    return sd


def detect_variant(sd: dict) -> str:
    if "cnn.layer1.0.conv1.weight" in sd:
        return "resnet"
    return "se" if "context_bilstm.weight_ih_l0" in sd else "vgg"


def load_checkpoint(path) -> dict:
    """Return {name: np.float32 array}.  `.npz` -> compact fixture (fp16/fp32 arrays);
    anything else -> torch.load of a reference checkpoint (predictor.py:38-40)."""
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(f"Model not found at {path}")
    if path.suffix == ".npz":
        with np.load(path) as z:
            return {k: np.ascontiguousarray(z[k], dtype=np.float32) for k in z.files}
    import torch
    ckpt = torch.load(path, map_location="cpu")
    state = ckpt.get("model_state_dict", ckpt)
    out = {}
    for k, v in state.items():
        if k.endswith("num_batches_tracked"):
            continue
        out[k] = np.ascontiguousarray(v.detach().to(torch.float32).numpy())
    return out


def validate_state_dict(sd: dict, variant: str | None = None) -> tuple[str, int, int, int]:
    """Check names/shapes against the spec; returns (variant, emb_dim, max_global_len, dec_max_len)."""
    variant = variant or detect_variant(sd)
    if "global_pos" not in sd or "dec.pos_emb" not in sd:
        raise KeyError("checkpoint lacks global_pos / dec.pos_emb")
    max_global_len, emb_dim = sd["global_pos"].shape
    dec_max_len = sd["dec.pos_emb"].shape[0]
    spec = state_dict_spec(variant, emb_dim, max_global_len, sd["dec.tok_emb.weight"].shape[0], dec_max_len)
    for name, shape in spec.items():
        if name not in sd:
            raise KeyError(f"checkpoint lacks {name}")
        if tuple(sd[name].shape) != tuple(shape):
            raise ValueError(f"{name}: shape {tuple(sd[name].shape)} != expected {shape}")
    return variant, int(emb_dim), int(max_global_len), int(dec_max_len)
