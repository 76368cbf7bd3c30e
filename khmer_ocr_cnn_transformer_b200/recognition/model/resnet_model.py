"""Model-class marker for the ResNet-Transformer baseline (BasicBlocks with 1x1-conv shortcuts, no SE, no BiLSTM;
reference: model/resnet_model.py:5-91,218-240, selected by "resnet" in the checkpoint name, recognize_text.py:41-42).
The residual blocks run on the same tcgen05 implicit-GEMM kernel: the shortcut is the GEMM's fp32 addend."""


class KhmerOCR:
    variant = "resnet"

    def __init__(self, vocab_size, pad_idx=0, emb_dim=256, max_global_len=4096):
        self.vocab_size = vocab_size
        self.pad_idx = pad_idx
        self.emb_dim = emb_dim
        self.max_global_len = max_global_len
