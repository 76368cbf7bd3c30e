"""Model-class marker for the baseline VGG-Transformer (no SE, no BiLSTM, conv7 without BN/ReLU;
reference: model/vgg_model.py:5-59,198-212).  Same kernels as the SE model with the flags off."""


class KhmerOCR:
    variant = "vgg"

    def __init__(self, vocab_size, pad_idx=0, emb_dim=256, max_global_len=4096):
        self.vocab_size = vocab_size
        self.pad_idx = pad_idx
        self.emb_dim = emb_dim
        self.max_global_len = max_global_len
