"""Model-class marker for the SE-VGG-Transformer recogniser (the "proposed model").

The reference passes an nn.Module class to `OCRPredictor(..., model_class=...)`
(recognize_text.py:39-44, predictor.py:26-31; class at model/se_model.py:210-238).  Here the class
only records the constructor arguments and names the variant; the math lives in libkocr_b200.so."""


class KhmerOCR:
    variant = "se"

    def __init__(self, vocab_size, pad_idx=0, emb_dim=256, max_global_len=4096):
        self.vocab_size = vocab_size
        self.pad_idx = pad_idx
        self.emb_dim = emb_dim
        self.max_global_len = max_global_len
