"""Public API: recognize / recognize_batch / CLI, same names, defaults, singleton and error
behaviour as the reference (netra_ocr/recognition/recognize_text.py:23-128)."""
import argparse
import os
import sys

from .config import OCRConfig
from .utils import setup_logging, autodetect_config
from .tokenizer import Tokenizer
from .predictor import OCRPredictor
from .model.se_model import KhmerOCR as SE_KhmerOCR
from .model.vgg_model import KhmerOCR as VGG_KhmerOCR
from .model.resnet_model import KhmerOCR as ResNet_KhmerOCR

CURRENT_DIR = os.path.dirname(os.path.abspath(__file__))
DEFAULT_MODEL_PATH = os.path.join(CURRENT_DIR, "weight", "khmerocr_se_transformer.pth")
DEFAULT_VOCAB_PATH = os.path.join(CURRENT_DIR, "char2idx.json")

_PREDICTOR_INSTANCE = None


def _get_predictor(model_path=None, vocab_path=None):
    """Load the model once (process-global singleton; later model_path arguments are ignored once
    cached, exactly like the reference, recognize_text.py:46-47)."""
    global _PREDICTOR_INSTANCE
    model_path = model_path or DEFAULT_MODEL_PATH
    vocab_path = vocab_path or DEFAULT_VOCAB_PATH
    if "vgg" in str(model_path).lower():
        model = VGG_KhmerOCR
    elif "resnet" in str(model_path).lower():
        model = ResNet_KhmerOCR
    else:
        model = SE_KhmerOCR
    if _PREDICTOR_INSTANCE is not None:
        return _PREDICTOR_INSTANCE
    try:
        detected_cfg = autodetect_config(model_path)
        config = OCRConfig(**detected_cfg)
        tokenizer = Tokenizer(vocab_path)
        _PREDICTOR_INSTANCE = OCRPredictor(model_path=model_path, tokenizer=tokenizer, config=config,
                                           model_class=model)
        return _PREDICTOR_INSTANCE
    except Exception as e:
        print(f"Failed to load model: {e}")
        sys.exit(1)


def recognize(image_input, beam_width: int = 3, model_path=None, vocab_path=None) -> str:
    predictor = _get_predictor(model_path, vocab_path)
    try:
        return predictor.predict(image_input, beam_width=beam_width)
    except Exception as e:
        print(f"Prediction error: {e}")
        return ""


def recognize_batch(image_list: list, beam_width: int = 1, batch_size: int = 8, model_path=None,
                    vocab_path=None) -> list:
    if not image_list:
        return []
    predictor = _get_predictor(model_path, vocab_path)
    try:
        return predictor.predict_batch(image_list, beam_width=beam_width, batch_size=batch_size)
    except Exception as e:
        print(f"Batch prediction error: {e}")
        return [recognize(img, beam_width, model_path, vocab_path) for img in image_list]


def main():
    setup_logging()
    parser = argparse.ArgumentParser(description="Khmer OCR Inference Pipeline")
    parser.add_argument("--image", type=str, required=True, help="Path to input image")
    parser.add_argument("--model", type=str, default=DEFAULT_MODEL_PATH, help="Path to .pth")
    parser.add_argument("--vocab", type=str, default=DEFAULT_VOCAB_PATH, help="Path to vocab json")
    parser.add_argument("--beam", type=int, default=3, help="Beam width (1 for greedy)")
    parser.add_argument("--output", type=str, help="Save result to text file")
    args = parser.parse_args()
    text = recognize(args.image, args.beam, args.model, args.vocab)
    print("\n" + "=" * 40)
    print(f"RESULT: {text}")
    print("=" * 40 + "\n")
    if args.output:
        with open(args.output, "w", encoding="utf-8") as f:
            f.write(text)
        print(f"Saved to {args.output}")


if __name__ == "__main__":
    main()
