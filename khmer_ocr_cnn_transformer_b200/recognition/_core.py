"""Access to the parent package (checkpoint / weights / _native / scheduling / textline_crops) that works under BOTH
import conventions:

  * `khmer_ocr_cnn_transformer_b200.recognition....`  (this repository used as a normal package), and
  * `recognition....` as a TOP-LEVEL package - the reference's own convention: `netra_ocr/ocr_engine.py:6-10` puts its
    package directory on `sys.path` and then does `from recognition.recognize_text import recognize_batch`.

A relative `from .. import x` fails in the second case ("attempted relative import beyond top-level package"), so the
modules of this directory reach their siblings' parent through the names exported here."""
import importlib
import sys
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent.parent            # .../khmer_ocr_cnn_transformer_b200

if __name__.count(".") >= 2:                                   # <parent package>.recognition._core
    _ROOT = __name__.rsplit(".", 2)[0]
else:                                                          # recognition._core: the parent is not on the import chain
    if str(_PKG_DIR.parent) not in sys.path:
        sys.path.insert(0, str(_PKG_DIR.parent))
    _ROOT = _PKG_DIR.name


def _sub(name):
    return importlib.import_module(_ROOT + "." + name)


checkpoint = _sub("checkpoint")
weights = _sub("weights")
native = _sub("_native")
scheduling = _sub("scheduling")


def textline_crops():            # imported lazily (predict_page only)
    return _sub("textline_crops")


def pipeline():
    return _sub("pipeline")
