"""Which side does the REAL reference (torch CPU, baseline/_ref) take on the lines where the CUDA path and the numpy oracle
disagree?  Reads profiles/r01/parity_c3full.json (mismatching lines), decodes those lines with the unmodified reference and
reports, per line, whether its tokens equal the oracle's, the CUDA path's, or neither.
  python tests/parity/parity_reference_check.py [parity json] [cuda tokens npz]"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from baseline import run_reference as R
from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint
from khmer_ocr_cnn_transformer_b200.recognition.tokenizer import Tokenizer
from workloads import synth

pj = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "profiles/r01/parity_c3full.json"
cz = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "profiles/r01/c3full_cuda_tokens.npz"
par = json.loads(pj.read_text())
lines = [m["line"] for m in par["mismatches"]]
o = np.load(ROOT / "tests/golden/oracle_tokens_c3full.npz")
g = np.load(cz)
imgs, _ = synth.make_lines(max(lines) + 1, 200, 1600, seed=3)
pred = R.load_predictor(load_checkpoint(ROOT / "tests/golden/fixture_se_ckpt.npz"), "cpu")
tok = Tokenizer(ROOT / "khmer_ocr_cnn_transformer_b200/recognition/char2idx.json")
dec = lambda z, i: tok.decode([int(t) for t in z["tokens"][i, :z["lengths"][i]]])
out = []
for i in lines:
    _, _, texts = R.time_predict_batch(pred, [imgs[i]])
    out.append({"line": i, "reference_equals_oracle": texts[0] == dec(o, i), "reference_equals_cuda": texts[0] == dec(g, i)})
    print(out[-1], flush=True)
(ROOT / "profiles" / "r02").mkdir(exist_ok=True)
(ROOT / "profiles/r02/parity_c3full_reference_check.json").write_text(json.dumps(
    {"what": "unmodified reference (torch CPU fp32) on the lines where CUDA and the numpy oracle disagree", "source": str(pj.name), "lines": out}, indent=1))
