"""Where does the time of a 256-line step go when several batches are in flight?  (diagnosis tool, GPU box)

  python tools/inflight_probe.py [in_flight=12] [steps=36]

Three schedules over the same c2 batches, S host threads / handles / streams each:
  heavy   stages 1-5a only (gather, SE-VGG, encoder, BiLSTM, cross K/V)
  decode  the greedy decode loop only (memory of the batch computed once beforehand)
  full    the whole step (what bench.py times)
If heavy + decode ~ full the GPU is the limit and the two parts do not overlap; if full ~ max(heavy, decode) they do.
"""
import sys
import threading
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from khmer_ocr_cnn_transformer_b200 import _native, weights
from workloads import synth
from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint

ROOT = Path(__file__).resolve().parent.parent
S = int(sys.argv[1]) if len(sys.argv) > 1 else 12
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 36
OPTS = dict(a.split("=") for a in sys.argv[3:])          # e.g. use_pdl=0 dec_wide=0 straggler_threshold=8 compact_rows=1

blob = weights.pack_blob(load_checkpoint(ROOT / "tests/golden/fixture_se_ckpt.npz"))


class W:
    def __init__(self, i):
        self.imgs = synth.make_lines(256, 400, 800, seed=i)[0]
        self.batch = _native.LineBatch(self.imgs)
        self.rec = _native.Recognizer(blob, max_lines=256, max_chunks=2816)
        self.rec.set_option("dec_wide", int(OPTS.get("dec_wide", 0)))
        self.rec.set_option("big_gemm_sms", int(OPTS.get("big_gemm_sms", 0)))
        self.rec.set_option("straggler_threshold", int(OPTS.get("straggler_threshold", 8)))
        self.rec.set_option("blocking_wait", 1 if S > 1 else 0)
        for k in ("use_graphs", "use_pdl", "dec_cross_impl", "compact_rows", "kv_split", "dec_skip", "dec_fused", "dec_lookahead"):
            if k in OPTS:
                self.rec.set_option(k, int(OPTS[k]))
        self.stream = torch.cuda.Stream()
        self.tok = np.zeros((256, _native.TOKENS_LD), np.int32)
        self.ln = np.zeros(256, np.int32)

    def heavy(self):
        st = self.stream.cuda_stream
        self.rec.gather_chunks(self.batch, stream=st)
        self.rec.sevgg_encoder_forward(stream=st)
        self.rec.merge_bilstm_forward(stream=st)
        self.stream.synchronize()

    def decode(self):
        _native.check(self.rec.lib.kocr_decode_greedy(self.rec._h, int(OPTS.get("max_steps", 0)), self.tok.ctypes.data, self.ln.ctypes.data, None))

    def full(self):
        self.rec.recognize_lines(self.batch, tokens_out=self.tok, lengths_out=self.ln)


workers = [W(i) for i in range(S)]
for w in workers:
    w.full()
if "forced" in OPTS:        # every line stays active for exactly max_steps positions (work attribution with dec_skip)
    for w in workers:
        forced = np.full((256, _native.TOKENS_LD), 5, np.int32); forced[:, 0] = 2
        w.rec.set_forced_tokens(forced)
        w.rec.set_option("force_tokens", 1)
        w.rec.set_option("straggler_threshold", 0)
KINDS = OPTS.get("kinds", "heavy,decode,full").split(",")
torch.cuda.synchronize()


def run(kind, steps):
    nxt = {"i": 0}
    lock = threading.Lock()

    def loop(w):
        fn = getattr(w, kind)
        while True:
            with lock:
                if nxt["i"] >= steps:
                    return
                nxt["i"] += 1
            fn()
    ts = [threading.Thread(target=loop, args=(w,)) for w in workers]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / steps


for kind in (KINDS if STEPS > 0 else ()):
    run(kind, S)
    ms = run(kind, STEPS)
    hl = sum(int(w.rec.debug_read("host_launch_us")) for w in workers) / 1e3
    hw = sum(int(w.rec.debug_read("host_wait_us")) for w in workers) / 1e3
    print(f"in_flight={S} {kind:7s} {ms:7.3f} ms/step  ({256 / ms * 1e3:8.0f} lines/s)  opts={OPTS}  "
          f"cumulative host ms in decode: launching {hl:.0f}, waiting {hw:.0f}", flush=True)
