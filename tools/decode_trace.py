"""Per-kernel GPU time of the decode loop (or the whole step) with several batches in flight, from CUPTI through
torch.profiler (no nsys in the image).  Diagnosis tool, GPU box:

  python tools/decode_trace.py [in_flight=12] [kind=decode|full|heavy] [steps=24] [opt=value ...]
"""
import sys
import threading
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import torch
from torch.profiler import profile, ProfilerActivity

S = int(sys.argv[1]) if len(sys.argv) > 1 else 12
KIND = sys.argv[2] if len(sys.argv) > 2 else "decode"
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 24
opts = sys.argv[4:]
sys.argv = [sys.argv[0], str(S), "0"] + opts          # inflight_probe parses argv at import: S workers, no timed steps
import inflight_probe as P                             # builds the workers and runs the three warm schedules (0 steps)

for k in ("heavy", "decode", "full"):
    P.run(k, S)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    P.run(KIND, STEPS)
    wall = (time.perf_counter() - t0) * 1e3
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = e.cuda_time_total
    if t > 0:
        rows.append((t / 1e3, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"kind={KIND} in_flight={S} steps={STEPS} wall {wall:.1f} ms = {wall / STEPS:.3f} ms/step; sum of kernel durations {tot:.1f} ms "
      f"= {tot / STEPS:.3f} ms/step  opts={opts}")
for t, c, k in rows[:16]:
    print(f"  {t / STEPS:8.3f} ms/step  {c / STEPS:8.1f} launches/step  avg {t * 1e3 / c:7.1f} us  {k[:90]}")
