#!/usr/bin/env python
"""Benchmark of the text-line recognition hot path (BASELINE.json metric: text-lines/s at 1/2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W [--config c3|c2|c4|c5] [--impl reference]

Workloads (BASELINE.json `configs`):
  c3 (default)  8192 synthetic lines of resized width 200-1600 px (3-20 chunks), the SAME fixed set at every N, sharded over
                the N ranks by `scheduling.shard_lines` (balanced by chunk count); a step = the whole set recognised once,
                every rank's ids gathered on rank 0 over NCCL and restored to input order INSIDE the timed region.
                Strong scaling: this is the configuration the metric "lines/s at 1/2/4/8 GPUs" is quoted on.
  c2            256 lines of width 400-800 px per GPU per step (weak scaling; the round-1 headline, also measured as a
                side object `c2` of the default run so the two rounds stay comparable)
  c4            96 lines of width 2400 px (29 chunks, T = 928) per GPU per step: long merged sequences (BiLSTM, decoder)
  c5            VGG-Transformer baseline (no SE, no BiLSTM), 1024 scene-text-like lines (width 100-320) per GPU per step

A step streams through `LinePipeline` (khmer_ocr_cnn_transformer_b200/pipeline.py, the engine behind
`OCRPredictor.predict_batch`): `--in-flight` device passes per GPU, each a C-ABI call `kocr_recognize_lines` on its own
handle + stream + host thread; consecutive steps overlap like a stream of requests.  No collective on the data path; one
NCCL gather of the decoded ids per step.

  value        lines/s over all ranks with the grey line images already resident in HBM
  e2e          same metric through the C-ABI call with HOST buffers: pinned H2D of the pixels and D2H of the ids inside
  roofline     dominant kernel = the tcgen05 implicit-GEMM conv6 (gemm_tc_kernel<256, a16, column-fused>): achieved algorithmic
               TFLOP/s from CUDA events around its launches on the launching stream (separate instrumented pass, one batch in
               flight), against the measured bf16 peak of MEASURED_PEAKS.json (burst; the sustained fraction is given too)
  cpu_baseline the UNMODIFIED reference (baseline/_ref, torch CPU fp32) on a bounded sample, rank 0, N=1
               (falls back to the numpy oracle port when the reference is not installed, and says which ran)
"""
from __future__ import annotations

import argparse
import json
import os
import queue
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

FLOP_PER_CHUNK = 2_337_054_720          # stages 2-4, algorithmic (2*MACs); the VGG baseline lacks the SE terms
FLOP_SE = 3_686_400
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
WORKLOADS = {
    "c2": dict(lines=256, lo=400, hi=800, seed=0, scaling="weak", variant="se",
               text="c2: 256 synthetic Khmer text lines per GPU per step, resized width 400-800 px"),
    "c3": dict(lines=8192, lo=200, hi=1600, seed=3, scaling="strong", variant="se",
               text="c3: the same 8192 synthetic Khmer text lines (resized width 200-1600 px, 3-20 chunks) per step at every N, "
                    "sharded over the ranks by chunk count, ids gathered on rank 0 and restored to input order per step"),
    "c4": dict(lines=96, lo=2400, hi=2400, seed=4, scaling="weak", variant="se",
               text="c4: 96 long lines of width ~2400 px (29 chunks, T = 928) per GPU per step"),
    "c5": dict(lines=1024, lo=100, hi=320, seed=55, scaling="weak", variant="vgg",
               text="c5: VGG-Transformer baseline (no SE, no BiLSTM), 1024 short lines (width 100-320 px) per GPU per step"),
}


def load_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            d["_source"] = "measured (MEASURED_PEAKS.json)"
            return d
        except Exception:
            pass
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback (B200_PROFILING.md)"
    return d


def load_state_dict(variant="se"):
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint, seeded_state_dict
    ck = REPO / "tests" / "golden" / ("fixture_se_ckpt.npz" if variant == "se" else "fixture_vgg_ckpt.npz")
    if ck.exists():
        return load_checkpoint(ck), f"{ck.name} (reference architecture trained on synthetic lines)"
    return seeded_state_dict(variant, 0, max_global_len=1024), "seeded random init"


def make_lines(cfg, seed_offset=0):
    from workloads import synth
    return synth.make_lines(cfg["lines"], cfg["lo"], cfg["hi"], seed=cfg["seed"] + seed_offset)[0]


class stdout_to_stderr:
    """fd-level redirect: whatever C libraries print to stdout inside the block (NCCL's "NCCL version ..." banner at
    communicator creation) goes to stderr, so that stdout carries exactly ONE line, the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)
        return False


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(pynvml, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            while not self.stop_flag:
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
                time.sleep(0.02)
        except Exception as e:  # NVML unavailable: report that instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------------
# The comparison arm: the reference's own CPU implementation of the path
# ------------------------------------------------------------------------------------------------------------------
def reference_runner(sd, variant):
    """Returns (run(images) -> (lines/s, seconds), kind, note): the unmodified reference when baseline/_ref holds it,
    else the numpy oracle port (and the note says why)."""
    from baseline import run_reference as R
    ok, why = R.available()
    if ok:
        try:
            pred = R.load_predictor(sd, "cpu", variant)

            def run(images):
                v, dt, _ = R.time_predict_batch(pred, images, batch_size=8)
                return v, dt
            import torch
            return run, "reference", ("unmodified reference OCRPredictor.predict_batch(images, beam_width=1, batch_size=8), torch "
                                      f"{torch.__version__} CPU fp32, {torch.get_num_threads()} intra-op threads")
        except Exception as e:      # pragma: no cover
            why = f"reference failed to load: {type(e).__name__}: {e}"
    try:        # torchrun exports OMP_NUM_THREADS=1; give BLAS every host core back
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    from oracle import recognizer_np as O

    def run_port(images):
        t0 = time.perf_counter()
        O.recognise_lines(sd, images, variant, batch_size=8)
        dt = time.perf_counter() - t0
        return len(images) / dt, dt
    return run_port, "port", f"numpy oracle port ({why})"


def run_reference(args, rank):
    """--impl reference: bounded samples of the same workload through the reference's own CPU path, all host cores.
    Under torchrun only rank 0 runs."""
    if rank != 0:
        return
    cfg = WORKLOADS[args.config]
    sd, wname = load_state_dict(cfg["variant"])
    run, kind, note = reference_runner(sd, cfg["variant"])
    per_step = 16
    pool_cfg = dict(cfg, lines=min(cfg["lines"], per_step * 8))
    imgs = make_lines(pool_cfg)
    take = lambda i: [imgs[(i * per_step + j) % len(imgs)] for j in range(per_step)]
    for i in range(args.warmup):
        run(take(i))
    t0 = time.perf_counter()
    for i in range(args.steps):
        run(take(args.warmup + i))
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "text_lines_per_s", "value": value, "unit": "lines/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": cfg["text"] + f"; the reference arm runs a bounded sample of {per_step} of those lines per step "
                                             "on the host cores", "weights": wname, "how": note},
        "cpu_baseline": {"value": value, "unit": "lines/s", "cores": cores, "kind": kind,
                         "sample": f"{per_step} lines/step x {args.steps} steps of the {args.config} line set; {note}"},
        "e2e": {"value": value, "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if kind == "reference" and not args.no_incumbent:
        # the same unmodified call on the B200 itself (device='cuda' is the reference's default: config.py:13): the
        # incumbent GPU number.  fp32 eager torch, one line at a time, a host sync per generated token.
        try:
            import torch
            if torch.cuda.is_available():
                from baseline import run_reference as R
                pred = R.load_predictor(sd, "cuda", cfg["variant"])
                R.time_predict_batch(pred, take(0)[:4])
                v, dt2, _ = R.time_predict_batch(pred, take(1))
                line["incumbent_gpu"] = {"value": v, "unit": "lines/s", "device": torch.cuda.get_device_name(0),
                                         "sample": f"{per_step} lines, reference OCRPredictor.predict_batch at device='cuda' (fp32 eager)"}
        except Exception as e:      # pragma: no cover
            line["incumbent_gpu"] = {"unavailable": f"{type(e).__name__}: {e}"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# Own arm
# ------------------------------------------------------------------------------------------------------------------
class Workload:
    """The lines of one step that THIS rank processes, pre-packed into pipeline jobs (host-pinned and device-resident
    pixel buffers), plus the bookkeeping to put gathered ids back into input order."""

    def __init__(self, name, pipe, rank, world, torch):
        from khmer_ocr_cnn_transformer_b200 import _native
        from khmer_ocr_cnn_transformer_b200.scheduling import shard_lines, chunks_for
        cfg = WORKLOADS[name]
        self.name, self.cfg, self.scaling = name, cfg, cfg["scaling"]
        if cfg["scaling"] == "strong":
            imgs = make_lines(cfg)                                  # the same fixed set on every rank
            shapes = [im.shape for im in imgs]
            self.shards = shard_lines(shapes, world, pipe.max_seq_len)
            self.global_lines = len(imgs)
            self.total_pixel_bytes = int(sum(im.size for im in imgs))
            mine = self.shards[rank]
            self.images = [imgs[i] for i in mine]
            self.chunks_per_rank = [int(sum(chunks_for(*shapes[i], pipe.max_seq_len) for i in s)) for s in self.shards]
        else:
            self.images = make_lines(cfg, seed_offset=rank)         # rank 0 holds the parity-test batch of the config
            self.shards = [list(range(r * cfg["lines"], (r + 1) * cfg["lines"])) for r in range(world)]
            self.global_lines = cfg["lines"] * world
            self.chunks_per_rank = None
        self.n_local = len(self.images)
        self.cap = max(len(s) for s in self.shards)
        self.plan = pipe.plan([im.shape for im in self.images])
        self.protos = []
        for ids in self.plan:
            b = _native.LineBatch([self.images[i] for i in ids])
            host = torch.from_numpy(b.pixels).pin_memory()
            bh = _native.LineBatch.__new__(_native.LineBatch)
            bh.__dict__.update(b.__dict__)
            bh.pixels = host.numpy()
            self.protos.append((ids, bh, host, host.cuda()))
        self.pixel_bytes = int(sum(p[1].pixel_bytes for p in self.protos))
        if cfg["scaling"] != "strong":
            self.total_pixel_bytes = self.pixel_bytes * world        # (every rank's own lines have about the same size)
        self.n_chunks = int(sum(chunks_for(im.shape[0], im.shape[1], pipe.max_seq_len) for im in self.images))

    def jobs(self, kind, step):
        from khmer_ocr_cnn_transformer_b200.pipeline import Job
        out = []
        for ids, bh, _, dev in self.protos:
            out.append(Job(ids, batch=bh, dev_ptr=dev.data_ptr() if kind == "resident" else None, tag=step))
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-side-c2", action="store_true", help="skip the c2 side measurement of the default c3 run")
    ap.add_argument("--no-api", action="store_true", help="skip the OCRPredictor.predict_batch measurement (N = 1 only)")
    ap.add_argument("--no-incumbent", action="store_true", help="reference arm: skip the device='cuda' run of the reference")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="time budget of the bounded cpu_baseline sample")
    ap.add_argument("--straggler-threshold", type=int, default=8,
                    help="a pass returns once <= this many of its lines (per 256) are still decoding; they are pooled")
    ap.add_argument("--in-flight", type=int, default=12,
                    help="device passes in flight per GPU (one handle + stream + host thread each)")
    ap.add_argument("--lines-per-pass", type=int, default=0, help="line capacity of a device pass (0: 256; c5: 1024)")
    ap.add_argument("--max-chunks", type=int, default=0, help="chunk capacity of a device pass (0: 2816; c5: 4096)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from khmer_ocr_cnn_transformer_b200 import _native, weights
    from khmer_ocr_cnn_transformer_b200.pipeline import LinePipeline, TOKENS_LD
    from khmer_ocr_cnn_transformer_b200.distributed import gather_ids

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()          # NCCL creates its communicator (and prints its banner) on the first collective

    cfg = WORKLOADS[args.config]
    sd, wname = load_state_dict(cfg["variant"])
    blob = weights.pack_blob(sd)
    S = max(1, args.in_flight)
    max_lines = args.lines_per_pass or (1024 if args.config == "c5" else 256)
    max_chunks = args.max_chunks or (4096 if args.config == "c5" else 2816)
    pipe = LinePipeline(blob, device=local_rank, in_flight=S, max_lines=max_lines, max_chunks=max_chunks,
                        straggler_per_256=args.straggler_threshold)
    pipe._ensure(S)      # every handle (weights + workspace) exists before anything is timed, also when a warm-up has fewer jobs than S

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(wl, kind, steps, results=None):
        """`steps` steps of workload `wl` streamed through the pipeline; every step's ids are gathered on rank 0 (NCCL)
        and put back into input order.  results (rank 0): list that receives (tokens, lengths) of every step."""
        toks = [np.zeros((wl.n_local, TOKENS_LD), np.int32) for _ in range(steps)]
        lens = [np.zeros(wl.n_local, np.int32) for _ in range(steps)]
        remaining = [len(wl.protos)] * steps
        ready: queue.Queue = queue.Queue()
        lock = threading.Lock()

        def on_done(job):
            with lock:
                remaining[job.tag] -= 1
                fin = remaining[job.tag] == 0
            if fin:
                ready.put(job.tag)

        gather_err = []

        def gatherer():
            try:
                torch.cuda.set_device(local_rank)
                done, nxt = set(), 0
                while nxt < steps:
                    while nxt not in done:
                        done.add(ready.get())
                    got = gather_ids(toks[nxt], lens[nxt], wl.shards, rank, world, device=torch.device("cuda", local_rank))
                    if rank == 0 and results is not None:
                        results.append(got)
                    nxt += 1
            except Exception as e:      # pragma: no cover
                gather_err.append(e)

        # one run_jobs call per step would drain the pipeline between steps: all steps go into ONE queue
        jobs = []
        for k in range(steps):
            for j in wl.jobs(kind, k):
                jobs.append(j)
        gt = threading.Thread(target=gatherer, daemon=True)      # (daemon: a failing pass must not leave the process hanging on it)
        gt.start()
        # the pipeline indexes ONE pair of output arrays: stack the steps (row = step * n_local + local id)
        big_tok = np.zeros((steps * wl.n_local, TOKENS_LD), np.int32)
        big_len = np.zeros(steps * wl.n_local, np.int32)
        for j in jobs:
            j.ids = [j.tag * wl.n_local + i for i in j.ids]

        def done_and_copy(job):
            k = job.tag
            loc = [i - k * wl.n_local for i in job.ids]
            toks[k][loc] = big_tok[job.ids]
            lens[k][loc] = big_len[job.ids]
            on_done(job)

        pipe.run_jobs(jobs, big_tok, big_len, image_of=lambda i: wl.images[i % wl.n_local], on_done=done_and_copy)
        gt.join()
        if gather_err:
            raise gather_err[0]

    def timed(wl, kind, steps, results=None):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        run_steps(wl, kind, steps, results)
        b.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([a.elapsed_time(b)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), wall_ms

    def measure(wl, steps, warmup, want_results=False):
        run_steps(wl, "resident", max(1, warmup))
        run_steps(wl, "e2e", max(1, warmup))
        launches0 = _native.launch_count()
        ms_res, wall_res = timed(wl, "resident", steps)
        launches = _native.launch_count() - launches0
        results = [] if want_results else None
        ms_e2e, wall_e2e = timed(wl, "e2e", steps, results)
        return dict(ms_res=ms_res, ms_e2e=ms_e2e, wall_res=wall_res, wall_e2e=wall_e2e, launches=launches, results=results)

    wl = Workload(args.config, pipe, rank, world, torch)
    if world > 1:       # NCCL creates its communicator lazily on the first collective: do that outside the timed region
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    m = measure(wl, args.steps, args.warmup, want_results=True)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    stats_main = dict(pipe.stats)

    total_lines = wl.global_lines
    value = total_lines * args.steps / (m["ms_res"] * 1e-3)
    e2e = total_lines * args.steps / (m["ms_e2e"] * 1e-3)

    # ---- correctness of what was timed: the gathered ids of the last step against the committed oracle tokens
    identity = None
    if rank == 0 and m["results"]:
        tokens, lengths = m["results"][-1]
        p = REPO / "tests" / "golden" / f"oracle_tokens_{'c3full' if args.config == 'c3' else args.config}.npz"
        if p.exists() and args.config in ("c2", "c3"):
            o = np.load(p)
            same = int(sum(lengths[i] == o["lengths"][i] and np.array_equal(tokens[i, :lengths[i]], o["tokens"][i, :lengths[i]])
                           for i in range(min(total_lines, o["tokens"].shape[0]))))
            identity = {"identical_lines": same, "of": int(min(total_lines, o["tokens"].shape[0])),
                        "against": f"tests/golden/{p.name} (numpy oracle = the fp32 reference, see tests/test_oracle_golden.py)"}
        mean_len = float(lengths.mean())
    else:
        mean_len = None

    # ---- side measurement: the round-1 headline workload (c2, weak scaling) on the same handles
    side_c2 = None
    if args.config == "c3" and not args.no_side_c2:
        wl2 = Workload("c2", pipe, rank, world, torch)
        steps2 = max(4 * S, 24)
        m2 = measure(wl2, steps2, S)
        side_c2 = {"workload": WORKLOADS["c2"]["text"], "scaling": "weak", "steps": steps2,
                   "value": wl2.global_lines * steps2 / (m2["ms_res"] * 1e-3), "unit": "lines/s",
                   "e2e": wl2.global_lines * steps2 / (m2["ms_e2e"] * 1e-3), "chunks_per_gpu": wl2.n_chunks,
                   "round1_value": 21340, "round1_e2e": 21565}

    # ---- single pass alone (latency, for context): the first job of the workload, full-length decode
    rec = pipe.recs[0]
    rec.set_option("blocking_wait", 0)
    rec.set_option("straggler_threshold", 0)
    ids0, bh0, _, dev0 = wl.protos[0]
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        rec.recognize_lines(bh0)
    lat_ms = (time.perf_counter() - t0) * 1e3 / 3
    decode_steps = int(rec.debug_read("last_steps"))

    # ---- instrumented pass: CUDA events around every launch of stages 2-5a on the launching stream (roofline evidence),
    #      one batch in flight so that the per-launch times are not perturbed by other streams
    n_chunks0 = int(rec.gather_chunks(bh0, pixels_dev_ptr=dev0.data_ptr()).sum())
    reps = min(max(args.steps, 8), 24)
    rec.set_option("kernel_timing", 1)
    for _ in range(reps):
        rec.recognize_lines(bh0, pixels_dev_ptr=dev0.data_ptr())
    kt = rec.kernel_timing()
    rec.set_option("kernel_timing", 0)
    peaks = load_peaks()
    peak_burst = float(peaks["bf16_tflops"])
    peak_sust = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    dom = kt.get("conv6", {"ms": 0.0, "launches": 0, "flops": 0.0})
    achieved_tf = dom["flops"] / (dom["ms"] * 1e-3) / 1e12 if dom["ms"] > 0 else 0.0
    # algorithmic bytes of one conv6 launch: read conv5's output (150 px x 512 ch), write the pooled rows (75 px x 512) and
    # the fp32 column means (25 x 512), all per chunk, plus the weights once
    alg_bytes = n_chunks0 * (150 * 512 * 2 + 75 * 512 * 2 + 25 * 512 * 4) + 512 * 4608 * 2
    traffic, traffic_note = None, None
    try:        # DRAM bytes of this launch from the committed `ncu --set full` capture of the same launch shape
        cap = json.loads((REPO / "profiles" / "r02" / "conv6_ncu.json").read_text())
        traffic = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) * n_chunks0 / cap["chunks"]
        traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from " + cap["source"] +
                        f" ({cap['chunks']} chunks), scaled to this launch's {n_chunks0} chunks")
    except Exception:
        pass
    stage_sites = [s for s in kt if s.startswith(("conv", "se", "pool", "final_pool", "patch_proj", "enc_", "res"))]
    stage_ms = sum(kt[s]["ms"] for s in stage_sites) / reps
    stage_chunks_per_s = n_chunks0 / (stage_ms * 1e-3) if stage_ms > 0 else 0.0
    flop_chunk = FLOP_PER_CHUNK - (FLOP_SE if cfg["variant"] != "se" else 0)
    per_site = {s: {"ms_per_pass": kt[s]["ms"] / reps,
                    "tflops": (kt[s]["flops"] / (kt[s]["ms"] * 1e-3) / 1e12) if kt[s]["ms"] > 0 else 0.0}
                for s in kt}

    cpu = None
    if rank == 0 and world == 1 and args.cpu_seconds > 0:
        run, kind, note = reference_runner(sd, cfg["variant"])
        imgs = wl.images
        run(imgs[:2])
        done_lines, t_used, i = 0, 0.0, 0
        while t_used < args.cpu_seconds and done_lines < 256:
            batch = [imgs[(i + j) % len(imgs)] for j in range(16)]
            _, dt = run(batch)
            t_used += dt; done_lines += 16; i += 16
        cpu = {"value": done_lines / t_used, "unit": "lines/s", "cores": os.cpu_count(), "kind": kind,
               "sample": f"{done_lines} lines of this rank's {args.config} set in {t_used:.1f} s; {note}"}

    # ---- the drop-in API itself: OCRPredictor.predict_batch on a list of grey images (host packing, H2D, recognition, D2H
    #      and Tokenizer.decode to str inside), same engine, own handles (the bench's are released first)
    api = None
    if rank == 0 and world == 1 and not args.no_api and cfg["variant"] == "se":
        pipe.close()
        from khmer_ocr_cnn_transformer_b200.recognition.predictor import OCRPredictor
        from khmer_ocr_cnn_transformer_b200.recognition.tokenizer import Tokenizer
        from khmer_ocr_cnn_transformer_b200.recognition.config import OCRConfig
        from khmer_ocr_cnn_transformer_b200.recognition.utils import autodetect_config
        from khmer_ocr_cnn_transformer_b200.recognition.model.se_model import KhmerOCR
        ck = REPO / "tests" / "golden" / "fixture_se_ckpt.npz"
        pkg = REPO / "khmer_ocr_cnn_transformer_b200" / "recognition"
        pred = OCRPredictor(ck, Tokenizer(pkg / "char2idx.json"), OCRConfig(**autodetect_config(ck)), KhmerOCR,
                            max_lines=max_lines, max_chunks=max_chunks, in_flight=S)
        imgs = wl.images if wl.scaling == "strong" else [im for k in range(16) for im in make_lines(cfg, seed_offset=100 + k)]
        pred.predict_batch(imgs[:min(len(imgs), 12 * max_lines)])           # creates the handles, captures the decode graphs
        t0 = time.perf_counter()
        texts = pred.predict_batch(imgs, beam_width=1, batch_size=8)
        dt = time.perf_counter() - t0
        api = {"value": len(imgs) / dt, "unit": "lines/s", "lines": len(imgs), "seconds": dt,
               "call": "OCRPredictor.predict_batch(list of grey uint8 arrays, beam_width=1, batch_size=8) -> list[str]",
               "frac_of_e2e": (len(imgs) / dt) / e2e, "non_empty": int(sum(1 for t in texts if t))}
        nb = min(len(imgs), 1024)
        pred.predict_batch(imgs[:nb], beam_width=3)                          # (allocates the beam caches)
        t0 = time.perf_counter()
        pred.predict_batch(imgs[:nb], beam_width=3)
        api["beam3"] = {"value": nb / (time.perf_counter() - t0), "unit": "lines/s", "lines": nb,
                        "call": "OCRPredictor.predict_batch(..., beam_width=3): kocr_beam_search per batch of max_lines // 3 lines"}
        pred.close()

    if rank == 0:
        line = {
            "metric": "text_lines_per_s", "value": value, "unit": "lines/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["ms_res"] / args.steps, "higher_is_better": True,
            "scaling": wl.scaling, "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": {"workload": cfg["text"] + "; SE-VGG-Transformer, greedy decode" if cfg["variant"] == "se" else cfg["text"] + "; greedy decode",
                       "weights": wname, "lines_per_step": total_lines, "lines_this_rank": wl.n_local,
                       "chunks_this_rank": wl.n_chunks, "chunks_per_rank": wl.chunks_per_rank,
                       "device_passes_per_step_this_rank": len(wl.protos), "mean_decoded_len": mean_len,
                       "decode_positions_first_pass": decode_steps, "in_flight_device_passes": S,
                       "lines_per_device_pass": max_lines, "chunks_per_device_pass": max_chunks,
                       "straggler_threshold_per_256": args.straggler_threshold, "pipeline_stats": stats_main,
                       "single_pass_alone_ms": lat_ms, "wall_ms_resident": m["wall_res"], "wall_ms_e2e": m["wall_e2e"],
                       "gather": "one NCCL gather of the ids per step, inside the timed region" if world > 1 else "single rank: no collective",
                       "l2": "per-pass working set (~1.5 MB of activations per chunk, >3 GB per pass) exceeds the 126 MB L2",
                       "parallelism": f"lines sharded over {world} GPU(s) by chunk count, no data-path collective"},
            "e2e": {"value": e2e, "unit": "lines/s", "ms_per_step": m["ms_e2e"] / args.steps,
                    "h2d_bytes_per_step": int(wl.total_pixel_bytes), "h2d_bytes_per_step_rank0": int(wl.pixel_bytes),
                    "d2h_bytes_per_step": int(total_lines * (TOKENS_LD + 1) * 4)},
            "gpu_launches": int(m["launches"]),
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor", "kernel": f"gemm_tc_kernel<256, a16, column-fused> @ conv6 (implicit GEMM by TMA im2col, "
                                                      f"M = {n_chunks0} chunks x 150 px, N = 512, K = 4608; epilogue: (2,1) max-pool + SE column means)",
                         "achieved": achieved_tf, "peak": peak_burst, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_burst if peak_burst else None,
                         "frac_of_sustained_peak": achieved_tf / peak_sust if peak_sust else None,
                         "peak_source": peaks["_source"] + ": burst figure (kernel timed with one pass in flight); sustained peak "
                                        f"{peak_sust} TFLOP/s given for the in-step view",
                         "traffic": traffic, "traffic_unit": "bytes", "traffic_source": traffic_note,
                         "algorithmic_bytes": alg_bytes,
                         "launches_timed": dom["launches"], "ms_per_launch": dom["ms"] / max(dom["launches"], 1)},
            "sevgg_encoder_stage": {"chunks_per_s": stage_chunks_per_s, "ms_per_pass": stage_ms, "chunks_per_pass": n_chunks0,
                                    "tflops_algorithmic": stage_chunks_per_s * flop_chunk / 1e12,
                                    "frac_of_bf16_burst_peak": stage_chunks_per_s * flop_chunk / 1e12 / peak_burst,
                                    "frac_of_bf16_sustained_peak": stage_chunks_per_s * flop_chunk / 1e12 / peak_sust,
                                    "how": "sum of the per-launch CUDA-event times of every stage 2-4 kernel, one pass in flight"},
            "kernels": per_site,
        }
        if identity is not None:
            line["token_identity"] = identity
        if side_c2 is not None:
            line["c2"] = side_c2
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if api is not None:
            line["api"] = api
        print(json.dumps(line), flush=True)
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
