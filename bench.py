#!/usr/bin/env python
"""Benchmark of the text-line recognition hot path (BASELINE.json metric: text-lines/s).

  python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one pass of the whole hot path (resize + chunk gather -> SE-VGG -> patch projection ->
encoder -> merge + BiLSTM -> greedy decode) over one batch of 256 synthetic lines of resized width
400-800 px (BASELINE.json configs[1]) per GPU.  `--in-flight` (default 12) such passes run concurrently, each on
its own handle + stream + host thread, because one decode chain alone leaves most SMs idle; `--coalesce k` puts k
batches into one C-ABI call instead (K steps = K x 256 lines processed, whatever the grouping).  Weak scaling: every rank owns its own batch (lines are
independent); no collective on the data path, one gather of the decoded ids at the end
of each end-to-end step.  Prints ONE JSON line on rank 0.

  value        lines/s over all ranks with the grey line images already resident in HBM
  e2e          same metric through the public C-ABI call with HOST buffers: pinned H2D of the pixels
               and D2H of the token ids inside the timed region
  roofline     dominant kernel = the tcgen05 implicit-GEMM conv (gemm_tc_kernel) of conv6; achieved
               algorithmic TFLOP/s from CUDA events around its launches (a second, instrumented pass
               over the same steps), against the measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline the numpy oracle (a port of the reference's CPU path) on a bounded sample, rank 0, N=1
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

LINES_PER_STEP = 256
WIDTH_LO, WIDTH_HI = 400, 800
FLOP_PER_CHUNK = 2_337_054_720          # stages 2-4, algorithmic (2*MACs)
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


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


def load_state_dict():
    from khmer_ocr_cnn_transformer_b200.checkpoint import load_checkpoint, seeded_state_dict
    ck = REPO / "tests" / "golden" / "fixture_se_ckpt.npz"
    if ck.exists():
        return load_checkpoint(ck), "fixture_se_ckpt.npz (reference model trained on synthetic lines)"
    return seeded_state_dict("se", 0, max_global_len=1024), "seeded random init"


def make_batch(rank: int, n_batches: int = 1):
    """`n_batches` c2 batches of 256 lines (seeds rank, rank+1000, ...); seed 0 == the parity-test batch."""
    from workloads import synth
    imgs = []
    for b in range(n_batches):
        imgs += synth.make_lines(LINES_PER_STEP, WIDTH_LO, WIDTH_HI, seed=rank + 1000 * b)[0]
    return imgs


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


def cpu_oracle_lines_per_s(sd, imgs, n_lines):
    """Times the numpy oracle (port of the reference CPU path) on `n_lines` lines of the workload."""
    from oracle import recognizer_np as O
    t0 = time.perf_counter()
    O.recognise_lines(sd, imgs[:n_lines], "se", batch_size=8)
    dt = time.perf_counter() - t0
    return n_lines / dt, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the Python
    reference cannot travel to the GPU box), all host threads numpy/BLAS can use, same workload."""
    if rank != 0:
        return
    try:        # torchrun exports OMP_NUM_THREADS=1; give BLAS every host core back
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    sd, wname = load_state_dict()
    imgs = make_batch(0)
    per_step = 2
    for i in range(args.warmup):
        cpu_oracle_lines_per_s(sd, imgs[i * per_step:], per_step)
    t0 = time.perf_counter()
    for i in range(args.steps):
        lo = (args.warmup + i) * per_step % (LINES_PER_STEP - per_step)
        cpu_oracle_lines_per_s(sd, imgs[lo:], per_step)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "text_lines_per_s", "value": value, "unit": "lines/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "c2: 256 synthetic lines, width 400-800 px, greedy decode; reference arm runs a "
                               f"bounded sample of {per_step} lines per step", "weights": wname},
        "cpu_baseline": {"value": value, "unit": "lines/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} lines/step x {args.steps} steps of the c2 batch (numpy oracle, BLAS threads)"},
        "e2e": {"value": value, "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=96)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-lines", type=int, default=8, help="lines of the bounded cpu_baseline sample")
    ap.add_argument("--dec-wide", type=int, default=-1,
                    help="decode GEMM shape: 1 = split-K over many CTAs (latency), 0 = few CTAs (throughput), -1 = auto")
    ap.add_argument("--no-pdl", action="store_true", help="disable programmatic dependent launch in the decode loop")
    ap.add_argument("--straggler-threshold", type=int, default=8,
                    help="a step returns once <= this many of its 256 lines are still decoding; they are pooled")
    ap.add_argument("--big-gemm-sms", type=int, default=0,
                    help="persistent grid size of the large GEMMs when several batches are in flight (0 = all SMs; "
                         "reserving SMs for the small decode kernels measured no gain: tools/inflight_probe.py)")
    ap.add_argument("--in-flight", type=int, default=12,
                    help="device passes in flight per GPU (one handle + stream + host thread each)")
    ap.add_argument("--coalesce", type=int, default=1,
                    help="256-line batches (steps) coalesced into one device pass / C-ABI call.  Measured on B200 "
                         "(profiles/r01/README): 1 x 12 in flight, 4 x 4 and 8 x 3 all give 20-22 k lines/s - the decode "
                         "loop is bound by per-line attention work, not by launch count - so the default keeps one "
                         "256-line batch per call")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from khmer_ocr_cnn_transformer_b200 import _native, weights

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":     # keeps "NCCL version ..." out of stdout (one JSON line)
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sd, wname = load_state_dict()
    blob = weights.pack_blob(sd)
    S = max(1, args.in_flight)

    KB = max(1, args.coalesce)
    LPC = LINES_PER_STEP * KB                                   # lines per device pass

    pool, pool_lock = [], threading.Lock()      # stragglers handed back by the in-flight passes

    class Worker:
        """One in-flight device pass: its own handle (weights + workspace + stream), its own pinned buffers.
        A pass covers `KB` steps (256-line batches); `sub[k]` is the same data cut down to k batches for the tail."""

        def __init__(self, w):
            self.imgs = make_batch(rank * 64 + w, KB)       # rank 0 / worker 0 starts with the parity-test batch (seed 0)
            self.rec = _native.Recognizer(blob, device=local_rank, max_lines=LPC, max_chunks=LPC * 11)
            self.sub = {}
            for k in sorted({KB, 1} | set(range(1, KB))):
                b = _native.LineBatch(self.imgs[:k * LINES_PER_STEP])
                host = torch.from_numpy(b.pixels).pin_memory()
                bh = _native.LineBatch.__new__(_native.LineBatch)
                bh.__dict__.update(b.__dict__)
                bh.pixels = host.numpy()
                self.sub[k] = (b, bh, host, host.cuda())
            self.batch = self.sub[KB][0]
            self.tok_host = torch.zeros((LPC, _native.TOKENS_LD), dtype=torch.int32).pin_memory()
            self.len_host = torch.zeros(LPC, dtype=torch.int32).pin_memory()
            self.tok_np, self.len_np = self.tok_host.numpy(), self.len_host.numpy()
            self.n_stragglers, self.n_flushes = 0, 0
            self.rec.set_option("dec_wide", 1 if (args.dec_wide == 1 or (args.dec_wide < 0 and S * KB <= 4)) else 0)
            # programmatic dependent launch shortens ONE decode chain (latency); with several passes in flight the early-
            # resident dependents only hold SM slots while they wait, which costs ~7 % of throughput (tools/inflight_probe.py)
            if args.no_pdl or S > 1:
                self.rec.set_option("use_pdl", 0)
            if S > 1:       # a dozen host threads per GPU (x 8 ranks per host) must not spin inside cudaStreamSynchronize
                self.rec.set_option("blocking_wait", 1)
            if args.big_gemm_sms > 0 and S > 1:
                self.rec.set_option("big_gemm_sms", args.big_gemm_sms)
            self.n_chunks = int(self.rec.gather_chunks(self.sub[1][0], pixels_dev_ptr=self.sub[1][3].data_ptr()).sum())

        # Long tail: a pass returns once <= 8 lines per 256 are still decoding; those stragglers go to a pool SHARED by the
        # in-flight passes and are decoded to the end in passes of up to LPC lines, inside the timed region.  Same
        # results, see predictor.py.  The last passes of a run (`final`) decode every line in place instead, so that the
        # run does not end with a lone, latency-bound straggler pass.
        def _collect(self, k):
            todo = np.nonzero(self.rec.unfinished(k * LINES_PER_STEP))[0]
            self.n_stragglers += len(todo)
            part = None
            with pool_lock:
                pool.extend(self.imgs[i] for i in todo)
                if len(pool) >= LPC:
                    part = pool[:LPC]
                    del pool[:LPC]
            if part:
                self._decode_pool(part)

        def _decode_pool(self, part):
            self.rec.set_option("straggler_threshold", 0)
            self.rec.recognize_lines(_native.LineBatch(part))
            self.n_flushes += 1

        def flush(self):
            while True:
                with pool_lock:
                    part = pool[:LPC]
                    del pool[:LPC]
                if not part:
                    return
                self._decode_pool(part)

        # The same pass as two calls (stages 1-5a, then the decode loop) for the phased schedule below.
        def step_heavy(self, host, k=None):
            k = k or KB
            self.rec.set_option("straggler_threshold", args.straggler_threshold * k)
            b, bh, _, dev = self.sub[k]
            if host:
                self.rec.gather_chunks(bh)
            else:
                self.rec.gather_chunks(b, pixels_dev_ptr=dev.data_ptr())
            self.rec.sevgg_encoder_forward()
            self.rec.merge_bilstm_forward()

        def step_decode(self, k=None):
            check = _native.check
            check(self.rec.lib.kocr_decode_greedy(self.rec._h, 0, self.tok_np.ctypes.data, self.len_np.ctypes.data, None))
            self._collect(k or KB)

        def step_resident(self, k=None, final=False):
            k = k or KB
            self.rec.set_option("straggler_threshold", 0 if final else args.straggler_threshold * k)
            b, _, _, dev = self.sub[k]
            self.rec.recognize_lines(b, pixels_dev_ptr=dev.data_ptr(), tokens_out=self.tok_np, lengths_out=self.len_np)
            self._collect(k)

        def step_e2e(self, k=None, final=False):   # H2D of the pixels (pinned) ... D2H of the ids, all inside the C-ABI call
            k = k or KB
            self.rec.set_option("straggler_threshold", 0 if final else args.straggler_threshold * k)
            self.rec.recognize_lines(self.sub[k][1], tokens_out=self.tok_np, lengths_out=self.len_np)
            self._collect(k)

    workers = [Worker(w) for w in range(S)]
    n_chunks = workers[0].n_chunks
    batch = workers[0].sub[1][0]            # one 256-line step (byte counts are quoted per step)
    rec = workers[0].rec

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(kind, steps):
        """`steps` 256-line batches, `KB` of them per device pass, at most S passes in flight (one host thread each)."""
        counter = {"next": 0}
        lock = threading.Lock()
        errors = []

        def loop(wk):
            try:
                torch.cuda.set_device(local_rank)
                fn = wk.step_resident if kind == "resident" else wk.step_e2e
                while True:
                    with lock:
                        i = counter["next"]
                        if i >= steps:
                            break
                        k = min(KB, steps - i)
                        counter["next"] = i + k
                    fn(k, final=(steps - i) <= n_threads * KB)      # the last pass of every worker: no hand-off
                wk.flush()          # decode this worker's pooled stragglers to the end (inside the timed region)
            except Exception as e:  # surface worker failures instead of hanging
                errors.append(e)

        n_threads = max(1, min(S, (steps + KB - 1) // KB))
        threads = [threading.Thread(target=loop, args=(wk,)) for wk in workers[:n_threads]]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]

    gathered = [torch.zeros((LINES_PER_STEP, _native.TOKENS_LD), dtype=torch.int32, device="cuda")
                for _ in range(world)] if (world > 1 and rank == 0) else None

    def timed(kind, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        run_steps(kind, steps)
        if world > 1 and kind == "e2e":   # the only collective: decoded ids to rank 0 (NCCL gather over NVLink)
            dist.gather(workers[0].tok_host[:LINES_PER_STEP].cuda(non_blocking=True), gathered, dst=0)
        b.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([a.elapsed_time(b)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), wall_ms

    run_steps("resident", max(args.warmup, S * KB))
    run_steps("e2e", max(args.warmup, S * KB))
    if world > 1:       # NCCL creates its communicator lazily on the first collective: do that outside the timed region
        dist.gather(workers[0].tok_host[:LINES_PER_STEP].cuda(non_blocking=True), gathered, dst=0)
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _native.launch_count()
    ms_res, wall_res = timed("resident", args.steps)
    launches = _native.launch_count() - launches0
    ms_e2e, wall_e2e = timed("e2e", args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    for wk in workers:          # one plain full-length step each for the length statistics
        wk.rec.set_option("straggler_threshold", 0)
    workers[0].rec.recognize_lines(workers[0].sub[1][0], tokens_out=workers[0].tok_np, lengths_out=workers[0].len_np)
    mean_len = float(workers[0].len_np[:LINES_PER_STEP].mean())
    decode_steps = int(rec.debug_read("last_steps"))

    # ---- single in-flight latency of one step (for context)
    workers[0].rec.set_option("blocking_wait", 0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):          # full-length decode of every line (no straggler hand-off)
        workers[0].rec.recognize_lines(workers[0].sub[1][1], tokens_out=workers[0].tok_np, lengths_out=workers[0].len_np)
    lat_ms = (time.perf_counter() - t0) * 1e3 / 3

    # ---- instrumented pass: CUDA events around every launch of stages 2-5a (roofline evidence),
    #      one batch in flight so that the per-launch times are not perturbed by other streams
    rec.set_option("kernel_timing", 1)
    for _ in range(args.steps):
        workers[0].step_resident(1)            # one 256-line batch per pass: per-launch times of the c2 batch itself
    kt = rec.kernel_timing()
    rec.set_option("kernel_timing", 0)
    peaks = load_peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    dom = kt.get("conv6", {"ms": 0.0, "launches": 0, "flops": 0.0})
    traffic, traffic_note = None, None
    try:        # DRAM bytes of this launch from the committed `ncu --set full` capture of the same workload
        cap = json.loads((REPO / "profiles" / "r01" / "conv6_ncu_v4.json").read_text())
        if cap.get("chunks") == n_chunks:
            traffic = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) * 1e6
            traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, " + cap["source"] +
                            "; algorithmic bytes per launch = chunks * (2 * 182*512*2 B activations) + 4.7 MB weights")
    except Exception:
        pass
    achieved_tf = dom["flops"] / (dom["ms"] * 1e-3) / 1e12 if dom["ms"] > 0 else 0.0
    gemm_sites = ["conv2", "conv3", "conv4", "conv5", "conv6", "conv7", "patch_proj", "enc_qkv", "enc_out_proj",
                  "enc_ffn1", "enc_ffn2"]
    # every launch of stages 2-4: the GEMMs, conv1, the pools, the SE blocks (fused kernels, or - VGG baseline / A-B
    # option - the squeeze / FC / apply kernels), per-chunk attention and LayerNorm
    stage_sites = gemm_sites + ["conv1_pool1", "pool2", "enc_attention", "enc_layernorm"] + \
        [s for s in kt if s.startswith("se")]
    stage_ms = sum(kt[s]["ms"] for s in stage_sites if s in kt) / max(args.steps, 1)
    stage_chunks_per_s = n_chunks / (stage_ms * 1e-3) if stage_ms > 0 else 0.0
    per_site = {s: {"ms_per_step": kt[s]["ms"] / args.steps,
                    "tflops": (kt[s]["flops"] / (kt[s]["ms"] * 1e-3) / 1e12) if kt[s]["ms"] > 0 else 0.0}
                for s in kt}

    total_lines = LINES_PER_STEP * world
    value = total_lines * args.steps / (ms_res * 1e-3)
    e2e = total_lines * args.steps / (ms_e2e * 1e-3)

    cpu = None
    if rank == 0 and world == 1:
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=os.cpu_count())
        except Exception:
            pass
        v, dt = cpu_oracle_lines_per_s(sd, workers[0].imgs, args.cpu_lines)
        cpu = {"value": v, "unit": "lines/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"first {args.cpu_lines} lines of the c2 batch through the numpy oracle ({dt:.1f} s)"}

    if rank == 0:
        line = {
            "metric": "text_lines_per_s", "value": value, "unit": "lines/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": {"workload": f"c2: {LINES_PER_STEP} synthetic Khmer text lines per GPU, resized width "
                                   f"{WIDTH_LO}-{WIDTH_HI} px ({n_chunks} chunks of 48x100), SE-VGG-Transformer, greedy decode",
                       "weights": wname, "lines_per_gpu": LINES_PER_STEP, "chunks_per_gpu": n_chunks,
                       "mean_decoded_len": mean_len, "decode_steps": decode_steps, "in_flight_device_passes": S, "big_gemm_sms": args.big_gemm_sms,
                       "batches_coalesced_per_device_pass": KB, "lines_per_device_pass": LPC,
                       "straggler_threshold": args.straggler_threshold,
                       "stragglers_pooled": int(sum(wk.n_stragglers for wk in workers)),
                       "straggler_batches": int(sum(wk.n_flushes for wk in workers)),
                       "single_in_flight_e2e_ms_per_step": lat_ms, "wall_ms_resident": wall_res, "wall_ms_e2e": wall_e2e,
                       "l2": "per-step working set (~1.6 MB of activations per chunk, >3 GB per step) exceeds the 126 MB L2",
                       "parallelism": f"lines sharded over {world} GPU(s), no data-path collective"},
            "e2e": {"value": e2e, "unit": "lines/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(batch.pixel_bytes) * world,
                    "d2h_bytes_per_step": int(LINES_PER_STEP * (_native.TOKENS_LD + 1) * 4) * world},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel<256> @ conv6 (implicit GEMM, M=chunks*182, N=512, K=4608)",
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                         "traffic_unit": "bytes", "traffic_source": traffic_note,
                         "algorithmic_bytes": n_chunks * 2 * 182 * 512 * 2 + 512 * 4608 * 2,
                         "peak_source": peaks["_source"] + " (sustained figure: kernel timed inside a long step)",
                         "launches_timed": dom["launches"], "ms_per_launch": dom["ms"] / max(dom["launches"], 1)},
            "sevgg_encoder_stage": {"chunks_per_s": stage_chunks_per_s, "ms_per_step": stage_ms,
                                    "tflops_algorithmic": stage_chunks_per_s * FLOP_PER_CHUNK / 1e12,
                                    "frac_of_bf16_burst_peak": stage_chunks_per_s * FLOP_PER_CHUNK / 1e12 / float(peaks["bf16_tflops"]),
                                    "frac_of_bf16_sustained_peak": stage_chunks_per_s * FLOP_PER_CHUNK / 1e12 / peak_tf},
            "kernels": per_site,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    for wk in workers:
        wk.rec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
