"""Turn the raw ncu outputs of round 2 (tools/profile_r02.sh, run on the GPU box) into the committed summaries:
    python tools/summarize_r02.py        -> profiles/r02/ncu_full_all_kernels.md, ncu_launch_list.csv, conv6_ncu.json"""
import csv
import json
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles" / "r02"
SITES = ["resize_h", "resize_vcoef", "resize_v_chunk", "conv1+pool1 (mma.sync)", "conv2 (im2col gemm)", "pool2", "conv3 (im2col gemm)",
         "conv4 (im2col gemm, pool + SE means epilogue)", "se3 excite", "conv5 (im2col gemm)", "conv6 (im2col gemm, pool + SE means epilogue)",
         "se4 excite", "conv7 (im2col gemm, row-bin + SE means epilogue)", "se5 excite + final pool", "patch proj (gemm)",
         "enc0 qkv (gemm)", "enc0 attention (mma.sync)", "enc0 out_proj (gemm + residual)", "enc0 ln1", "enc0 ffn1 (gemm)", "enc0 ffn2 (gemm + residual)", "enc0 ln2",
         "enc1 qkv (gemm)", "enc1 attention (mma.sync)", "enc1 out_proj (gemm + residual)", "enc1 ln1", "enc1 ffn1 (gemm)", "enc1 ffn2 (gemm + residual)",
         "enc1 ln2 + global_pos", "lstm in split3", "lstm in_proj (gemm, K = 1152)", "bilstm recurrence", "cross K/V split3",
         "cross K/V proj (gemm, K = 1152)"]


def short(name):
    n = name.replace("void ", "").replace("kocr::", "").replace("(anonymous namespace)::", "")
    n = n.replace("(int)", "").replace("(bool)", "")
    return re.sub(r"\(.*", "", n).strip()


def full_table():
    rows = list(csv.reader(open(ROOT / "gpurun_out" / "prof_full_raw.csv")))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    f = lambda r, k: float(r[col[k]].replace(",", "")) if r[col[k]] not in ("", "n/a") else float("nan")
    seen, lines, conv6 = {}, [], None
    lines.append("# `ncu --set full --clock-control none` of ONE pass over the c2 batch (256 lines, 1885 chunks): one row per distinct kernel configuration\n")
    lines.append("Command: `ncu --set full --clock-control none --profile-from-start off -c 100 python tools/profile_step.py 256 3` "
                 "(tools/profile_r02.sh; CUDA graphs off so every kernel shows by name; cold-cache, serialised). "
                 "First launch of every (kernel, grid) pair of the pass; `site` = position in the launch sequence.\n")
    lines.append("| # | site | kernel | grid | time us | tensor pipe % | DRAM read MB | DRAM write MB | DRAM % | L2 (lts) % | L2 hit % | L2->SM GB | SM % | warps active % | issue active % | regs | top stalls (warps per issue) |")
    lines.append("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    idx = 0
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        name = short(r[col["Kernel Name"]])
        site = SITES[idx] if idx < len(SITES) else "decode position"
        key = (name, r[col["launch__grid_size"]], site if idx < len(SITES) else "")
        idx += 1
        if key in seen:
            continue
        seen[key] = True
        st = sorted(((f(r, h), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stall), reverse=True)[:3]
        row = {"site": site, "kernel": name, "grid": r[col["launch__grid_size"]], "us": f(r, "gpu__time_duration.sum") * 1e3,
               "tensor": f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
               "rd": f(r, "dram__bytes_read.sum"), "wr": f(r, "dram__bytes_write.sum"),
               "dram": f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "lts": f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
               "hit": f(r, "lts__t_sector_hit_rate.pct"), "xbar": f(r, "l1tex__m_xbar2l1tex_read_bytes.sum"),
               "sm": f(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"), "warps": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
               "issue": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), "regs": r[col["launch__registers_per_thread"]]}
        # unit fix-ups: ncu picks units per column (Mbyte / Gbyte)
        lines.append(f"| {idx - 1} | {site} | `{name}` | {row['grid']} | {row['us']:.1f} | {row['tensor']:.1f} | {row['rd']:.1f} | {row['wr']:.1f} | "
                     f"{row['dram']:.1f} | {row['lts']:.1f} | {row['hit']:.1f} | {row['xbar']:.2f} | {row['sm']:.1f} | {row['warps']:.1f} | "
                     f"{row['issue']:.1f} | {row['regs']} | " + ", ".join(f"{n} {v:.1f}" for v, n in st) + " |")
        if site.startswith("conv6"):
            conv6 = row
    units = rows[1]
    lines.append(f"\nUnits as reported by ncu: DRAM read {units[col['dram__bytes_read.sum']]}, write {units[col['dram__bytes_write.sum']]}, "
                 f"L2->SM {units[col['l1tex__m_xbar2l1tex_read_bytes.sum']]}, time {units[col['gpu__time_duration.sum']]} (x1000 -> us).")
    (OUT / "ncu_full_all_kernels.md").write_text("\n".join(lines) + "\n")
    if conv6:
        unit_r = units[col["dram__bytes_read.sum"]]
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[unit_r]
        scale_w = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[col["dram__bytes_write.sum"]]]
        json.dump({"kernel": "gemm_tc_kernel<256, a16, column-fused> @ conv6", "chunks": 1885, "time_us": conv6["us"],
                   "dram_bytes_read": conv6["rd"] * scale, "dram_bytes_write": conv6["wr"] * scale_w,
                   "tensor_pipe_active_pct": conv6["tensor"], "lts_throughput_pct": conv6["lts"], "l2_hit_rate_pct": conv6["hit"],
                   "source": "profiles/r02/ncu_full_all_kernels.md (ncu --set full --clock-control none, tools/profile_r02.sh)"},
                  open(OUT / "conv6_ncu.json", "w"), indent=1)
    print("\n".join(lines[:60]))


def launch_list():
    rows = [r for r in csv.reader(open(ROOT / "gpurun_out" / "launches_r02.csv")) if len(r) > 10]
    hdr = next(r for r in rows if r[0] == "ID")
    data = [r for r in rows if r[0].isdigit()]
    ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    us = lambda r: float(r[iv].replace(",", "")) / 1e3
    one = data[:len(SITES)]
    tot = sum(us(r) for r in one)
    gemm = sum(us(r) for r in one if "gemm_tc" in r[ik])
    stage = [r for i, r in enumerate(one) if 3 <= i <= 28]
    stot = sum(us(r) for r in stage)
    out = ["# ncu launch list (gpu__time_duration.sum, --clock-control none) of ONE pass of the hot path, round-2 kernels",
           "# command: ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python tools/profile_step.py 256 3",
           "# workload: c2 batch, 256 lines = 1885 chunks; plain launches (CUDA graphs off so every kernel shows by name); cold-cache, serialised: compare SHARES",
           f"# one-time stages (1-5a): {len(one)} launches, {tot / 1e3:.3f} ms; gemm_tc_kernel share = {gemm / tot:.3f}; stage 2-4 (launches 3-28) = {stot / 1e3:.3f} ms",
           "index,site,kernel,us,share_of_one_time_stages,share_of_stage_2_4,grid"]
    for i, r in enumerate(one):
        out.append(f'{i},{SITES[i]},{short(r[ik])},{us(r):.1f},{us(r) / tot:.4f},{(us(r) / stot if 3 <= i <= 28 else 0):.4f},"{r[ig]}"')
    dec = data[len(SITES):]
    agg = {}
    for r in dec:
        a = agg.setdefault(short(r[ik]), [0, 0.0])
        a[0] += 1
        a[1] += us(r)
    dtot = sum(a[1] for a in agg.values())
    npos = sum(1 for r in dec if "dec_argmax" in r[ik] or "dec_out_argmax" in r[ik])
    out.append(f"\n# decode loop: {len(dec)} launches captured = {npos} positions; {dtot / max(npos, 1):.0f} us per position (cold-cache, serialised)")
    out.append("kernel,launches,total_us,share_of_decode,avg_us")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k},{a[0]},{a[1]:.1f},{a[1] / dtot:.4f},{a[1] / a[0]:.1f}")
    (OUT / "ncu_launch_list.csv").write_text("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    launch_list()
    full_table()
