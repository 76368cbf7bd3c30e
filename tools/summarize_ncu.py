"""Turn the raw ncu outputs of one round into the committed summaries under profiles/r01/ (run in the build container).

  python tools/summarize_ncu.py launches gpurun_out/launches_v5.csv profiles/r01/ncu_launch_list_v5.csv
  python tools/summarize_ncu.py full gpurun_out/prof_conv6_v4.ncu-rep profiles/r01/conv6_ncu_v4.json
  python tools/summarize_ncu.py fullmd gpurun_out/prof_misc_v4.ncu-rep profiles/r01/ncu_full_misc_v4.md
"""
import csv
import io
import json
import subprocess
import sys

ONE_TIME = ["resize_h", "resize_vcoef", "resize_v_chunk", "conv1+pool1 (mma.sync)", "conv2 (gemm)", "pool2", "conv3 (gemm)",
            "conv4 (gemm)", "se3 fused + pool3", "conv5 (gemm)", "conv6 (gemm)", "se4 fused + pool4", "conv7 (gemm)",
            "se5 fused + final pool", "patch proj (gemm)",
            "enc0 qkv (gemm)", "enc0 attention (mma.sync)", "enc0 out_proj (gemm)", "enc0 ln1", "enc0 ffn1 (gemm)", "enc0 ffn2 (gemm)", "enc0 ln2",
            "enc1 qkv (gemm)", "enc1 attention (mma.sync)", "enc1 out_proj (gemm)", "enc1 ln1", "enc1 ffn1 (gemm)", "enc1 ffn2 (gemm)",
            "enc1 ln2+global_pos", "lstm in split3", "lstm in_proj (gemm, K = 1152)", "bilstm recurrence", "cross K/V split3",
            "cross K/V proj (gemm, K = 1152)"]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("kocr::", "").strip()


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    hdr = next(r for r in csv.reader(open(src)) if len(r) > 10 and r[0] == "ID")
    ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    one = rows[:len(ONE_TIME)]
    tot = sum(float(r[iv]) for r in one) / 1e3
    gemm = sum(float(r[iv]) for r in one if "gemm_tc" in r[ik]) / 1e3
    out = io.StringIO()
    out.write("# ncu launch list (gpu__time_duration.sum, --clock-control none) of ONE pass of the hot path, final round-1 kernels\n")
    out.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none -s 470 -c 126 --csv python tools/profile_step.py 256 24\n")
    out.write("# workload: c2 batch, 256 lines = 1885 chunks; plain launches (CUDA graphs off so every kernel shows by name); "
              "cold-cache, serialised: compare SHARES\n")
    out.write(f"# one-time stages (1-5a): {len(one)} launches, {tot / 1e3:.3f} ms; gemm_tc_kernel share = {gemm / tot:.3f}\n")
    out.write("index,site,kernel,us,share_of_one_time_stages,grid\n")
    for i, r in enumerate(one):
        us = float(r[iv]) / 1e3
        out.write(f'{i},{ONE_TIME[i]},{short(r[ik])},{us:.1f},{us / tot:.4f},"{r[ig]}"\n')
    dec = rows[len(ONE_TIME):]
    agg = {}
    for r in dec:
        k = short(r[ik])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv]) / 1e3
    dtot = sum(a[1] for a in agg.values())
    npos = sum(1 for r in dec if "dec_embed" in r[ik])
    out.write(f"\n# decode loop: {len(dec)} launches captured = {npos} positions x 25 kernels; {dtot / max(npos, 1):.0f} us per position (cold-cache, serialised)\n")
    out.write("kernel,launches,total_us,share_of_decode,avg_us\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"{k},{a[0]},{a[1]:.1f},{a[1] / dtot:.4f},{a[1] / a[0]:.1f}\n")
    open(dst, "w").write(out.getvalue())
    print(out.getvalue())


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
           "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__grid_size", "launch__block_size"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def full(rep, dst):
    hdr, units, data = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    r = data[0]
    def val(k):
        v, u = float(r[idx[k]]), units[idx[k]]
        return v, u
    def mb(k):
        v, u = val(k)
        return v * {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}[u]
    t, tu = val("gpu__time_duration.sum")
    t_ms = t * {"ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(tu, 1.0)
    out = {"kernel": r[idx["Kernel Name"]], "site": "conv6 (implicit GEMM, M = 1885*182, N = 512, K = 4608)", "chunks": 1885,
           "source": "ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 246 -c 1 python tools/profile_step.py 256 24",
           "duration_ms_under_ncu": t_ms, "dram_bytes_read": mb("dram__bytes_read.sum"), "dram_bytes_write": mb("dram__bytes_write.sum"),
           "dram_unit": "MB",
           "tensor_pipe_active_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")[0],
           "lts_throughput_pct": val("lts__throughput.avg.pct_of_peak_sustained_elapsed")[0],
           "l2_hit_rate_pct": val("lts__t_sector_hit_rate.pct")[0],
           "l2_to_sm_read_MB": mb("l1tex__m_xbar2l1tex_read_bytes.sum"),
           "registers_per_thread": val("launch__registers_per_thread")[0]}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


def fullmd(rep, dst):
    hdr, units, data = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["# ncu --set full summary (final round-1 kernels, c2 batch: 256 lines / 1885 chunks)", "",
             "command: `ncu --set full --clock-control none --import-source on -k \"regex:dec_cross_attn_flash|dec_self_attn|conv1_pool_mma|se_fused\" "
             "-s 46 -c 6 python tools/profile_step.py 256 24` (times are cold-cache and serialised)", ""]
    keys = [k for k in METRICS if k in idx]
    lines.append("| kernel | " + " | ".join(k.replace(".avg.pct_of_peak_sustained", "%").replace("_elapsed", "").replace("_active", "(act)") for k in keys) + " |")
    lines.append("|---|" + "---|" * len(keys))
    for r in data:
        cells = []
        for k in keys:
            v = r[idx[k]]
            try:
                cells.append(f"{float(v):.4g} {units[idx[k]]}".strip())
            except ValueError:
                cells.append(v)
        lines.append("| " + short(r[idx["Kernel Name"]]) + " | " + " | ".join(cells) + " |")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    {"launches": launches, "full": full, "fullmd": fullmd}[sys.argv[1]](sys.argv[2], sys.argv[3])
