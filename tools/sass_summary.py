"""SASS evidence for the tcgen05 / TMEM / TMA claims: per kernel of libkocr_b200.so, the count of the instruction
mnemonics that prove them (cuobjdump -sass; no GPU needed).    python tools/sass_summary.py > profiles/r02/sass_summary.md
  UTCHMMA / UTCQMMA...  tcgen05.mma (5th-gen tensor core, accumulators in TMEM)      UTMALDG   TMA tensor load (cp.async.bulk.tensor)
  UTCBAR               tcgen05.commit -> mbarrier                                    LDTM/STTM tcgen05.ld / tcgen05.st (TMEM <-> registers)
  HMMA                 warp-level mma.sync (conv1, per-chunk attention, SE FCs, BiLSTM recurrence)"""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "khmer_ocr_cnn_transformer_b200" / "libkocr_b200.so"
PAT = OrderedDict([("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("UTMALDG", r"\bUTMALDG"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"),
                   ("STTM", r"\bSTTM"), ("HMMA", r"\bHMMA"), ("SYNCS", r"\bSYNCS"), ("total", r"^\s+/\*[0-9a-f]{4,6}\*/")])


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    kernels, cur = OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {k: 0 for k in PAT}
            continue
        if cur is None:
            continue
        for k, p in PAT.items():
            if re.search(p, line):
                kernels[cur][k] += 1
    print("# SASS instruction counts per kernel of `libkocr_b200.so` (sm_100a)\n")
    print("`python tools/sass_summary.py` = `cuobjdump -sass khmer_ocr_cnn_transformer_b200/libkocr_b200.so`, counted per `Function :` block.")
    print("UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load (tiled and im2col), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st,")
    print("HMMA = mma.sync, SYNCS = mbarrier operations.\n")
    print("| kernel | " + " | ".join(PAT) + " |")
    print("|---|" + "---|" * len(PAT))
    tot = {k: 0 for k in PAT}
    for name, c in kernels.items():
        d = demangle(name)
        d = d.replace("(int)", "").replace("(bool)", "")
        d = re.sub(r"\(.*", "", d).replace("void ", "").replace("kocr::", "").replace("(anonymous namespace)::", "")
        print(f"| `{d}` | " + " | ".join(str(c[k]) for k in PAT) + " |")
        for k in PAT:
            tot[k] += c[k]
    print("| **all kernels** | " + " | ".join(str(tot[k]) for k in PAT) + " |")


if __name__ == "__main__":
    main()
