#!/usr/bin/env python
"""SASS mnemonic summary of the built library (no GPU needed): proves which instruction families the
kernels really contain (B200_PROFILING.md: tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM, TMA = UTMALDG /
UTMASTG / UBLKCP).

    python tools/sass_summary.py profiles/rNN_sass_mnemonics.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "assembled_cnn_b200", "libacnn.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR",
             "UTCATOMSWS", "SYNCS", "HMMA", "HGMMA", "REDG", "ATOMG", "FHADD", "FHFMA", "HSET2"]
NOTES = {
    "REDG": "fp32 `red.global.add`: the split-K epilogues of wgrad and of the small SK / SE GEMMs",
    "ATOMG": "the arrival counter of the SGD kernel's ordered L2 sum",
    "FHADD": "`add.f32.bf16`: fp32 accumulator += bf16 operand (conv epilogue add / statistics)",
    "FHFMA": "`fma.rn.f32.bf16`: the sum of squares of the batch-norm statistics",
    "HSET2": "`set.gt.u32.bf16x2`: the ReLU mask of the dgrad epilogue on packed halves",
}


def main(dst):
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    total = collections.Counter()
    fam = collections.defaultdict(collections.Counter)
    inst = collections.Counter()
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"<.*", "", name).replace("void ", "").replace("acnn::", "")
            cur = re.sub(r"\(.*", "", cur)
            inst[cur] += 1
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and cur:
            op = m.group(1)
            for k in MNEMONICS:
                if op == k or op.startswith(k):
                    total[k] += 1
                    fam[cur][k] += 1
                    break
    with open(dst, "w") as fh:
        fh.write("# SASS mnemonic summary of assembled_cnn_b200/libacnn.so (sm_100a)\n\n")
        fh.write("`python tools/sass_summary.py` = `cuobjdump -sass assembled_cnn_b200/libacnn.so`, instruction "
                 "counts over all kernel instantiations (B200_PROFILING.md: `tcgen05.mma` = `UTC*MMA`, "
                 "`tcgen05.ld` = `LDTM`, TMA = `UTMALDG` / `UTMASTG` / `UBLKCP`; no legacy `HMMA` / Hopper "
                 "`HGMMA`).\n\n| mnemonic | total |\n|---|---:|\n")
        for k in MNEMONICS:
            note = " (%s)" % NOTES[k] if k in NOTES else ""
            fh.write("| `%s`%s | %d |\n" % (k, note, total[k]))
        cols = [k for k in MNEMONICS if total[k]]
        fh.write("\n| kernel family | instantiations | " + " | ".join("`%s`" % c for c in cols) + " |\n")
        fh.write("|---|---:|" + "---:|" * len(cols) + "\n")
        for name, c in sorted(fam.items(), key=lambda kv: -sum(kv[1].values())):
            fh.write("| `%s` | %d | " % (name, inst[name]) + " | ".join(str(c[k]) for k in cols) + " |\n")
    print("wrote", dst, dict(total))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_mnemonics.md"))
