"""Instruction mix and stall hot spots of one kernel from an `ncu --set full --import-source on` report (SASS view).
  python scripts/ncu_sass_mix.py <file.ncu-rep> <out.txt>
Classifies every SASS instruction (FP64 / shared ld-st / global ld-st / integer+address / control / other), sums executed
warp instructions and stall samples per class, and lists the instructions with the most stall samples."""
import collections
import csv
import io
import re
import subprocess
import sys


def classify(op):
    if op.startswith(("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")):
        return "fp64"
    if op.startswith(("LDS", "STS", "LDSM")):
        return "shared ld/st"
    if op.startswith(("LDG", "STG", "LD.", "ST.", "LDGSTS", "LDGDEPBAR", "DEPBAR", "RED", "ATOM")):
        return "global ld/st"
    if op.startswith(("LDC", "ULDC", "S2R", "S2UR", "CS2R")):
        return "const/special"
    if op.startswith(("BAR", "BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "NANOSLEEP", "CALL", "RET", "YIELD", "BREAK")):
        return "control/barrier"
    if op.startswith(("IMAD", "IADD", "LEA", "SHF", "LOP", "ISETP", "MOV", "UMOV", "SEL", "UIADD", "ULEA", "UIMAD", "USHF", "ULOP",
                      "IABS", "I2F", "F2I", "IMNMX", "PRMT", "R2UR", "UISETP", "PLOP", "USEL", "VIADD", "VIMNMX", "SGXT", "UFLO",
                      "FLO", "POPC", "BREV", "R2P", "P2R", "I2I", "F2F", "MUFU", "FMUL", "FFMA", "FADD", "I2FP", "UP2UR", "UPLOP3")):
        return "integer/address/move"
    return "other"


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = txt.splitlines()
    kname = lines[0]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[1:]))))
    mix = collections.OrderedDict()
    tot_i = tot_s = 0
    hot = []
    for r in rows:
        src = r["Source"].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        op = m.group(2) if m else "?"
        ni, ns = int(r["Instructions Executed"] or 0), int(r["# Samples"] or 0)
        c = classify(op)
        a = mix.setdefault(c, [0, 0])
        a[0] += ni
        a[1] += ns
        tot_i += ni
        tot_s += ns
        hot.append((ns, ni, src))
    with open(out, "w") as f:
        f.write(f"# {rep}\n# {kname[:200]}\n")
        f.write(f"# {len(rows)} SASS instructions, {tot_i} warp instructions executed, {tot_s} stall samples\n")
        f.write(f"{'class':24s} {'warp instr':>14s} {'share':>7s} {'samples':>10s} {'share':>7s}\n")
        for c, (ni, ns) in sorted(mix.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{c:24s} {ni:14d} {100.0 * ni / max(1, tot_i):6.1f}% {ns:10d} {100.0 * ns / max(1, tot_s):6.1f}%\n")
        f.write("\n# instructions with the most stall samples (samples, warp instr executed, SASS)\n")
        for ns, ni, src in sorted(hot, reverse=True)[:30]:
            f.write(f"{ns:8d} {ni:12d}  {src[:110]}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
