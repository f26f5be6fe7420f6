#!/usr/bin/env python
"""tools/sass_stats.py OBJ [filter] -- static SASS instruction mix per kernel instance (analysis tool).

Counts the instructions of each smfft_tile_kernel instance in an object file by class and prints them per
point of one tile iteration (straight-line unrolled body, so static count / points-per-thread is a good proxy
for issue slots per point).  Usage: python tools/sass_stats.py smfft_b200/lib/obj/inst_e10.o
"""
import collections
import re
import subprocess
import sys

CLASSES = [("fp", r"^(FFMA|FADD|FMUL|FFMA2|FADD2|FMUL2)"), ("lds", r"^LDS"), ("sts", r"^STS"), ("ldg", r"^LDG"),
           ("stg", r"^STG"), ("bar", r"^(BAR|SYNCS|WARPSYNC)"), ("tma", r"^(UTMA|UBLK)"), ("mufu", r"^MUFU"),
           ("mov", r"^(MOV|IMAD\.MOV|UMOV|SEL|FSEL|PRMT|SHFL)"), ("int", r"^(IMAD|IADD|LOP|SHF|LEA|ISETP|UIADD|ULOP|USHF|ULEA|UISETP|VIADD|UIMAD|R2UR|S2R|S2UR|CS2R)")]


def main():
    obj = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name, counts = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            counts[name]["total"] += 1
            for cls, pat in CLASSES:
                if re.match(pat, op):
                    counts[name][cls] += 1
                    break
            else:
                counts[name]["other"] += 1
    print(f"{'E B F D R TW':14s} mode io st reps | per point: total fp lds sts ldg stg int mov other")
    for name, c in counts.items():
        m = re.search(r"BlockCfgILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)E.*?EEELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(-?\d+|n\d+)E", name)
        if not m:
            continue
        e, b, f, d, r, tw, mode, io, st, reps, minb, pf = m.groups()
        key = f"{e} {b} {f} {d} {r} {tw}  m{mode} io{io} st{st} x{reps}"
        if flt and not re.search(flt, key):
            continue
        pts = 1 << int(b)
        print(f"{key:32s} | {c['total'] / pts:6.1f} " + " ".join(f"{c[k] / pts:5.1f}" for k in ("fp", "lds", "sts", "ldg", "stg", "int", "mov", "other")) + f"  bar={c['bar']} tma={c['tma']}")


main()
