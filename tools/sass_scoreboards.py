"""Offline look at the scoreboards of a gemm_tc_kernel variant (no GPU needed).

ptxas gives every global load of the epilogue warps ONE hardware scoreboard (a counter): an instruction that waits on
it waits for everything in flight, including a register prefetch issued just before, and constant-bank reads (LDC /
LDCU) can land on the same scoreboard. This script decodes the control bits of `cuobjdump -sass` (write barrier =
bits 110-112, wait mask = bits 116-121 of each 128-bit instruction) and lists, per matching variant, every non-load
instruction that waits on a scoreboard used by global loads, every non-load instruction that signals one, and the
LDTM positions (= chunk boundaries). It is how the out-projection epilogue was analysed in profiles/r01_summary.md
section 3.   Usage: python tools/sass_scoreboards.py [lib.so | file.o] [substring of the mangled kernel name,
e.g. attn_tc_bwd_dq_kernel; default: the K-major CTA-pair store-epilogue GEMM]"""
import re
import subprocess
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "neurovit_b200/libneurovit_b200.so"
want = sys.argv[2] if len(sys.argv) > 2 else "ILi256ELi6ELb0ELb0ELi2ELi0ELb0EE"
text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout.split("\n")
starts = [i for i, l in enumerate(text) if "Function :" in l]
for si, st in enumerate(starts):
    if want not in text[st]:
        continue
    body = text[st:(starts[si + 1] if si + 1 < len(starts) else len(text))]
    ins, i = [], 0
    while i < len(body) - 1:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", body[i])
        m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", body[i + 1]) if m else None
        if m and m2:
            hi = int(m2.group(1), 16)
            ins.append((m.group(1), m.group(2), (hi >> 46) & 7, (hi >> 52) & 0x3F))   # write barrier, wait mask
            i += 2
        else:
            i += 1
    ldg_sb = {wb for _, t, wb, _ in ins if re.match(r"(@!?U?P\d+\s+)?(LDG|LD)\.E", t) and wb != 7}
    print(text[st].strip()[:150])
    print("  instructions", len(ins), " scoreboards used by global loads:", sorted(ldg_sb))
    for k, (addr, t, wb, wait) in enumerate(ins):
        is_ld = re.match(r"(@!?U?P\d+\s+)?(LDG|LD)\.E", t) is not None
        waits = any(wait & (1 << sb) for sb in ldg_sb)
        if (waits and not is_ld) or (wb in ldg_sb and not is_ld) or "LDTM" in t:
            tag = "WAITS" if waits else ("signals" if wb in ldg_sb else "")
            print(f"  {k:5d} {addr} {t[:70]:70s} wb{wb if wb != 7 else '-'} wait{wait:06b} {tag}")
