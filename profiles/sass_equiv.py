"""Instruction-by-instruction comparison of the scan kernels of two cubins (cuobjdump -sass), branch and return targets taken
relative to the instruction so that code that merely moved compares equal.  Used once, at the end of round 2, to show that the
templated read-2 scan (k_scan<MODE, OffT, DEFER, SHORT>) is the code of the two libraries that were measured and parity-tested
on the GPU (-DSSD_SCAN_SHORT_FAST=0 / =1 builds of commit 75545af): output in profiles/r02_short/sass_equivalence.txt.

    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -cubin -o /tmp/new.cubin csrc/ssd_b200.cu
    (the same from a checkout of 75545af, without and with -DSSD_SCAN_SHORT_FAST=1 -> /tmp/ref_base.cubin, /tmp/ref_v1.cubin)
    python profiles/sass_equiv.py [-v]
"""
import re, subprocess, sys, difflib
def sass(cubin, fn):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, cubin], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;?\s*(/\*.*\*/)?\s*$", line)
        if m and m.group(2) and not m.group(2).startswith("/*"):
            ins.append((int(m.group(1), 16), m.group(2)))
    # branch targets -> relative
    res = []
    for addr, t in ins:
        t2 = re.sub(r"0x([0-9a-f]+)", lambda mm: ("T%+d" % (int(mm.group(1), 16) - addr)) if re.search(r"\b(BRA|BSSY|CALL|BRX|JMP|WARPSYNC|BSYNC|BREAK|RET|LEPC)\b", t) else mm.group(0), t)
        res.append(t2)
    return res
def compare(a_cubin, a_fn, b_cubin, b_fn):
    A, B = sass(a_cubin, a_fn), sass(b_cubin, b_fn)
    sm = difflib.SequenceMatcher(None, A, B, autojunk=False)
    ops = [o for o in sm.get_opcodes() if o[0] != "equal"]
    nd = sum(max(o[2]-o[1], o[4]-o[3]) for o in ops)
    print("%-60s %5d vs %5d instructions, %d differing in %d places" % (b_fn[:60], len(A), len(B), nd, len(ops)))
    return A, B, ops
if __name__ == "__main__":
    pairs = []
    for off in "jm":
        for mode in (0, 1, 2):
            for d in (0, 1):
                old = "_ZN3ssd6k_scanILi%dE%sLb%dEEEvNS_8ScanArgsIT0_EENS_6DevCfgE" % (mode, off, d)
                new = "_ZN3ssd6k_scanILi%dE%sLb%dELb0EEEvNS_8ScanArgsIT0_EENS_6DevCfgE" % (mode, off, d)
                pairs.append(("/tmp/ref_base.cubin", old, "/tmp/new.cubin", new))
                if mode == 1:
                    new1 = "_ZN3ssd6k_scanILi%dE%sLb%dELb1EEEvNS_8ScanArgsIT0_EENS_6DevCfgE" % (mode, off, d)
                    pairs.append(("/tmp/ref_v1.cubin", old, "/tmp/new.cubin", new1))
    for p in pairs:
        A, B, ops = compare(*p)
        if "-v" in sys.argv:
            for tag, i1, i2, j1, j2 in ops[:12]:
                print("   ", tag, A[i1:i2][:6], "->", B[j1:j2][:6])
