#!/usr/bin/env python
"""Static instruction mix of the RK step loop of lens_seg_kernel<false> for a set of -D flags (no GPU needed):

    python profiles/sass_mix.py [-DFLAG ...]

Compiles a stub that instantiates only that kernel to a cubin, disassembles it and reports the innermost
loop that contains the MUFU.RSQ64H seeds: instructions per trip by opcode, split into the part executed on
the common path and the out-of-line fallback block."""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "centrex-molecule-trajectories_b200" / "csrc"
STUB = """#include "cmt_kernels.cuh"
void *cmt_stub_keep() { return (void *)cmt::lens_seg_kernel<%s, %s>; }
"""


def main():
    flags = [a for a in sys.argv[1:] if a.startswith("-D") or a.startswith("-maxrreg")]
    contract = "true" if "--contracted" in sys.argv else "false"
    with tempfile.TemporaryDirectory() as td:
        src, cubin = Path(td) / "stub.cu", Path(td) / "stub.cubin"
        copies = next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("--copies=")), "8")
        src.write_text(STUB % (contract, copies))
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-cubin", "-I", str(CSRC),
               "-I", str(ROOT / "include"), "-Xptxas", "-v", *flags, "-o", str(cubin), str(src)]
        out = subprocess.run(cmd, capture_output=True, text=True)
        if out.returncode:
            print(out.stderr)
            return 1
        for line in out.stderr.splitlines():
            if "registers" in line or "spill" in line:
                print(line.strip())
        sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True).stdout
    ins = []
    inside = False
    for l in sass.splitlines():
        if "Function :" in l:
            inside = "lens_seg_kernel" in l
            continue
        if not inside:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    if "--dump" in sys.argv:
        for a, t in ins:
            print(hex(a), t)
    # loops = backward branches; pick the smallest one that holds >= 4 MUFU.RSQ64H
    best = None
    for a, t in ins:
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                lo = int(m.group(1), 16)
                body = [(x, y) for x, y in ins if lo <= x <= a]
                if sum("MUFU.RSQ64H" in y for _, y in body) >= 4 and (best is None or len(body) < len(best)):
                    best = body
    if best is None:
        print("no step loop found")
        return 1
    # the fallback block is the stretch around CALL that the common path jumps over
    call = [i for i, (_, t) in enumerate(best) if t.startswith("CALL") or " CALL" in t]
    skip = set()
    for i, (a, t) in enumerate(best):
        m = re.search(r"BRA (?:P\d, )?0x([0-9a-f]+)", t)
        if m and t.startswith("@") and call:
            tgt = int(m.group(1), 16)
            if a < best[call[0]][0] < tgt:
                skip = {x for x, _ in best if a < x < tgt}
    mix, mix_fb = collections.Counter(), collections.Counter()
    for a, t in best:
        t = re.sub(r"^@!?U?P\d\s+", "", t)
        op = t.split()[0].split(".")[0]
        (mix_fb if a in skip else mix)[op] += 1
    fp64 = sum(v for k, v in mix.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(f"loop {hex(best[0][0])}..{hex(best[-1][0])}: {sum(mix.values())} instructions on the common path "
          f"({fp64} DFMA/DMUL/DADD/DSETP), {sum(mix_fb.values())} in the fallback block")
    print("  " + ", ".join(f"{k} {v}" for k, v in mix.most_common()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
