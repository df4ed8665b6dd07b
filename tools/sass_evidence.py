"""profiles/rN_sass_evidence.txt: per kernel of libgdr.so the counts of the mnemonics that prove the design claims, and
an excerpt of the inner loop of both blend kernels so the placement of UBLKCP / SYNCS / FFMA2 can be read.

    python tools/sass_evidence.py > profiles/r2_sass_evidence.txt
"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "generativedensification_b200", "libgdr.so")
KEYS = ["UBLKCP", "UBLKRED", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU.EX2", "LDS.128", "LDG.128", "REDG", "RED", "ATOMG",
        "REDUX", "SHFL", "VOTE", "MATCH", "ACQBULK", "PREEXIT", "FENCE.VIEW.ASYNC"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((m.group(1), re.sub(r"\s+", " ", m.group(2)).strip()))


def short(name):
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\(anonymous namespace\)::", "", out).split("(")[0]


print("# SASS evidence from generativedensification_b200/libgdr.so (cuobjdump -sass)")
print("# UBLKCP = cp.async.bulk (TMA engine, 1-D; .G.S = shared->global), UBLKRED = cp.reduce.async.bulk (bulk add at the L2);")
print("# SYNCS = mbarrier ops; FFMA2/FMUL2/FADD2 = packed FP32 (sm_100); REDG/RED = red.global; ATOMG = atomics with return")
print("# (the slot claims of the binning); REDUX = warp-wide integer reduce; ACQBULK = griddepcontrol.wait, PREEXIT =")
print("# griddepcontrol.launch_dependents (programmatic dependent launch); FENCE.VIEW.ASYNC = fence.proxy.async")
print("\n## mnemonic counts per kernel")
for name, ins in sorted(funcs.items(), key=lambda kv: short(kv[0])):
    c = Counter()
    for _, t in ins:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                c[k] += 1
    print(f"{short(name):60s} {len(ins):5d} instr  " + " ".join(f"{k}={v}" for k, v in sorted(c.items())))


def excerpt(match, title, anchor, before, after):
    name = next(n for n in funcs if match in n)
    ins = funcs[name]
    idx = [i for i, (_, t) in enumerate(ins) if anchor in t]
    print(f"\n## {title}\n# {short(name)}: {len(ins)} instructions; lines around the first `{anchor}` that follows a UBLKCP")
    first_copy = next((i for i, (_, t) in enumerate(ins) if "UBLKCP" in t), 0)
    i0 = next((i for i in idx if i > first_copy), idx[0] if idx else 0)
    for a, t in ins[max(0, i0 - before):i0 + after]:
        print(f"  /*{a}*/ {t}")


# forward: the chunk-issue (UBLKCP + SYNCS.ARRIVE.TRANS64), the wait (SYNCS...TRYWAIT) and the packed pair evaluation
excerpt("blend_forward_kernel", "forward blend: ring issue / wait and the head of a (warp, record) visit", "FFMA2", 60, 50)
excerpt("blend_backward_kernelILb1", "backward blend (all gradients): head of a (warp, record) visit", "FFMA2", 50, 60)
