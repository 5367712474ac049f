"""SASS mnemonic counts of the built library (cuobjdump -sass): the evidence lines of B200_PROFILING.md (UTCHMMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor loads / stores, ...), whole library and per kernel.
Usage: python scripts/sass_counts.py > profiles/rNN_sass_counts.txt"""
import re
import subprocess
import sys

sys.path.insert(0, ".")
from cacophony_b200 import _lib as L

sass = subprocess.run(["cuobjdump", "-sass", L.LIB_PATH], capture_output=True, text=True).stdout
WANT = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDGSTS", "FFMA2", "FADD2", "FMNMX3", "MUFU.EX2",
        "MUFU.TANH", "MUFU.SQRT", "MUFU.LG2", "F2FP.SATFINITE", "USETMAXREG", "RED.E", "SYNCS"]
print("# SASS mnemonic counts of cacophony_b200/libcaco_b200.so (cuobjdump -sass), final code of the round")
print("# built from the committed sources with: python -m cacophony_b200.build --force  (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)")
for w in WANT:
    print(f"{w:16s} {len(re.findall(r'\b' + re.escape(w), sass))}")
print("\n# per kernel: UTCHMMA / LDTM / UTMALDG+UBLKCP (kernels with none of them omitted)")
kernels = re.split(r"\n\s*Function : ", sass)[1:]
for k in sorted(kernels, key=lambda t: t.split("\n", 1)[0]):
    name = k.split("\n", 1)[0].strip()
    a, b, c = len(re.findall(r"\bUTCHMMA", k)), len(re.findall(r"\bLDTM", k)), len(re.findall(r"\bUTMALDG|\bUBLKCP", k))
    if a or b or c:
        print(f"{name.replace('_ZN4caco', '')} UTCHMMA={a} LDTM={b} TMA={c}")
print(f"\n# kernels in the library:\n{len(kernels)}")
for k in kernels:
    if "l2norm_scatter_kernel" in k.split("\n", 1)[0]:
        print("# l2norm_scatter_kernel (the fused exchange: rows stored to every peer's mapping, system-scope fence, release-stored flags):")
        print(f"  ST.E.128 (row stores to generic = peer addresses, unrolled over peers)={len(re.findall(r'ST\.E\.128', k))}  "
              f"STG.E.STRONG.SYS (flag release stores)={len(re.findall(r'STG\.E\.STRONG\.SYS', k))}  "
              f"MEMBAR.ALL.SYS={len(re.findall(r'MEMBAR\.ALL\.SYS', k))}  MEMBAR.SC.SYS (__threadfence_system)={len(re.findall(r'MEMBAR\.SC\.SYS', k))}  "
              f"ATOMG.E.ADD (block ticket)={len(re.findall(r'ATOMG\.E\.ADD', k))}")
