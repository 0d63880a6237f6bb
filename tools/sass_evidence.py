"""Per-kernel SASS instruction counts of the built library (cuobjdump -sass): the mnemonics that prove the tcgen05 / TMA /
multimem paths are what the .so contains.  usage: python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kokoro_ruslan_b200", "libkokoro_b200.so")
COLS = ["UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "LDGMC", "MUFU.EX2", "REDG", "ATOMG"]


def main() -> None:
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    rows: "OrderedDict[str, dict]" = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            rows[cur] = {c: 0 for c in COLS}
            rows[cur]["n"] = 0
            continue
        if cur is None or "/*" not in line:
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        rows[cur]["n"] += 1
        for c in COLS:
            if op == c or op.startswith(c + ".") or (c == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                rows[cur][c] += 1
    print("SASS evidence (cuobjdump -sass kokoro_ruslan_b200/libkokoro_b200.so, sm_100a): instruction counts per kernel")
    print("UTCHMMA = tcgen05.mma (kind::f16), UTMALDG = cp.async.bulk.tensor (TMA load), LDTM / STTM = tcgen05.ld / tcgen05.st,")
    print("UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, LDGMC = multimem.ld_reduce through the NVSwitch multicast address (multimem.st is")
    print("emitted as STG.E.128.STRONG.SYS on that address), MUFU.EX2 = ex2.approx, REDG / ATOMG = global reductions / atomics;")
    print("n = SASS instructions of the kernel\n")
    print(f"{'kernel':72s}" + "".join(f"{c:>10s}" for c in COLS) + f"{'n':>8s}")
    for k, r in sorted(rows.items(), key=lambda kv: re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]{8}\d+", "", kv[0])):
        name = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]{8}\d+", "", k)
        print(f"{name[:72]:72s}" + "".join(f"{r[c]:10d}" for c in COLS) + f"{r['n']:8d}")
    print(f"\n{len(rows)} kernels; {sum(1 for r in rows.values() if r['UTCHMMA'])} issue tcgen05.mma, "
          f"{sum(1 for r in rows.values() if r['UTMALDG'])} load through TMA, {sum(1 for r in rows.values() if r['LDGMC'])} use multimem")


if __name__ == "__main__":
    sys.exit(main())
