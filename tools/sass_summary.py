"""Per-kernel counts of the Blackwell-native SASS mnemonics in nbmf_b200/libnbmf_cuda.so (profiles/*_sass_*.txt).
    python tools/sass_summary.py > profiles/r2_sass_tp_pass.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "nbmf_b200", "libnbmf_cuda.so")
WANT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "UTCATOMSWS", "USETMAXREG", "SYNCS", "DMMA", "FFMA2", "FMUL2",
        "FADD2", "FSET", "MUFU", "ATOMS", "REDG", "STL", "LDL"]
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
kern = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        name = m.group(1)
        kern[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        kern[name]["instr"] += 1
        op = m.group(1)
        if op in WANT:
            kern[name][op] += 1
dem = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass nbmf_b200/libnbmf_cuda.so (sm_100a), built from the committed sources (tools/sass_summary.py).")
print("# Per kernel: instruction count and the counts of the Blackwell-native mnemonics -- UTCHMMA = tcgen05.mma kind::f16,")
print("# UTCQMMA = tcgen05.mma kind::f8f6f4, LDTM / STTM = tcgen05.ld / tcgen05.st, UTMALDG = cp.async.bulk.tensor (TMA),")
print("# UTCBAR = tcgen05.commit -> mbarrier, USETMAXREG = setmaxnreg, DMMA = mma.sync f64, FFMA2 / FMUL2 / FADD2 = packed fp32,")
print("# FSET = set.le.f32 (1.0 / 0.0), ATOMS = shared-memory atomic, STL / LDL = register spills.")
print("# tp_pass_kernel<KP, ROWS, WANT_LOG, HAS_NA, USE_LO, LO8>; the bench times <128,1,1,0,0,0> + <128,0,0,0,0,0> (hi only) or")
print("# <128,1,1,0,1,1> + <128,0,0,0,1,1> (hi + e5m2 lo).\n")
for n, d in sorted(zip(kern.values(), dem), key=lambda kv: kv[1]):
    short = re.sub(r"\(.*", "", d).replace("void ", "").replace("(bool)", "").replace("(int)", "")
    if not any(n[w] for w in WANT):
        continue
    print(f"{short:62s} instr {n['instr']:6d}  " + "  ".join(f"{w} {n[w]}" for w in WANT if n[w]))
