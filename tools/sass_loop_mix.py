"""Static instruction mix of the evaluation loops of an LM kernel: the SASS between each `DEPBAR.LE SB0, 0x1`
(cp.async.wait_group 1 at the top of the loop body) and the following `DEPBAR.LE SB0, 0x0` (after the loop).
usage: python tools/sass_loop_mix.py mdrp_b200/csrc/repose_lm.o '_ZN2rp14lm_warp_kernelILi1ELi9ELi1EEEvNS_6LMArgsE'"""
import re, subprocess, sys
from collections import Counter
obj, fun = sys.argv[1], sys.argv[2]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ins = [m.group(1) for m in (re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l) for l in sass.splitlines()) if m]
regions, cur = [], None
for t in ins:
    if "DEPBAR.LE SB0, 0x1" in t: cur = []
    elif "DEPBAR.LE SB0, 0x0" in t and cur is not None: regions.append(cur); cur = None
    elif cur is not None: cur.append(t)
for r in regions:
    c = Counter()
    for t in r:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        c[t.split()[0].split(".")[0]] += 1
    n = len(r)
    fp64 = sum(c[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
    print(f"loop body: {n} instr, FP64-pipe {fp64}, LDL {c['LDL']} STL {c['STL']} LDS {c['LDS']} | " + ", ".join(f"{k} {v}" for k, v in c.most_common(10)), flush=True)
