"""Write one gzip'ed `cuobjdump -sass` listing per live kernel class (J+K variant) of
libjoltqc_b200.so into profiles/sass/r2/, plus an index with registers / stack / opcode counts.
The routing mirrors the engine: brick_shape().fits -> jk_brick_kernel, measured table ->
jk_bwarp_kernel, else jk_warp_kernel (li <= 3) / jk_1q1t_kernel_large.
usage: dump_sass.py [lmax=3]"""
import gzip, os, re, subprocess, sys
from collections import Counter
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "joltqc_b200", "libjoltqc_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass", "r2")
lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 3
nf = lambda l: (l + 1) * (l + 2) // 2
sel = open(os.path.join(ROOT, "joltqc_b200", "csrc", "jk_class_select.h")).read()
bw = set(re.findall(r"case (\d+): return true", sel))

def brick_fits(li, lj, lk, ll):
    n = nf(li) * nf(lj) * nf(lk) * nf(ll)
    nki, njkl = nf(li) * (nf(lk) + nf(ll)), nf(lk) * nf(ll)
    nroots = (li + lj + lk + ll) // 2 + 1
    live = n + nki + njkl
    acc, di = live > 80 and nki > 12, nki <= 48
    dlk = 0 if njkl <= 3 else (1 if njkl <= 9 else 2)
    slots = (nki if acc else 0) + (nki if di else 0) + (njkl if dlk == 1 else 0)
    regs = 128 if live <= 12 else (168 if live <= 36 else 255)   # JQC_BRICK_LIVE128 / JQC_BRICK_LIVE168
    minb = 65536 // (regs * 128)
    smem = nroots * (14 + 2 * nroots) * 15 * 16 + 4 * 32 * slots * 8
    return n <= 108 and smem * minb <= 216 * 1024

names = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = re.findall(r"Function : (\S+)", names)
os.makedirs(OUT, exist_ok=True)
index = ["class,kernel,registers,stack_bytes,sass_lines,DFMA,DMUL,DADD,LDS,STS,LDG,REDG,SHFL,file,fp32_twin_registers,fp32_twin_stack_bytes"]
res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
usage = {m[0]: (m[1], m[2]) for m in re.findall(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+)", res)}
blocks = names.split("\t\tFunction : ")
body = {b.split("\n", 1)[0].strip(): b for b in blocks[1:]}
for li in range(lmax + 1):
    for lj in range(li + 1):
        for lk in range(li + 1):
            for ll in range(lk + 1):
                key = "%d%d%d%d" % (li, lj, lk, ll)
                if brick_fits(li, lj, lk, ll):
                    pat = "jk_brick_kernelIdLi%dELi%dELi%dELi%dELb1ELb1EEE" % (li, lj, lk, ll)
                elif key in bw:
                    pat = "jk_bwarp_kernelIdLi%dELi%dELi%dELi%dELb1ELb1E" % (li, lj, lk, ll)
                else:
                    pat = "jk_warp_kernelILi%dELi%dELi%dELi%dELb1ELb1E" % (li, lj, lk, ll)
                f = [x for x in funcs if pat in x]
                if not f:
                    f = [x for x in funcs if ("jk_1q1t_kernel_smallILi%dELi%dELi%dELi%dELb1ELb1E" % (li, lj, lk, ll)) in x]
                if not f:
                    continue
                fn = f[0]
                txt = body[fn]
                ops = Counter(m.group(1).split(".")[0] for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", txt))
                fname = "%s_%s.sass.gz" % (key, re.search(r"jk_[a-z0-9_]+kernel(_small)?", fn).group(0))
                with gzip.open(os.path.join(OUT, fname), "wt") as g:
                    g.write("\t\tFunction : " + txt)
                reg, stack = usage.get(fn, ("?", "?"))
                index.append(",".join(["(%d%d|%d%d)" % (li, lj, lk, ll), re.search(r"jk_[a-z0-9_]+kernel(_small)?", fn).group(0), reg, stack,
                                       str(txt.count("\n"))] + [str(ops[o]) for o in ("DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "REDG", "SHFL")] + [fname] +
                                      list(usage.get(fn.replace("kernelId", "kernelIf"), ("", "")) if "kernelId" in fn else ("", "")) ))
open(os.path.join(OUT, "INDEX.csv"), "w").write("\n".join(index) + "\n")
print(len(index) - 1, "listings written to", OUT)
