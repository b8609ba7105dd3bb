"""Writes profiles/r02_sass_hot_loops.txt: per-kernel SASS instruction census of libsktt_b200.so (cuobjdump -sass) and the
densest DMMA window of the hot kernels, as evidence of what the tensor / TMA / barrier instruction mix really is.
    python tools/sass_excerpt.py"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "scikit_tt_b200", "lib", "libsktt_b200.so")
OUT = os.path.join(ROOT, "profiles", "r02_sass_hot_loops.txt")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs, cur, name = {}, None, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = funcs.setdefault(name, [])
        continue
    if cur is not None and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        cur.append(line.rstrip())
KEYS = ["DMMA", "DFMA", "UBLKCP", "UTMALDG", "SYNCS", "LDS", "LDG", "STG", "BAR", "LDGSTS", "ARRIVES", "REDUX", "ATOM", "RED", "SHFL", "MEMBAR", "UCGABAR", "CCTL"]
HOT = ["pcg_persistent_kernel", "stack_nat_kernel<false>", "lu_fused_kernel(", "stack_persistent_kernel", "gemm_dmma_kernel<128, 128", "gemm_dmma_kernel<64, 64", "mv_stage23_kernel",
       "mv_stage1_kernel", "beig_kernel<cplx>", "bsvd_kernel", "cholqr_kernel<double", "lu_panel_cluster_kernel<double>"]
def op(line):
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    return m.group(1) if m else ""
with open(OUT, "w") as f:
    f.write("# SASS census of scikit_tt_b200/lib/libsktt_b200.so (sm_100a), cuobjdump -sass; produced by tools/sass_excerpt.py\n")
    f.write("# columns: instructions per kernel by mnemonic prefix.  fp64 has no tcgen05 kind: the tensor path is DMMA.8x8x4;\n")
    f.write("# UBLKCP = 1-D TMA bulk copy (cp.async.bulk), SYNCS = mbarrier ops, LDGSTS = cp.async, UCGABAR = cluster barrier.\n\n")
    f.write(f"{'kernel':70s} {'total':>7s} " + " ".join(f"{k:>7s}" for k in KEYS) + "\n")
    tot = collections.Counter()
    for name, lines in sorted(funcs.items(), key=lambda kv: -len(kv[1])):
        cnt = collections.Counter()
        for l in lines:
            o = op(l)
            for k in KEYS:
                if o.startswith(k):
                    cnt[k] += 1
        tot.update(cnt)
        if any(h in name for h in HOT) or cnt["DMMA"] or cnt["UBLKCP"]:
            f.write(f"{name[:70]:70s} {len(lines):7d} " + " ".join(f"{cnt[k]:7d}" for k in KEYS) + "\n")
    f.write(f"{'ALL KERNELS':70s} {sum(len(v) for v in funcs.values()):7d} " + " ".join(f"{tot[k]:7d}" for k in KEYS) + "\n")
    for h in HOT[:4]:
        for name, lines in funcs.items():
            if h in name:
                idx = [i for i, l in enumerate(lines) if op(l).startswith("DMMA")]
                if not idx:
                    continue
                # densest window of 48 instructions
                best, bi = -1, 0
                for s in range(0, max(1, len(lines) - 48)):
                    c = sum(1 for i in idx if s <= i < s + 48)
                    if c > best:
                        best, bi = c, s
                f.write(f"\n## {name}\n## densest 48-instruction window: {best} DMMA ({len(idx)} DMMA in the kernel)\n")
                for l in lines[bi:bi + 48]:
                    f.write(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l) + "\n")
                break
print("wrote", OUT)
