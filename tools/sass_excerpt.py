"""SASS evidence for the hot kernels: `cuobjdump -sass` of libnoize_b200.so, per kernel
  * the instruction mix of the whole kernel and of its hottest loop (the largest backward-branch region),
  * the loop body itself (mnemonics + operands, encodings stripped),
so that claims such as "FFMA2/FMUL2/FADD2 in the hot loops", "LDGSTS (cp.async) row feeds" and "no UTMALDG (TMA)" can be
checked against the machine code instead of against prose.  Runs without a GPU.

    python tools/sass_excerpt.py [out_dir=profiles/sass]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "noize-job_b200", "libnoize_b200.so")
# (file stem, regex on the demangled name)
KERNELS = [
    ("fbm_simplex_pair_kernel", r"fbm_simplex_pair_kernel\("),
    ("fbm_psr_pair_kernel", r"fbm_psr_pair_kernel"),
    # the default filter kernel holds two bodies (border items first, interior CTAs after: MERGED) whose blocks ptxas
    # interleaves, so its largest backward-branch region spans both; the interior body alone is the unmerged instantiation
    ("sep_walk_kernel_R2_T4_merged", r"sep_walk_kernel<2, 4, false, 16, false, 1, true>"),
    ("sep_walk_kernel_R2_T4_skewed_interior", r"sep_walk_kernel<2, 4, false, 16, false, 1, false>"),
    ("sep_walk_kernel_R2_T3_skewed_interior", r"sep_walk_kernel<2, 3, false, 16, false, 1, false>"),
    ("sep_walk_kernel_R2_T4_chained_interior", r"sep_walk_kernel<2, 4, false, 16, false, 0, false>"),
    ("sep_ring_kernel_R2", r"sep_ring_kernel<2,"),
    ("flow_walk_kernel_I5", r"flow_walk_kernel<5,"),
    ("flow_group_kernel_I5_NW4", r"flow_group_kernel<5, 4,"),
    ("min_walk_kernel_5", r"min_walk_kernel<5>"),
    ("mesh_kernel_overshoot", r"mesh_kernel<1>"),
    ("thermal_tile_kernel", r"thermal_tile_kernel"),
]
INTERESTING = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "FMNMX", "LDS", "STS", "LDG", "STG", "LDGSTS", "LDGDEPBAR", "DEPBAR",
               "SHFL", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "BAR", "MUFU", "FRND", "F2I", "I2F", "IMAD", "IADD3", "LOP3", "LEA",
               "VIADD", "VIADDMNMX", "ISETP", "FSETP", "SEL", "FSEL", "PRMT", "MOV", "BRA", "CALL", "FCHK"]


def sass():
    out = subprocess.run(["cuobjdump", "-sass", LIB], check=True, stdout=subprocess.PIPE, text=True).stdout
    funcs, name, body = {}, None, []
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            if name:
                funcs[name] = body
            name, body = m.group(1), []
        elif name is not None:
            body.append(line)
    if name:
        funcs[name] = body
    names = list(funcs)
    dem = subprocess.run(["c++filt"], input="\n".join(names), check=True, stdout=subprocess.PIPE, text=True).stdout.splitlines()
    return {d: funcs[n] for n, d in zip(names, dem)}


INS = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*")


def parse(body):
    """[(address, text)] of the instructions of one function."""
    out = []
    for line in body:
        m = INS.search(line)
        if m:
            out.append((int(m.group(1), 16), re.sub(r"\s+", " ", m.group(2).strip())))
    return out


def mnemonic(text):
    t = text.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0] if t else ""


def loops(ins):
    """Backward-branch regions that are not nested in another one, largest first: [(first index, last index)]."""
    addr = {a: i for i, (a, _) in enumerate(ins)}
    regs = []
    for i, (a, t) in enumerate(ins):
        if mnemonic(t) == "BRA":
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) in addr and int(m.group(1), 16) <= a:
                regs.append((addr[int(m.group(1), 16)], i))
    outer = [r for r in regs if not any(o != r and o[0] <= r[0] and r[1] <= o[1] for o in regs)]
    return sorted(set(outer), key=lambda r: r[0] - r[1])


def hottest_loop(ins):
    ls = loops(ins)
    return ls[0] if ls else None


def mix(ins):
    c = collections.Counter(mnemonic(t) for _, t in ins)
    return c


def fmt_mix(c, total):
    keys = [k for k in INTERESTING if c.get(k)]
    rest = total - sum(c[k] for k in keys)
    return "  ".join(f"{k} {c[k]}" for k in keys) + f"  (other {rest})"


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass")
    os.makedirs(out_dir, exist_ok=True)
    funcs = sass()
    summary = ["# SASS summary of the hot kernels (tools/sass_excerpt.py; cuobjdump -sass libnoize_b200.so, sm_100a)", ""]
    for stem, pat in KERNELS:
        hits = [n for n in funcs if re.search(pat, n)]
        if not hits:
            continue
        name = hits[0]
        ins = parse(funcs[name])
        loop = hottest_loop(ins)
        whole = mix(ins)
        lines = [f"kernel: {name}", f"instructions: {len(ins)}", "whole kernel: " + fmt_mix(whole, len(ins))]
        body = ins
        if loop:
            body = ins[loop[0]:loop[1] + 1]
            lm = mix(body)
            lines.append(f"hottest loop: {len(body)} instructions at 0x{ins[loop[0]][0]:x}..0x{ins[loop[1]][0]:x}: " + fmt_mix(lm, len(body)))
        tma = sum(whole.get(k, 0) for k in ("UTMALDG", "UTMASTG", "UBLKCP"))
        lines.append(f"TMA instructions (UTMALDG/UTMASTG/UBLKCP): {tma};  cp.async (LDGSTS): {whole.get('LDGSTS', 0)};  "
                     f"packed FP32 (FFMA2+FMUL2+FADD2): {whole.get('FFMA2', 0) + whole.get('FMUL2', 0) + whole.get('FADD2', 0)}")
        summary += lines + [""]
        with open(os.path.join(out_dir, stem + ".txt"), "w") as f:
            f.write("\n".join(lines) + "\n\n--- loop body (address: instruction) ---\n")
            shown = body if len(body) <= 700 else body[:350] + [(-1, f"... {len(body) - 700} instructions elided ...")] + body[-350:]
            for a, t in shown:
                f.write((f"{a:05x}: " if a >= 0 else "       ") + t + "\n")
    with open(os.path.join(out_dir, "SUMMARY.md"), "w") as f:
        f.write("\n".join(summary) + "\n")
    print("\n".join(summary))


if __name__ == "__main__":
    main()
