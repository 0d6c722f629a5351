#!/usr/bin/env python3
"""Symbolic dataflow view of floating-point SASS (test/analysis tooling, not product code).

Usage: cuobjdump -sass -fun <mangled> obj.o | python tools/sass_trace.py [start_addr end_addr]

Every FP-producing instruction gets an SSA name tN; operands are printed as the SSA name that
last wrote the register (linear scan, so only trustworthy inside straight-line regions).
Division / sqrt / rcp fast paths emitted by ptxas are collapsed to DIV()/SQRT()/RCP() so the
rounding-relevant structure (which mul/add pairs were fused into FFMA) can be read off directly.
Used to pin the as-compiled arithmetic of the reference build (oracle/_ref) -- see DESIGN.md.
"""
import re, sys

FP_OPS = ("FADD", "FMUL", "FFMA", "MUFU", "FMNMX", "FSEL", "I2FP", "I2F", "F2I", "F2F", "FSETP", "FCHK",
          "FRND", "TEX", "LDG", "LD", "LDL", "LDS", "LDC", "MOV", "IMAD.MOV", "SEL", "FSET", "SHFL", "LOP3", "SHF", "IADD3", "IMAD")

line_re = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);")

def main():
    lo = int(sys.argv[1], 16) if len(sys.argv) > 1 else 0
    hi = int(sys.argv[2], 16) if len(sys.argv) > 2 else 1 << 60
    verbose_int = "--int" in sys.argv
    reg = {}
    n = 0
    for raw in sys.stdin:
        m = line_re.search(raw)
        if not m:
            continue
        addr = int(m.group(1), 16)
        pred = (m.group(2) or "").strip()
        op = m.group(3)
        args = [a.strip() for a in m.group(4).split(",")] if m.group(4) else []
        base = op.split(".")[0]

        def res(a):
            a0 = a
            neg = a.startswith("-")
            a = a.lstrip("-")
            ab = a.startswith("|")
            a = a.strip("|")
            a = a.replace(".reuse", "")
            core = a.split(".")[0]
            s = reg.get(core, core) if re.fullmatch(r"U?R\d+", core) else a
            if a != core and re.fullmatch(r"U?R\d+", core):
                s += a[len(core):]
            if ab:
                s = "|" + s + "|"
            if neg:
                s = "-" + s
            return s

        if base in ("BRA", "CALL", "RET", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "BAR", "NOP", "ST", "STG", "STL", "STS"):
            if lo <= addr <= hi and base in ("BRA", "CALL", "RET", "EXIT", "STG", "ST", "STL"):
                print(f"{addr:04x} {pred:7s} {op} " + ", ".join(res(a) if i or base.startswith('ST') else a for i, a in enumerate(args)))
            continue
        if not args:
            continue
        dst = args[0]
        srcs = [res(a) for a in args[1:]]
        is_fp = base in ("FADD", "FMUL", "FFMA", "MUFU", "FMNMX", "FSEL", "I2FP", "I2F", "F2I", "F2F", "FRND", "TEX", "FSETP", "FCHK", "FSET")
        is_ld = base in ("LDG", "LD", "LDL", "LDS", "LDC", "ULDC", "TLD")
        if base in ("FSETP", "FCHK", "ISETP"):
            if lo <= addr <= hi and (base != "ISETP" or verbose_int):
                print(f"{addr:04x} {pred:7s} {op} {', '.join(args[:2])} <- " + ", ".join(res(a) for a in args[2:]))
            continue
        if is_fp or is_ld or verbose_int:
            n += 1
            name = f"t{n}"
            if lo <= addr <= hi:
                print(f"{addr:04x} {pred:7s} {name} = {op}(" + ", ".join(srcs) + f")   [{dst}]")
            if base == "TEX":
                # TEX.LL RZ, Rdst, Rcoord, ...  -> second arg is destination
                d2 = args[1].split(".")[0]
                reg[d2] = name
            elif re.fullmatch(r"U?R\d+", dst.split(".")[0]):
                if pred:
                    reg[dst] = f"({name}|{reg.get(dst, dst)})"
                else:
                    reg[dst] = name
            continue
        # other integer / move ops: track moves so names propagate
        if base in ("MOV", "IMAD") and op in ("MOV", "IMAD.MOV.U32", "IMAD.MOV"):
            src = args[-1]
            if re.fullmatch(r"U?R\d+", dst):
                val = res(src)
                reg[dst] = f"({val}|{reg.get(dst, dst)})" if pred else val
            continue
        if re.fullmatch(r"U?R\d+", dst.split(".")[0]):
            reg[dst.split(".")[0]] = f"{dst}@{addr:04x}"

if __name__ == "__main__":
    main()
