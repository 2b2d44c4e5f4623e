#!/usr/bin/env python
"""Disassemble ALOHA 96-bit R-type HE instruction words (24 hex digits per line).

Encoding restated from the reference decoder src/vp/sequncer/expander.v:65-107,123-130."""
import sys

F6 = {0x04: "VSETVL", 0x08: "VSETQ", 0x0c: "VSETIQ", 0x10: "BREAK", 0x00: "NOP", 0x01: "VFQMUL",
      0x05: "VFQADD", 0x09: "VFQSUB", 0x0d: "VFQMOD", 0x11: "VCPY", 0x15: "VAUT", 0x19: "VROLI",
      0x02: "VNTT", 0x06: "VINTT", 0x03: "VLE", 0x07: "VSE"}
F3 = {0: ".vv", 1: ".vs", 2: ".sv"}


def disasm(word: str) -> str:
    inst, imm = int(word[:8], 16), int(word[8:], 16)
    f6, vs2, vs1, f3, vd = inst >> 26, (inst >> 20) & 31, (inst >> 15) & 31, (inst >> 12) & 7, (inst >> 7) & 31
    name = F6.get(f6, f"?{f6:02x}")
    if name in ("VSETVL", "VSETQ", "VSETIQ"):
        return f"{name} 0x{imm:x}"
    if name in ("BREAK", "NOP"):
        return name
    if name == "VLE":
        return f"VLE v{vd} <- base{imm >> 48}[row {(imm >> 10) & 0xffff}]"
    if name == "VSE":
        return f"VSE base{imm >> 48}[row {(imm >> 10) & 0xffff}] <- v{vs1}"
    if name in ("VFQMUL", "VFQADD", "VFQSUB"):
        if f3 == 0:
            return f"{name}.vv v{vd} <- v{vs1}, v{vs2}"
        if f3 == 1:
            return f"{name}.vs v{vd} <- v{vs1}, 0x{imm:x}"
        return f"{name}.sv v{vd} <- 0x{imm:x}, v{vs2}"
    if name in ("VAUT", "VROLI"):
        return f"{name} v{vd} <- v{vs1}, imm={imm}"
    return f"{name} v{vd} <- v{vs1}"


if __name__ == "__main__":
    for i, line in enumerate(open(sys.argv[1]).read().split()):
        print(f"{i:4d}  {line}  {disasm(line)}")
