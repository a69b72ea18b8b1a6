"""Count SASS instructions per issue pipe for one kernel of a cubin / .so (offline cost model).
usage: python tools/sass/pipecount.py file.cubin [kernel-name-substring]
Pipe classes follow B300_MICROARCH.md (fma: IMAD/FFMA..., alu: IADD3/LOP3/SHF/SEL/ISETP/MOV/PRMT/IMNMX)."""
import re, subprocess, sys, collections
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IDP")
ALU = ("IADD", "LOP3", "SHF", "SEL", "ISETP", "MOV", "PRMT", "IMNMX", "VIMNMX", "LEA", "IABS", "FMNMX", "FSEL", "PLOP3", "SGXT", "BMSK", "VIADD", "FSETP", "P2R", "R2P", "POPC", "FLO", "CS2R")
def classify(op):
    base = op.split(".")[0]
    if base == "IMAD":
        if "WIDE" in op: return "fma_wide"
        if ".HI" in op: return "fma_hi"
        return "fma"
    if base.startswith(FMA): return "fma"
    if base.startswith(("LDG", "STG", "LDS", "STS", "LD", "ST", "ATOM", "RED", "LDC")): return "lsu"
    if base.startswith(("BAR", "BRA", "EXIT", "BSSY", "BSYNC", "CALL", "RET", "WARPSYNC", "NOP", "YIELD")): return "ctl"
    if base.startswith(("S2R", "S2UR", "SHFL", "MUFU", "I2F", "F2I")): return "xu"
    if base.startswith("U") or base.startswith("R2UR"): return "uniform"
    if base.startswith(ALU): return "alu"
    return "other:" + base
def main():
    f = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["cuobjdump", "-sass", f], capture_output=True, text=True).stdout
    cur = None; counts = {}
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m: cur = m.group(1); counts[cur] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur: counts[cur][classify(m.group(1))] += 1
    for k, c in counts.items():
        if pat in k:
            tot = sum(c.values())
            print(k[:110]); print("   total %d  " % tot + "  ".join("%s=%d" % kv for kv in sorted(c.items())))
main()
