"""Static view of a kernel's SASS (no GPU needed): straight-line segments between control-flow instructions, with
their opcode class counts.  For an issue-bound kernel the instruction count of the steady-state segments is the
figure of merit.   usage: python scripts/sass_blocks.py <obj.o> <kernel-name-substring> [min_len]"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
body = next(f for f in funcs if pat in f.split("\n")[0])
ins = []
for line in body.split("\n"):
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        s = m.group(2).strip()
        mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
        ins.append((int(m.group(1), 16), mm.group(2)))
print("total instructions", len(ins))
CLS = [("FP64", r"^(DADD|DMUL|DFMA|DSETP)"), ("XU", r"^(F2F|MUFU|I2F|F2I)"), ("FP32", r"^(FMUL|FADD|FFMA|FSETP|FMNMX)"),
       ("SEL", r"^(SEL|FSEL)"), ("MOV", r"^(MOV|IMAD\.MOV|UMOV)"), ("SHFL", r"^SHFL"), ("LDST", r"^(LD|ST|ATOM|RED)"),
       ("INT", r".*")]
def cls(op):
    for n, r in CLS:
        if re.match(r, op):
            return n
seg = []
def flush():
    if len(seg) >= min_len:
        c = collections.Counter(cls(o) for _, o in seg)
        print(f"{seg[0][0]:#07x}-{seg[-1][0]:#07x} n={len(seg):4d} ", " ".join(f"{k}={c[k]}" for k, _ in CLS if c[k]))
for a, o in ins:
    seg.append((a, o))
    if re.match(r"^(BRA|EXIT|CALL|RET|BSYNC|BSSY|WARPSYNC\.ALL|BRX|JMP)", o):
        flush(); seg = []
flush()
