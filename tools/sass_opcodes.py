"""Opcode histogram of the shipped library, per kernel family: the evidence that the hot kernels are Blackwell-native
(UTCIMMA = tcgen05.mma.kind::i8, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy,
no HMMA / IMMA legacy tensor path).   python tools/sass_opcodes.py > profiles/sass_opcodes_r2.txt   (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "protoquant_b200", "libprotoquant_b200.so")
INTEREST = ("UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACMDFLUSH", "UBLKCP", "UBLKPF",
            "SYNCS", "HMMA", "IMMA", "QGMMA", "HGMMA", "IGMMA", "LDGSTS", "LDG", "STG", "LDS", "STS", "ATOM", "ATOMG", "RED", "MULTIMEM",
            "ACQBULK", "FENCE", "MEMBAR", "ERRBAR", "CCTL", "NANOSLEEP", "BAR", "WARPSYNC", "ELECT", "VIMNMX", "VIMNMX3", "PRMT", "FFMA", "FMUL", "FADD", "I2F", "F2F", "F2FP", "MUFU")


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    arch = set()
    for line in out.splitlines():
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("unnamed>::", "")
            fam = re.sub(r"<.*", "", name.split("(")[0]).split("::")[-1].split()[-1]
            cur = per.setdefault(fam, {"n": 0, "ops": collections.Counter()})
            cur["n"] += 1
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur["ops"][m.group(1)] += 1
            full = m.group(1) + m.group(2)
            if m.group(1) in ("UTCIMMA", "UTMALDG", "UTMASTG", "LDTM", "UBLKCP", "MULTIMEM", "UTCBAR"):
                cur["ops"][full] += 1
    print(f"# cuobjdump -sass protoquant_b200/libprotoquant_b200.so ; architectures: {sorted(arch)}")
    print("# kernel family : instantiations : opcode counts summed over the instantiations (selected opcodes; dotted = exact variants)")
    tot = collections.Counter()
    for fam, d in per.items():
        tot.update(d["ops"])
        sel = {k: v for k, v in d["ops"].items() if k.split(".")[0] in INTEREST}
        txt = "  ".join(f"{k}={v}" for k, v in sorted(sel.items(), key=lambda kv: (-kv[1], kv[0])))
        print(f"{fam} : {d['n']} : total_instructions={sum(v for k, v in d['ops'].items() if '.' not in k)}  {txt}")
    legacy = {k: tot[k] for k in ("HMMA", "IMMA", "HGMMA", "IGMMA", "QGMMA") if tot[k]}
    print(f"# legacy tensor opcodes in the whole library (HMMA / IMMA / *GMMA): {legacy or 'none'}")
    print(f"# tcgen05 / TMA totals: UTCIMMA={tot['UTCIMMA']} LDTM={tot['LDTM']} UTMALDG={tot['UTMALDG']} UTMASTG={tot['UTMASTG']} UBLKCP={tot['UBLKCP']} UTCBAR={tot['UTCBAR']} MULTIMEM={tot['MULTIMEM']}")


if __name__ == "__main__":
    sys.exit(main())
