"""Per-kernel Blackwell opcode census of a `cuobjdump -sass` listing.   cuobjdump -sass _C.so | python tools/sass_summary.py"""
import collections
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "SYNCS", "MULTIMEM",
        "LDGMC", "STGMC", "REDGMC", "ACQBULK", "UCGABAR", "HMMA", "LDGSTS", "REDUX", "ERRBAR", "CCTL"]


def main():
    fn, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in sys.stdin:
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            total[fn] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[fn][k] += 1
    names = {}
    try:
        dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.split("\n")
        names = dict(zip(counts, dem))
    except Exception:
        pass
    print(f"{'SASS instr':>10}  kernel : Blackwell-specific opcodes (static counts)")
    for fn, c in counts.items():
        nm = re.sub(r"\(.*$", "", names.get(fn, fn)).replace("void ", "").replace("nxdi::", "")
        print(f"{total[fn]:>10}  {nm} : " + (", ".join(f"{k} x{v}" for k, v in c.items()) or "-"))


if __name__ == "__main__":
    main()
