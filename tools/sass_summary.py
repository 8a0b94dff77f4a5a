"""Per-kernel SASS evidence of the Blackwell-native path: counts of tcgen05 MMAs (UTC*MMA), TMEM loads / stores (LDTM / STTM),
bulk async copies (UBLKCP = cp.async.bulk, the TMA engine), mbarrier ops (SYNCS) and legacy tensor instructions (HMMA).

  python tools/sass_summary.py [nerf_keras_b200/libnerf_b200.so] > profiles/rN_sass_summary.txt"""
import collections
import re
import subprocess
import sys

PATTERNS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "RED.E", "ATOMG",
            "MUFU.EX2", "F2FP", "BAR.SYNC", "LDG", "STG", "LDS", "STS"]


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*(REG:\d+ STACK:\d+ SHARED:\d+[^\n]*)", res):
        usage[m.group(1)] = m.group(2)
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_instr"] += 1
        for p in PATTERNS:
            if op.startswith(p):
                kernels[cur][p] += 1
                if p == "UTCHMMA" and ".2CTA" in op:
                    kernels[cur]["UTCHMMA.2CTA"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS summary of {path} (cuobjdump -sass, sm_100a); instruction counts are static occurrences per kernel")
    tot = collections.Counter()
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", dn)[:110]
        keys = [k for k in PATTERNS + ["UTCHMMA.2CTA"] if c[k]]
        print(f"{short}\n    instr {c['_instr']:6d} | " + " ".join(f"{k}={c[k]}" for k in keys) + (f"\n    {usage[name]}" if name in usage else ""))
        tot.update(c)
    print("\nTOTAL " + " ".join(f"{k}={tot[k]}" for k in PATTERNS + ["UTCHMMA.2CTA"] if tot[k]))
    print("tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk -> UBLKCP; HMMA (legacy mma.sync) must be 0.")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "nerf_keras_b200/libnerf_b200.so")
