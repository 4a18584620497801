"""Static SASS statistics of one kernel of a built library (no GPU needed):

    python tools/sass_stats.py /tmp/librdfk_x.so pair_prune_kernelILi1ELb0 [--dump out.sass]

prints registers / spills (cuobjdump --dump-resource-usage), the instruction
count and the opcode histogram; with --dump the cleaned listing is written out.
"""
import collections
import re
import subprocess
import sys


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib],
                         capture_output=True, text=True).stdout.splitlines()
    for i, l in enumerate(res):
        if "Function" in l and pat in l:
            print(l.strip()[:100])
            print(res[i + 1].strip())
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True,
                          text=True).stdout.splitlines()
    on = False
    ops = collections.Counter()
    lines = []
    for l in sass:
        if "Function :" in l:
            on = pat in l
            continue
        if not on:
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if not m:
            continue
        ins = m.group(2).strip()
        lines.append("/*%s*/ %s" % (m.group(1), ins))
        t = ins.split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0] if not op.startswith(("LDS", "STS", "MUFU")) else op] += 1
    print("instructions:", len(lines))
    try:
        print(" ".join("%s=%d" % kv for kv in ops.most_common(40)))
    except BrokenPipeError:
        pass
    if dump:
        open(dump, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
