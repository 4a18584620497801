"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): instruction
mix by opcode and the hottest address ranges."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ci, si, sa, ti = (h.index('Instructions Executed'), h.index('Source'),
                  h.index('# Samples'), h.index('Thread Instructions Executed'))
ins = []
for r in rows[2:]:
    if len(r) <= ci:
        continue
    ins.append((r[si].strip(), float(r[ci] or 0), float(r[sa] or 0), float(r[ti] or 0)))
tot = sum(i[1] for i in ins); tots = sum(i[2] for i in ins)
print("static instr %d  executed %.4g  samples %.0f  avg threads %.1f" %
      (len(ins), tot, tots, sum(i[3] for i in ins) / tot))
op = collections.defaultdict(lambda: [0, 0])
for s, c, smp, _ in ins:
    t = s.split()
    k = t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '?')
    k = k.split('.')[0] + ('.' + k.split('.')[1] if k.startswith(('LDS', 'STS', 'LDG', 'ATOMS', 'MUFU', 'LDL', 'STL', 'SHFL', 'BAR', 'WARPSYNC')) and '.' in k else '')
    op[k][0] += c; op[k][1] += smp
print("--- opcode mix (share of executed / share of stall samples)")
for k, (c, smp) in sorted(op.items(), key=lambda x: -x[1][0])[:28]:
    print("  %-16s %5.1f%%  %5.1f%%" % (k, c / tot * 100, smp / tots * 100))
# hot ranges: split into runs of instructions with similar execution count
print("--- address ranges (run of >=8 instr with count within 2x)")
i = 0
runs = []
while i < len(ins):
    j = i; c0 = max(ins[i][1], 1)
    while j + 1 < len(ins) and 0.5 < max(ins[j + 1][1], 1) / c0 < 2:
        j += 1
    runs.append((i, j, sum(x[1] for x in ins[i:j + 1]), sum(x[2] for x in ins[i:j + 1])))
    i = j + 1
for a, b, c, smp in sorted(runs, key=lambda x: -x[2])[:14]:
    sample = "; ".join(x[0].split()[0] if not x[0].startswith('@') else x[0].split()[1] for x in ins[a:min(b + 1, a + 10)])
    print("  [%5d..%5d] n=%4d exec/instr %.3g  inst %5.1f%% samp %5.1f%%  thr %.1f | %s" %
          (a, b, b - a + 1, c / (b - a + 1), c / tot * 100, smp / tots * 100,
           sum(x[3] for x in ins[a:b + 1]) / max(c, 1), sample))
