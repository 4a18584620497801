"""registers / spills / stack of every kernel in rdfk.cu (nvcc -Xptxas -v);
extra arguments are passed to nvcc (e.g. -DRDFK_QS_TRI=48)"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
       "-std=c++17", "--fmad=false", "-Xptxas", "-v", "-I",
       os.path.join(ROOT, "include"), "-c",
       os.path.join(ROOT, "pmda_b200", "csrc", "rdfk.cu"), "-o", "/dev/null"] + sys.argv[1:]
out = subprocess.run(cmd, stderr=subprocess.PIPE, stdout=subprocess.PIPE, text=True).stderr
name = None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        stack = m.groups()
        continue
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        print("%-60s regs %3s  stack %4s  spill st %4s ld %4s" % (name[:60], m.group(1), *stack))
        name = None
