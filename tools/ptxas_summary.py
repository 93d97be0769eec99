#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: registers / spills / smem per kernel (demangled template args)."""
import re, subprocess, sys

txt = open(sys.argv[1] if len(sys.argv) > 1 else "sctl_b200/csrc/ptxas_report.txt").read()
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:.*\n)*?ptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers(.*)")
rows = []
for m in pat.finditer(txt):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(.*", "", name).replace("sctl_cuda::", "").replace("void ", "")
    rows.append((name, int(m.group(5)), int(m.group(2)), int(m.group(3)), m.group(6).strip()))
for r in sorted(rows):
    print("%-90s regs=%3d stack=%d spill=%d %s" % r)
