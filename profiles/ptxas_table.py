#!/usr/bin/env python
"""Registers / stack / spills of every kernel from the committed ptxas logs (sde_burst_coast_b200/csrc/*.ptxas.log).
usage: python profiles/ptxas_table.py [substring]"""
import glob, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                 r"(\d+) bytes spill loads\n.*?Used (\d+) registers")


def table(sub=None):
    rows = []
    for f in sorted(glob.glob(os.path.join(ROOT, 'sde_burst_coast_b200', 'csrc', '*.ptxas.log'))):
        for m in PAT.finditer(open(f).read()):
            dem = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            dem = dem.replace('(anonymous namespace)::', '')
            if 'cub::' in dem or (sub and sub not in dem):
                continue
            rows.append((os.path.basename(f)[:-10], dem.split('(')[0], int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4))))
    return rows


if __name__ == '__main__':
    for r in table(sys.argv[1] if len(sys.argv) > 1 else None):
        print('%-14s %-60s regs=%3d stack=%3d spill=%d/%d' % r)
