#!/bin/bash
# registers / spills of the sweep and coop kernels: tools/ptxas_info.sh [extra nvcc flags]
cd "$(dirname "$0")/.."
nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -Xcompiler -fPIC -shared -Xptxas -v "$@" -o /tmp/libptx_$$.so nmpc_nav_control_b200/csrc/rti_kernels.cu 2> /tmp/ptxas_$$.log || { tail -20 /tmp/ptxas_$$.log; exit 1; }
python - /tmp/ptxas_$$.log <<'P'
import re,sys,subprocess
t=open(sys.argv[1]).read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?",t):
    n=m.group(1)
    if 'k_sweep' in n or 'k_ipm' in n or 'solo' in n:
        d=subprocess.run(['c++filt',n],capture_output=True,text=True).stdout.strip()
        d=re.sub(r'\(.*','',d)
        print(f"{d:60s} regs {m.group(5):>3s} stack {m.group(2):>5s} spill st/ld {m.group(3):>5s}/{m.group(4):>5s} smem {m.group(6)}")
P
rm -f /tmp/libptx_$$.so /tmp/ptxas_$$.log
