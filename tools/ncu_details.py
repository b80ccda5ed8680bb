"""condense `ncu -i X.ncu-rep --page details` to the lines the DESIGN numbers come from
    ncu -i gpurun_out/sweeps_full.ncu-rep --page details | python tools/ncu_details.py "header line" > profiles/....txt"""
import re
import sys

KEEP = re.compile(r"^\s{2}\S|Section:|Memory Throughput|DRAM Throughput|Duration|Compute \(SM\) Throughput|Executed Ipc|Issue Slots Busy|"
                  r"Mem Busy|Hit Rate|Mem Pipes Busy|No Eligible|Eligible Warps|Warp Cycles Per Issued|Avg. Active Threads|"
                  r"Registers Per Thread|Shared Memory Per Block|Waves Per SM|Block Limit|Occupancy|Branch Efficiency|"
                  r"Local Speedup|stalled waiting|This stall type|fused and|FP64|Local Memory|dram__bytes|uncoalesced|bank conflicts")

if __name__ == "__main__":
    print(sys.argv[1] if len(sys.argv) > 1 else "ncu details")
    for ln in sys.stdin:
        if KEEP.search(ln):
            print(ln.rstrip()[:200])
