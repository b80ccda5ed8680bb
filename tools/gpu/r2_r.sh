# round 2: phase cycles of the block-per-instance kernel (a -DNMPC_SOLO_PROF build)
NMPC_B200_LIB=$PWD/build/var/lib_soloprof.so timeout 120 python tools/solo_prof.py ${1:-diff} 2>&1 | tail -26
