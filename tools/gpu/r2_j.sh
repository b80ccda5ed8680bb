# round 2, call J: factorising-sweep variants (build/var/lib_v*.so) A/B on the three models + parity of the default build
mkdir -p gpurun_out
for v in v0 v1 v2 v3; do
  for m in diff tric omni4; do
    echo -n "$v $m: "; NMPC_B200_LIB=$PWD/build/var/lib_$v.so timeout 200 python tools/prof_k3.py 65536 $m 4 2>&1 | tail -1
  done
done > gpurun_out/j_variants.log 2>&1
cat gpurun_out/j_variants.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_schedules.py -x -q 2>&1 | tail -5 > gpurun_out/j_tests.log; cat gpurun_out/j_tests.log
