mkdir -p gpurun_out
for v in default w6_b2 w4_b3 w8_b1; do
  if [ $v = default ]; then unset TRACY_B200_LIB; else export TRACY_B200_LIB=$PWD/build_variants/lib_$v.so; fi
  echo "== $v"; PP_PROBE_QUICK=1 timeout 300 python profiles/pp_probe.py 2>&1 | grep "gcups\|checksum\|Error\|error" 
done | tee gpurun_out/r02_pp_variants.txt
