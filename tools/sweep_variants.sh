# usage: bash tools/sweep_variants.sh  -- times bench.py with each prebuilt libmgatk2_b200_*.so variant
cd mgatk2_b200
cp libmgatk2_b200.so /tmp/lib_default.so
for v in ${VARIANTS:-default}; do
  if [ $v = default ]; then cp /tmp/lib_default.so libmgatk2_b200.so; else cp libmgatk2_b200_$v.so libmgatk2_b200.so; fi
  touch libmgatk2_b200.so
  (cd .. && python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$v',d['ms_per_step'],d['roofline']['stage_ms'])")
done
cp /tmp/lib_default.so libmgatk2_b200.so
