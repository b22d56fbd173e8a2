for u in 200 300 380 440 480; do
  MGATK_UNIT_READS=$u python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('unit_reads',$u,d['ms_per_step'],d['roofline']['stage_ms']['pileup'])"
done
