for s in 64 16 0; do for u in 409 450 490; do
  MGATK_OVERFLOW_SLACK=$s MGATK_UNIT_READS=$u python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('slack',$s,'unit_reads',$u,round(d['ms_per_step'],4),round(d['roofline']['stage_ms']['pileup'],4))"
done; done
