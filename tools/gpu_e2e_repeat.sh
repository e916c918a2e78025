for i in 1 2 3; do TPT_BENCH_DEBUG=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/e2e_dbg_$i.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['seconds_per_step'])"; grep "e2e iter" gpurun_out/e2e_dbg_$i.err; done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --kernel mega 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mega value', round(d['value']), 'e2e', round(d['e2e']['value']))"
