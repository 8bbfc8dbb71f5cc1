# bench.py at N GPUs for a list of "exchange:chunks:warps" settings
N=$1; shift
for cfg in "$@"; do IFS=: read ex ch w nf <<< "$cfg"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $N --no-cpu --no-e2e --steps 20 --exchange $ex --chunks ${ch:-1} --phase1-warps ${w:-0} ${nf:+--no-filter} > gpurun_out/r2_b${N}_$cfg.json 2> gpurun_out/r2_b${N}_$cfg.err; python -c "
import json,sys
j=json.loads(open('gpurun_out/r2_b${N}_$cfg.json').read().strip().splitlines()[-1]); print('N=$N $cfg', round(j['ms_per_step'],4), 'eval %.3f star %.3f pool %.3f' % (j['stages_ms']['eval'], j['stages_ms']['star'], j['stages_ms']['pool']), j['pool_sha16'], j['nfev'])
" || tail -5 gpurun_out/r2_b${N}_$cfg.err; done
