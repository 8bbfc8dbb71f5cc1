# strong-scaling sweep of bench.py on one box: N = 1 2 4 8 (as the driver launches it), one JSON line each
# usage: bash tools/scale_run.sh [extra bench.py flags, e.g. --no-e2e]
out=gpurun_out/r2_scaling_strong_2p28.jsonl; : > $out
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu "$@" >> $out 2>/dev/null
for N in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu "$@" >> $out 2> gpurun_out/scale_$N.err || tail -5 gpurun_out/scale_$N.err
done
python - <<'PY'
import json
rows=[json.loads(l) for l in open("gpurun_out/r2_scaling_strong_2p28.jsonl") if l.startswith("{")]
t1=rows[0]["ms_per_step"]
for j in rows:
    n=j["n_gpus"]
    e=j.get("e2e",{}).get("value",0)
    print("N=%d  %.3f ms/step  %.3e points/s  efficiency %.3f  stages %s  e2e %.3e  pool_sha16 %s nfev %d" % (
        n, j["ms_per_step"], j["value"], t1/(n*j["ms_per_step"]), {k: round(v,3) for k,v in j["stages_ms"].items() if k!="note"},
        e, j["pool_sha16"], j["nfev"]))
PY
