TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for o in "xchg_split=1" "xchg_split=2"; do
$TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-sub --no-lanczos --opt $o > gpurun_out/r2p_n2_$o.json 2> gpurun_out/r2p_n2.err; python - <<P
import json
for line in open("gpurun_out/r2p_n2_$o.json"):
    if line.startswith("{"):
        d=json.loads(line); print("$o", round(d["ms_per_step"],3), d["gpu_launches"], {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()}, d["checksum"])
P
done
tail -3 gpurun_out/r2p_n2.err
