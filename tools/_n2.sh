TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time python -m pytest tests/test_gpu_spmd.py -x -q -m gpu ) > gpurun_out/r2m_spmd_pytest.log 2>&1; tail -8 gpurun_out/r2m_spmd_pytest.log
$TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-sub > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; python - <<P
import json
for line in open("gpurun_out/r2m_bench_n2.json"):
    if line.startswith("{"):
        d=json.loads(line); print(round(d["ms_per_step"],3), d["gpu_launches"], {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()}, d["gs_lanczos"]["seconds"], d["checksum"])
P
tail -3 gpurun_out/r2m_bench_n2.err
