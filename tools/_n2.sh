set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests/test_gpu_spmd.py -x -q -m gpu > gpurun_out/r2g_spmd_pytest.log 2>&1; tail -15 gpurun_out/r2g_spmd_pytest.log
for o in "" "--opt xchg_chunks=1" "--opt xchg_chunks=8" "--opt use_ipc=2" "--no-ipc"; do
  $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e $o > gpurun_out/r2g_n2_$(echo $o | tr -d ' =-').json 2> gpurun_out/r2g_n2.err; python - <<P
import json,sys
d=json.load(open("gpurun_out/r2g_n2_$(echo $o | tr -d ' =-').json"))
print("$o", d["ms_per_step"], d["kernels"], d.get("gs_lanczos"))
P
done
