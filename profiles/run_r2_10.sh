cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -20
python - <<'PY'
import pynvml, os
pynvml.nvmlInit()
n = pynvml.nvmlDeviceGetCount()
print("gpus", n, "cpus", os.cpu_count(), "affinity", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)))
for i in range(n):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    w = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count()+63)//64)
    cpus=[c for c in range(os.cpu_count()) if (w[c//64]>>(c%64))&1]
    print(i, "cpu affinity", cpus[0], "-", cpus[-1], len(cpus))
PY
numactl -H 2>/dev/null | head -8; lscpu | grep -i "numa\|socket\|model name" | head
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; tail -3 gpurun_out/r2_bench_2gpu.err; cat gpurun_out/r2_bench_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong > gpurun_out/r2_bench_2gpu_strong.json 2> gpurun_out/r2_bench_2gpu_strong.err; tail -3 gpurun_out/r2_bench_2gpu_strong.err; cat gpurun_out/r2_bench_2gpu_strong.json
