cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m 2>&1 | head -12 | cut -c1-160
for sc in weak strong; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 --scaling $sc > gpurun_out/r2_bench_${N}gpu_$sc.json 2> gpurun_out/r2_bench_${N}gpu_$sc.err; tail -2 gpurun_out/r2_bench_${N}gpu_$sc.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${N}gpu_$sc.json"))
print("$sc", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "rand", d["random_order"]["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "agg GB/s", d["e2e"]["all_ranks_h2d_plus_d2h_gbs"], d["e2e"]["host_binding"])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 5 --warmup 3 --config c5 > gpurun_out/r2_bench_${N}gpu_c5.json 2> gpurun_out/r2_bench_${N}gpu_c5.err; tail -2 gpurun_out/r2_bench_${N}gpu_c5.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${N}gpu_c5.json"))
print("c5", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "rand", d["random_order"]["value"], "e2e", d["e2e"]["value"])
PY
