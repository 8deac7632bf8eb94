# Multi-GPU bench lines of the final code: bash tools/capture_multi.sh N "C3 C4 C5" (under gpurun --gpus N); outputs under gpurun_out/
set -u
N=$1; WL=${2:-C3}; O=gpurun_out; mkdir -p $O
if [ "$N" = "2" ]; then timeout 400 python -m pytest tests -m gpu -k two_rank -q 2>&1 | tail -5 > $O/r02f_two_rank_tests.log; fi
for w in $WL; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --workload $w --no-cpu-baseline --no-from-structs > $O/r02f_bench_${w}_${N}gpu.json 2> $O/r02f_bench_${w}_${N}gpu.err
done
echo done
