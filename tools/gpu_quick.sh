set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=r1g
for K in 20 200; do
timeout 150 python bench.py --steps $K --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_k$K.json 2> $OUT/${TAG}_bench_k$K.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('$OUT/${TAG}_bench_k$K.json'));print($K, d['value'],d['ms_per_step'],d['e2e'],d['clocks'])"
done
