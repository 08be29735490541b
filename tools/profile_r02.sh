#!/bin/bash
# Round-2 ncu evidence for profiles/ (run on the GPU box through gpurun; every ncu command runs only after the same
# command line exited 0 without ncu).  Outputs land in gpurun_out/; tools/ncu_summary.py / launch_list_summary.py reduce
# them to the files committed under profiles/.
set -u
O=gpurun_out
mkdir -p $O
BENCH_ARGS="--steps 1 --warmup 3 --no-cpu-baseline --no-variants --inflight 1 --e2e-steps 1"
# 1. launch list of the headline workload (G = (64, 256): escape-heavy y strings)
python bench.py $BENCH_ARGS > $O/r02_ll_plain.json 2> $O/r02_ll_plain.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12000 --csv \
    --log-file $O/r02_launches.csv python bench.py $BENCH_ARGS > $O/r02_ll_ncu.json 2> $O/r02_ll_ncu.err
echo "launch list rc=$?"
# 2. transform kernels, --set full: 128->128 k5 s2 + GDN (conv_gemm_kernel), first layer + GDN (conv_tma_kernel), deconv + IGDN
for kind in conv conv1 deconv; do
  python tools/conv_probe.py 32 $kind > $O/r02_probe_$kind.txt 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"conv_gemm|conv_tma" --launch-skip 4 -c 1 -f \
      -o $O/r02_$kind python tools/conv_probe.py 32 $kind > $O/r02_ncu_$kind.log 2>&1
  echo "$kind rc=$?"
done
# 3. coder on the headline stream (one micro-batch: 32 strings x 294,912 symbols, 42 rows, 35 % escapes)
CODER="--B 32 --n 294912 --max-idx 42 --indep-sigma 3.3 --no-quant --iters 1"
python tools/coder_microbench.py $CODER > $O/r02_coder_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"rans_(en|de)code" --launch-skip 1 -c 2 -f \
    -o $O/r02_rans_headline python tools/coder_microbench.py $CODER > $O/r02_ncu_coder.log 2>&1
echo "coder rc=$?"
ls -la $O | tail -20
