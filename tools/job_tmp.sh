O=gpurun_out
timeout 900 python -m pytest tests/test_ar_gpu.py tests/test_conv_tma_gpu.py tests/test_transforms_gpu.py -m gpu -x -q > $O/t_ar_tma.log 2>&1; echo "tests rc=$?"
tail -8 $O/t_ar_tma.log | cut -c1-300
python tools/layer_times.py 32 5 > $O/lt_ew16.txt 2>&1; CAI_TMA_EPI_WARPS=8 python tools/layer_times.py 32 5 > $O/lt_ew8.txt 2>&1
grep -E "tma|total|stack" $O/lt_ew16.txt; echo ---; grep -E "tma|total|stack" $O/lt_ew8.txt
for ew in 16 8; do
  CAI_TMA_EPI_WARPS=$ew timeout 300 python bench.py --steps 12 --warmup 3 --no-variants --no-cpu-baseline --e2e-steps 5 > $O/ew_$ew.json 2> $O/ew_$ew.err
  python - <<PY
import json
d=json.loads([l for l in open("$O/ew_$ew.json") if l.startswith("{")][-1])
print("epi warps $ew value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "conv union ms", round(d["roofline"]["kernel_ms_per_step"],1), "frac", round(d["roofline"]["frac"],4))
PY
done
