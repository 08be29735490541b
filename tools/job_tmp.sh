O=gpurun_out
timeout 900 python -m pytest tests/test_conv_tma_gpu.py tests/test_transforms_gpu.py -m gpu -x -q > $O/t_teams.log 2>&1; echo "tests rc=$?"
tail -3 $O/t_teams.log | cut -c1-300
python tools/layer_times.py 32 5 > $O/lt_teams2b.txt 2>&1
grep -E "tma|total|stack" $O/lt_teams2b.txt
