O=gpurun_out
for k in conv deconv; do for d in 0 1 2 3 4; do echo "== $k debug=$d"; CAI_CONV_DEBUG=$d python tools/conv_probe.py 32 $k 2>&1 | tail -1; done; done > $O/ablate_r02.txt 2>&1
cat $O/ablate_r02.txt
